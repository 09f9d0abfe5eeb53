// Host side of the implicit-GEMM engine: tensor-map construction (driver entry point resolved at run time, so the
// library does not link libcuda), tile-shape selection and launch of either engine.
#include "gemm.cuh"
#include "conv_pair.cuh"
#include "lin_mc.cuh"

#include <cudaTypedefs.h>
#include <stdio.h>
#include <stdlib.h>

namespace dexb {

// ------------------------------------------------------------------------------------------------
// driver entry point
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

static int resolve_encode() {
  if (g_encode != nullptr) return 0;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_last_error("cuTensorMapEncodeTiled not available: %s", cudaGetErrorString(e));
    return -2;
  }
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return 0;
}

static int encode_bf16(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_b,
                       const cuuint32_t* box, const cuuint32_t* estr, const char* what) {
  DEXB_TRY(resolve_encode());
  DEXB_CHECK((reinterpret_cast<uintptr_t>(base) & 15) == 0, "tensor map %s: base not 16 B aligned", what);
  for (int i = 0; i + 1 < rank; ++i)
    DEXB_CHECK(strides_b[i] % 16 == 0, "tensor map %s: stride %d (%llu B) not a multiple of 16", what, i,
               (unsigned long long)strides_b[i]);
  CUresult r = g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_b,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DEXB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(%s) failed with CUresult %d", what, (int)r);
  return 0;
}

// ------------------------------------------------------------------------------------------------
// plan
// ------------------------------------------------------------------------------------------------
static int pick_block_n(int N) {
  static int bn_max = -1;
  if (bn_max < 0) {
    const char* e = getenv("DEXB_BN_MAX");
    bn_max = (e != nullptr) ? atoi(e) : 128;   // 128: three 64 KiB stages hide the TMA latency; 256 only has room for two
  }
  if (N <= 32) return 32;
  if (N <= 64) return 64;
  if (N <= 128 || bn_max <= 128) return 128;
  if (N % 256 == 0 || N > 512) return 256;
  return 128;
}

int gemm_plan_init(GemmPlan* gp, const GemmParams& p, int n_img_a, long b_rows, int n_bmat) {
  gp->p = p;
  GemmParams& q = gp->p;
  DEXB_CHECK(q.BH * q.BW == 128, "gemm: tile %d x %d is not 128 pixels", q.BH, q.BW);
  DEXB_CHECK(q.K % 16 == 0 && q.K > 0, "gemm: K = %d must be a positive multiple of 16", q.K);
  DEXB_CHECK(q.nheads >= 1 && q.nz >= 1 && q.nz % q.nheads == 0, "gemm: nz %d / nheads %d", q.nz, q.nheads);
  q.TH = cdiv(q.CH, q.BH);
  q.TW = cdiv(q.CW, q.BW);
  gp->block_n = (q.block_n_hint == 64 || q.block_n_hint == 128) ? q.block_n_hint : pick_block_n(q.N);
  gp->n_img_a = n_img_a;
  gp->tc_ok = (q.K % kTcBlockK == 0) && (q.a_row_stride % 8 == 0) && (q.b_row_stride % 8 == 0) &&
              (q.in_stride == 1 || q.in_stride == 2) && (q.BW * q.in_stride <= 256) && (q.BH * q.in_stride <= 256) &&
              ((q.a_hi | q.a_lo | q.b_hi | q.b_lo | q.a_head_stride | q.b_head_stride) % 8 == 0);
  if (!gp->tc_ok) return 0;
  {
    // halo mode (gemm.cuh): 3x3 stride-1 convolution on 1 x 128 pixel tiles, shared weights, stacked-N tile widths
    static int halo_mode = -1;
    if (halo_mode < 0) {
      const char* e = getenv("DEXB_HALO");
      halo_mode = (e != nullptr) ? atoi(e) : 1;
    }
    gp->halo = halo_mode != 0 && q.KH == 3 && q.KW == 3 && q.offH == -1 && q.offW == -1 && q.tap_sw == 1 && q.in_stride == 1 &&
               q.out_scale == 1 && q.BW == 128 && q.BH == 1 && q.nheads == 1 && q.b_mode == 0 && q.nsplit == 3 && q.dbg == 0 &&
               gp->block_n <= 128 && q.N % gp->block_n == 0;
  }
  {
    const cuuint64_t dims[4] = {(cuuint64_t)q.a_row_stride, (cuuint64_t)q.W, (cuuint64_t)q.H, (cuuint64_t)n_img_a};
    const cuuint64_t str[3] = {(cuuint64_t)q.a_row_stride * 2, (cuuint64_t)q.a_row_stride * 2 * q.W,
                               (cuuint64_t)q.a_row_stride * 2 * q.W * q.H};
    const cuuint32_t box[4] = {(cuuint32_t)kTcBlockK, (cuuint32_t)(gp->halo ? kTcHaloRows : q.BW * q.in_stride),
                               (cuuint32_t)(q.BH * q.in_stride), 1};
    const cuuint32_t es[4] = {1, (cuuint32_t)q.in_stride, (cuuint32_t)q.in_stride, 1};
    DEXB_TRY(encode_bf16(&gp->tmA, q.A, 4, dims, str, box, es, "A"));
  }
  {
    const cuuint64_t dims[3] = {(cuuint64_t)q.b_row_stride, (cuuint64_t)b_rows, (cuuint64_t)n_bmat};
    const cuuint64_t str[2] = {(cuuint64_t)q.b_row_stride * 2,
                               (cuuint64_t)(n_bmat > 1 ? q.b_mat_stride : q.b_row_stride * b_rows) * 2};
    const cuuint32_t box[3] = {(cuuint32_t)kTcBlockK, (cuuint32_t)gp->block_n, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    DEXB_TRY(encode_bf16(&gp->tmB, q.Bw, 3, dims, str, box, es, "B"));
    if (gp->halo) {
      const cuuint32_t boxh[3] = {(cuuint32_t)kTcBlockK, (cuuint32_t)(gp->block_n / 2), 1};
      DEXB_TRY(encode_bf16(&gp->tmBh, q.Bw, 3, dims, str, boxh, es, "B half"));
    }
  }
  return 0;
}

// ------------------------------------------------------------------------------------------------
// launch
// ------------------------------------------------------------------------------------------------
static int g_num_sms = 148;
static long g_generic_epilogues = 0;          // launches that could not take the FAST epilogue (diagnostic)
long gemm_generic_epilogue_launches() { return g_generic_epilogues; }

int gemm_global_init() {
  int dev = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<32, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<64, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<128, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<256, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<64, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<128, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<256, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<64, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<128, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(lin_mc_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(lin_mc_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, kTcSmemMax));
  DEXB_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, cp_smem_bytes(64)));
  DEXB_CUDA_OK(cudaFuncSetAttribute(conv_pair_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, cp_smem_bytes(128)));
  return resolve_encode();
}

// Resident-B eligibility: shared weights, full split product, stacked-N tile widths, the whole weight tile (all taps and
// k-chunks, hi + lo) plus at least `min_stages` A stages fit in shared memory, and every CTA gets at least two tiles.
static int rb_stages_for(const GemmParams& p, int BN, long m_tiles, int ntn) {
  static int mode = -1, min_stages = 3;
  if (mode < 0) {
    const char* e = getenv("DEXB_RB");
    mode = (e != nullptr) ? atoi(e) : 1;
    const char* m = getenv("DEXB_RB_MIN_STAGES");
    if (m != nullptr) min_stages = atoi(m);
  }
  if (mode == 0 || p.b_mode != 0 || p.nheads != 1 || p.nsplit != 3 || BN > 128 || p.dbg != 0) return 0;
  const int nk = p.KH * p.KW * (p.K / kTcBlockK);
  int rb = (kTcSmemMax - tc_rb_bytes(BN, nk, 0)) / (2 * kTcBlockM * kTcBlockK * 2);
  if (rb > kTcMaxStages) rb = kTcMaxStages;
  if (rb < min_stages) return 0;
  const int ctas_per_n = g_num_sms / ntn;
  if (ctas_per_n < 1 || m_tiles < 2L * ctas_per_n) return 0;
  return rb;
}

// The FAST epilogue instantiation (gemm.cuh: epi_apply) assumes the widest vector access everywhere.
static bool epi_fast_ok(const GemmParams& p) {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("DEXB_EPI_FAST");
    mode = (e != nullptr) ? atoi(e) : 1;
  }
  if (mode == 0) return false;
  if (p.dbg != 0 || p.nsplit != 3) return false;          // the FAST kernel also drops the tuning / single-product branches
  const EpiParams& e = p.epi;
  auto al = [](const void* q, uintptr_t a) { return (reinterpret_cast<uintptr_t>(q) & (a - 1)) == 0; };
  if (p.N % 32 != 0 || e.out_vt != nullptr || e.colmean != nullptr || e.o_head_stride % 16 != 0) return false;
  if (e.bias != nullptr && !(al(e.bias, 16) && e.bias_zstride % 4 == 0 && e.bias_head_stride % 4 == 0)) return false;
  if (e.gate != nullptr && !al(e.gate, 16)) return false;
  if (e.resid_f32 != nullptr && !(al(e.resid_f32, 32) && e.resid_f32_stride % 8 == 0)) return false;
  if (e.resid_s != nullptr && !(al(e.resid_s, 16) && e.resid_s_stride % 8 == 0 && e.resid_s_hi % 8 == 0 && e.resid_s_lo % 8 == 0))
    return false;
  if (e.out_f32 != nullptr && !(al(e.out_f32, 32) && e.out_f32_stride % 8 == 0 && e.out_f32_col % 8 == 0)) return false;
  if (e.out_s != nullptr && !(al(e.out_s, 32) && e.out_s_stride % 16 == 0 && e.out_s_hi % 16 == 0 && e.out_s_lo % 16 == 0 &&
                              e.out_s_ncols >= p.N))
    return false;
  if (e.out_s_gshift > 0 && !(e.out_s_gshift >= 4 && e.out_s_gpitch % 16 == 0)) return false;   // a 16-column chunk stays in one group
  return true;
}

template <int BN, bool FAST, bool GNF = false>
static int launch_tc(const GemmPlan& gp, const GemmParams& p, cudaStream_t st, const typename GnFuseArg<GNF>::type& gf = {}) {
  const int ntn = cdiv(p.N, BN);
  const long m_tiles = (long)p.nz * p.TH * p.TW;
  const long total = m_tiles * ntn;
  if constexpr (FAST && !GNF && (BN == 64 || BN == 128)) {
    // CTA pairs (conv_pair.cuh): two m-tiles of one n-tile per cluster, half of the weight tile per CTA.  DEXB_PAIR=0: single CTAs.
    static int pair_mode = -1;
    if (pair_mode < 0) { const char* e = getenv("DEXB_PAIR"); pair_mode = (e != nullptr) ? atoi(e) : 1; }
    if (gp.halo && pair_mode != 0 && m_tiles % 2 == 0 && g_num_sms >= 2) {
      const long pairs = (m_tiles / 2) * ntn;
      const int grid = 2 * (int)(pairs < g_num_sms / 2 ? pairs : g_num_sms / 2);
      launch_pdl(conv_pair_kernel<BN>, dim3(grid), dim3(kTcThreads), cp_smem_bytes(BN), st, gp.tmA, gp.tmBh, p, (int)pairs, ntn);
      return 0;
    }
  }
  if (gp.halo && BN <= 128) {                          // the plan encoded 136-row A boxes: halo mode is the only valid launch
    int nb = (kTcSmemMax - tc_halo_bytes(BN, 0)) / (2 * BN * kTcBlockK * 2);
    if (nb > kTcMaxStages) nb = kTcMaxStages;
    static int grouped = -1;                             // DEXB_HALO_GROUP=0: one barrier round per tap (the round-1 scheme)
    if (grouped < 0) { const char* e = getenv("DEXB_HALO_GROUP"); grouped = (e != nullptr) ? atoi(e) : 1; }
    const int mode = (grouped != 0 && nb >= 6) ? -(nb + 16) : -nb;      // grouped rounds need two rounds of weight slots
    const int grid = (int)(total < g_num_sms ? total : g_num_sms);
    launch_pdl(gemm_tc_kernel<BN, FAST, GNF>, dim3(grid), dim3(TcEpi<GNF>::kBlock), tc_halo_bytes(BN, nb), st, gp.tmA, gp.tmB, p, (int)total, ntn, mode, gf);
    return 0;
  }
  const int rb = rb_stages_for(p, BN, m_tiles, ntn);
  if constexpr (FAST && !GNF && (BN == 64 || BN == 128)) {
    // resident weights + the activation stream shared by the two CTAs of a cluster (lin_mc.cuh): parity-green, measured neutral
    // (these GEMMs are bound by their stores, not by the activation stream) -- opt-in with DEXB_LINMC=1.
    static int mc_mode = -1;
    if (mc_mode < 0) { const char* e = getenv("DEXB_LINMC"); mc_mode = (e != nullptr) ? atoi(e) : 0; }
    const int nk = p.K / kTcBlockK;
    if (rb > 0 && mc_mode != 0 && ntn % 2 == 0 && p.N % BN == 0 && p.KH == 1 && p.KW == 1 && p.in_stride == 1 && lm_stages(BN, nk) >= 3) {
      const int npairs = ntn / 2;
      const int ncl = (g_num_sms / 2 / npairs) * npairs;       // clusters: a multiple of the n-tile pairs
      const int nms = ncl / npairs;
      if (nms >= 1 && m_tiles >= 2L * nms) {
        launch_pdl(lin_mc_kernel<BN>, dim3(2 * ncl), dim3(kTcThreads), lm_smem_bytes(BN, nk), st, gp.tmA, gp.tmB, p, (int)m_tiles, ntn, nms);
        return 0;
      }
    }
  }
  if (rb > 0) {
    const int nk = p.KH * p.KW * (p.K / kTcBlockK);
    const int grid = (g_num_sms / ntn) * ntn;            // multiple of ntn: a CTA never changes its n-tile
    launch_pdl(gemm_tc_kernel<BN, FAST, GNF>, dim3(grid), dim3(TcEpi<GNF>::kBlock), tc_rb_bytes(BN, nk, rb), st, gp.tmA, gp.tmB, p, (int)total, ntn, rb, gf);
    return 0;
  }
  const int grid = (int)(total < g_num_sms ? total : g_num_sms);
  launch_pdl(gemm_tc_kernel<BN, FAST, GNF>, dim3(grid), dim3(TcEpi<GNF>::kBlock), TcSmem<BN>::kBytes, st, gp.tmA, gp.tmB, p, (int)total, ntn, 0, gf);
  return 0;
}

// GroupNorm-apply can ride in the convolution kernel (gemm.cuh, GNF) when the tcgen05 engine runs the FAST epilogue with deferred
// GroupNorm sums (one n-tile of 64 or 128 channels, no heads) and a thread of the 512 epilogue threads keeps one channel octet.
bool gemm_can_fuse_gn(const GemmPlan& gp, const GemmParams& p, int engine) {
  static int mode = -1;
  if (mode < 0) { const char* e = getenv("DEXB_GN_FUSE"); mode = (e != nullptr) ? atoi(e) : 0; }
  if (mode == 0 || engine != 0 || !gp.tc_ok || !epi_fast_ok(p)) return false;
  if (gp.block_n != 64 && gp.block_n != 128) return false;
  return p.epi.gn_stats != nullptr && p.N <= gp.block_n && p.nheads == 1 && (p.N == 64 || p.N == 128) && p.epi.out_f32 != nullptr;
}

int gemm_launch(const GemmPlan& gp, const GemmParams& p_in, int engine, cudaStream_t st, const GnFuse* gf) {
  static int late_wait = -1;
  if (late_wait < 0) { const char* e = getenv("DEXB_EARLY_WAIT"); late_wait = (e != nullptr && e[0] == '0') ? 1 : 0; }
  GemmParams p = p_in;
  p.late_wait = late_wait;
  if (gf != nullptr) {
    DEXB_CHECK(gemm_can_fuse_gn(gp, p, engine) && gf->a.raw == p.epi.out_f32 && gf->a.C == p.N && gf->a.B == p.nz &&
                   gf->a.P == p.OH * p.OW && gf->done != nullptr,
               "gemm_launch: this convolution cannot carry a fused GroupNorm-apply");
    if (gp.block_n == 64) DEXB_TRY((launch_tc<64, true, true>(gp, p, st, *gf)));
    else DEXB_TRY((launch_tc<128, true, true>(gp, p, st, *gf)));
    DEXB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  if (engine == 0 && gp.tc_ok) {
    if (epi_fast_ok(p)) {
      switch (gp.block_n) {
        case 32: DEXB_TRY((launch_tc<32, true>(gp, p, st))); break;
        case 64: DEXB_TRY((launch_tc<64, true>(gp, p, st))); break;
        case 128: DEXB_TRY((launch_tc<128, true>(gp, p, st))); break;
        default: DEXB_TRY((launch_tc<256, true>(gp, p, st))); break;
      }
    } else {
      ++g_generic_epilogues;
      switch (gp.block_n) {
        case 32: DEXB_TRY((launch_tc<32, false>(gp, p, st))); break;
        case 64: DEXB_TRY((launch_tc<64, false>(gp, p, st))); break;
        case 128: DEXB_TRY((launch_tc<128, false>(gp, p, st))); break;
        default: DEXB_TRY((launch_tc<256, false>(gp, p, st))); break;
      }
    }
  } else {
    dim3 grid((unsigned)(p.nz * p.TH * p.TW), (unsigned)cdiv(p.N, 64));
    gemm_simt_kernel<<<grid, kSimtThreads, 0, st>>>(p);
  }
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dexb
