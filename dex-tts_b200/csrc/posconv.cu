// Grouped 16x16 positional convolution of the DiT patch embedding (make_conv_pos, DEX-TTS/model/dit.py:75-90,444-446) as a
// weight-stationary tensor-core kernel.
//
//   out[b][y][x][g*32+n] = GELU( bias + sum_{ky,kx,ci} in[b][y+ky-8][x+kx-8][g*32+ci] * W[g*32+n][ci][ky][kx] )     (SamePad crop)
//
// The generic implicit-GEMM engine runs this with N = 32 output channels per group and one K = 64 chunk per tap pair: every
// MMA re-reads a 4 KiB activation tile for 16 cycles of math, and each tap pair re-streams the activations (36x redundancy).
// Here the roles are swapped:
//   * A (M = 128) = the weights of FOUR x-taps x 32 output channels of one group for one ky, resident in TENSOR MEMORY
//     (split hi | lo, 32 columns), staged by the accumulation warps with tcgen05.st;
//   * B (N = 96)  = one row of input pixels [x0-8, x0+88) of that group, K = 32 input channels, stored [hi(32) | lo(32)] = one
//     128 B row per pixel, so the split product is one chain of six K=16 MMAs over the same shared-memory tile
//     (W_hi x in_hi, W_hi x in_lo, W_lo x in_hi); one TMA row tile serves all 16 x-taps;
//   * D[(tap j, n)][input column c] lands in TMEM; warp j reads ITS lane group shifted by its tap (columns [kx, kx+72)) --
//     the x-shift of the convolution is a warp-uniform column offset of tcgen05.ld -- and accumulates 72 outputs in registers
//     over all (ky, tap group) steps; the four taps of a group are summed through shared memory at the end of the tile.
// Rows of the kernel that fall outside the image (zero padding) are skipped (20 % of the steps for a 20-row grid).
// Warps: 0 = TMA, 1 = MMA issuer, 2..5 = weight staging + accumulation + epilogue.  Two CTAs per SM (256 TMEM columns each).
#include "posconv.cuh"

#include <cudaTypedefs.h>
#include <string.h>

#include "ptx.cuh"

namespace dexb {

constexpr int kPcXT = 72;                 // output pixels per tile
constexpr int kPcNB = 96;                 // input pixels per row tile = MMA N (XT + 15 rounded up to 16)
constexpr int kPcStages = 3;
constexpr int kPcRowBytes = kPcNB * 128;  // 12 KiB
constexpr int kPcWOff = kPcStages * kPcRowBytes;                   // 2 x 16 KiB weight blocks (TMA, 128B swizzle)
constexpr int kPcWBytes = 128 * 128;
constexpr int kPcRedOff = kPcWOff + 2 * kPcWBytes;                 // [4][72][33] fp32
constexpr int kPcRedBytes = 4 * kPcXT * 33 * 4;
constexpr int kPcBarOff = kPcRedOff + kPcRedBytes;
constexpr int kPcSmem = kPcBarOff + 256 + 1024;
constexpr int kPcThreads = 192;
constexpr uint32_t kTmD = 0;              // 2 x 96 fp32 accumulator columns
constexpr uint32_t kTmW = 192;            // 2 x (16 hi + 16 lo) weight columns

struct PcTile {
  int b, g, y, x0, ky_lo, ky_hi;
};
__device__ __forceinline__ PcTile pc_decode(const PosConvParams& p, int t) {
  PcTile r;
  const int xt = t % p.XTn; t /= p.XTn;
  r.g = t % p.G; t /= p.G;
  r.y = t % p.Fq; t /= p.Fq;
  r.b = t;
  r.x0 = xt * kPcXT;
  r.ky_lo = max(0, p.KS / 2 - r.y);                       // rows y + ky - KS/2 inside [0, Fq)
  r.ky_hi = min(p.KS - 1, p.Fq - 1 + p.KS / 2 - r.y);
  return r;
}

__global__ void __launch_bounds__(kPcThreads, 2)
posconv_kernel(const __grid_constant__ CUtensorMap tmIn, const __grid_constant__ CUtensorMap tmW, const PosConvParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kPcBarOff);
  uint64_t* in_full = bars;            // 4
  uint64_t* in_empty = bars + 4;       // 4
  uint64_t* w_full = bars + 8;         // 2: weights staged in tensor memory
  uint64_t* d_full = bars + 10;        // 2
  uint64_t* ws_full = bars + 14;       // 2: weight block landed in shared memory
  uint64_t* ws_empty = bars + 16;      // 2
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmIn);
    ptx::prefetch_tmap(&tmW);
    for (int s = 0; s < kPcStages; ++s) { ptx::mbar_init(&in_full[s], 1); ptx::mbar_init(&in_empty[s], 1); }
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&w_full[s], 128); ptx::mbar_init(&d_full[s], 1);
      ptx::mbar_init(&ws_full[s], 1); ptx::mbar_init(&ws_empty[s], 128);
    }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<256>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  pdl_wait();
  const int KG = p.KS / 4;                                  // tap groups of 4 along x

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer: one input row tile per valid ky,
    //                                                              one 128 x 64 weight block per (ky, tap group) step
    if (ptx::elect_one()) {
      uint32_t r = 0, s = 0;
      for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
        const PcTile tl = pc_decode(p, t);
        for (int ky = tl.ky_lo; ky <= tl.ky_hi; ++ky, ++r) {
          const int st = r % kPcStages;
          ptx::mbar_wait(&in_empty[st], ((r / kPcStages) & 1) ^ 1);
          ptx::mbar_expect_tx(&in_full[st], kPcRowBytes);
          ptx::tma_load_4d(smem + st * kPcRowBytes, &tmIn, &in_full[st], 0, tl.x0 - p.KS / 2, tl.y + ky - p.KS / 2,
                           tl.b * p.G + tl.g);
          for (int kg = 0; kg < KG; ++kg, ++s) {
            const int bf = s & 1;
            ptx::mbar_wait(&ws_empty[bf], ((s >> 1) & 1) ^ 1);
            ptx::mbar_expect_tx(&ws_full[bf], kPcWBytes);
            ptx::tma_load_2d(smem + kPcWOff + bf * kPcWBytes, &tmW, &ws_full[bf], 0, ((tl.g * p.KS + ky) * KG + kg) * 128);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc = ptx::make_idesc_bf16(128, kPcNB);
    uint32_t r = 0, s = 0;
    for (int t = blockIdx.x; t < p.total_tiles; t += gridDim.x) {
      const PcTile tl = pc_decode(p, t);
      for (int ky = tl.ky_lo; ky <= tl.ky_hi; ++ky, ++r) {
        const int st = r % kPcStages;
        ptx::mbar_wait(&in_full[st], (r / kPcStages) & 1);
        for (int kg = 0; kg < KG; ++kg, ++s) {
          const int bf = s & 1;
          const uint32_t ph = (s >> 1) & 1;
          // ONE wait per step: w_full(s) is only complete once all four accumulation warps have staged the weights of step s, and each
          // of them has read D(s - 2) one loop iteration earlier (it reads D(s - 2), fences, then stages W(s)) -- so the accumulator
          // buffer of this step is free as well.  A second wait would cost the issuing thread ~240 cycles per step against ~350
          // cycles of MMAs (six N = 96 TS-mode instructions): r01 waited on both and ran 61 % tensor-pipe active.
          ptx::mbar_wait(&w_full[bf], ph);
          ptx::tc_fence_after();
          if (ptx::elect_one()) {
            const uint32_t in_base = ptx::smem_u32(smem + st * kPcRowBytes);
            const uint32_t d = tmem + kTmD + (uint32_t)(bf * kPcNB);
            const uint32_t wh = tmem + kTmW + (uint32_t)(bf * 32), wl = wh + 16;
            // in row = [hi ci 0..31 | lo ci 0..31]: k16 slab q of the tile at byte offset 32 q
            ptx::mma_bf16_ts(d, wh, ptx::make_desc_k128(in_base), idesc, 0u);            // W_hi[0:16]  x in_hi[0:16]
            ptx::mma_bf16_ts(d, wh + 8, ptx::make_desc_k128(in_base + 32), idesc, 1u);   // W_hi[16:32] x in_hi[16:32]
            ptx::mma_bf16_ts(d, wh, ptx::make_desc_k128(in_base + 64), idesc, 1u);       // W_hi[0:16]  x in_lo[0:16]
            ptx::mma_bf16_ts(d, wh + 8, ptx::make_desc_k128(in_base + 96), idesc, 1u);   // W_hi[16:32] x in_lo[16:32]
            ptx::mma_bf16_ts(d, wl, ptx::make_desc_k128(in_base), idesc, 1u);            // W_lo[0:16]  x in_hi[0:16]
            ptx::mma_bf16_ts(d, wl + 8, ptx::make_desc_k128(in_base + 32), idesc, 1u);   // W_lo[16:32] x in_hi[16:32]
            ptx::mma_commit(&d_full[bf]);
            if (kg == KG - 1) ptx::mma_commit(&in_empty[st]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ------------------------------------------------------------ weight staging + accumulation + epilogue
    const int j = warp & 3;                                  // TMEM lane group of this warp = x-tap inside the tap group
    const int row = j * 32 + lane;                           // A row (tap j, output channel n = lane)
    const uint32_t tl_lane = tmem + ((uint32_t)(j * 32) << 16);
    float* red = reinterpret_cast<float*>(smem + kPcRedOff);
    // move this thread's row of weight block `sn` (hi | lo, 128 B, 128B-swizzled by the TMA) from shared to tensor memory
    const uint32_t wrow = ptx::smem_u32(smem + kPcWOff) + (uint32_t)row * 128u;
    auto stage_w = [&](uint32_t sn) {
      const int bf = sn & 1;
      ptx::mbar_wait(&ws_full[bf], (sn >> 1) & 1);
      uint32_t h[16], l[16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const uint4 a = ptx::lds_b128(wrow + bf * kPcWBytes + (uint32_t)((i ^ (row & 7)) << 4));
        const uint4 c = ptx::lds_b128(wrow + bf * kPcWBytes + (uint32_t)(((4 + i) ^ (row & 7)) << 4));
        h[i * 4] = a.x; h[i * 4 + 1] = a.y; h[i * 4 + 2] = a.z; h[i * 4 + 3] = a.w;
        l[i * 4] = c.x; l[i * 4 + 1] = c.y; l[i * 4 + 2] = c.z; l[i * 4 + 3] = c.w;
      }
      ptx::mbar_arrive(&ws_empty[bf]);
      ptx::tmem_st16(tl_lane + kTmW + bf * 32, h);
      ptx::tmem_st16(tl_lane + kTmW + bf * 32 + 16, l);
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(&w_full[bf]);
    };
    uint32_t s = 0;
    int t = blockIdx.x;
    bool have = t < p.total_tiles;
    PcTile tl;
    if (have) {
      tl = pc_decode(p, t);
      stage_w(0);
    }
    while (have) {
      float acc[kPcXT];
#pragma unroll
      for (int i = 0; i < kPcXT; ++i) acc[i] = 0.f;
      const int tn = t + gridDim.x;
      const bool have_n = tn < p.total_tiles;
      PcTile tln;
      if (have_n) tln = pc_decode(p, tn);
      for (int ky = tl.ky_lo; ky <= tl.ky_hi; ++ky) {
#pragma unroll 1
        for (int kg = 0; kg < KG; ++kg, ++s) {
          // weights of the NEXT step (the buffer it uses was last read by step s-1, whose MMAs completed before
          // d_full(s-1) was observed in the previous iteration)
          if (kg + 1 < KG || ky < tl.ky_hi || have_n) stage_w(s + 1);
          const int bf = s & 1;
          ptx::mbar_wait(&d_full[bf], (s >> 1) & 1);
          ptx::tc_fence_after();
          // output pixel xr of this tap reads input column xr + kx:  columns [kx, kx + 72) of this warp's lanes
          const uint32_t col = tl_lane + kTmD + (uint32_t)(bf * kPcNB + kg * 4 + j);
          float v0[32], v1[32], v2[8];
          ptx::tmem_ld32(col, v0);
          ptx::tmem_ld32(col + 32, v1);
          {
            uint32_t* q = reinterpret_cast<uint32_t*>(v2);
            asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                         : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                         : "r"(col + 64)
                         : "memory");
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
          }
          ptx::tc_fence_before();                                // orders the tensor-memory reads before this thread's next arrival (w_full)
#pragma unroll
          for (int i = 0; i < 32; ++i) { acc[i] += v0[i]; acc[32 + i] += v1[i]; }
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[64 + i] += v2[i];
        }
      }
      // ---- tile epilogue: sum the four taps of a group through shared memory, bias, GELU, store
      asm volatile("bar.sync 1, 128;" ::: "memory");          // previous tile's readers are done with `red`
#pragma unroll
      for (int i = 0; i < kPcXT; ++i) red[(j * kPcXT + i) * 33 + lane] = acc[i];
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const float bias = __ldg(p.bias + tl.g * 32 + lane);
      for (int xr = j; xr < kPcXT; xr += 4) {
        const int x = tl.x0 + xr;
        if (x >= p.Wq) break;
        float v = ((red[(0 * kPcXT + xr) * 33 + lane] + red[(1 * kPcXT + xr) * 33 + lane]) +
                   (red[(2 * kPcXT + xr) * 33 + lane] + red[(3 * kPcXT + xr) * 33 + lane])) + bias;
        v = gelu_fast(v);
        p.out[(((long)tl.b * p.Fq + tl.y) * p.Wq + x) * p.D + tl.g * 32 + lane] = v;
      }
      t = tn; have = have_n; tl = tln;
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<256>(tmem);
}

// ------------------------------------------------------------------------------------------------
// packers
// ------------------------------------------------------------------------------------------------
// xe fp32 [b][y][x][D] -> pin [b][g][y][x][hi(32) | lo(32)]  (one 128 B row per pixel and group)
__global__ void __launch_bounds__(256) k_posconv_pack_in(const float* __restrict__ xe, bf16* __restrict__ pin, int B, int Fq,
                                                         int Wq, int D, int G) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * Fq * Wq * (D / 8);
  if (i >= total) return;
  const int c0 = (int)(i % (D / 8)) * 8;
  const long pix = i / (D / 8);
  const int x = (int)(pix % Wq), y = (int)((pix / Wq) % Fq), b = (int)(pix / ((long)Wq * Fq));
  const int g = c0 / 32, ci = c0 % 32;
  const float4 r0 = *reinterpret_cast<const float4*>(xe + pix * D + c0);
  const float4 r1 = *reinterpret_cast<const float4*>(xe + pix * D + c0 + 4);
  const float v[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  bf16* row = pin + ((((long)b * G + g) * Fq + y) * Wq + x) * 64;
  store_split8(row + ci, row + 32 + ci, v);
}
void launch_posconv_pack_in(const float* xe, bf16* pin, int B, int Fq, int Wq, int D, int G, cudaStream_t st) {
  launch_pdl(k_posconv_pack_in, dim3((unsigned)(cdiv((long)B * Fq * Wq * (D / 8), 256))), dim3(256), 0, st, xe, pin, B, Fq, Wq, D, G);
}

// W [Co = G*32][Cg = 32][KS][KS] -> pw [g][ky][kg][row = j*32 + n][hi(32 ci) | lo(32 ci)],  kx = 4 kg + j
__global__ void k_posconv_pack_w(const float* __restrict__ w, bf16* __restrict__ pw, int G, int KS) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)G * 32 * 32 * KS * KS;
  if (i >= total) return;
  long t = i;
  const int kx = (int)(t % KS); t /= KS;
  const int ky = (int)(t % KS); t /= KS;
  const int ci = (int)(t % 32); t /= 32;
  const int co = (int)t;
  const int g = co / 32, n = co % 32, kg = kx / 4, j = kx % 4;
  bf16* row = pw + ((((long)g * KS + ky) * (KS / 4) + kg) * 128 + (j * 32 + n)) * 64;
  split2(w[i], row[ci], row[32 + ci]);
}
void launch_posconv_pack_w(const float* w, bf16* pw, int G, int KS, cudaStream_t st) {
  k_posconv_pack_w<<<cdiv((long)G * 32 * 32 * KS * KS, 256), 256, 0, st>>>(w, pw, G, KS);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
bool posconv_supported(int hidden, int groups, int ks) { return hidden / groups == 32 && ks == 16; }

int posconv_global_init() {
  DEXB_CUDA_OK(cudaFuncSetAttribute(posconv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kPcSmem));
  return 0;
}

int posconv_plan_init(PosConvPlan* pp, const bf16* pin, const bf16* pw, const float* bias, float* out, int B, int Fq, int Wq,
                      int D, int G, int KS) {
  DEXB_CHECK(posconv_supported(D, G, KS), "weight-stationary pos-conv is instantiated for 32 channels per group and a 16x16 kernel");
  PosConvParams& p = pp->p;
  memset(&p, 0, sizeof(p));
  p.B = B; p.Fq = Fq; p.Wq = Wq; p.D = D; p.G = G; p.KS = KS;
  p.XTn = (Wq + kPcXT - 1) / kPcXT;
  p.total_tiles = B * Fq * G * p.XTn;
  p.w = pw; p.bias = bias; p.out = out;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  DEXB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  DEXB_CHECK(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
  auto enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  const cuuint64_t dims[4] = {64, (cuuint64_t)Wq, (cuuint64_t)Fq, (cuuint64_t)B * G};
  const cuuint64_t str[3] = {128, 128ull * Wq, 128ull * Wq * Fq};
  const cuuint32_t box[4] = {64, (cuuint32_t)kPcNB, 1, 1};
  const cuuint32_t es[4] = {1, 1, 1, 1};
  CUresult r = enc(&pp->tmIn, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<bf16*>(pin), dims, str, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DEXB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pos-conv input) failed with CUresult %d", (int)r);
  const cuuint64_t wdims[2] = {64, (cuuint64_t)G * KS * (KS / 4) * 128};
  const cuuint64_t wstr[1] = {128};
  const cuuint32_t wbox[2] = {64, 128};
  r = enc(&pp->tmW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<bf16*>(pw), wdims, wstr, wbox, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DEXB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(pos-conv weights) failed with CUresult %d", (int)r);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  pp->grid = p.total_tiles < 2 * sms ? p.total_tiles : 2 * sms;
  return 0;
}

int posconv_launch(const PosConvPlan& pp, cudaStream_t st) {
  launch_pdl(posconv_kernel, dim3(pp.grid), dim3(kPcThreads), kPcSmem, st, pp.tmIn, pp.tmW, pp.p);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

double posconv_flop(const PosConvPlan& pp) {
  const PosConvParams& p = pp.p;
  return 2.0 * p.B * p.Fq * p.Wq * (double)p.D * 32 * p.KS * p.KS;
}

}  // namespace dexb
