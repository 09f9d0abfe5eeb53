// Token GEMMs with resident weights on CTA pairs that SHARE THE ACTIVATION STREAM (TMA multicast) -- the resident-B mode of
// gemm_tc_kernel (gemm.cuh) for the DiT linears (qkv, proj, fc1, fc2: K = 256 / 512, N = 256 ... 768).
//
// In resident-B mode a CTA keeps all K chunks of ONE n-tile of the weights in shared memory and streams every m-tile of the activations
// past them, so the activation matrix is read once per n-tile: 6 x 21 MB for qkv, and the kernel is bound by that stream plus its own
// stores (tools/gemm_bench.py: TMA only 16.9 us, TMA + epilogue 32.0 us, everything 33.4 us -- the MMAs are free).  Here the two CTAs of
// a cluster own two adjacent n-tiles and walk the same m-tiles in lockstep; every A stage is loaded ONCE for both: the even CTA fetches
// the hi tile, the odd CTA the lo tile, each with `.multicast::cluster` into the same ring slot of both CTAs (the data and the
// complete_tx arrive at the same CTA-relative offsets in both).  L2 -> SM traffic of the activations halves.
//
// RESULT (B200, profiles/r02_lin_mc.md): parity-green and bit-identical to the engine, but neutral -- qkv 33.4 -> 32.9 us, fc1 24.9 -> 24.6,
// fc2 21.5 -> 21.7; without the epilogue's stores the kernels do gain (qkv 23.9 -> 21.8 us), i.e. the stores, not the activation stream,
// set their time.  Shipped OFF (DEXB_LINMC=1 enables it).
//
// Protocol: full[s] (per CTA, 1 arrival with expect_tx of the whole stage, armed by the CTA's own producer) receives the bytes of both
// multicasts; empty[s] (per CTA, 2 arrivals) is signalled by BOTH CTAs' MMA threads with a multicast commit -- a producer may only
// overwrite a slot that both consumers have retired, because its load lands in both.  MMAs, tensor memory and the epilogue are per CTA
// (cta_group::1), exactly as in the engine: stacked-N split product, double-buffered accumulator, epi_apply<16, true>.
#pragma once
#include "conv_pair.cuh"

namespace dexb {
namespace ptx {

__device__ __forceinline__ void tma_load_4d_mc(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3, uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4, %5, %6}], [%2], %7;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "h"(mask)
      : "memory");
}
// cta_group::1 MMAs of this thread retired -> arrive on the barrier at this offset in every CTA of `mask`
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"(mask)
               : "memory");
}

}  // namespace ptx

constexpr int kLmMaxStages = 6;
__host__ __device__ constexpr int lm_b_bytes(int block_n, int nk) { return nk * 2 * block_n * kTcBlockK * 2; }
__host__ __device__ constexpr int lm_stages(int block_n, int nk) {
  const int n = (kTcSmemMax - 1024 - 512 - lm_b_bytes(block_n, nk)) / (2 * kTcBlockM * kTcBlockK * 2);
  return n > kLmMaxStages ? kLmMaxStages : n;
}
__host__ __device__ constexpr int lm_smem_bytes(int block_n, int nk) {
  return lm_b_bytes(block_n, nk) + lm_stages(block_n, nk) * 2 * kTcBlockM * kTcBlockK * 2 + 1024 + 512;
}

// grid = 2 x (#clusters): cluster c owns the n-tiles 2 (c % (ntn / 2)) + {0, 1} and the m-tiles c / (ntn / 2), + nms, ... (nms = clusters per
// n-tile pair).  Requires ntn even, N % BLOCK_N == 0, one tap, shared weights (host: lin_mc_ok).
template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
lin_mc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p, const int m_tiles,
              const int ntn, const int nms) {
  static_assert(BLOCK_N == 64 || BLOCK_N == 128, "multicast linear kernel: n-tiles of 64 or 128 columns");
  constexpr int ABYTES = kTcBlockM * kTcBlockK * 2;                // 16 KiB per (hi | lo) tile
  constexpr int BB2 = 2 * BLOCK_N * kTcBlockK * 2;                 // one resident weight chunk: [B_hi | B_lo]
  constexpr int ACC_COLS = 2 * BLOCK_N;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int cl = (int)(blockIdx.x >> 1);
  const int npairs = ntn >> 1;
  const int n_tile = 2 * (cl % npairs) + (int)rank, ms = cl / npairs;
  const int nk = p.K / kTcBlockK;
  const int STAGES = (kTcSmemMax - 1024 - 512 - nk * BB2) / (2 * ABYTES) > kLmMaxStages ? kLmMaxStages
                                                                                          : (kTcSmemMax - 1024 - 512 - nk * BB2) / (2 * ABYTES);
  uint8_t* ring = smem + nk * BB2;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + STAGES * 2 * ABYTES);   // [STAGES]
  uint64_t* empty = full + kLmMaxStages;                                      // [STAGES]
  uint64_t* acc_full = empty + kLmMaxStages;                                  // [2]
  uint64_t* acc_empty = acc_full + 2;                                         // [2]
  uint64_t* b_full = acc_empty + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(b_full + 1);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full[s], 1); ptx::mbar_init(&empty[s], 2); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], kTcEpiThreads); }
    ptx::mbar_init(b_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                         // the peer's barriers exist before a multicast can signal them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---------------- TMA producer: resident weights of the own n-tile, then HALF of every A stage for both CTAs ----------------
    if (ptx::elect_one()) {
      ptx::mbar_expect_tx(b_full, (uint32_t)(nk * BB2));
      for (int kc = 0; kc < nk; ++kc) {
        ptx::tma_load_3d(smem + kc * BB2, &tmB, b_full, p.b_hi + kc * kTcBlockK, n_tile * BLOCK_N, 0);
        ptx::tma_load_3d(smem + kc * BB2 + BB2 / 2, &tmB, b_full, p.b_lo + kc * kTcBlockK, n_tile * BLOCK_N, 0);
      }
      const int a_col = rank ? p.a_lo : p.a_hi;
      uint32_t s = 0, ph = 0;
      for (int m = ms; m < m_tiles; m += nms) {
        const TcTile tl = tc_decode_tile(p, m, 1, BLOCK_N);
        const int ax = tl.cw0 + p.offW, ay = tl.ch0 + p.offH;
        for (int kc = 0; kc < nk; ++kc) {
          ptx::mbar_wait(&empty[s], ph ^ 1);                       // both CTAs have retired this slot
          uint8_t* st = ring + s * (2 * ABYTES);
          uint64_t* fb = &full[s];
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
          ptx::mbar_expect_tx(fb, 2 * ABYTES);                     // own barrier: the bytes of both multicasts land here
          ptx::tma_load_4d_mc(st + rank * ABYTES, &tmA, fb, a_col + kc * kTcBlockK, ax, ay, tl.img_a, (uint16_t)3);
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer (per CTA) ----------------
    if (ptx::elect_one()) {
      constexpr uint32_t idesc = ptx::make_idesc_bf16(kTcBlockM, BLOCK_N);
      constexpr uint32_t idesc2 = ptx::make_idesc_bf16(kTcBlockM, 2 * BLOCK_N);
      constexpr uint64_t kDescBase = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t ring_u = ptx::smem_u32(ring) >> 4, bres_u = ptx::smem_u32(smem) >> 4;
      uint32_t s = 0, ph = 0;
      int li = 0;
      bool pre = false;
      ptx::mbar_wait(b_full, 0);
      for (int m = ms; m < m_tiles; m += nms, ++li) {
        const int buf = li & 1;
        ptx::mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
        for (int kc = 0; kc < nk; ++kc) {
          if (!pre) ptx::mbar_wait(&full[s], ph);
          pre = false;
          ptx::tc_fence_after();
          const uint32_t a_hi = ring_u + s * (uint32_t)(2 * ABYTES >> 4);
          const uint32_t a_lo = a_hi + (ABYTES >> 4);
          const uint32_t b_hi = bres_u + kc * (uint32_t)(BB2 >> 4);
#pragma unroll
          for (int kk = 0; kk < kTcBlockK / 16; ++kk) {
            const uint32_t ko = kk * 2;
            const uint64_t dbh = kDescBase + (b_hi + ko);
            ptx::mma_bf16_ss(tacc, kDescBase + (a_hi + ko), dbh, idesc2, (kc > 0 || kk > 0) ? 1u : 0u);    // [A_hi B_hi | A_hi B_lo]
            ptx::mma_bf16_ss(tacc, kDescBase + (a_lo + ko), dbh, idesc, 1u);                                //  + A_lo B_hi
          }
          uint64_t* eb = &empty[s];
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
          if (kc == nk - 1) ptx::mma_commit(&acc_full[buf]);
          else { ptx::mbar_wait(&full[s], ph); pre = true; }       // the next stage of this tile, before the commit
          ptx::mma_commit_mc(eb, (uint16_t)3);                     // slot retired here: tell both producers
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (per CTA): TMEM -> registers -> global ----------------
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lg * 32 + lane;
    constexpr int CW = kTcEpiCW;
    static_assert(CW == 16, "multicast linear kernel: 16-column epilogue chunks");
    constexpr int PLG = kTcEpiPerLG;
    constexpr int MAXCH = BLOCK_N / CW / PLG;
    const int n0 = n_tile * BLOCK_N;
    int li = 0;
    for (int m = ms; m < m_tiles; m += nms, ++li) {
      const TcTile tl = tc_decode_tile(p, m, 1, BLOCK_N);
      const int buf = li & 1;
      const int ch = tl.ch0 + r / p.BW, cw = tl.cw0 + r % p.BW;
      const bool valid = (ch < p.CH) && (cw < p.CW);
      const int oh = ch * p.out_scale + p.out_offh, ow = cw * p.out_scale + p.out_offw;
      const uint32_t tacc = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * ACC_COLS);
#pragma unroll 1
      for (int k = 0; k < MAXCH; ++k) {
        const int c = half + PLG * k;
        const int n0c = n0 + c * CW;
        float rpre[CW];
        epi_load_resid<CW, true>(p.epi, p.N, tl.z, p.nheads, oh, ow, p.OH, p.OW, valid, n0c, rpre);
        if (k == 0) {
          ptx::mbar_wait_backoff(&acc_full[buf], (li >> 1) & 1, 128);
          ptx::tc_fence_after();
        }
        float v[CW], v2[CW];
        ptx::tmem_ld16_nowait(tacc + (uint32_t)(c * CW), v);
        ptx::tmem_ld16_nowait(tacc + (uint32_t)(BLOCK_N + c * CW), v2);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] += v2[i];
        if (k == MAXCH - 1) {
          ptx::tc_fence_before();
          ptx::mbar_arrive(&acc_empty[buf]);
        }
        epi_apply<CW, true>(p.epi, p.N, tl.z, p.nheads, oh, ow, p.OH, p.OW, valid, n0c, v, rpre, nullptr);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                         // the peer may still multicast into this CTA's ring / signal its barriers
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

}  // namespace dexb
