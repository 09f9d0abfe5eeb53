// 3x3 stride-1 convolution on CTA PAIRS (tcgen05 cta_group::2): the halo mode of gemm_tc_kernel (gemm.cuh) with the weight stream split
// between the two SMs of a TPC.
//
// Why (profiles/r02_issue_bench.md, profiles/r02_cta2_bench.md): the split-bf16 stage mix [N = 2 BN | N = BN] x 4 retires at the
// tensor floor when nothing else touches shared memory (449 / 769 cycles for BN = 64 / 128), but in the real kernel the TMA fill of
// the same stage -- an input row slab every third stage plus a [B_hi | B_lo] weight tile EVERY stage -- goes through the same
// shared-memory port, and the convolutions ran at 50 % tensor-pipe activity.  A CTA pair computes two 128-pixel tiles (M = 256) of
// the same n-tile with ONE weight tile: every CTA loads and reads half of it.
//
// Operand layout (tools/cta2_bench.cu checks the semantics): an N-wide cta_group::2 MMA takes rows [0, N/2) of B from the even CTA
// and rows [N/2, N) from the odd CTA, both at the same shared-memory offset; D (its own 128 rows x N columns) lands in each CTA's
// tensor memory.  With h = BN / 2 a weight slot holds
//      even CTA:  [ B_hi[0:h]  ; B_lo[h:BN] ]        odd CTA:  [ B_hi[h:BN] ; B_lo[0:h] ]
// so that   MMA 1 (A_hi, N = 2 BN, whole slots)  -> columns [ A_hi B_hi[0:h] | A_hi B_lo[h:BN] | A_hi B_hi[h:BN] | A_hi B_lo[0:h] ]
//           MMA 2 (A_lo, N = BN, first h rows)   -> columns [ A_lo B_hi[0:h] | A_lo B_hi[h:BN] ]  accumulated onto the first two blocks,
// i.e. every column block only ever holds terms of ONE output column range: out[o] = T[o] + T[o < h ? 3h + o : h + o].
// Each CTA loads BN weight rows per (tap, k-chunk) instead of 2 BN, and no row is loaded twice.
//
// Protocol: TMA loads of both CTAs (cp.async.bulk.tensor ... cta_group::2) complete on the EVEN CTA's `full` barrier (2 arrivals
// with expect_tx, one per producer); the even CTA's elected thread issues all MMAs and releases slabs / weight slots / accumulators
// in both CTAs with multicast commits; the epilogue warps of both CTAs hand the accumulator back on the even CTA's `acc_empty`
// (one arrival per warp, remote for the odd CTA).  Everything else -- tile geometry, epilogue, deferred GroupNorm sums -- is the
// engine's (GemmParams / epi_apply), so the kernel is a drop-in for halo-mode launches with the FAST epilogue.
#pragma once
#include "gemm.cuh"

namespace dexb {
namespace ptx {

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// (default semantics -- release at CTA scope -- as for the local arrivals: an explicit .release.cluster puts a MEMBAR in front of every
//  arrival, which waits for the warp's outstanding global stores: 3.1 stalled warps per issue in the first version of the pair kernel)
__device__ __forceinline__ void mbar_expect_tx_cluster(uint32_t bar_cluster_addr, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cluster.b64 _, [%0], %1;" ::"r"(bar_cluster_addr), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster_addr) : "memory");
}
// TMA loads of a CTA pair: the destination is this CTA's shared memory, the barrier may live in the peer CTA
__device__ __forceinline__ void tma2_load_3d(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tma2_load_4d(void* dst, const CUtensorMap* m, uint32_t bar_cluster_addr, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
__device__ __forceinline__ void mma2_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on the barrier at this offset in BOTH CTAs of the pair once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void mma2_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}

}  // namespace ptx

// One ring of ROUNDS: a slot holds the input row slab (hi | lo) of one (dy, k-chunk) and this CTA's halves of the three weight tiles of its
// dx taps -- one `full` wait, 24 MMAs and ONE multicast commit per round in the issuing thread (a multicast commit is not free: with one
// commit per weight tile plus one per slab the 64-channel convolutions were 12-18 % slower than with this scheme).
__host__ __device__ constexpr int cp_round_bytes(int block_n) { return 2 * kTcHaloSlab + 3 * block_n * kTcBlockK * 2; }
__host__ __device__ constexpr int cp_slots(int block_n) { return (kTcSmemMax - 1024 - 1024) / cp_round_bytes(block_n); }
// + 1024 alignment slack + 1024: barriers (256 B) and the GroupNorm reduction scratch (16 warps x up to 8 floats)
__host__ __device__ constexpr int cp_smem_bytes(int block_n) { return cp_slots(block_n) * cp_round_bytes(block_n) + 1024 + 1024; }

// Deferred GroupNorm sums of a CTA leaving image `img`: warp sums -> shared memory -> the four lane-group warps of a column chunk are
// added by one of them -> one double atomic per (CTA, 8-column group, sum | sum of squares): 16 / 32 instead of 64 / 128 per CTA.
// (All CTAs leave an image at about the same time; the atomics on its 8 x 2 x kGnRep addresses serialise in L2.)  Called by all 16
// epilogue warps at the same tile.
template <int MAXCH>
__device__ __forceinline__ void cp_flush_gn(const EpiParams& e, int N, int img, int we, int lane, float (&gacc)[MAXCH][8], float* scratch) {
  constexpr int NV = MAXCH * 4;                              // values per warp: chunk k, group g, (sum, sumsq)
#pragma unroll
  for (int k = 0; k < MAXCH; ++k)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float v = warp_sum(gacc[k][j]);
      gacc[k][j] = 0.f;
      if (lane == 0) scratch[we * NV + k * 4 + j] = v;
    }
  asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
  if ((we & 3) == 0 && lane < NV) {
    const int half = we >> 2, k = lane >> 2, j = lane & 3;
    const float* sp = scratch + (half * 4) * NV + lane;
    const float v = (sp[0] + sp[NV]) + (sp[2 * NV] + sp[3 * NV]);
    const int n0 = (half + kTcEpiPerLG * k) * kTcEpiCW + (j >> 1) * 8;
    const int gs = e.gn_gs;
    if (n0 < N) atomicAdd(e.gn_stats + (((long)img * kGnRep + (blockIdx.x & (kGnRep - 1))) * (N / gs) + n0 / gs) * 2 + (j & 1), (double)v);
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kTcEpiThreads) : "memory");
}

// grid = 2 x min(#tile pairs, #SMs / 2); pair q walks u = q, q + #pairs, ...: n-tile u % ntn of the m-tiles 2 (u / ntn) + {0, 1}
template <int BLOCK_N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kTcThreads, 1)
conv_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmBh, const GemmParams p,
                 const int total_pairs, const int ntn) {
  static_assert(BLOCK_N == 64 || BLOCK_N == 128, "pair kernel: n-tiles of 64 or 128 channels");
  constexpr int AS = cp_slots(BLOCK_N);
  constexpr int H2 = BLOCK_N / 2;                                  // rows of one weight load
  constexpr int BSLOT = BLOCK_N * kTcBlockK * 2;                   // this CTA's half of a [B_hi | B_lo] tile
  constexpr int RB = cp_round_bytes(BLOCK_N);                      // a multiple of 1024
  constexpr int ACC_COLS = 2 * BLOCK_N;
  constexpr int TMEM_COLS = 2 * ACC_COLS;
  static_assert(AS >= 2 && RB % 1024 == 0, "pair kernel: ring geometry");
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = ptx::cluster_ctarank();
  const int pair = (int)(blockIdx.x >> 1), npairs = (int)(gridDim.x >> 1);
  const int kchunks = p.K / kTcBlockK;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + AS * RB);    // [AS]  (even CTA's copy is the live one)
  uint64_t* empty = full + AS;                                     // [AS]
  uint64_t* acc_full = empty + AS;                                 // [2]
  uint64_t* acc_empty = acc_full + 2;                              // [2]  (even CTA's copy is the live one)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_empty + 2);
  float* gn_scratch = reinterpret_cast<float*>(smem + AS * RB + 256);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmBh);
    for (int s = 0; s < AS; ++s) { ptx::mbar_init(&full[s], 2); ptx::mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], 2 * (kTcEpiThreads / 32)); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc2<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                         // the peer's barriers exist before anything arrives on them
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ---------------- TMA producer (both CTAs: own input slab, own half of the weights) ----------------
    if (ptx::elect_one()) {
      const uint32_t full_r0 = ptx::mapa_u32(ptx::smem_u32(&full[0]), 0);    // the even CTA's barriers, 8 B apart
      const int row_hi = (int)rank * H2, row_lo = (1 - (int)rank) * H2;
      uint32_t sa = 0, pa = 0;
      for (int u = pair; u < total_pairs; u += npairs) {
        const TcTile tl = tc_decode_tile(p, (2 * (u / ntn) + (int)rank) * ntn + u % ntn, ntn, BLOCK_N);
        int brow = tl.n0;
        for (int ty = 0; ty < 3; ++ty) {
          for (int kc = 0; kc < kchunks; ++kc) {
            ptx::mbar_wait(&empty[sa], pa ^ 1);
            uint8_t* sl = smem + sa * RB;
            const uint32_t fr = full_r0 + sa * 8u;
            if (++sa == (uint32_t)AS) { sa = 0; pa ^= 1; }
            ptx::mbar_expect_tx_cluster(fr, (uint32_t)RB);
            ptx::tma2_load_4d(sl, &tmA, fr, p.a_hi + kc * kTcBlockK, tl.cw0 - 1, tl.ch0 + ty - 1, tl.img_a);
            ptx::tma2_load_4d(sl + kTcHaloSlab, &tmA, fr, p.a_lo + kc * kTcBlockK, tl.cw0 - 1, tl.ch0 + ty - 1, tl.img_a);
#pragma unroll
            for (int tx = 0; tx < 3; ++tx) {
              uint8_t* st = sl + 2 * kTcHaloSlab + tx * BSLOT;
              const int br = brow + tx * p.b_rows_per_tap;
              ptx::tma2_load_3d(st, &tmBh, fr, p.b_hi + kc * kTcBlockK, br + row_hi, 0);
              ptx::tma2_load_3d(st + H2 * kTcBlockK * 2, &tmBh, fr, p.b_lo + kc * kTcBlockK, br + row_lo, 0);
            }
          }
          brow += 3 * p.b_rows_per_tap;
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ---------------- MMA issuer: one thread of the even CTA ----------------
    if (rank == 0 && ptx::elect_one()) {
      constexpr uint32_t idesc1 = ptx::make_idesc_bf16(256, BLOCK_N);
      constexpr uint32_t idesc2 = ptx::make_idesc_bf16(256, 2 * BLOCK_N);
      constexpr uint64_t kDescBase = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t ring_u = ptx::smem_u32(smem) >> 4;
      uint32_t sa = 0, pa = 0;
      bool prewaited = false;
      int li = 0;
      const int na = 3 * kchunks;
      for (int u = pair; u < total_pairs; u += npairs, ++li) {
        const int buf = li & 1;
        ptx::mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1);
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
        for (int ia = 0; ia < na; ++ia) {
          if (!prewaited) ptx::mbar_wait(&full[sa], pa);
          prewaited = false;
          ptx::tc_fence_after();
          const uint32_t a_base = ring_u + sa * (uint32_t)(RB >> 4);
#pragma unroll
          for (int tx = 0; tx < 3; ++tx) {
            const uint32_t a_hi = a_base + (uint32_t)(tx * 8);     // tap dx: the slab shifted by dx rows of 128 B
            const uint32_t a_lo = a_hi + (kTcHaloSlab >> 4);
            const uint32_t b0 = a_base + (uint32_t)((2 * kTcHaloSlab + tx * BSLOT) >> 4);
#pragma unroll
            for (int kk = 0; kk < kTcBlockK / 16; ++kk) {
              const uint32_t ko = kk * 2;
              const uint64_t db = kDescBase + (b0 + ko);
              ptx::mma2_bf16_ss(tacc, kDescBase + (a_hi + ko), db, idesc2, (ia > 0 || tx > 0 || kk > 0) ? 1u : 0u);
              ptx::mma2_bf16_ss(tacc, kDescBase + (a_lo + ko), db, idesc1, 1u);
            }
          }
          const uint32_t sa0 = sa;
          if (++sa == (uint32_t)AS) { sa = 0; pa ^= 1; }
          if (ia == na - 1) {
            ptx::mma2_commit(&acc_full[buf]);                      // never delay the epilogue behind the next tile's operands
          } else {
            ptx::mbar_wait(&full[sa], pa);                         // the next round's operands, before this round's commit
            prewaited = true;
          }
          ptx::mma2_commit(&empty[sa0]);
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue (both CTAs, own tile): TMEM -> registers -> global ----------------
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lg * 32 + lane;
    constexpr int CW = kTcEpiCW;
    static_assert(CW == 16, "pair kernel: 16-column epilogue chunks");
    constexpr int PLG = kTcEpiPerLG;
    constexpr int MAXCH = BLOCK_N / CW / PLG;
    const bool defer_gn = p.epi.gn_stats != nullptr && ntn == 1 && p.nheads == 1;
    const uint32_t acc_empty_r0 = ptx::mapa_u32(ptx::smem_u32(&acc_empty[0]), 0), acc_empty_r1 = ptx::mapa_u32(ptx::smem_u32(&acc_empty[1]), 0);
    float gacc[MAXCH][8];
#pragma unroll
    for (int k = 0; k < MAXCH; ++k)
#pragma unroll
      for (int g = 0; g < 8; ++g) gacc[k][g] = 0.f;
    int gn_img = -1;
    int li = 0;
    for (int u = pair; u < total_pairs; u += npairs, ++li) {
      const TcTile tl = tc_decode_tile(p, (2 * (u / ntn) + (int)rank) * ntn + u % ntn, ntn, BLOCK_N);
      const int buf = li & 1;
      if (defer_gn && tl.z != gn_img) {
        if (gn_img >= 0) cp_flush_gn<MAXCH>(p.epi, p.N, gn_img, warp - 2, lane, gacc, gn_scratch);
        gn_img = tl.z;
      }
      const int ch = tl.ch0 + r / p.BW, cw = tl.cw0 + r % p.BW;
      const bool valid = (ch < p.CH) && (cw < p.CW);
      const int oh = ch * p.out_scale + p.out_offh, ow = cw * p.out_scale + p.out_offw;
      const uint32_t tacc = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * ACC_COLS);
#pragma unroll                                                     // static gacc indices: the sums stay in registers (no L1 here: a
      for (int k = 0; k < MAXCH; ++k) {                            // stack slot is an L2 round trip)
        const int c = half + PLG * k;
        const int o = c * CW;                                      // output column inside the n-tile
        const int n0c = tl.n0 + o;
        float rpre[CW];
        epi_load_resid<CW, true>(p.epi, p.N, tl.z, p.nheads, oh, ow, p.OH, p.OW, valid, n0c, rpre);
        if (k == 0) {
          ptx::mbar_wait_backoff(&acc_full[buf], (li >> 1) & 1, 128);
          ptx::tc_fence_after();
        }
        float v[CW], v2[CW];
        ptx::tmem_ld16_nowait(tacc + (uint32_t)o, v);
        ptx::tmem_ld16_nowait(tacc + (uint32_t)(o < H2 ? 3 * H2 + o : H2 + o), v2);
        ptx::tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < CW; ++i) v[i] += v2[i];
        if (k == MAXCH - 1) {                                      // last chunk read: hand the buffer back (one arrival per warp)
          ptx::tc_fence_before();
          __syncwarp();
          if (lane == 0) ptx::mbar_arrive_cluster(buf ? acc_empty_r1 : acc_empty_r0);
        }
        epi_apply<CW, true>(p.epi, p.N, tl.z, p.nheads, oh, ow, p.OH, p.OW, valid, n0c, v, rpre, defer_gn ? &gacc[k][0] : nullptr);
      }
    }
    if (defer_gn && gn_img >= 0) cp_flush_gn<MAXCH>(p.epi, p.N, gn_img, warp - 2, lane, gacc, gn_scratch);
  }
  ptx::tc_fence_before();
  __syncthreads();
  ptx::cluster_sync_all();                         // neither CTA leaves (or frees tensor memory) while the peer may still signal it
  if (warp == 1) ptx::tmem_dealloc2<TMEM_COLS>(tmem_base);
}

}  // namespace dexb
