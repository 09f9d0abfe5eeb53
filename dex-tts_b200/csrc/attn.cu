// Fused DiT self-attention (timm Attention inside DiTBlock, DEX-TTS/model/dit.py:270,282): softmax(q k^T / sqrt(hd)) v
// for one (sample, head) and 128 queries per CTA, no key mask (the reference has none), split-bf16 x3 precision.
//
// One pass over the keys with LAZY rescaling (r02):
//   S = Q K^T (3 MMAs per product) -> tile row maximum (the two column halves of a row exchange it through shared memory) ->
//   the reference maximum m of a row only moves when a tile exceeds it by more than 2^8 in the exponent (then l and the row of
//   the O accumulator in tensor memory are multiplied by exp2(m_old - m_new): between the wait for P(j-1) V(j-1) and the hand-over
//   of P(j), when O is stable; after the first tile that is rare) -> P = exp2((S - m) * scale * log2e) <= 256 -> O += P V.
//   softmax is shift-invariant, so any m gives the same result; the split-bf16 P keeps its relative precision up to 2^8.
//   (r01 ran an exact two-pass softmax -- row maximum first, from S = Qhi Khi^T only: 1/7 more MMA work and K streamed twice.)
// S (2 x 64 columns) and O (128 columns) live in TMEM; K and V^T tiles of 64 keys stream through two separate 3-slot TMA rings:
// a K slot is released when S(j) has retired, a V slot when P(j) V(j) has (one tile later), which gives both streams ~2.5 tiles of
// lead over the TMA round trip (one ring of (K, V) stages had one tile; measured neutral -- see below).
// What bounds it (profiles/r02_attn.md): after these changes the issuing thread hardly ever waits (1.4 polls of p_full per tile, its
// UTCHMMA slots stall on a full MMA queue) and the softmax warps idle 40 % of the time waiting for S: the tile time is the tensor
// pipe's own occupancy -- 36 TS-mode MMAs per 64-key tile (24 of N = 64 at 75 % efficiency, 12 of N = 128 at 86 %), 57-60 % math
// active.  Wider instructions would need a 128-key S tile, which does not fit tensor memory next to O, Q and P (640 > 512 columns).
// warp 0 = TMA, warp 1 = tcgen05.mma issuer, warps 2..9 = softmax (one query row x half of the columns per thread) + epilogue.
// S(j+1) is issued before P(j) V(j) so the tensor pipe works while the softmax warps exponentiate.
#include "attn.cuh"

#include <cudaTypedefs.h>
#include <stdlib.h>
#include <string.h>

#include "ptx.cuh"

namespace dexb {

constexpr int kAtBM = 128;          // queries per CTA
constexpr int kAtBN = 64;           // keys per iteration
constexpr int kAtHD = 128;          // head dim
constexpr int kAtThreads = 320;     // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two column halves x four lane groups)
constexpr int kAtStages = 3;                              // slots of the K ring and of the V ring
constexpr int kKBytes = 4 * 8192;                         // [hi kc0][hi kc1][lo kc0][lo kc1], 64 rows x 128 B each
constexpr int kVBytes = 2 * 16384;                        // [hi][lo], 128 d-rows x 128 B (64 keys) each
constexpr int kVOff = kAtStages * kKBytes;                // V ring behind the K ring
constexpr int kBarOff = kAtStages * (kKBytes + kVBytes);  // 192 KiB
constexpr int kAtSmem = kBarOff + 256 + 2048 + 1024;      // barriers (256 B) + max/sum exchange (2 x 1 KiB) + alignment slack
constexpr float kAtTau = 8.f;       // one-pass softmax: the reference maximum moves when a tile exceeds it by 2^kAtTau
// tensor-memory columns (512 allocated): both MMA A operands (Q and P) live here, so shared memory only carries the
// streamed K / V^T tiles -- with split-bf16 every A tile would otherwise be re-read from shared memory three times per
// k-step and the kernel is shared-memory-bandwidth bound (profiles/r01_ncu_v4.md)
constexpr uint32_t kTmS = 0;        // 2 x 64  fp32 score tiles
constexpr uint32_t kTmO = 128;      // 128     fp32 output accumulator
constexpr uint32_t kTmQh = 256;     // 64      Q hi: 128 bf16 per row, two per column
constexpr uint32_t kTmQl = 320;     // 64      Q lo
constexpr uint32_t kTmPh = 384;     // 32      P hi: 64 bf16 per row
constexpr uint32_t kTmPl = 416;     // 32      P lo

// Packed fp32 pairs (sm_100: FFMA2 / FADD2) and the three-input maximum (FMNMX3): the softmax warps -- not the tensor pipe -- set the
// tile time of this kernel (profiles/r02_attn.md: 395 instructions per warp and 64-key tile, 75 % of their time busy), so the inner
// loop is written for instruction count.
__device__ __forceinline__ uint64_t f2_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void f2_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t f2_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t f2_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t f2_sub(uint64_t a, uint64_t b) { uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ float f_max3(float a, float b, float c) { float r; asm("max.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c)); return r; }

__global__ void __launch_bounds__(kAtThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kBarOff);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // 3
  uint64_t* k_empty = bars + 4;       // 3
  uint64_t* s_full = bars + 7;        // 2
  // (bars + 9, + 10: formerly s_empty -- not needed, see issue_s)
  uint64_t* p_full = bars + 11;       // 1
  uint64_t* p_empty = bars + 12;      // 1
  uint64_t* o_full = bars + 13;       // 1
  uint64_t* v_full = bars + 14;       // 3
  uint64_t* v_empty = bars + 17;      // 3
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // DiT / TV: blockIdx.x = query tile, all key tiles.  Split-KV mode (linear-attention context, one 128-row query tile):
  // blockIdx.x = key split, the CTA covers key tiles [t0, t0 + nt) and writes un-normalised partials.  Tail-split mode (DiT):
  // 1-D grid of work items, the last ones cover a key range of their tile only.
  int split = 0, m0, z, t0 = 0, nt = p.nt, omode = p.out_mode;
  long part_idx = 0;
  if (p.tail_splits > 1) {
    const int item = (int)blockIdx.x;
    int tile = item;
    if (item >= p.tail_first) {
      const int k = item - p.tail_first;
      tile = p.tail_first + k / p.tail_splits;
      split = k % p.tail_splits;
      t0 = split * p.tail_tps;
      nt = max(0, min(p.nt, t0 + p.tail_tps) - t0);
      omode = 2;
      part_idx = k;
    }
    z = tile / p.q_tiles;
    m0 = (tile % p.q_tiles) * kAtBM;
  } else if (p.kv_splits > 1) {
    split = (int)blockIdx.x;
    m0 = 0;
    z = blockIdx.y;
    t0 = split * p.tiles_per_split;
    nt = max(0, min(p.nt, t0 + p.tiles_per_split) - t0);
    part_idx = (long)(z / p.nheads) * p.kv_splits + split;
  } else {
    m0 = (int)blockIdx.x * kAtBM;
    z = blockIdx.y;
  }
  const int b = z / p.nheads, head = z % p.nheads;
  const int kch = p.kchunks;

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmK); ptx::prefetch_tmap(&tmV);
    ptx::mbar_init(q_full, 256);
    for (int s = 0; s < 2; ++s) {
      ptx::mbar_init(&s_full[s], 1);
    }
    for (int s = 0; s < kAtStages; ++s) {
      ptx::mbar_init(&k_full[s], 1); ptx::mbar_init(&k_empty[s], 1);
      ptx::mbar_init(&v_full[s], 1); ptx::mbar_init(&v_empty[s], 1);
    }
    ptx::mbar_init(p_full, 256); ptx::mbar_init(p_empty, 1); ptx::mbar_init(o_full, 1);
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<512>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tm_o = tmem + kTmO;
  pdl_wait();                                      // the setup above overlaps the tail of the previous kernel

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (ptx::elect_one()) {
      const int kcol = p.k_hi + head * kAtHD, lo = p.k_lo - p.k_hi;
      auto load_k = [&](int i) {                           // K tile i of this CTA's key range -> K ring
        const int s = i % kAtStages, j = t0 + i;
        ptx::mbar_wait(&k_empty[s], ((i / kAtStages) & 1) ^ 1);
        uint8_t* st = smem + s * kKBytes;
        ptx::mbar_expect_tx(&k_full[s], 2 * kch * 8192);
        for (int part = 0; part < 2; ++part)
          for (int kc = 0; kc < kch; ++kc)
            ptx::tma_load_3d(st + (part * 2 + kc) * 8192, &tmK, &k_full[s], part * lo + kcol + kc * 64, j * kAtBN, b);
      };
      auto load_v = [&](int i) {                           // V tile i -> V ring
        const int s = i % kAtStages, j = t0 + i;
        ptx::mbar_wait(&v_empty[s], ((i / kAtStages) & 1) ^ 1);
        uint8_t* st = smem + kVOff + s * kVBytes;
        ptx::mbar_expect_tx(&v_full[s], 2 * p.vchunks * 8192);
        if (p.v_mn) {                                      // row-major V: [part][d chunk of 64] boxes of 64 keys x 128 B
          for (int part = 0; part < 2; ++part)
            for (int c = 0; c < p.vchunks; ++c)
              ptx::tma_load_3d(st + part * 16384 + c * 8192, &tmV, &v_full[s], (part ? p.v_lo : p.v_hi) + head * kAtHD + c * 64,
                               j * kAtBN, b);
        } else {
          for (int part = 0; part < 2; ++part)
            ptx::tma_load_3d(st + part * 16384, &tmV, &v_full[s], part * p.KP + j * kAtBN, head * kAtHD, b);
        }
      };
      // issue order = the order in which the tensor pipe frees the slots: S(0) S(1) PV(0) S(2) PV(1) S(3) PV(2) ...
      for (int i = 0; i < kAtStages && i < nt; ++i) load_k(i);
      for (int i = 0; i < kAtStages && i < nt; ++i) load_v(i);
      if (kAtStages < nt) load_k(kAtStages);
      for (int i = 0; kAtStages + i < nt; ++i) {
        if (kAtStages + 1 + i < nt) load_k(kAtStages + 1 + i);
        load_v(kAtStages + i);
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer
    constexpr uint32_t idesc_s = ptx::make_idesc_bf16(128, kAtBN);
    const uint32_t idesc_o = ptx::make_idesc_bf16(128, 64 * p.vchunks, p.v_mn);
    ptx::mbar_wait(q_full, 0);
    ptx::tc_fence_after();
    // S[it & 1] = Q K^T for key tile `it` (Q from tensor memory); the K slot is free again once these MMAs have retired
    auto issue_s = [&](int it) {
      const int st = it % kAtStages, sb = it & 1;
      ptx::mbar_wait(&k_full[st], (it / kAtStages) & 1);
      // S buffer `sb` was last used by tile it - 2, whose P the issuer has already waited for (p_full(it - 2) is only complete once
      // all softmax warps have read that S tile) -- no `s_empty` barrier: every mbarrier wait costs this thread ~240 cycles that the
      // tensor pipe does not hide (profiles/r02_issue_bench.md)
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t k_base = ptx::smem_u32(smem + st * kKBytes);
        const uint32_t d = tmem + kTmS + (uint32_t)(sb * kAtBN);
#pragma unroll
        for (int kc = 0; kc < 2; ++kc)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            if (kc >= kch) continue;
            const uint32_t qh = tmem + kTmQh + (uint32_t)(kc * 32 + kk * 8);
            const uint32_t ql = tmem + kTmQl + (uint32_t)(kc * 32 + kk * 8);
            const uint64_t kh = ptx::make_desc_k128(k_base + kc * 8192 + kk * 32);
            const uint64_t kl = ptx::make_desc_k128(k_base + (2 + kc) * 8192 + kk * 32);
            ptx::mma_bf16_ts(d, qh, kh, idesc_s, (kc | kk) ? 1u : 0u);
            ptx::mma_bf16_ts(d, qh, kl, idesc_s, 1u);
            ptx::mma_bf16_ts(d, ql, kh, idesc_s, 1u);
          }
        ptx::mma_commit(&k_empty[st]);
        ptx::mma_commit(&s_full[sb]);
      }
      __syncwarp();
    };
    if (nt > 0) issue_s(0);
    for (int j = 0; j < nt; ++j) {
      if (j + 1 < nt) issue_s(j + 1);
      ptx::mbar_wait(&v_full[j % kAtStages], (j / kAtStages) & 1);        // landed ~2 tiles ago: before the (long) wait for P
      ptx::mbar_wait(p_full, j & 1);
      ptx::tc_fence_after();
      if (ptx::elect_one()) {
        const uint32_t v_base = ptx::smem_u32(smem + kVOff + (j % kAtStages) * kVBytes);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ph_ = tmem + kTmPh + (uint32_t)(kk * 8);
          const uint32_t pl_ = tmem + kTmPl + (uint32_t)(kk * 8);
          // K-major V^T tile: 16 keys = 32 B inside the swizzle span; MN-major V tile: 16 keys = 16 rows = 2048 B
          const uint64_t vh = p.v_mn ? ptx::make_desc_mn128(v_base + kk * 2048, 8192, 1024) : ptx::make_desc_k128(v_base + kk * 32);
          const uint64_t vl = p.v_mn ? ptx::make_desc_mn128(v_base + 16384 + kk * 2048, 8192, 1024)
                                     : ptx::make_desc_k128(v_base + 16384 + kk * 32);
          ptx::mma_bf16_ts(tm_o, ph_, vh, idesc_o, (j | kk) ? 1u : 0u);
          ptx::mma_bf16_ts(tm_o, ph_, vl, idesc_o, 1u);
          ptx::mma_bf16_ts(tm_o, pl_, vh, idesc_o, 1u);
        }
        ptx::mma_commit(p_empty);
        ptx::mma_commit(&v_empty[j % kAtStages]);
        if (j == nt - 1) ptx::mma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    // ------------------------------------------------------------ softmax + epilogue
    // 8 warps: lane group lg = warp & 3 (the TMEM lanes a warp may touch), column half hf = (warp - 2) >> 2.
    // Thread (r, hf) owns row r and 32 of the 64 key columns of every S tile / 64 of the 128 output columns.
    const int lg = warp & 3;
    const int hf = (warp - 2) >> 2;
    const int r = lg * 32 + lane;
    const uint32_t tl = tmem + ((uint32_t)(lg * 32) << 16);
    float* xch = reinterpret_cast<float*>(smem + kBarOff + 256);        // [2][2][128] exchange of row max / row sum halves
    {
      // stage this thread's half of the query row (64 of the 128 head dims, hi and lo) into tensor memory
      uint32_t qh[32], ql[32];
      const int row = m0 + r;
      if (row < p.NQ && hf < kch) {
        const bf16* qp = p.q + ((long)b * p.q_img_rows + row) * p.q_stride + head * kAtHD + hf * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = *reinterpret_cast<const uint4*>(qp + p.q_hi + i * 8);
          const uint4 c = *reinterpret_cast<const uint4*>(qp + p.q_lo + i * 8);
          qh[i * 4] = a.x; qh[i * 4 + 1] = a.y; qh[i * 4 + 2] = a.z; qh[i * 4 + 3] = a.w;
          ql[i * 4] = c.x; ql[i * 4 + 1] = c.y; ql[i * 4 + 2] = c.z; ql[i * 4 + 3] = c.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) { qh[i] = 0u; ql[i] = 0u; }
      }
      ptx::tmem_st32(tl + kTmQh + hf * 32, qh);
      ptx::tmem_st32(tl + kTmQl + hf * 32, ql);
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(q_full);
    }
    // keys that take part: all NK, or only the visible ones of this image (TV adaptor)
    const int nkv = (p.vis_len != nullptr) ? min(p.NK, p.vis_len[b] + 1) : p.NK;
    const float* kb = (p.kbias != nullptr) ? p.kbias + (long)b * p.kbias_stride : nullptr;
    float v[32];
    float m = -INFINITY;                                     // reference maximum of this row (shared by its two column halves)
    const float sl2 = p.scale_log2e;
    float msl = m * sl2;
    float l = 0.f;
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      ptx::mbar_wait(&s_full[s], (j >> 1) & 1);
      ptx::tc_fence_after();
      ptx::tmem_ld32(tl + kTmS + s * kAtBN + hf * 32, v);
      ptx::tc_fence_before();                                    // orders the S read before this thread's p_full arrival
      const int nvalid = nkv - (t0 + j) * kAtBN - hf * 32;
      if (kb != nullptr) {
#pragma unroll
        for (int c = 0; c < 32; ++c) v[c] += __ldg(kb + (t0 + j) * kAtBN + hf * 32 + c);
      }
      const bool fast = nvalid >= 32;                        // every column of this thread's half is a real key (all but the last tile)
      float resc = 1.f;                                      // factor the O row has to be multiplied with (1 = none)
      {
        float mt = -INFINITY;
        if (fast) {
#pragma unroll
          for (int c = 0; c < 32; c += 2) mt = f_max3(mt, v[c], v[c + 1]);
        } else {
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c < nvalid) mt = fmaxf(mt, v[c]);
        }
        float* xb = xch + (j & 1) * 256;                     // double-buffered: one barrier per tile is enough
        xb[hf * 128 + r] = mt;
        asm volatile("bar.sync 1, 256;" ::: "memory");
        mt = fmaxf(mt, xb[(hf ^ 1) * 128 + r]);
        if (mt * sl2 > msl + kAtTau) {                       // also the first tile (m = -inf); both halves of a row decide alike
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(resc) : "f"(msl - mt * sl2));       // exp2(-inf) = 0 on the first tile
          l *= resc;
          m = mt;
          msl = m * sl2;
          if (j == 0) resc = 1.f;                            // O is empty: the first P V overwrites it
        }
      }
      uint32_t hi2[16], lo2[16];
      if (fast) {
        // two columns per instruction: P = exp2(S * scale*log2e - m * scale*log2e), row sum, hi / lo split (same roundings as below)
        const uint64_t sc2 = f2_pack(sl2, sl2), nm2 = f2_pack(-msl, -msl);
        uint64_t ls2 = f2_pack(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 32; c += 2) {
          float a0, a1, e0, e1;
          f2_unpack(f2_fma(f2_pack(v[c], v[c + 1]), sc2, nm2), a0, a1);
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
          const uint64_t e2 = f2_pack(e0, e1);
          ls2 = f2_add(ls2, e2);
          const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
          const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
          float d0, d1;
          f2_unpack(f2_sub(e2, f2_pack(__uint_as_float(hb << 16), __uint_as_float(hb & 0xffff0000u))), d0, d1);
          const __nv_bfloat162 l2 = __floats2bfloat162_rn(d0, d1);
          hi2[c >> 1] = hb;
          lo2[c >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
        }
        float s0, s1;
        f2_unpack(ls2, s0, s1);
        l += s0 + s1;
      } else {
#pragma unroll
      for (int c = 0; c < 32; c += 2) {
        float e0, e1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(fmaf(v[c], sl2, -msl)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(fmaf(v[c + 1], sl2, -msl)));
        if (c >= nvalid) e0 = 0.f;
        if (c + 1 >= nvalid) e1 = 0.f;
        l += e0 + e1;
        const __nv_bfloat162 h2 = __floats2bfloat162_rn(e0, e1);
        const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
        const __nv_bfloat162 l2 = __floats2bfloat162_rn(e0 - __uint_as_float(hb << 16), e1 - __uint_as_float(hb & 0xffff0000u));
        hi2[c >> 1] = hb;
        lo2[c >> 1] = *reinterpret_cast<const uint32_t*>(&l2);
      }
      }
      ptx::mbar_wait(p_empty, (j & 1) ^ 1);               // P(j-1) V(j-1) has consumed the previous P
      ptx::tc_fence_after();
      if (__any_sync(0xffffffffu, resc != 1.f)) {  // O is stable here: P(j-1) V(j-1) has retired, P(j) V(j) waits for p_full
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          if (hf >= p.vchunks) break;
          float o[32];
          ptx::tmem_ld32(tl + kTmO + hf * 64 + c * 32, o);
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] *= resc;
          ptx::tmem_st32(tl + kTmO + hf * 64 + c * 32, *reinterpret_cast<uint32_t(*)[32]>(o));
        }
      }
      ptx::tmem_st16(tl + kTmPh + hf * 16, hi2);
      ptx::tmem_st16(tl + kTmPl + hf * 16, lo2);
      ptx::tmem_wait_st();
      ptx::tc_fence_before();
      ptx::mbar_arrive(p_full);
    }
    asm volatile("bar.sync 1, 256;" ::: "memory");          // everyone has read the max exchange
    xch[hf * 128 + r] = l;                                  // (the two halves of a row hold the same m)
    asm volatile("bar.sync 1, 256;" ::: "memory");
    l += xch[(hf ^ 1) * 128 + r];
    const int row = m0 + r;
    if (omode == 2) {
      // split-KV partials: O (un-normalised, fp32), row sum l and row max m of this key range
      float* po = p.part_o + (part_idx * kAtBM + r) * kAtHD;
      if (hf == 0) {
        p.part_l[part_idx * kAtBM + r] = l;
        p.part_m[part_idx * kAtBM + r] = m;
      }
      if (nt > 0) {
        ptx::mbar_wait(o_full, 0);
        ptx::tc_fence_after();
      }
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        if (hf >= p.vchunks) break;                           // output narrower than 128 columns (linear attention, C = 64)
        float o[32];
        if (nt > 0) ptx::tmem_ld32(tl + kTmO + hf * 64 + c * 32, o);
        else {
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = 0.f;
        }
#pragma unroll
        for (int i = 0; i < 32; i += 8) st256_f32(po + hf * 64 + c * 32 + i, &o[i]);
      }
    } else {
    ptx::mbar_wait(o_full, 0);
    ptx::tc_fence_after();
    const float inv = 1.f / l;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      float o[32];
      ptx::tmem_ld32(tl + kTmO + hf * 64 + c * 32, o);
      if (row < p.NQ) {
        const long grow = (long)b * p.NQ + row;
        const int col = head * kAtHD + hf * 64 + c * 32;
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] *= inv;
        if (omode == 0) {
          bf16* op = p.out + grow * p.out_stride + col;
#pragma unroll
          for (int i = 0; i < 32; i += 16) store_split16(op + p.out_hi + i, op + p.out_lo + i, &o[i]);
        } else {
          const bf16* qp = p.q + grow * p.q_stride + col;          // residual = the query row itself
          const float rm = (p.rowmask != nullptr) ? p.rowmask[(long)b * p.rowmask_w + row % p.rowmask_w] : 1.f;
#pragma unroll
          for (int i = 0; i < 32; i += 8) {
            float x[8];
            load_split8(qp + p.q_hi + i, qp + p.q_lo + i, x);
#pragma unroll
            for (int k = 0; k < 8; ++k) o[i + k] = (x[k] + o[i + k]) * rm;
          }
          float* of = p.out_f + grow * p.out_f_stride + col;
#pragma unroll
          for (int i = 0; i < 32; i += 8) st256_f32(of + i, &o[i]);
        }
      }
    }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<512>(tmem);
}

// ------------------------------------------------------------------------------------------------
// host
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_enc = nullptr;

static int enc3(CUtensorMap* tm, const void* base, cuuint64_t d0, cuuint64_t d1, cuuint64_t d2, cuuint64_t s1_bytes,
                cuuint64_t s2_bytes, cuuint32_t b0, cuuint32_t b1, const char* what) {
  if (g_enc == nullptr) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    DEXB_CUDA_OK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
    DEXB_CHECK(q == cudaDriverEntryPointSuccess && fn != nullptr, "cuTensorMapEncodeTiled not available");
    g_enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  const cuuint64_t dims[3] = {d0, d1, d2};
  const cuuint64_t str[2] = {s1_bytes, s2_bytes};
  const cuuint32_t box[3] = {b0, b1, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, str, box, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  DEXB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(attention %s) failed with CUresult %d", what, (int)r);
  return 0;
}

int attn_global_init() {
  DEXB_CUDA_OK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAtSmem));
  return 0;
}

bool attn_supported(int hd) { return hd == kAtHD; }


int attn_plan_init(AttnPlan* ap, const bf16* qkv, const bf16* vT, bf16* out, int B, int N, int NP, int heads, int hid) {
  DEXB_CHECK(hid / heads == kAtHD, "fused attention is instantiated for head dim %d", kAtHD);
  ap->B = B;
  AttnParams& p = ap->p;
  memset(&p, 0, sizeof(p));
  p.NQ = N; p.NK = N; p.KP = NP; p.nheads = heads;
  p.nt = (N + kAtBN - 1) / kAtBN;
  p.q = qkv; p.q_stride = 6L * hid; p.q_hi = 0; p.q_lo = 3 * hid; p.q_img_rows = N;
  p.kchunks = 2; p.vchunks = 2; p.kv_splits = 1; p.tiles_per_split = p.nt;
  p.q_tiles = (N + kAtBM - 1) / kAtBM; p.tail_first = B * heads * p.q_tiles; p.tail_splits = 1; p.tail_tps = p.nt;
  p.k_hi = hid; p.k_lo = 4 * hid;
  p.scale_log2e = (1.f / sqrtf((float)kAtHD)) * 1.4426950408889634f;
  p.out_mode = 0;
  p.out = out; p.out_stride = 2L * hid; p.out_hi = 0; p.out_lo = hid;
  const cuuint64_t qrow = 6ull * hid * 2;
  DEXB_TRY(enc3(&ap->tmK, qkv, 6ull * hid, (cuuint64_t)N, (cuuint64_t)B, qrow, qrow * N, 64, kAtBN, "K"));
  if (vT == nullptr) {                                  // V read row-major straight from the qkv rows (MN-major B operand)
    p.v_mn = 1; p.v_hi = 2 * hid; p.v_lo = 5 * hid;
    DEXB_TRY(enc3(&ap->tmV, qkv, 6ull * hid, (cuuint64_t)N, (cuuint64_t)B, qrow, qrow * N, 64, kAtBN, "V rows"));
    return 0;
  }
  const cuuint64_t vrow = 2ull * NP * 2;
  DEXB_TRY(enc3(&ap->tmV, vT, 2ull * NP, (cuuint64_t)hid, (cuuint64_t)B, vrow, vrow * hid, 64, kAtHD, "V"));
  return 0;
}

int attn_plan_init_tv(AttnPlan* ap, const bf16* x, long x_stride, int x_hi, int x_lo, const bf16* kq, const float* sbias,
                      const bf16* vlt, const int* sty_len, float* out, const float* mask, int mask_w, int B, int P, int NK, int KP,
                      int C) {
  DEXB_CHECK(C == kAtHD, "fused TV attention is instantiated for %d channels", kAtHD);
  DEXB_CHECK(KP % kAtBN == 0 && NK <= KP, "TV attention: key padding %d / %d", NK, KP);
  ap->B = B;
  AttnParams& p = ap->p;
  memset(&p, 0, sizeof(p));
  p.NQ = P; p.NK = NK; p.KP = KP; p.nheads = 1;
  p.nt = (NK + kAtBN - 1) / kAtBN;
  p.q = x; p.q_stride = x_stride; p.q_hi = x_hi; p.q_lo = x_lo; p.q_img_rows = P;
  p.kchunks = 2; p.vchunks = 2; p.kv_splits = 1; p.tiles_per_split = p.nt;
  p.k_hi = 0; p.k_lo = C;
  p.scale_log2e = 1.4426950408889634f;              // 1/sqrt(C) is folded into the key matrix (k_tv_fold)
  p.kbias = sbias; p.kbias_stride = KP; p.vis_len = sty_len;
  p.out_mode = 1;
  p.out_f = out; p.out_f_stride = C; p.rowmask = mask; p.rowmask_w = mask_w;
  const cuuint64_t krow = 2ull * C * 2;
  DEXB_TRY(enc3(&ap->tmK, kq, 2ull * C, (cuuint64_t)KP, (cuuint64_t)B, krow, krow * KP, 64, kAtBN, "TV K"));
  const cuuint64_t vrow = 2ull * KP * 2;
  DEXB_TRY(enc3(&ap->tmV, vlt, 2ull * KP, (cuuint64_t)C, (cuuint64_t)B, vrow, vrow * C, 64, kAtHD, "TV V"));
  return 0;
}

int attn_plan_init_la(AttnPlan* ap, const bf16* wk, const bf16* x, long x_stride, int x_hi, int x_lo, float* part_o,
                      float* part_l, float* part_m, int B, int P, int PP, int C, int splits) {
  DEXB_CHECK(C == 64 || C == 128, "linear-attention context: C must be 64 or 128 (got %d)", C);
  DEXB_CHECK(PP % kAtBN == 0 && PP >= P && splits >= 1, "linear-attention context: bad padding / split");
  ap->B = B;
  AttnParams& p = ap->p;
  memset(&p, 0, sizeof(p));
  p.NQ = 128; p.NK = P; p.KP = PP; p.nheads = 1;
  p.nt = (P + kAtBN - 1) / kAtBN;
  p.kv_splits = splits;
  p.tiles_per_split = (p.nt + splits - 1) / splits;
  p.kchunks = C / 64;
  p.vchunks = C / 64;
  p.q = wk; p.q_stride = 2L * C; p.q_hi = 0; p.q_lo = C; p.q_img_rows = 0;      // the k rows of to_qkv, shared by all images
  p.k_hi = x_hi; p.k_lo = x_lo;
  p.scale_log2e = 1.4426950408889634f;
  p.out_mode = 2;
  p.part_o = part_o; p.part_l = part_l; p.part_m = part_m;
  const cuuint64_t xrow = (cuuint64_t)x_stride * 2;
  DEXB_TRY(enc3(&ap->tmK, x, (cuuint64_t)x_stride, (cuuint64_t)P, (cuuint64_t)B, xrow, xrow * P, 64, kAtBN, "LA x"));
  // "values" = the same activation rows (MN-major B operand, 64 pixels x 64 channels per box)
  p.v_mn = 1; p.v_hi = x_hi; p.v_lo = x_lo;
  DEXB_TRY(enc3(&ap->tmV, x, (cuuint64_t)x_stride, (cuuint64_t)P, (cuuint64_t)B, xrow, xrow * P, 64, kAtBN, "LA x as V"));
  return 0;
}

// Merge of the tail-split partials: out = sum_s f_s O_s / sum_s f_s l_s,  f_s = exp2((m_s - max_s m_s) * scale*log2e)  -> split rows.
// One block per split tile, thread = (row, half of the 128 output columns).
__global__ void __launch_bounds__(256) k_attn_tail_merge(const AttnParams p) {
  pdl_wait();
  const int k = blockIdx.x, tile = p.tail_first + k;
  const int z = tile / p.q_tiles, m0 = (tile % p.q_tiles) * kAtBM;
  const int b = z / p.nheads, head = z % p.nheads;
  const int r = threadIdx.x >> 1, hf = threadIdx.x & 1;
  const int row = m0 + r;
  if (row >= p.NQ) return;
  const int S = p.tail_splits;
  // (a chain of L2 round trips on 40 blocks: the maxima of all splits (S <= 8) and the partials of two splits are in flight together)
  float M = -INFINITY;
#pragma unroll 8
  for (int s = 0; s < S; ++s) M = fmaxf(M, p.part_m[((long)k * S + s) * kAtBM + r]);
  float o[64];
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] = 0.f;
  float l = 0.f;
#pragma unroll 2
  for (int s = 0; s < S; ++s) {
    const long pi = ((long)k * S + s) * kAtBM + r;
    const float f = exp2f((p.part_m[pi] - M) * p.scale_log2e);      // exp2(-inf) = 0 for an empty split
    l = fmaf(f, p.part_l[pi], l);
    const float* po = p.part_o + pi * kAtHD + hf * 64;
#pragma unroll
    for (int i = 0; i < 64; i += 8) {
      float v[8];
      ld256_f32(po + i, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[i + j] = fmaf(f, v[j], o[i + j]);
    }
  }
  const float inv = 1.f / l;
#pragma unroll
  for (int i = 0; i < 64; ++i) o[i] *= inv;
  bf16* op = p.out + ((long)b * p.NQ + row) * p.out_stride + head * kAtHD + hf * 64;
#pragma unroll
  for (int i = 0; i < 64; i += 16) store_split16(op + p.out_hi + i, op + p.out_lo + i, &o[i]);
}

static int attn_sm_count() {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}
// tiles beyond the last full wave, and how many key ranges each is cut into so that they all run at once
static void attn_tail_shape(int total_tiles, int nt, int* first, int* splits, int* tps) {
  const int sms = attn_sm_count();
  const int rem = total_tiles % sms;
  *first = total_tiles; *splits = 1; *tps = nt;
  if (rem == 0 || total_tiles < sms) return;
  int S = sms / rem;
  if (S > 8) S = 8;
  if (S > nt) S = nt;
  if (S < 2) return;
  const int t = (nt + S - 1) / S;
  *first = total_tiles - rem; *splits = (nt + t - 1) / t; *tps = t;     // no empty ranges
  if (*splits < 2) { *first = total_tiles; *splits = 1; *tps = nt; }
}
long attn_tail_scratch_floats(int B, int N, int heads) {
  const int qt = (N + kAtBM - 1) / kAtBM, nt = (N + kAtBN - 1) / kAtBN;
  int first, S, tps;
  attn_tail_shape(B * heads * qt, nt, &first, &S, &tps);
  if (S < 2) return 0;
  return (long)(B * heads * qt - first) * S * (kAtBM * kAtHD + 2 * kAtBM);
}
void attn_plan_set_tail(AttnPlan* ap, float* scratch) {
  AttnParams& p = ap->p;
  const int total = ap->B * p.nheads * p.q_tiles;
  int first, S, tps;
  attn_tail_shape(total, p.nt, &first, &S, &tps);
  const char* e = getenv("DEXB_ATTN_TAIL");
  if (S < 2 || scratch == nullptr || (e != nullptr && e[0] == '0')) return;
  p.tail_first = first; p.tail_splits = S; p.tail_tps = tps;
  const long n = (long)(total - first) * S;
  p.part_o = scratch; p.part_l = scratch + n * kAtBM * kAtHD; p.part_m = p.part_l + n * kAtBM;
}

int attn_launch(const AttnPlan& ap, cudaStream_t st) {
  const AttnParams& p = ap.p;
  if (p.tail_splits > 1) {
    const int total = ap.B * p.nheads * p.q_tiles, rem = total - p.tail_first;
    launch_pdl(attn_fwd_kernel, dim3(p.tail_first + rem * p.tail_splits), dim3(kAtThreads), kAtSmem, st, ap.tmK, ap.tmV, p);
    launch_pdl(k_attn_tail_merge, dim3((unsigned)(rem)), dim3(256), 0, st, p);
    DEXB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  dim3 grid((unsigned)(ap.p.kv_splits > 1 ? ap.p.kv_splits : (ap.p.NQ + kAtBM - 1) / kAtBM), (unsigned)(ap.B * ap.p.nheads));
  launch_pdl(attn_fwd_kernel, grid, dim3(kAtThreads), kAtSmem, st, ap.tmK, ap.tmV, ap.p);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}
int attn_launch_count(const AttnPlan& ap) { return ap.p.tail_splits > 1 ? 2 : 1; }

double attn_flop(const AttnPlan& ap) {
  return 2.0 * ap.B * ap.p.nheads * (double)ap.p.NQ * ap.p.NK * 64.0 * (ap.p.kchunks + ap.p.vchunks);
}

}  // namespace dexb
