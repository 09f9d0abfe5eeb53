// Implicit-GEMM engine for every contraction on the reverse-diffusion path.
//
//   D[z][pixel][n] = sum_{tap, k}  A[img(z)][pixel + off(tap)][k] * Bw[tap][n][k]        (+ fused epilogue)
//
// A is a split-bf16 ("S") activation image [img][H][W][row]; a tap is a spatial offset (3x3 convs: 9 taps, 1x1 /
// linear layers: 1 tap, pos-conv 16x16: 256 taps, ConvTranspose phases: 2x2 taps).  Bw are pre-packed split-bf16
// weights [tap][n][hi(K)|lo(K)] (optionally one matrix per sample / per (sample, head)).
// Precision: bf16x3 -- Ah*Bh + Ah*Bl + Al*Bh accumulated in fp32 (TMEM) -- which reproduces the fp32 product to
// ~2^-17 and keeps the 50-step trajectory inside the 1e-3 parity bound (plain bf16 / tf32 do not).
//
// Two interchangeable engines consume the same GemmParams and the same epilogue:
//   * gemm_tc_kernel  : TMA (tiled, OOB zero-fill = conv padding) -> 128B-swizzled smem -> tcgen05.mma, fp32
//                       accumulators in TMEM, warp-specialised (1 TMA warp, 1 MMA warp, 4 epilogue warps).
//   * gemm_simt_kernel: plain CUDA-core fallback / on-device cross-check (also covers strided input).
#pragma once
#include <cuda.h>

#include "common.cuh"
#include "gemm_host.cuh"
#include "gn_apply.cuh"
#include "ptx.cuh"

namespace dexb {

// ------------------------------------------------------------------------------------------------
// Epilogue shared by both engines: one output row, NV consecutive columns starting at n0.
// Must be called by all 32 lanes of a warp with the same n0 (it shuffles); `valid` masks stores.
// ------------------------------------------------------------------------------------------------
// `gacc` (optional): per-thread GroupNorm partial sums [NV/8][2] that the caller keeps across tiles of one image and
// flushes with epi_flush_gn -- 17x fewer double atomics (and no shuffles per tile) than accumulating per tile.
//
// FAST = true is the instantiation every shipped shape runs: the host has verified (epi_fast_ok) that N is a multiple of 32 and
// that every pointer / stride / column offset allows the widest vector access, so the scalar fallbacks, the transposed store and
// the column-mean atomics are compiled out.  With them the kernel was 10 800 SASS instructions (170 KB, far beyond the
// instruction caches; 17 % of the warp stall samples were instruction fetches -- profiles/r01_ncu_gemm_epilogue.md).
//
// The residual rows are the only operand that needs a full L2 round trip and does not depend on the accumulator:
// epi_load_resid issues those loads BEFORE the caller waits for the tile / the tensor-memory load, so the latencies overlap.
template <int NV, bool FAST>
__device__ __forceinline__ void epi_load_resid(const EpiParams& e, int N, int z, int nheads, int oh, int ow, int OH, int OW,
                                               bool valid, int n0, float (&r)[NV]) {
#pragma unroll
  for (int i = 0; i < NV; ++i) r[i] = 0.f;
  if (!valid || (e.resid_f32 == nullptr && e.resid_s == nullptr)) return;
  const int head = z % nheads;
  const int img = e.o_by_z ? z : z / nheads;
  const long row = ((long)img * OH + oh) * OW + ow;
  const int nc0 = n0 + head * e.o_head_stride;
  const bool full = FAST || (n0 + NV <= N);
  if (e.resid_f32 != nullptr) {
    const float* rp = e.resid_f32 + row * e.resid_f32_stride + nc0;
    if (FAST || (NV % 8 == 0 && full && ((reinterpret_cast<uintptr_t>(rp) & 31) == 0))) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) ld256_f32(rp + i, &r[i]);
    } else if constexpr (!FAST) {
      if (full && ((reinterpret_cast<uintptr_t>(rp) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < NV; i += 4) {
          const float4 r4 = *reinterpret_cast<const float4*>(rp + i);
          r[i] = r4.x; r[i + 1] = r4.y; r[i + 2] = r4.z; r[i + 3] = r4.w;
        }
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) if (n0 + i < N) r[i] = rp[i];
      }
    }
  }
  if (e.resid_s != nullptr) {
    const bf16* rp = e.resid_s + row * e.resid_s_stride + nc0;
    if (FAST || (full && (((reinterpret_cast<uintptr_t>(rp + e.resid_s_hi) | reinterpret_cast<uintptr_t>(rp + e.resid_s_lo)) & 15) == 0))) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) {
        float t[8];
        load_split8(rp + e.resid_s_hi + i, rp + e.resid_s_lo + i, t);
#pragma unroll
        for (int j = 0; j < 8; ++j) r[i + j] += t[j];
      }
    } else if constexpr (!FAST) {
#pragma unroll
      for (int i = 0; i < NV; ++i) if (n0 + i < N) r[i] += join2(rp[e.resid_s_hi + i], rp[e.resid_s_lo + i]);
    }
  }
}

template <int NV, bool FAST = false>
__device__ __forceinline__ void epi_apply(const EpiParams& e, int N, int z, int nheads, int oh, int ow, int OH, int OW,
                                          bool valid, int n0, float (&v)[NV], const float (&r)[NV], float* gacc = nullptr) {
  // Every branch below is warp-uniform and taken once per NV-column chunk; the per-element loops are branch-free
  // (a first version tested the optional features per element and was instruction-issue bound: ~45 SASS instructions
  // per output element, 20 % of the stalls on instruction fetch -- profiles/r01_ncu_gemm_v1.md).
  const int head = z % nheads;
  const int img = e.o_by_z ? z : z / nheads;
  const long row = ((long)img * OH + oh) * OW + ow;
  const int nc0 = n0 + head * e.o_head_stride;     // output column of v[0]
  const bool full = FAST || (n0 + NV <= N);
  if (e.alpha != 1.f) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= e.alpha;
  }
  if (e.bias != nullptr) {
    const float* bp = e.bias + (long)z * e.bias_zstride + head * e.bias_head_stride + n0;
    if (FAST || (full && ((reinterpret_cast<uintptr_t>(bp) & 15) == 0))) {
#pragma unroll
      for (int i = 0; i < NV; i += 4) {
        const float4 b4 = __ldg(reinterpret_cast<const float4*>(bp + i));
        v[i] += b4.x; v[i + 1] += b4.y; v[i + 2] += b4.z; v[i + 3] += b4.w;
      }
    } else if constexpr (!FAST) {
#pragma unroll
      for (int i = 0; i < NV; ++i) if (n0 + i < N) v[i] += bp[i];
    }
  }
  if constexpr (!FAST) {
    if (!full) {
#pragma unroll
      for (int i = 0; i < NV; ++i) if (n0 + i >= N) v[i] = 0.f;
    }
  }
  if (e.act == 1) {
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] = gelu_fast(v[i]);       // gelu(0) == 0 keeps the out-of-range columns at zero
  }
  if (valid && (e.resid_f32 != nullptr || e.resid_s != nullptr || e.gate != nullptr)) {
    if (e.gate != nullptr) {
      const float* gp = e.gate + n0;
      if (FAST || (full && ((reinterpret_cast<uintptr_t>(gp) & 15) == 0))) {
#pragma unroll
        for (int i = 0; i < NV; i += 4) {
          const float4 g4 = __ldg(reinterpret_cast<const float4*>(gp + i));
          v[i] = fmaf(g4.x, v[i], r[i]); v[i + 1] = fmaf(g4.y, v[i + 1], r[i + 1]);
          v[i + 2] = fmaf(g4.z, v[i + 2], r[i + 2]); v[i + 3] = fmaf(g4.w, v[i + 3], r[i + 3]);
        }
      } else if constexpr (!FAST) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = fmaf((n0 + i < N) ? __ldg(gp + i) : 0.f, v[i], r[i]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] += r[i];
    }
  }
  if (e.rowmask != nullptr && valid) {
    const float rm = e.rowmask[(long)img * e.rowmask_stride + ow];
#pragma unroll
    for (int i = 0; i < NV; ++i) v[i] *= rm;
  }
  if (e.gn_stats != nullptr && gacc != nullptr) {
    if (valid) {
#pragma unroll
      for (int g0 = 0; g0 < NV; g0 += 8) {
        float s = 0.f, ss = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += v[g0 + i]; ss = fmaf(v[g0 + i], v[g0 + i], ss); }
        gacc[(g0 >> 3) * 2] += s;
        gacc[(g0 >> 3) * 2 + 1] += ss;
      }
    }
  } else if (e.gn_stats != nullptr) {
    // per-(image, group) sum / sum-of-squares: thread-local over its channels, warp-shuffle over the 32 rows
    // of this warp (all rows of a tile belong to one image), one double atomic per (warp, group).
    const int gs = e.gn_gs;
#pragma unroll
    for (int g0 = 0; g0 < NV; g0 += 8) {
      float s = 0.f, ss = 0.f;
      if (valid) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { s += v[g0 + i]; ss += v[g0 + i] * v[g0 + i]; }
      }
      s = warp_sum(s);
      ss = warp_sum(ss);
      if ((threadIdx.x & 31) == 0 && n0 + g0 < N) {
        double* dst = e.gn_stats + (((long)img * kGnRep + (blockIdx.x & (kGnRep - 1))) * (N / gs) + (n0 + g0) / gs) * 2;
        atomicAdd(dst, (double)s);
        atomicAdd(dst + 1, (double)ss);
      }
    }
  }
  if (!valid || e.dbg_nostore) return;
  if (e.out_f32 != nullptr) {
    float* op = e.out_f32 + row * e.out_f32_stride + e.out_f32_col + nc0;
    if (FAST || (NV % 8 == 0 && n0 + NV <= N && ((reinterpret_cast<uintptr_t>(op) & 31) == 0))) {
#pragma unroll
      for (int i = 0; i < NV; i += 8) st256_f32(op + i, &v[i]);
    } else if constexpr (!FAST) {
      if (n0 + NV <= N && ((reinterpret_cast<uintptr_t>(op) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < NV; i += 4) *reinterpret_cast<float4*>(op + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i) if (n0 + i < N) op[i] = v[i];
      }
    }
  }
  if (e.out_s != nullptr && n0 < e.out_s_ncols) {
    if (e.s_lrelu != 0.f) {
#pragma unroll
      for (int i = 0; i < NV; ++i) v[i] = v[i] > 0.f ? v[i] : v[i] * e.s_lrelu;
    }
    const int cs = (e.out_s_gshift > 0) ? (nc0 >> e.out_s_gshift) * e.out_s_gpitch + (nc0 & ((1 << e.out_s_gshift) - 1)) : nc0;
    bf16* hp = e.out_s + row * e.out_s_stride + e.out_s_hi + cs;
    bf16* lp = e.out_s + row * e.out_s_stride + e.out_s_lo + cs;
    if (FAST || (NV % 16 == 0 && n0 + NV <= N && n0 + NV <= e.out_s_ncols && ((reinterpret_cast<uintptr_t>(hp) & 31) == 0) &&
                 ((reinterpret_cast<uintptr_t>(lp) & 31) == 0))) {
#pragma unroll
      for (int i = 0; i < NV; i += 16) store_split16(hp + i, lp + i, &v[i]);
    } else if constexpr (!FAST) {
      if (n0 + NV <= N && n0 + NV <= e.out_s_ncols && ((reinterpret_cast<uintptr_t>(hp) & 15) == 0) &&
          ((reinterpret_cast<uintptr_t>(lp) & 15) == 0)) {
#pragma unroll
        for (int i = 0; i < NV; i += 8) store_split8(hp + i, lp + i, &v[i]);
      } else {
#pragma unroll
        for (int i = 0; i < NV; ++i)
          if (n0 + i < N && n0 + i < e.out_s_ncols) split2(v[i], hp[i], lp[i]);
      }
    }
  }
  if constexpr (!FAST) {
    if (e.out_vt != nullptr && n0 + NV > e.out_s_ncols) {
      // transposed store (V of the attention): consecutive lanes hold consecutive tokens -> coalesced per d
      const long token = (long)oh * OW + ow;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const int n = n0 + i;
        if (n >= e.out_s_ncols && n < N) {
          const int c = n - e.out_s_ncols;
          const int hd_i = c / e.out_vt_hd, d = c % e.out_vt_hd;
          bf16* base = e.out_vt + ((long)img * e.out_vt_heads + hd_i) * e.out_vt_zstride + (long)d * e.out_vt_rstride + token;
          split2(v[i], base[0], base[e.out_vt_lo]);
        }
      }
    }
    if (e.colmean != nullptr) {
      float* cp = e.colmean + ((long)img * OW + ow) * e.colmean_ld + nc0;
#pragma unroll
      for (int i = 0; i < NV; ++i)
        if (n0 + i < N) atomicAdd(cp + i, v[i] * e.colmean_scale);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Note on store coalescing (measured, profiles/r01_epilogue_experiments.md): tcgen05.ld hands every thread one ROW of
// the tile, so a warp-level float4 store touches 32 different cache lines.  Transposing each 32x32 chunk through shared
// memory so that lane = column (128 contiguous bytes per warp store) was 4-5x SLOWER: it needs 4x as many store
// instructions, and the epilogue is bound by the number of store instructions a warp can have in flight, not by
// sectors.  The engine therefore keeps the row-per-thread 16 B stores and doubles the number of epilogue warps instead.
// ------------------------------------------------------------------------------------------------
// V of the attention, stored transposed (vT[img][head][d][token]) from the row-per-thread registers: consecutive lanes
// hold consecutive tokens, so every store of the warp is one contiguous 64 B run.  A 32-column chunk never straddles heads.
template <int NV>
__device__ __forceinline__ void epi_store_vt(const EpiParams& e, int N, int z, int nheads, long token, bool valid, int n0,
                                             const float (&v)[NV]) {
  if (!valid) return;
  const int img = z / nheads;
  const int c0 = n0 - e.out_s_ncols;
  const int hd_i = c0 / e.out_vt_hd, d0 = c0 % e.out_vt_hd;
  bf16* base = e.out_vt + ((long)img * e.out_vt_heads + hd_i) * e.out_vt_zstride + (long)d0 * e.out_vt_rstride + token;
  const float* bp = (e.bias != nullptr) ? e.bias + n0 : nullptr;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (n0 + i < N) {
      const float x = v[i] * e.alpha + (bp != nullptr ? __ldg(bp + i) : 0.f);
      bf16* q = base + (long)i * e.out_vt_rstride;
      split2(x, q[0], q[e.out_vt_lo]);
    }
  }
}

// warp 0: TMA, warp 1: MMA (+TMEM alloc), warps 2..: epilogue.  kTcEpiPerLG warps share each TMEM lane group and take every
// kTcEpiPerLG-th 32-column chunk.  The epilogue is latency-bound (each warp issues one instruction per ~9 cycles: dependent
// chains of the activation / split conversion, tensor-memory and store latencies), and for the K = 256 GEMMs it -- not the MMA
// stream -- sets the tile time, so four warps per lane group (16 epilogue warps, <= 112 registers) beat two.
constexpr int kTcEpiPerLG = 4;
constexpr int kTcEpiCW = (kTcEpiPerLG == 4) ? 16 : 32;      // columns per epilogue chunk: 16 keeps 16 warps inside 96 registers
constexpr int kTcEpiThreads = 128 * kTcEpiPerLG;
constexpr int kTcThreads = 64 + kTcEpiThreads;

// Flush the deferred GroupNorm partial sums of one warp: gacc[k][g][2] for the warp's k-th chunk (columns n0 = 64 k +
// 32 half when there is a single n-tile) -> warp reduction -> one double atomic per (warp, 8-column group).
template <int MAXCH>
__device__ __forceinline__ void epi_flush_gn(const EpiParams& e, int N, int img, int half, int nper, float (&gacc)[MAXCH][8]) {
  const int gs = e.gn_gs;
#pragma unroll
  for (int k = 0; k < MAXCH; ++k) {
    const int n0 = (half + nper * k) * kTcEpiCW;
#pragma unroll
    for (int g = 0; g < kTcEpiCW / 8; ++g) {
      const float s = warp_sum(gacc[k][g * 2]);
      const float ss = warp_sum(gacc[k][g * 2 + 1]);
      gacc[k][g * 2] = 0.f;
      gacc[k][g * 2 + 1] = 0.f;
      if ((threadIdx.x & 31) == 0 && n0 + g * 8 < N) {
        double* dst = e.gn_stats + (((long)img * kGnRep + (blockIdx.x & (kGnRep - 1))) * (N / gs) + (n0 + g * 8) / gs) * 2;
        atomicAdd(dst, (double)s);
        atomicAdd(dst + 1, (double)ss);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// tcgen05 engine
// ------------------------------------------------------------------------------------------------
constexpr int kTcBlockM = 128;

template <int BLOCK_N>
struct TcSmem {
  static constexpr int kABytes = kTcBlockM * kTcBlockK * 2;       // 16 KiB per (hi | lo) tile
  static constexpr int kBBytes = BLOCK_N * kTcBlockK * 2;
  static constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;
  static constexpr int kStages = (BLOCK_N >= 256) ? 2 : (BLOCK_N >= 128 ? 3 : 4);
  // +1024 for manual 1024 B alignment (SWIZZLE_128B atoms), +256 for barriers / tmem pointer
  static constexpr int kBytes = kStages * kStageBytes + 1024 + 256;
};
// Resident-B mode ("rb" > 0 = number of A stages): all taps x k-chunks of the CTA's weight tile are loaded ONCE into shared memory
// and the ring only carries A (32 KiB per stage instead of 48-64).  The grid is a multiple of the n-tile count, so t += gridDim.x
// keeps a CTA on one n-tile for its whole life.  Small-K GEMMs (DiT linears: K = 256) are bound by the SM's L2->SMEM ingest
// (~64 B/clk: `neither` column of profiles/r01_epilogue_experiments.md), and half of their ingest was the re-streamed weight tile.
constexpr int kTcMaxStages = 8;
// Halo mode (rb < 0, -rb = number of weight stages): 3x3 stride-1 convolutions on tiles of 1 x 128 pixels.  One staged input
// row slab of 136 pixels (x0 - 1 ... x0 + 134) serves the three dx taps of its dy: the A descriptor of tap dx simply starts dx
// rows (dx * 128 B) into the slab -- SWIZZLE_128B is a function of the absolute shared-memory address, so a row-shifted start
// is legal (tools/mma_bench.cu, `row-shifted A descriptor`).  A traffic drops from 9 to 3.2 tile-loads per tile; with the ring
// at 192 KiB the operand stream of the plain path is latency-bound (bytes in flight / TMA latency ~ 64 B/clk per SM).
constexpr int kTcHaloRows = 136;                                   // slab rows: 128 + 2 halo, rounded up to 8
constexpr int kTcHaloSlab = kTcHaloRows * kTcBlockK * 2;           // 17 408 B per (hi | lo) slab, a multiple of 1024
constexpr int kTcHaloASlots = 3;
__host__ __device__ constexpr int tc_halo_bytes(int block_n, int nb) {
  return kTcHaloASlots * 2 * kTcHaloSlab + nb * 2 * block_n * kTcBlockK * 2 + 1024 + 256;
}
constexpr int kTcSmemMax = 232448;             // 227 KiB: the most a CTA can opt in to
__host__ __device__ constexpr int tc_rb_bytes(int block_n, int nk, int rb) {
  return nk * 2 * block_n * kTcBlockK * 2 + rb * 2 * kTcBlockM * kTcBlockK * 2 + 1024 + 256;
}

// Persistent: grid = min(#tiles, #SMs); every CTA walks tiles t = blockIdx.x, +gridDim.x, ... (n-tile fastest, so CTAs
// that run concurrently share the A tile in L2).  The smem ring and both TMEM accumulator buffers stay live across
// tiles, so the MMA warp starts tile i+1 while the epilogue warps drain tile i (acc_full / acc_empty barriers).
struct TcTile {
  int z, head, img_a, ch0, cw0, n0;
};
__device__ __forceinline__ TcTile tc_decode_tile(const GemmParams& p, int t, int ntn, int block_n) {
  TcTile r;
  r.n0 = (t % ntn) * block_n; t /= ntn;
  const int tw = t % p.TW; t /= p.TW;
  const int th = t % p.TH; t /= p.TH;
  r.z = t;
  r.head = r.z % p.nheads;
  r.img_a = p.a_by_z ? r.z : r.z / p.nheads;
  r.ch0 = th * p.BH; r.cw0 = tw * p.BW;
  return r;
}

// GroupNorm-apply fused into the producing convolution (GNF = true; convolutions whose epilogue defers the GroupNorm sums, i.e. one
// n-tile).  Tiles are walked image-major, so all CTAs finish image i at about the same time.  When the epilogue warps of a CTA move
// on to a new image they (1) flush their GroupNorm partial sums of the image they leave and publish its tile count
// (__threadfence, named barrier of the 16 epilogue warps, ONE atomicAdd on done[img]), and (2) apply their share (1 / gridDim.x) of
// every image that lies at least two images back: its counter is long complete and its raw fp32 rows are still in the 126 MB L2
// (10.5 MB per level-0 image), so the stand-alone k_gn_apply pass -- an 84 MB HBM read + one launch per convolution, 10.6 % of a
// network call -- is replaced by L2 reads issued in the idle time of the epilogue warps (a convolution tile is 9 x K/64 stages of
// MMAs but only one 16-column chunk per epilogue warp).  The last images are applied after the CTA's last tile, waiting for the
// other CTAs' counters.  No deadlock: the grid is <= #SMs with one CTA per SM (all co-resident), a CTA only waits for images it
// has left behind, and the CTAs still holding tiles of the smallest unfinished image never wait for it.
struct GnFuseNone {};
template <bool GNF> struct GnFuseArg { using type = GnFuseNone; };
template <> struct GnFuseArg<true> { using type = GnFuse; };

// A fused kernel runs 2 epilogue warps per TMEM lane group instead of 4 (320 threads): a convolution tile is 9 x K/64 stages of MMAs
// against ONE 16-column chunk per warp, so half the warps drain it just as well, and the block may then use ~200 registers per
// thread -- the apply code (24 affine coefficients + four items in flight) lives beside the epilogue without a single spill.
// That matters more than usual here: the kernel takes the whole 227 KB shared-memory carve-out, which leaves no L1 -- every spill
// or stack access is an L2 round trip (a first version that called the apply as a __noinline__ function, and one that kept it
// inline at 96 registers, made the convolution itself 25-35 % slower and the apply 2-3x slower than the stand-alone pass).
constexpr int kTcEpiPerLGFused = 2;
template <bool GNF> struct TcEpi {
  static constexpr int kPerLG = GNF ? kTcEpiPerLGFused : kTcEpiPerLG;
  static constexpr int kThreads = 128 * kPerLG;
  static constexpr int kBlock = 64 + kThreads;
};

template <int NT>
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory"); }

// warp-uniform: has image `img` been published by every CTA that holds tiles of it?  (acquire load by lane 0, broadcast)
__device__ __forceinline__ bool gnf_image_ready(const unsigned* done, int img, unsigned target, bool block) {
  unsigned ok = 0;
  if ((threadIdx.x & 31) == 0) {
    unsigned spins = 0, v;
    for (;;) {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(done + img) : "memory");
      if (v >= target) { ok = 1; break; }
      if (!block) break;
      __nanosleep(200);
      if (++spins > (1u << 23)) __trap();
    }
  }
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}
// Publish `cnt` completed tiles of image `img`: the epilogue warps' raw rows and GroupNorm sums become visible at GPU scope through
// ONE fence (cumulative over the CTA barrier) and one relaxed atomic -- not one sequentially-consistent fence per thread.
template <int NT>
__device__ __forceinline__ void gnf_publish(unsigned* done, int img, int cnt, int epi_tid, int mode) {
  if (mode == 3) return;
  epi_bar_sync<NT>();
  if (epi_tid == 0) {
    // (__threadfence() is fence.sc.gpu = MEMBAR.SC.GPU + CCTL.IVALL; the release pattern only needs acq_rel)
    if (mode == 4) __threadfence();
    else asm volatile("fence.acq_rel.gpu;" ::: "memory");
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(done + img), "r"((unsigned)cnt) : "memory");
  }
}

// 576 threads: registers are allocated per 4-warp group, so the block counts as 20 warps and gets 96 registers per thread
// (112 fails to launch); the epilogue spills ~150 B per thread, which the extra warps more than pay for
template <int BLOCK_N, bool FAST, bool GNF = false>
__global__ void __launch_bounds__(TcEpi<GNF>::kBlock, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const GemmParams p, const int total_tiles, const int ntn, const int rb, const __grid_constant__ typename GnFuseArg<GNF>::type gf) {
  using SM = TcSmem<BLOCK_N>;
  const bool halo = rb < 0;
  // rb <= -16: halo mode with GROUPED rounds (STAGES = -rb - 16 >= 6 weight slots): the slab of a (dy, k-chunk) and the three weight
  // tiles of its dx taps land on ONE barrier (a_full), so the MMA-issuing thread waits once per 24 MMAs instead of four times.
  // Measured (tools/issue_bench.cu, profiles/r02_issue_bench.txt): an mbarrier wait costs the issuing thread ~240 cycles that do
  // not overlap with the tensor pipe; one round per 3 stages + waiting for the NEXT round before committing this one brings the
  // BLOCK_N = 64 stage from 987 to 589 cycles (tensor time 473).
  const bool hgrp = rb <= -16;
  const int STAGES = rb > 0 ? rb : (halo ? (hgrp ? -rb - 16 : -rb) : SM::kStages);
  constexpr int BB2 = 2 * SM::kBBytes;                               // one resident weight chunk: [B_hi | B_lo]
  // Stacked-N split product (BLOCK_N <= 128): B_hi and B_lo tiles are adjacent in shared memory, so ONE MMA with
  // N = 2*BLOCK_N computes A_hi*B_hi (columns [0, BN)) and A_hi*B_lo (columns [BN, 2BN)); a second MMA adds A_lo*B_hi to
  // the first half; the epilogue sums the halves.  An SS-mode MMA costs ~64-70 cycles for the 4 KiB A read regardless of
  // N <= 128 (measured: tools/gemm_bench.py `mma-only`), so two MMAs instead of three is ~25-30 % less tensor time.
  constexpr bool STACKED = BLOCK_N <= 128;
  constexpr int ACC_COLS = STACKED ? 2 * BLOCK_N : BLOCK_N;          // TMEM columns of one accumulator buffer
  constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int kchunks = p.K / kTcBlockK;
  const int nk = p.KH * p.KW * kchunks;
  // ring stage: [A_hi | A_lo | B_hi | B_lo], or [A_hi | A_lo] behind the resident weights
  // halo: [3 A slots of (hi slab | lo slab)] [STAGES weight slots of (B_hi | B_lo)]
  const int stage_bytes = rb > 0 ? 2 * SM::kABytes : (halo ? BB2 : SM::kStageBytes);
  uint8_t* ring = smem + (rb > 0 ? nk * BB2 : (halo ? kTcHaloASlots * 2 * kTcHaloSlab : 0));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ring + STAGES * stage_bytes);
  uint64_t* empty_bar = full_bar + kTcMaxStages;
  uint64_t* acc_full = empty_bar + kTcMaxStages;  // [2]
  uint64_t* acc_empty = acc_full + 2;             // [2]
  uint64_t* b_full = acc_empty + 2;               // resident weights landed
  uint64_t* a_full = b_full + 1;                  // [3] halo mode: input row slab landed / consumed
  uint64_t* a_empty = a_full + kTcHaloASlots;     // [3]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(a_empty + kTcHaloASlots);

  if (threadIdx.x == 0) {
    ptx::prefetch_tmap(&tmA);
    ptx::prefetch_tmap(&tmB);
    for (int s = 0; s < STAGES; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
    for (int b = 0; b < 2; ++b) { ptx::mbar_init(&acc_full[b], 1); ptx::mbar_init(&acc_empty[b], TcEpi<GNF>::kThreads); }
    ptx::mbar_init(b_full, 1);
    for (int s = 0; s < kTcHaloASlots; ++s) { ptx::mbar_init(&a_full[s], 1); ptx::mbar_init(&a_empty[s], 1); }
    ptx::fence_barrier_init();
  }
  if (warp == 1) ptx::tmem_alloc<TMEM_COLS>(tmem_slot);
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                                      // everything above overlaps the tail of the previous kernel

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    if (ptx::elect_one()) {
      const uint32_t tx_bytes = (rb > 0) ? (uint32_t)(2 * SM::kABytes)
                                         : ((p.nsplit == 3) ? (uint32_t)SM::kStageBytes : (uint32_t)(SM::kABytes + SM::kBBytes));
      if (rb > 0 && blockIdx.x < total_tiles) {
        // resident weights of this CTA's n-tile: every (tap, k-chunk), hi and lo, once
        const int n0 = ((int)blockIdx.x % ntn) * BLOCK_N;
        ptx::mbar_expect_tx(b_full, (uint32_t)(nk * BB2));
        for (int it = 0; it < nk; ++it) {
          const int tap = it / kchunks, kc = it % kchunks;
          const int brow = tap * p.b_rows_per_tap + n0;
          ptx::tma_load_3d(smem + it * BB2, &tmB, b_full, p.b_hi + kc * kTcBlockK, brow, 0);
          ptx::tma_load_3d(smem + it * BB2 + SM::kBBytes, &tmB, b_full, p.b_lo + kc * kTcBlockK, brow, 0);
        }
      }
      // The two single-thread loops (this one and the MMA issuer) are the serial resource of the kernel: every integer
      // division / modulo of the ring position or the tap index costs ~25 dependent instructions per stage, and the tensor
      // pipe retires a 64-wide stage in 473 cycles (tools/mma_bench.cu `kernel pattern`).  Stage index, phase, tap
      // coordinates and weight rows are therefore carried incrementally.
      uint32_t s = 0, ph = 0;                                // ring slot and its phase parity, continue across tiles
      if (halo) {
        uint32_t sa = 0, pa = 0;                             // A slab slot and phase
        for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
          const TcTile tl = tc_decode_tile(p, t, ntn, BLOCK_N);
          int brow = tl.n0;
          for (int ty = 0; ty < 3; ++ty) {
            for (int kc = 0; kc < kchunks; ++kc) {
              ptx::mbar_wait(&a_empty[sa], pa ^ 1);
              uint8_t* sl = smem + sa * (2 * kTcHaloSlab);
              if (hgrp) {
                // one barrier per round: slab + the three weight tiles.  a_full[sa] of round r is re-armed for round r + 3 only
                // after slab slot sa was released, i.e. after the issuer has passed the wait of round r.
                uint64_t* gb = &a_full[sa];
                ptx::mbar_expect_tx(gb, 2 * kTcHaloSlab + 3 * BB2);
                ptx::tma_load_4d(sl, &tmA, gb, p.a_hi + kc * kTcBlockK, tl.cw0 - 1, tl.ch0 + ty - 1, tl.img_a);
                ptx::tma_load_4d(sl + kTcHaloSlab, &tmA, gb, p.a_lo + kc * kTcBlockK, tl.cw0 - 1, tl.ch0 + ty - 1, tl.img_a);
                if (++sa == kTcHaloASlots) { sa = 0; pa ^= 1; }
                for (int tx = 0; tx < 3; ++tx) {
                  ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                  uint8_t* st = ring + s * BB2;
                  if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
                  const int br = brow + tx * p.b_rows_per_tap;
                  ptx::tma_load_3d(st, &tmB, gb, p.b_hi + kc * kTcBlockK, br, 0);
                  ptx::tma_load_3d(st + SM::kBBytes, &tmB, gb, p.b_lo + kc * kTcBlockK, br, 0);
                }
                continue;
              }
              ptx::mbar_expect_tx(&a_full[sa], 2 * kTcHaloSlab);
              ptx::tma_load_4d(sl, &tmA, &a_full[sa], p.a_hi + kc * kTcBlockK, tl.cw0 - 1, tl.ch0 + ty - 1, tl.img_a);
              ptx::tma_load_4d(sl + kTcHaloSlab, &tmA, &a_full[sa], p.a_lo + kc * kTcBlockK, tl.cw0 - 1, tl.ch0 + ty - 1, tl.img_a);
              if (++sa == kTcHaloASlots) { sa = 0; pa ^= 1; }
              for (int tx = 0; tx < 3; ++tx) {
                ptx::mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = ring + s * BB2;
                uint64_t* fb = &full_bar[s];
                if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
                ptx::mbar_expect_tx(fb, BB2);
                const int br = brow + tx * p.b_rows_per_tap;
                ptx::tma_load_3d(st, &tmB, fb, p.b_hi + kc * kTcBlockK, br, 0);
                ptx::tma_load_3d(st + SM::kBBytes, &tmB, fb, p.b_lo + kc * kTcBlockK, br, 0);
              }
            }
            brow += 3 * p.b_rows_per_tap;
          }
        }
      } else
      for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
        const TcTile tl = tc_decode_tile(p, t, ntn, BLOCK_N);
        const int bz = (p.b_mode == 0) ? 0 : (p.b_mode == 1 ? tl.z / p.nheads : tl.z);
        const int ax0 = tl.cw0 * p.in_stride + p.offW, ay0 = tl.ch0 * p.in_stride + p.offH;
        const int acol0 = tl.head * p.a_head_stride, bcol0 = tl.head * p.b_head_stride;
        int brow = tl.n0 + tl.head * p.b_head_rows;          // weight row of the current tap
        for (int ty = 0; ty < p.KH; ++ty) {
          for (int tx = 0; tx < p.KW; ++tx, brow += p.b_rows_per_tap) {
            const int ax = ax0 + tx * p.tap_sw, ay = ay0 + ty;
            for (int kc = 0; kc < kchunks; ++kc) {
              ptx::mbar_wait(&empty_bar[s], ph ^ 1);
              uint8_t* st = ring + s * stage_bytes;
              uint64_t* fb = &full_bar[s];
              if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
              if (p.dbg & 4) { ptx::mbar_arrive(fb); continue; }   // tuning aid: no operand traffic at all
              ptx::mbar_expect_tx(fb, tx_bytes);
              const int acol = kc * kTcBlockK + acol0;
              ptx::tma_load_4d(st, &tmA, fb, p.a_hi + acol, ax, ay, tl.img_a);
              if (rb > 0) {
                ptx::tma_load_4d(st + SM::kABytes, &tmA, fb, p.a_lo + acol, ax, ay, tl.img_a);
                continue;
              }
              const int bcol = kc * kTcBlockK + bcol0;
              ptx::tma_load_3d(st + 2 * SM::kABytes, &tmB, fb, p.b_hi + bcol, brow, bz);
              if (p.nsplit == 3) {
                ptx::tma_load_4d(st + SM::kABytes, &tmA, fb, p.a_lo + acol, ax, ay, tl.img_a);
                ptx::tma_load_3d(st + 2 * SM::kABytes + SM::kBBytes, &tmB, fb, p.b_lo + bcol, brow, bz);
              }
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer: one elected thread runs the whole loop ----------------
    constexpr uint32_t idesc = ptx::make_idesc_bf16(kTcBlockM, BLOCK_N);
    constexpr uint32_t idesc2 = ptx::make_idesc_bf16(kTcBlockM, STACKED ? 2 * BLOCK_N : BLOCK_N);
    if (ptx::elect_one()) {
      // K-major SWIZZLE_128B descriptors differ only in their 14-bit start-address field: desc(addr) = kDescBase + (addr >> 4)
      // (shared-memory addresses are < 256 KiB, so the field never carries)
      constexpr uint64_t kDescBase = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
      const uint32_t ring_u = ptx::smem_u32(ring) >> 4, stage_u = (uint32_t)stage_bytes >> 4;
      const uint32_t bres_u = ptx::smem_u32(smem) >> 4;
      uint32_t s = 0, ph = 0;
      int li = 0;                                            // local tile counter -> accumulator buffer li & 1
      const bool early_wait = p.late_wait == 0;              // tuning aid: DEXB_EARLY_WAIT=0 restores wait-at-the-top
      bool pre_g = false;
      if (rb > 0 && blockIdx.x < total_tiles) ptx::mbar_wait(b_full, 0);
      if constexpr (STACKED) {
        if (halo) {
          uint32_t sa = 0, pa = 0;
          bool prewaited = false;
          const uint32_t slab_u = ptx::smem_u32(smem) >> 4;
          for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
            const int buf = li & 1;
            ptx::mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1);
            ptx::tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
            const int na = 3 * kchunks;                      // (dy, k-chunk) slab steps of a tile
            if (hgrp) {
              for (int ia = 0; ia < na; ++ia) {
                if (!prewaited) ptx::mbar_wait(&a_full[sa], pa);
                prewaited = false;
                ptx::tc_fence_after();
                const uint32_t a_base = slab_u + sa * (uint32_t)(2 * kTcHaloSlab >> 4);
#pragma unroll
                for (int tx = 0; tx < 3; ++tx) {
                  const uint32_t a_hi = a_base + (uint32_t)(tx * 8);
                  const uint32_t a_lo = a_hi + (kTcHaloSlab >> 4);
                  const uint32_t b_hi = ring_u + s * stage_u;
#pragma unroll
                  for (int kk = 0; kk < kTcBlockK / 16; ++kk) {
                    const uint32_t ko = kk * 2;
                    const uint64_t dbh = kDescBase + (b_hi + ko);
                    ptx::mma_bf16_ss(tacc, kDescBase + (a_hi + ko), dbh, idesc2, (ia > 0 || tx > 0 || kk > 0) ? 1u : 0u);
                    ptx::mma_bf16_ss(tacc, kDescBase + (a_lo + ko), dbh, idesc, 1u);
                  }
                  ptx::mma_commit(&empty_bar[s]);            // weight slot free once these MMAs have retired
                  if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
                }
                const uint32_t sa0 = sa;
                if (++sa == kTcHaloASlots) { sa = 0; pa ^= 1; }
                if (ia == na - 1) {
                  ptx::mma_commit(&acc_full[buf]);           // never delay the epilogue behind the next tile's operands
                } else {
                  ptx::mbar_wait(&a_full[sa], pa);           // the next round's operands, BEFORE this round's slab commit
                  prewaited = true;
                }
                ptx::mma_commit(&a_empty[sa0]);
              }
              continue;
            }
            // per-tap rounds (weight tiles too large for two grouped rounds): the wait for the NEXT stage is issued before this
            // stage's commit (profiles/r02_issue_bench.txt: 1127 -> 924 cycles per BLOCK_N = 128 stage); never across a tile end
            bool pre_b = false;
            for (int ia = 0; ia < na; ++ia) {
              if (!prewaited) ptx::mbar_wait(&a_full[sa], pa);
              prewaited = false;
              const uint32_t a_base = slab_u + sa * (uint32_t)(2 * kTcHaloSlab >> 4);
              uint32_t san = sa + 1, pan = pa;
              if (san == kTcHaloASlots) { san = 0; pan ^= 1; }
              for (int tx = 0; tx < 3; ++tx) {
                if (!pre_b) ptx::mbar_wait(&full_bar[s], ph);
                pre_b = false;
                ptx::tc_fence_after();
                const uint32_t a_hi = a_base + (uint32_t)(tx * 8);          // tap dx: the slab shifted by dx rows of 128 B
                const uint32_t a_lo = a_hi + (kTcHaloSlab >> 4);
                const uint32_t b_hi = ring_u + s * stage_u;
#pragma unroll
                for (int kk = 0; kk < kTcBlockK / 16; ++kk) {
                  const uint32_t ko = kk * 2;
                  const uint64_t dbh = kDescBase + (b_hi + ko);
                  ptx::mma_bf16_ss(tacc, kDescBase + (a_hi + ko), dbh, idesc2, (ia > 0 || tx > 0 || kk > 0) ? 1u : 0u);
                  ptx::mma_bf16_ss(tacc, kDescBase + (a_lo + ko), dbh, idesc, 1u);
                }
                uint64_t* eb = &empty_bar[s];
                if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
                const bool last = (ia == na - 1) && (tx == 2);
                if (last) ptx::mma_commit(&acc_full[buf]);
                else if (early_wait) {
                  if (tx == 2) { ptx::mbar_wait(&a_full[san], pan); prewaited = true; }
                  ptx::mbar_wait(&full_bar[s], ph);
                  pre_b = true;
                }
                ptx::mma_commit(eb);
              }
              ptx::mma_commit(&a_empty[sa]);                 // slab free once the MMAs of its three taps have retired
              sa = san; pa = pan;
            }
          }
        }
      }
      for (int t = blockIdx.x; t < total_tiles && !halo; t += gridDim.x, ++li) {
        const int buf = li & 1;
        ptx::mbar_wait(&acc_empty[buf], ((li >> 1) & 1) ^ 1);  // epilogue has drained this buffer
        ptx::tc_fence_after();
        const uint32_t tacc = tmem_base + (uint32_t)(buf * ACC_COLS);
        uint32_t bres = bres_u;                              // resident weight chunk of this iteration
        for (int it = 0; it < nk; ++it, bres += (uint32_t)(BB2 >> 4)) {
          if (!pre_g) ptx::mbar_wait(&full_bar[s], ph);
          pre_g = false;
          ptx::tc_fence_after();
          const uint32_t a_hi = ring_u + s * stage_u;
          const uint32_t a_lo = a_hi + (SM::kABytes >> 4);
          const uint32_t b_hi = (rb > 0) ? bres : a_hi + (2 * SM::kABytes >> 4);
          const uint32_t b_lo = b_hi + (SM::kBBytes >> 4);
          if (FAST || !(p.dbg & 2)) {
#pragma unroll
            for (int kk = 0; kk < kTcBlockK / 16; ++kk) {
              const uint32_t ko = kk * 2;                    // 16 bf16 = 32 B inside the 128 B swizzle span
              const uint64_t dah = kDescBase + (a_hi + ko);
              const uint64_t dbh = kDescBase + (b_hi + ko);
              const uint32_t acc = (it > 0 || kk > 0) ? 1u : 0u;
              if (FAST || p.nsplit == 3) {
                const uint64_t dal = kDescBase + (a_lo + ko);
                if constexpr (STACKED) {
                  ptx::mma_bf16_ss(tacc, dah, dbh, idesc2, acc);            // [A_hi B_hi | A_hi B_lo]
                  ptx::mma_bf16_ss(tacc, dal, dbh, idesc, 1u);              //  + A_lo B_hi
                } else {
                  const uint64_t dbl = kDescBase + (b_lo + ko);
                  ptx::mma_bf16_ss(tacc, dah, dbh, idesc, acc);
                  ptx::mma_bf16_ss(tacc, dah, dbl, idesc, 1u);
                  ptx::mma_bf16_ss(tacc, dal, dbh, idesc, 1u);
                }
              } else {
                ptx::mma_bf16_ss(tacc, dah, dbh, idesc, acc);
              }
            }
          }
          uint64_t* eb = &empty_bar[s];
          if (++s == (uint32_t)STAGES) { s = 0; ph ^= 1; }
          if (it == nk - 1) ptx::mma_commit(&acc_full[buf]); // accumulator complete
          else if (early_wait) { ptx::mbar_wait(&full_bar[s], ph); pre_g = true; }   // next stage of this tile, before the commit
          ptx::mma_commit(eb);                               // frees the smem stage when these MMAs retire
        }
      }
    }
    __syncwarp();
  } else {
    // ---------------- epilogue: TMEM -> registers -> global ----------------
    // 8 warps: lane group lg = warp & 3 (the TMEM lanes a warp may touch); the two warps of a lane group take the
    // even / odd 32-column chunks, which doubles the number of global stores in flight per tile.
    const int lg = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r = lg * 32 + lane;                            // row of the tile
    constexpr int CW = (BLOCK_N > 0) ? kTcEpiCW : 0;          // value-dependent, so `if constexpr (CW == 16)` discards the other branch
    constexpr int PLG = TcEpi<GNF>::kPerLG;                  // epilogue warps per TMEM lane group
    constexpr int MAXCH = (BLOCK_N / CW + PLG - 1) / PLG;    // chunks one warp owns per tile
    const bool defer_gn = p.epi.gn_stats != nullptr && ntn == 1 && p.nheads == 1;
    float gacc[MAXCH][8];
#pragma unroll
    for (int k = 0; k < MAXCH; ++k)
#pragma unroll
      for (int g = 0; g < 8; ++g) gacc[k][g] = 0.f;
    int gn_img = -1;                                         // image the deferred sums belong to
    int li = 0;
    [[maybe_unused]] int gn_cnt = 0;                         // GNF: tiles of image gn_img this CTA has completed
    [[maybe_unused]] const int epi_tid = (int)threadIdx.x - 64;
    // GNF apply state: image ap_img, item round ap_j of ap_J (items gi = ap_first + epi_tid + j * NT, gi < ap_last)
    constexpr int NT = TcEpi<GNF>::kThreads;
    [[maybe_unused]] int ap_img = 0, ap_j = 0, ap_J = 0, ap_ipt = 0;
    [[maybe_unused]] unsigned ap_first = 0, ap_last = 0, gn_target = 0;
    if constexpr (GNF) {
      gn_target = (unsigned)(p.TH * p.TW);                   // tiles per image (one n-tile)
      const unsigned ngroups = (unsigned)gf.a.P * (unsigned)(gf.a.C >> 3);
      unsigned per = (ngroups + gridDim.x - 1) / gridDim.x;
      per = (per + NT - 1) / NT * NT;
      ap_first = blockIdx.x * per;
      ap_last = (ap_first + per < ngroups) ? ap_first + per : ngroups;
      ap_J = (ap_first < ngroups) ? (int)(per / NT) : 0;
      if (ap_J == 0) ap_img = p.nz;                          // this CTA has no share
      const int tiles_cta = total_tiles / (int)gridDim.x;    // >= 1: every CTA of a fused launch has at least one tile
      ap_ipt = (p.nz * ap_J + tiles_cta - 1) / (tiles_cta > 0 ? tiles_cta : 1) + 1;      // item rounds per tile: keeps up with margin
      ap_ipt = (ap_ipt + 3) & ~3;                            // whole groups of four (one L2 round trip each)
      if (ap_ipt > 8) ap_ipt = 8;
    }
    for (int t = blockIdx.x; t < total_tiles; t += gridDim.x, ++li) {
      const TcTile tl = tc_decode_tile(p, t, ntn, BLOCK_N);
      const int buf = li & 1;
      if (defer_gn && tl.z != gn_img) {
        if (gn_img >= 0) {
          epi_flush_gn<MAXCH>(p.epi, p.N, gn_img, half, PLG, gacc);
          if constexpr (GNF) { gnf_publish<NT>(gf.done, gn_img, gn_cnt, epi_tid, gf.mode); gn_cnt = 0; }
        }
        gn_img = tl.z;
      }
      if constexpr (GNF) {
        ++gn_cnt;
        // a slice of the GroupNorm-apply of an image `lag` behind the tile front: its raw rows are in L2, and the loads' round trip is
        // hidden behind the MMAs of this tile (the accumulator of the previous tile has been handed back already)
        int budget = (gf.mode == 0) ? ap_ipt : 0;
        while (budget > 0 && ap_img <= gn_img - gf.lag) {
          if (!gnf_image_ready(gf.done, ap_img, gn_target, false)) break;
          const int n = (budget < ap_J - ap_j) ? budget : ap_J - ap_j;
          gn_apply_items(gf.a, ap_img, ap_first + (unsigned)epi_tid + (unsigned)ap_j * NT, ap_last, NT, n);
          ap_j += n; budget -= n;
          if (ap_j >= ap_J) { ++ap_img; ap_j = 0; }
        }
      }
      const int ch = tl.ch0 + r / p.BW, cw = tl.cw0 + r % p.BW;
      const bool valid = (ch < p.CH) && (cw < p.CW);
      const int oh = ch * p.out_scale + p.out_offh, ow = cw * p.out_scale + p.out_offw;
      const uint32_t tacc = tmem_base + ((uint32_t)(lg * 32) << 16) + (uint32_t)(buf * ACC_COLS);
      constexpr int NCH = BLOCK_N / CW;
      // chunks this warp owns: c = half, half + kTcEpiPerLG, ... while n0 + 32 c < N
      int nmine = 0;
      for (int c = half; c < NCH && tl.n0 + c * CW < p.N; c += PLG) ++nmine;
      if (nmine == 0) {
        ptx::mbar_wait(&acc_full[buf], (li >> 1) & 1);       // never hand a buffer back before its MMAs have completed
        ptx::tc_fence_before();
        ptx::mbar_arrive(&acc_empty[buf]);
      }
#pragma unroll 1
      for (int k = 0; k < nmine; ++k) {
        const int c = half + PLG * k;
        const int n0c = tl.n0 + c * CW;
        const bool use_vt = !FAST && p.epi.out_vt != nullptr && n0c >= p.epi.out_s_ncols;
        // residual rows first: their L2 round trip overlaps the wait for the accumulator and the tensor-memory load
        float rpre[CW];
        if (!(p.dbg & 1) && !use_vt) epi_load_resid<CW, FAST>(p.epi, p.N, tl.z, p.nheads, oh, ow, p.OH, p.OW, valid, n0c, rpre);
        if (k == 0) {
          ptx::mbar_wait_backoff(&acc_full[buf], (li >> 1) & 1, 128);
          ptx::tc_fence_after();
        }
        float v[CW];
        if constexpr (CW == 16) {
          if (STACKED && p.nsplit == 3) {                    // add the A_hi * B_lo half
            float v2[16];
            ptx::tmem_ld16_nowait(tacc + (uint32_t)(c * 16), v);
            ptx::tmem_ld16_nowait(tacc + (uint32_t)(BLOCK_N + c * 16), v2);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += v2[i];
          } else {
            ptx::tmem_ld16_nowait(tacc + (uint32_t)(c * 16), v);
            ptx::tmem_wait_ld();
          }
        } else {
          if (STACKED && p.nsplit == 3) {                    // add the A_hi * B_lo half
            float v2[16];                                    // in two halves: 16 fewer live registers
            ptx::tmem_ld32_nowait(tacc + (uint32_t)(c * 32), v);
            ptx::tmem_ld16_nowait(tacc + (uint32_t)(BLOCK_N + c * 32), v2);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += v2[i];
            ptx::tmem_ld16_nowait(tacc + (uint32_t)(BLOCK_N + c * 32 + 16), v2);
            ptx::tmem_wait_ld();
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 + i] += v2[i];
          } else {
            ptx::tmem_ld32(tacc + (uint32_t)(c * 32), v);
          }
        }
        if (k == nmine - 1) {                                // last chunk read: hand the buffer back to the MMA warp
          ptx::tc_fence_before();
          ptx::mbar_arrive(&acc_empty[buf]);
        }
        if (!(p.dbg & 1)) {
          if constexpr (!FAST) {
            if (use_vt) {
              epi_store_vt<CW>(p.epi, p.N, tl.z, p.nheads, (long)oh * p.OW + ow, valid, n0c, v);
              continue;
            }
          }
          epi_apply<CW, FAST>(p.epi, p.N, tl.z, p.nheads, oh, ow, p.OH, p.OW, valid, n0c, v, rpre, defer_gn ? &gacc[k][0] : nullptr);
        }
      }
    }
    if (defer_gn && gn_img >= 0) epi_flush_gn<MAXCH>(p.epi, p.N, gn_img, half, PLG, gacc);
    if constexpr (GNF) {
      if (gn_img >= 0) gnf_publish<NT>(gf.done, gn_img, gn_cnt, epi_tid, gf.mode);
      // the images still open (at least the last `lag`): wait for the other CTAs' tiles
      for (; gf.mode != 2 && gf.mode != 3 && ap_img < p.nz; ++ap_img, ap_j = 0) {
        gnf_image_ready(gf.done, ap_img, gn_target, true);
        gn_apply_items(gf.a, ap_img, ap_first + (unsigned)epi_tid + (unsigned)ap_j * NT, ap_last, NT, ap_J - ap_j);
      }
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (warp == 1) ptx::tmem_dealloc<TMEM_COLS>(tmem_base);
}

// ------------------------------------------------------------------------------------------------
// CUDA-core engine (fallback / cross-check): 128 x 64 tile, 256 threads, 4 x 8 micro-tile
// ------------------------------------------------------------------------------------------------
constexpr int kSimtThreads = 256;

__global__ void __launch_bounds__(kSimtThreads)
gemm_simt_kernel(const GemmParams p) {
  __shared__ float As[16][128 + 4];
  __shared__ float Bs[16][64 + 4];
  int t = blockIdx.x;
  const int tw = t % p.TW; t /= p.TW;
  const int th = t % p.TH; t /= p.TH;
  const int z = t;
  const int head = z % p.nheads;
  const int img_a = p.a_by_z ? z : z / p.nheads;
  const int ch0 = th * p.BH, cw0 = tw * p.BW;
  const int n0 = blockIdx.y * 64;
  const int tid = threadIdx.x;
  const int tx = tid & 31, ty = tid >> 5;
  const int ntaps = p.KH * p.KW;
  const long bmat = (p.b_mode == 0) ? 0 : (p.b_mode == 1 ? (long)(z / p.nheads) : (long)z) * p.b_mat_stride;

  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  // loader roles
  const int a_r = tid >> 1, a_k = (tid & 1) * 8;          // A: row a_r, 8 consecutive k
  const int a_ch = ch0 + a_r / p.BW, a_cw = cw0 + a_r % p.BW;
  const int b_n = tid >> 2, b_k = (tid & 3) * 4;           // B: col b_n, 4 consecutive k

  for (int tap = 0; tap < ntaps; ++tap) {
    const int dy = tap / p.KW + p.offH, dx = (tap % p.KW) * p.tap_sw + p.offW;
    const int ih = a_ch * p.in_stride + dy, iw = a_cw * p.in_stride + dx;
    const bool a_ok = (a_ch < p.CH) && (a_cw < p.CW) && ih >= 0 && ih < p.H && iw >= 0 && iw < p.W;
    const bf16* arow = p.A + (((long)img_a * p.H + ih) * p.W + iw) * p.a_row_stride + head * p.a_head_stride;
    const bool b_ok = (n0 + b_n) < p.N;
    const bf16* brow = p.Bw + bmat + ((long)tap * p.b_rows_per_tap + n0 + b_n + head * p.b_head_rows) * p.b_row_stride +
                       head * p.b_head_stride;
    for (int k0 = 0; k0 < p.K; k0 += 16) {
      float av[8];
      if (a_ok) load_split8(arow + p.a_hi + k0 + a_k, arow + p.a_lo + k0 + a_k, av);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) av[i] = 0.f;
      }
      float bv[4];
      if (b_ok) {
        const uint2 hq = *reinterpret_cast<const uint2*>(brow + p.b_hi + k0 + b_k);
        const uint2 lq = *reinterpret_cast<const uint2*>(brow + p.b_lo + k0 + b_k);
        const bf16* h = reinterpret_cast<const bf16*>(&hq);
        const bf16* l = reinterpret_cast<const bf16*>(&lq);
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = join2(h[i], l[i]);
      } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) bv[i] = 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int i = 0; i < 8; ++i) As[a_k + i][a_r] = av[i];
#pragma unroll
      for (int i = 0; i < 4; ++i) Bs[b_k + i][b_n] = bv[i];
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 16; ++k) {
        const float4 a4 = *reinterpret_cast<const float4*>(&As[k][tx * 4]);
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[k][ty * 8]);
        const float4 b1 = *reinterpret_cast<const float4*>(&Bs[k][ty * 8 + 4]);
        const float a[4] = {a4.x, a4.y, a4.z, a4.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = tx * 4 + i;
    const int ch = ch0 + r / p.BW, cw = cw0 + r % p.BW;
    const bool valid = (ch < p.CH) && (cw < p.CW);
    const int oh = ch * p.out_scale + p.out_offh, ow = cw * p.out_scale + p.out_offw;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = acc[i][j];
    float rr[8];
    epi_load_resid<8, false>(p.epi, p.N, z, p.nheads, oh, ow, p.OH, p.OW, valid, n0 + ty * 8, rr);
    epi_apply<8, false>(p.epi, p.N, z, p.nheads, oh, ow, p.OH, p.OW, valid, n0 + ty * 8, v, rr);
  }
}

}  // namespace dexb
