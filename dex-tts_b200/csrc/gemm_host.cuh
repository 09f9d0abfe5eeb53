// Parameter blocks and host API of the implicit-GEMM engine (kernels: gemm.cuh, host code: gemm.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dexb {

// GroupNorm sums are accumulated into kGnRep replicas per (image, group), picked by the writing CTA's index, and summed by the readers:
// all CTAs of a convolution leave an image at about the same time, and 592 double atomics per address and image change (148 CTAs x 4
// lane-group warps) serialised in L2 -- the sums cost the 64-channel convolution 21 us of 79 (tools/pair_bench.py, out_mode 4).
constexpr int kGnRep = 4;

struct EpiParams {
  float alpha;                 // acc *= alpha (before bias)
  const float* bias;           // bias[z * bias_zstride + head * bias_head_stride + n] or null
  long bias_zstride;
  int bias_head_stride;
  int act;                     // 0 none, 1 exact GELU
  const float* gate;           // [N] or null: v = resid + gate * v
  const float* resid_f32;      // fp32 residual rows (same row geometry as the output), or null
  long resid_f32_stride;
  const bf16* resid_s;         // split residual, or null
  long resid_s_stride;
  int resid_s_hi, resid_s_lo;
  const float* rowmask;        // [img][W_out] multiplies the whole output row, or null
  long rowmask_stride;
  float* out_f32;              // fp32 output rows or null
  long out_f32_stride;
  int out_f32_col;
  bf16* out_s;                 // split output rows or null
  long out_s_stride;
  int out_s_hi, out_s_lo;
  int out_s_ncols;             // only columns n < out_s_ncols go to out_s (rest may go to out_vt)
  float s_lrelu;               // != 0: the split store holds leaky_relu(v, s_lrelu) -- the operand of the next convolution -- while
                               //       out_f32 keeps v itself (the residual stream of a HiFi-GAN ResBlock)
  int out_s_gshift, out_s_gpitch;   // gshift > 0: columns are grouped by 2^gshift and group g goes to column g * gpitch + (n mod 2^gshift)
                               //       (phase-stacked ConvTranspose1d: N = phases x channels, one split row per output time step)
  bf16* out_vt;                // transposed split store for columns n >= out_s_ncols:  vT[z'][d][token]
  long out_vt_zstride;         // elements between (img, head) matrices
  long out_vt_rstride;         // elements between d rows (= 2 * padded token count)
  int out_vt_lo;               // lo offset inside a row
  int out_vt_hd;               // head dim: column (n - out_s_ncols) -> head = /hd, d = %hd
  int out_vt_heads;
  double* gn_stats;            // [img][kGnRep][N/gs][2] (sum, sumsq) accumulated with atomics, or null
  int gn_gs;                   // channels per group (8 or 16)
  float* colmean;              // atomicAdd(colmean[(img*OW + ow)*colmean_ld + col], v * colmean_scale), or null
  float colmean_scale;
  int colmean_ld;
  int o_head_stride;           // output column offset per head (z % nheads)
  int o_by_z;                  // 1: output image index = z, 0: = z / nheads
  int dbg_nostore;             // tuning aid (dexb_gemm_bench bit 3): run the whole epilogue but skip the global stores
};

struct GemmParams {
  // grid decode
  int nz;                      // images * nheads
  int nheads;
  // A image geometry (input) and output geometry
  int H, W;                    // input image
  int OH, OW;                  // output image the rows are scattered into
  int TH, TW;                  // tile grid (in units of computed pixels)
  int CH, CW;                  // computed pixel grid (rows of the GEMM per image = CH*CW)
  int BH, BW;                  // tile shape, BH*BW == 128
  int in_stride;               // input pixel = computed pixel * in_stride + tap offset (SIMT engine only when != 1)
  int out_scale, out_offh, out_offw;   // output pixel = computed pixel * out_scale + off
  int KH, KW, offH, offW;      // taps: dy = ty + offH, dx = tx * tap_sw + offW
  int tap_sw;                  // x step between taps (1; 2 for the pair-packed pos-conv)
  int K;                       // contraction length per tap (multiple of 64)
  int N;                       // valid output columns
  int a_hi, a_lo;              // column offsets of hi / lo inside an A row
  int a_head_stride;           // extra A column offset per head
  int a_by_z;                  // 1: A image index = z, 0: = z / nheads
  int b_rows_per_tap;
  int b_hi, b_lo;              // column offsets of hi / lo inside a B row
  int b_head_stride;           // extra B column offset per head
  int b_head_rows;             // extra B row offset per head
  int b_mode;                  // 0: shared weights, 1: per image (z / nheads), 2: per z
  int block_n_hint;            // 0 = pick the n-tile width from N; 64 / 128 = the caller's choice (wave quantisation, see engine.cu)
  int nsplit;                  // 3 = bf16x3 (default), 1 = hi*hi only
  int late_wait;               // tuning aid (DEXB_EARLY_WAIT=0): the MMA issuer waits for a stage at its top instead of one stage ahead
  int dbg;                     // tuning aid (dexb_gemm_bench): bit 0 = skip the epilogue, bit 1 = skip the MMAs, bit 2 = skip TMA of A
  EpiParams epi;
  // raw views for the SIMT engine
  const bf16* A;
  long a_row_stride;           // elements per pixel row
  const bf16* Bw;
  long b_row_stride;
  long b_mat_stride;           // elements between per-image / per-z weight matrices
};

constexpr int kTcBlockK = 64;                  // 64 bf16 = 128 B = one swizzle span

struct GemmPlan {
  CUtensorMap tmA, tmB;        // 64 B aligned by the type's own alignment
  CUtensorMap tmBh;            // halo mode: weight boxes of block_n / 2 rows (CTA-pair kernel, conv_pair.cuh)
  GemmParams p;
  int block_n;
  int n_img_a;
  bool tc_ok;                  // shape is eligible for the tcgen05 engine
  bool halo;                   // 3x3 stride-1 convolution in halo mode: tmA boxes are 136-pixel row slabs (gemm.cuh)
};

// one-off kernel attribute setup + driver entry point resolution (call outside stream capture)
int gemm_global_init();
// Validate the problem, pick the tile shape and encode the TMA descriptors.  `n_img_a` = number of A images,
// `b_rows` = rows of one weight matrix (taps * rows_per_tap), `n_bmat` = number of weight matrices (1 if shared).
int gemm_plan_init(GemmPlan* gp, const GemmParams& p, int n_img_a, long b_rows, int n_bmat);
// Enqueue on `st`.  `p` is normally gp.p, possibly with per-step pointers (bias / gate / stats) patched.
// engine 0 = tcgen05 (falls back to the CUDA-core kernel only for shapes the plan marked ineligible), 1 = CUDA cores.
// `gf` != null: the GroupNorm-apply of the convolution's output runs inside the kernel (only if gemm_can_fuse_gn says so).
struct GnFuse;
bool gemm_can_fuse_gn(const GemmPlan& gp, const GemmParams& p, int engine);
int gemm_launch(const GemmPlan& gp, const GemmParams& p, int engine, cudaStream_t st, const GnFuse* gf = nullptr);
// number of tcgen05 launches so far that had to take the generic (scalar-fallback) epilogue instantiation
long gemm_generic_epilogue_launches();

}  // namespace dexb
