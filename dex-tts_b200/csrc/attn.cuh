// Fused attention kernel of the DiT blocks (attn.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dexb {

struct AttnParams {
  int NQ;                    // query rows per image
  int NK;                    // keys per image
  int KP;                    // padded key count: lo half of a V^T row starts at column KP
  int nheads;
  int nt;                    // key tiles of 64
  int kchunks;               // 64-wide chunks of the Q / K rows (head dim / 64): 2, or 1 for the C = 64 linear attention
  int vchunks;               // 64-wide chunks of the V rows (output width / 64): 2; C / 64 for the linear-attention context
  int kv_splits;             // > 1: split-KV mode -- blockIdx.x = key split, one 128-row query tile, partial outputs
  int tiles_per_split;
  int q_img_rows;            // Q rows per image (0: the same Q rows for every image)
  int v_mn;                  // 1: V tiles come row-major [key][d] (MN-major B operand) from tmV at columns v_hi / v_lo (+ head*128)
  int v_hi, v_lo;
  // Q rows (staged to tensor memory by the softmax warps): q + (b*NQ + row)*q_stride + head*128, hi at q_hi, lo at q_lo
  const bf16* q;
  long q_stride;
  int q_hi, q_lo;
  int k_hi, k_lo;            // columns of the K rows inside the tmK tensor (hi / lo), head h at + h*128
  float scale_log2e;         // score scale * log2(e)
  // optional (TV adaptor): additive per-key score bias [b][kbias_stride] and visible key count vis_len[b] + 1
  // (keys beyond it carry -1e4 in the reference, ref_encoder.py:171 -- exactly zero weight after the softmax)
  const float* kbias;
  long kbias_stride;
  const int* vis_len;
  // epilogue: mode 0 = split rows (head h at column h*128); mode 1 = fp32 rows (O/l + the Q row as residual) * rowmask;
  // mode 2 = split-KV partials: part_o [b][split][128][128] un-normalised, part_l / part_m [b][split][128]
  int out_mode;
  bf16* out;
  long out_stride;
  int out_hi, out_lo;
  float* out_f;
  long out_f_stride;
  const float* rowmask;      // [b][rowmask_w], column = row % rowmask_w
  int rowmask_w;
  float *part_o, *part_l, *part_m;
  // Tail split (DiT self-attention): the grid is 1-D over (image, head, query tile) work items.  Items [0, tail_first) run all key
  // tiles; each of the remaining tiles -- the ones that would form a partial last wave on the SMs -- is cut into tail_splits
  // key ranges of tail_tps tiles that run concurrently and write mode-2 partials, merged by k_attn_tail_merge.
  int q_tiles;               // query tiles per (image, head)
  int tail_first, tail_splits, tail_tps;
};

struct AttnPlan {
  CUtensorMap tmK, tmV;
  AttnParams p;
  int B;
};

int attn_global_init();
bool attn_supported(int hd);
// DiT self-attention over qkv rows [hi(3*hid) | lo(3*hid)] -> split rows `out`.  vT == nullptr: V is read row-major from the
// qkv rows (MN-major B operand); otherwise from the transposed copy V^T [b][hid][hi(NP)|lo(NP)].
int attn_plan_init(AttnPlan* ap, const bf16* qkv, const bf16* vT, bf16* out, int B, int N, int NP, int heads, int hid);
// TV adaptor cross-attention (single head, C = 128): queries = the split rows x (also the residual), keys kq [b][KP][hi(C)|lo(C)]
// with score bias sbias [b][KP], values vlt [b][C][hi(KP)|lo(KP)], output fp32 rows (x + attn) * mask
int attn_plan_init_tv(AttnPlan* ap, const bf16* x, long x_stride, int x_hi, int x_lo, const bf16* kq, const float* sbias,
                      const bf16* vlt, const int* sty_len, float* out, const float* mask, int mask_w, int B, int P, int NK, int KP,
                      int C);
// LinearAttention context (diffusion.py:82-95) as attention with the roles swapped: "queries" = the 128 k-channels (rows of the
// k part of to_qkv, split weights wk [128][hi(C)|lo(C)]), "keys" = the pixels x, softmax over ALL pixels of an image, "values" = v
// = the pixels x AGAIN (MN-major B operand straight from the activation rows): context = softmax(k)^T v = (softmax(k)^T x) W_v^T,
// so the kernel produces G = softmax(k)^T x (128 x C per image) and the small product with W_v is left to the merge kernel
// (launch_la_combine) -- v is never materialised.  Split over the pixels; partials [b][split][128][128 (first C used)].
int attn_plan_init_la(AttnPlan* ap, const bf16* wk, const bf16* x, long x_stride, int x_hi, int x_lo, float* part_o,
                      float* part_l, float* part_m, int B, int P, int PP, int C, int splits);
// Balance the last partial wave of the DiT attention (see AttnParams::tail_first).  Returns the number of floats of scratch the
// partials need (0: the tile count already fills the SMs evenly, nothing to do); call attn_plan_set_tail with that scratch.
long attn_tail_scratch_floats(int B, int N, int heads);
void attn_plan_set_tail(AttnPlan* ap, float* scratch);
int attn_launch(const AttnPlan& ap, cudaStream_t st);
int attn_launch_count(const AttnPlan& ap);      // kernels one attn_launch enqueues
double attn_flop(const AttnPlan& ap);

}  // namespace dexb
