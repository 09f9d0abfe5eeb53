// Launch wrappers of the non-GEMM kernels on the reverse-diffusion path (definitions in kernels_*.cu).
// All activations are NHWC; "S" = split-bf16 rows [hi(C)|lo(C)], "F" = fp32 rows.  Every wrapper only enqueues
// work on `st` (no allocation, no host sync) so the whole trajectory can be captured in a CUDA graph.
#pragma once
#include "common.cuh"
#include "gn_apply.cuh"

namespace dexb {

// one-off cudaFuncSetAttribute calls (must run outside stream capture)
int kernels_global_init();

void launch_fill_zero(void* p, size_t bytes, cudaStream_t st);

// x <- x * s   (latents * sigma_0, edm.py:184)
void launch_scale(float* x, long n, float s, cudaStream_t st);
// mask1[b][w] = mask[b][2w]
void launch_mask_down(const float* mask, float* mask1, int B, int T, int W1, cudaStream_t st);

// first conv of the U-Net: conv3x3(2 -> C) on stack[mu, c_in*x] * mask, + bias, raw fp32 out + GroupNorm partial sums
// spk_s != null: third input channel spk_s[b][h] (GeDEX-TTS speaker channel), w is [C][3][3][3]
void launch_conv_in(const float* x, const float* mu, const float* spk_s, const float* mask, const StepScalars* tab, int step,
                    const float* w /*[C][2][3][3]*/, const float* bias, float* raw /*F[M][C]*/, double* stats, int B,
                    int H, int W, int C, cudaStream_t st);

void launch_gn_apply(const GnApplyArgs& a, cudaStream_t st);

// final_block GroupNorm+Mish -> final_conv 1x1 (C->1) -> EDM preconditioning -> Euler update of x (in place);
// with den_out != null the denoised estimate D(x; sigma) is written there instead and x is left untouched
void launch_gn_final(const float* raw, int C, int G, const double* stats, const float* gamma, const float* beta,
                     const float* fc_w, const float* fc_b, const float* mask, float* x, float* den_out,
                     const StepScalars* tab, int step, int B, int H, int W, cudaStream_t st);

// LinearAttention pieces (kv = F[M][256] = [k(128) | v(128)])
void launch_la_colmax(const float* kv, unsigned* kmax_enc /*[B][128]*/, int B, int P, cudaStream_t st);
int la_ctx_blocks(int B, int P);                      // pixel blocks per image of launch_la_ctx (sizes `part`)
void launch_la_ctx(const float* kv, const unsigned* kmax_enc, float* part /*[B][blocks][4224] scratch*/,
                   float* ctx /*[B][4][32][32]*/, float* ssum /*[B][128]*/, int B, int P, cudaStream_t st);
// merge of the split-KV partials of the tensor-core context kernel -> ctx, ssum (see attn.cuh: attn_plan_init_la)
// (also applies W_v: the context kernel computes softmax(k)^T x, see attn.cuh: attn_plan_init_la); wv = v rows of to_qkv [128][C]
void launch_la_combine(const float* part_o, const float* part_l, const float* part_m, const float* wv, float* ctx, float* ssum,
                       int B, int S, int C, cudaStream_t st);
// W_eff[b] = I + g * W_out * ctxn^T * W_q  -> packed split weights [B][C][hi(C)|lo(C)], beff[b] = g * b_out
void launch_la_weff(const float* ctx, const float* ssum, const float* wq /*[128][C]*/, const float* wout /*[C][128]*/,
                    const float* bout, const float* g, bf16* weff, float* beff, int B,
                    int C, cudaStream_t st);

// per-(image, channel) sum / sumsq over all pixels of an S tensor (InstanceNorm2D statistics)
void launch_chan_stats_s(SView x, double* stats /*[B][C][2]*/, int B, int P, int C, cudaStream_t st);
void launch_chan_stats_f(const float* x, long stride, double* stats, int B, int P, int C, cudaStream_t st);

// TV adaptor: per-step fold of InstanceNorm into the key matrix, masked softmax over style tokens
void launch_tv_fold(const float* kw /*[B][NK-1][C] style-token rows, step-invariant*/, const float* kw0 /*[C] this step*/,
                    const double* stats, int P, bf16* kq /*[B][NKR][hi(C)|lo(C)]*/, float* sbias /*[B][NKR]*/, int B,
                    int NK, int NKR, int C, cudaStream_t st);
void launch_tv_vl0(const float* vl0 /*[C]*/, bf16* vlt /*[B][C][hi(KP)|lo(KP)]*/, int B, int C, int KP,
                   cudaStream_t st);
void launch_tv_softmax(const float* scores, long sstride, const int* sty_len, bf16* P_, int B, int Ppix, int NK, int KP,
                       cudaStream_t st);

void launch_freq_mean(const float* pg /*[B][Fq][Wq][D]*/, float* pe /*[B][Wq][D]*/, int B, int Fq, int Wq, int D, cudaStream_t st);
// AdaIN affine coefficients a, d [B][C] from the InstanceNorm sums and the pooled TIV scale / shift
void launch_tiv_affine(const double* stats, const float* sc, const float* sh, float* a_out, float* d_out, int B, int C, int P,
                       cudaStream_t st);
// TIV AdaIN (tiv_scale = a, tiv_shift = d from launch_tiv_affine) + DiT patch embed front: affine(InstanceNorm) -> zero pad -> depthwise conv pxp stride s -> SiLU -> S
void launch_dw_patch(const float* tv, const double* stats, const float* tiv_scale, const float* tiv_shift,
                     int use_tiv, const float* dw_w /*[C][p][p]*/, const float* dw_b, SView out, int B, int H, int W,
                     int C, int p, int s, int Fq, int Wq, cudaStream_t st);
// same front for GeDEX (no adaptors): input is an S tensor
void launch_dw_patch_s(SView in, const float* dw_w, const float* dw_b, SView out, int B, int H, int W, int C, int p,
                       int s, int Fq, int Wq, cudaStream_t st);

// tokens: x = xe + pe[b][w] + fpos[h]  (F), and LN+modulate -> S   (pe = launch_freq_mean of GELU(pos_conv))
void launch_tok_assemble(const float* xe, const float* pe, const float* fpos /*[Fq][D]*/, float* x, const float* shift,
                         const float* scale, SView out, int B, int Fq, int Wq, int D, cudaStream_t st);
void launch_ln_mod(const float* x, const float* shift, const float* scale, SView out, long M, int D, cudaStream_t st);
// attention row softmax: scores F[z][N][NS] -> P S[z][N][hi(NP)|lo(NP)] (zero padded)
void launch_attn_softmax(const float* scores, long NS, bf16* P_, long NP, long rows, int N, cudaStream_t st);
// v columns of the qkv rows -> transposed split operand vT[b][D][hi(NP)|lo(NP)] (D = heads*hd, head-major)
void launch_transpose_v(const bf16* qkv, long row_stride, int v_hi, int v_lo, bf16* vT, int B, int N, long NP, int D, int hd,
                        cudaStream_t st);
// unpatchify + crop + mask: y F[B][Fq*Wq][s*s*C] -> S[B][H][W][...]
void launch_unpatchify(const float* y, SView out, const float* mask1, int B, int Fq, int Wq, int s, int C, int H,
                       int W, cudaStream_t st);

// generic tiny linear for the per-step tables: out[r][n] = act_out( sum_k act_in(in[r][k]) * w[n][k] + b[n] )
// act codes: 0 none, 1 mish, 2 silu
void launch_small_linear(const float* in, long in_stride, const float* w, const float* b, float* out, long out_stride,
                         int R, int N, int K, int act_in, int act_out, cudaStream_t st);
// sinusoidal embeddings of the step table: mode 0 = SinusoidalPosEmb(scale*c_noise) [sin|cos], mode 1 = DiT
// timestep_embedding(c_noise) [cos|sin]
void launch_time_embed(const StepScalars* tab, int steps, float* out, int dim, float scale, int mode, cudaStream_t st);

// weight packing: fp32 [N][K] (row-major, row stride ld) -> split rows [hi(K)|lo(K)]
void launch_pack_split(const float* w, long ld, bf16* out, long out_stride, int lo_off, int N, int K, cudaStream_t st);
// conv weight [Co][Ci][KH][KW] -> [tap][Co][hi(Ci)|lo(Ci)]
void launch_pack_conv(const float* w, bf16* out, int Co, int Ci, int KH, int KW, cudaStream_t st);

// ---- once-per-call / once-per-load helpers (kernels_misc.cu) ----
void launch_ref_stats(const float* ref, float* mean, float* stdv, int B, int C, int Tr, int L, int l, cudaStream_t st);
void launch_tiv_sap(const float* t_tok, const float* rows, const float* W, const float* bias, float* out, int steps,
                    int B, int C, int L, cudaStream_t st);
void launch_transpose_scale(const float* in, float* out, int R, int Cc, float scale, cudaStream_t st);
void launch_bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t st);
void launch_tv_vlt_pack(const float* vl, bf16* vlt, int B, int Ts, int C, int KP, cudaStream_t st);
void launch_pair_pack(const float* e, bf16* pairs, int B, int Fq, int Wq, int D, int Cg, int PF, cudaStream_t st);
void launch_pack_posconv(const float* w, bf16* out, int Co, int Cg, int KP, int PF, cudaStream_t st);
void launch_pack_convT(const float* w, bf16* out, int Ci, int Co, cudaStream_t st);
int launch_stft_mel(const float* wav, int B, int S, const float* window, const float* mel_basis, int n_fft, int hop,
                    int n_mels, float* mel, float* energy /*(B, frames) or null*/, cudaStream_t st);
void launch_unpack_rows(const bf16* in, float* out, long rows, int K, cudaStream_t st);
void launch_pack_rows(const float* in, bf16* out, long rows, int K, cudaStream_t st);
// debug tap: S view (sp != null) or F rows (fp) of B images x P pixels x C channels -> fp32 (B, C, P)
void launch_tap_nchw(const bf16* sp, long s_stride, int hi, int lo, const float* fp, long f_stride, float* out, int B, int C, long P,
                     cudaStream_t st);

}  // namespace dexb
