/* dexb200 -- C ABI of the B200-native reverse-diffusion path of DEX-TTS / GeDEX-TTS.
 *
 * The reference has no FFI: its boundary for this path is the Python call
 *     model.decoder(z, mask, mu, [ref, ref_lengths, sty, sty_lengths,] n_timesteps=, infer=True, temperature=)
 * (DEX-TTS/model/diffusion.py:250-259, GeDEX-TTS/model/diffusion.py:220-229).  The drop-in `model` package in
 * dex-tts_b200/model binds the entry points below through ctypes.
 *
 * Conventions: every pointer named *_dev is a CUDA device pointer owned by the caller and only has to stay alive
 * for the duration of the call; `stream` is a cudaStream_t passed as void*; hot calls never allocate, never
 * create streams and never synchronise the host.  Return value 0 = success, negative = error
 * (message from dexb_last_error()).  One handle per (device, model); a handle is not thread-safe, distinct handles are.
 * sm_100a only: there is no CPU fallback.
 */
#ifndef DEXB200_H_
#define DEXB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dexb_handle dexb_handle;

/* Hyper-parameters of Diffusion / DiffusionDenoiser / DiTMask (DEX-TTS/config/VCTK/base.yaml:56-78). */
typedef struct dexb_config {
  int variant;         /* 1 = DEX-TTS (TV/TIV adaptors), 0 = GeDEX-TTS */
  int dim;             /* decoder.dim (64) */
  int hidden;          /* dit.hidden_size (256) */
  int depth;           /* dit.depth (4) */
  int heads;           /* dit.num_heads (2) */
  int mlp_hidden;      /* hidden * mlp_ratio (512) */
  int patch;           /* dit.patch_size (3 | 7) */
  int stride;          /* dit.stride_size (2 | 4) */
  int conv_pos;        /* 16 */
  int conv_pos_groups; /* 8 */
  int n_feats;         /* 80 */
  float pe_scale;      /* 1000 */
  int gemm_engine;     /* 0 = tcgen05 (product), 1 = CUDA-core cross-check engine */
  int nsplit;          /* 3 = bf16x3 split precision (parity-safe default), 1 = plain bf16 operands */
  int n_spks;          /* GeDEX-TTS: > 1 adds the speaker channel spk_mlp(spk) as third input (GeDEX-TTS/model/diffusion.py:132-134,170-175) */
  int spk_emb_dim;     /* 64 */
} dexb_config;

/* Conditioning that reaches the loop: DEX-TTS style / reference tensors (DEX-TTS/model/diffusion.py:190-196,220-221) or the
 * GeDEX-TTS speaker embedding (GeDEX-TTS/model/tts.py:30-31,53; diffusion.py:170-175). */
typedef struct dexb_cond {
  const float* sty_dev;          /* (B, 2*dim, Ts) fp32: `sty` of DiffusionDenoiser.forward */
  const int32_t* sty_len_dev;    /* (B) */
  const float* ref_skips_dev[6]; /* 6 x (B, 2*dim, Tr) fp32: `ref` skip tensors of the TIV encoder */
  int Tr;
  const float* spk_dev;          /* (B, spk_emb_dim) fp32: GeDEX-TTS with n_spks > 1 (the other fields are unused then) */
} dexb_cond;

const char* dexb_last_error(void);
int dexb_version(void);

/* replaces: Diffusion.__init__ (DEX-TTS/model/diffusion.py:239-247) */
int dexb_create(const dexb_config* cfg, dexb_handle** out);
void dexb_destroy(dexb_handle* h);

/* replaces: load_state_dict for the `decoder.denoise_fn.*` tensors (DEX-TTS/synthesize.py:68-72).
 * `name` is the reference key relative to `denoise_fn.` (e.g. "downs.0.0.block1.block.0.weight"); the tensor is
 * copied, so the caller may free it.  Call dexb_finalize_weights once after the last tensor. */
int dexb_load_weight(dexb_handle* h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dexb_finalize_weights(dexb_handle* h, void* stream);

/* Allocate the workspace (owned by the handle; its size is returned in *workspace_bytes), build the TMA descriptors,
 * the per-step scalar / embedding tables and the launch plan for a (B, T, Ts, Tr, n_steps) problem.
 * `sigmas_host` = the n_steps + 1 noise levels t_0 .. t_N (t_N = 0) of ablation_sampler's 'edm' discretisation
 * (DEX-TTS/model/edm.py:152,179-180), computed by the caller in fp32 exactly as the reference does.
 * T must be a multiple of 4 (model.utils.fix_len_compatibility).  Ts = length of `sty`, Tr = length of the 6 `ref` skip
 * tensors (both = the reference mel length in synthesize.py); ignored for GeDEX-TTS. */
int dexb_plan(dexb_handle* h, int B, int T, int Ts, int Tr, int n_steps, const float* sigmas_host, size_t* workspace_bytes);

/* replaces: Diffusion.forward(infer=True) after the Gaussian draw -- i.e. ablation_sampler(...)
 * (DEX-TTS/model/edm.py:104-211) over EDMPrecond (edm.py:88-98) over DiffusionDenoiser.forward
 * (DEX-TTS/model/diffusion.py:190-236).
 *   x_inout_dev (B, 80, T): in = z / temperature + mu, out = the generated mel
 *   mu_dev (B, 80, T), mask_dev (B, T) in {0,1}; cond = NULL for single-speaker GeDEX-TTS. */
int dexb_reverse_diffusion(dexb_handle* h, float* x_inout_dev, const float* mu_dev, const float* mask_dev,
                           const dexb_cond* cond, void* stream);

/* Same call with HOST buffers (pinned or pageable): copies in, runs, copies the mel back, synchronises `stream`. */
int dexb_reverse_diffusion_host(dexb_handle* h, float* x_inout_host, const float* mu_host, const float* mask_host,
                                const float* sty_host, const int32_t* sty_len_host, const float* const* ref_skips_host,
                                int Tr, const float* spk_host, void* stream);

/* One preconditioned network call D(x; sigma_step) written to out_dev (x untouched) -- unit parity of EDMPrecond. */
int dexb_denoise_once(dexb_handle* h, const float* x_dev, const float* mu_dev, const float* mask_dev,
                      const dexb_cond* cond, int step, float* out_dev, void* stream);

/* Unit parity of the implicit-GEMM engine: out[img][h][w][n] = sum_{tap,k} A[img][h+dy][w+dx][k] * Wt[tap][n][k] + bias[n]
 * with zero padding; input pixel = output pixel * in_stride + (dy, dx), output is ceil(H/in_stride) x ceil(W/in_stride);
 * engine 0 = tcgen05, 1 = CUDA cores.  All fp32 device buffers.  Allocates scratch and synchronises (test entry). */
int dexb_gemm_test(int engine, int nsplit, const float* a_dev, int nimg, int H, int W, int K, const float* w_dev, int N,
                   int KH, int KW, int offH, int offW, int in_stride, const float* bias_dev, float* out_dev, void* stream);

/* Tuning aid: time `iters` launches of the tcgen05 implicit-GEMM engine on a synthetic problem of the given shape
 * (out_mode 0 = fp32 rows, 1 = split-bf16 rows; dbg bit 0 skips the epilogue, bit 1 the MMAs).  Allocates, synchronises. */
int dexb_gemm_bench(int nsplit, int nimg, int H, int W, int K, int N, int KH, int KW, int offH, int offW, int in_stride,
                    int out_mode, int dbg, int iters, float* ms_out);

/* Unit parity of the fused attention kernel (timm Attention inside DiTBlock, DEX-TTS/model/dit.py:270,282):
 * qkv_dev (B, N, 3*hid) fp32 = output of the qkv Linear -> out_dev (B, N, hid) = softmax(q k^T / sqrt(hd)) v, heads
 * concatenated (before the proj Linear).  Allocates scratch and synchronises (test entry). */
int dexb_attn_test(const float* qkv_dev, int B, int N, int heads, int hid, float* out_dev, void* stream);

/* replaces: TacotronSTFT.mel_spectrogram (DEX-TTS/audio/stft.py:159-178) for n_fft 1024 / hop 256 and up to 128 mel rows.
 * wav_dev (B, S) in [-1, 1]; mel_dev (B, n_mels, 1 + S/256) log-mel; mel_basis_dev (n_mels, 513), window_dev (1024);
 * energy_dev (B, 1 + S/256) = L2 norm of the magnitude spectrum of every frame (stft.py:176), or NULL (the callers discard it). */
int dexb_stft_mel(const float* wav_dev, int B, int S, const float* window_dev, const float* mel_basis_dev, int n_fft,
                  int hop, int n_mels, float* mel_dev, float* energy_dev, void* stream);

/* Test aid (failure localisation against the reference's forward hooks, oracle/make_golden.py): copies one internal activation of
 * the LAST dexb_denoise_once call as fp32 (B, C, H, W) into out_dev and reports C, H, W (out_dev == NULL: sizes only).  Names are
 * the reference modules whose output the buffer holds (DEX-TTS/model/diffusion.py:204-232): "d00", "d01" (downs.0.0/1),
 * "skip" (downs.1.2, masked), "tv_out" (tv_adaptor), "dit_out" (vit), "u00", "u01" (ups.0.0/1), "up_out" (ups.0.3). */
int dexb_debug_tap(dexb_handle* h, const char* name, float* out_dev, int* C, int* H, int* W, void* stream);

/* Profiling aid: runs network call `step` once, un-graphed, with CUDA events around every launch, on the inputs staged
 * by the last dexb_reverse_diffusion; writes one "tag<TAB>ms<TAB>gflop" line per launch into buf.  Synchronises. */
int dexb_profile_step(dexb_handle* h, int step, char* buf, size_t buflen, void* stream);

/* Number of kernels (graph nodes) launched by the last dexb_reverse_diffusion on this handle (bench.py's gpu_launches). */
long dexb_last_launch_count(const dexb_handle* h);
/* Number of GEMMs per step that the tcgen05 engine could not take (shape ineligible) and ran on CUDA cores instead. */
int dexb_simt_fallbacks(const dexb_handle* h);

/* ---- TIV encoder (the once-per-utterance stage that feeds the loop) ---------------------------------
 * replaces: TIVEncoder (DEX-TTS/model/ref_encoder.py:83-107; attached as DeXTTS.tiv_encoder, DEX-TTS/model/tts.py:28,50) in
 * eval mode: in_conv -> num_layer x { residual conv block, skip, InstanceNorm1D } -> out_conv, BatchNorm1d on running statistics.
 * Its six skip tensors are the `ref_skips_dev` of dexb_cond. */
typedef struct dexb_tiv dexb_tiv;
#define DEXB_TIV_MAX_LAYERS 16

/* replaces: TIVEncoder.__init__(c_in, c_out, num_layer, c_h) (ref_encoder.py:84-93).  c_h % 64 == 0, c_out % 32 == 0. */
int dexb_tiv_create(int c_in, int c_h, int c_out, int num_layer, dexb_tiv** out);
void dexb_tiv_destroy(dexb_tiv* h);

/* replaces: load_state_dict for the `tiv_encoder.*` tensors.  `name` is the reference key relative to `tiv_encoder.`
 * ("in_conv.conv.weight", "in_conv.bn.running_var", "conv_blocks.3.conv_block.1.conv.weight", ...); the tensor is copied.
 * Call dexb_tiv_finalize_weights once after the last tensor (packs the split-bf16 weights, folds the BatchNorm affine). */
int dexb_tiv_load_weight(dexb_tiv* h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dexb_tiv_finalize_weights(dexb_tiv* h, void* stream);

/* replaces: TIVEncoder.forward(x, mask) (ref_encoder.py:95-107).
 *   ref_dev (B, c_in, T) fp32 reference mel, mask_dev (B, T) in {0,1}
 *   skips_dev: HOST array of num_layer device pointers, each (B, c_h, T) fp32 -- the `skips` list
 *   out_dev (B, c_out, T) fp32 -- the first return value (unused by DeXTTS.forward); may be NULL to skip out_conv.
 * The first call for a new (B, T) allocates the workspace and encodes the TMA descriptors; later calls with the same (B, T)
 * only enqueue kernels on `stream` (no allocation, no host synchronisation). */
int dexb_tiv_forward(dexb_tiv* h, const float* ref_dev, const float* mask_dev, int B, int T, float* out_dev,
                     float* const* skips_dev, void* stream);
/* Number of kernels the last dexb_tiv_forward enqueued. */
long dexb_tiv_last_launch_count(const dexb_tiv* h);

/* ---- TV encoder (second piece) ---------------------------------------------------------------------
 * replaces: TVEncoder (DEX-TTS/model/ref_encoder.py:109-140; attached as DeXTTS.tv_encoder, DEX-TTS/model/tts.py:26,43) in eval
 * mode: in_conv -> num_layer residual conv blocks (channel LayerNorm) -> out_conv -> VQEmbeddingEMA nearest-code search ->
 * Projection (proj_0) -> proj_1.  z_dec is what DeXTTS.forward turns into the loop's `sty` (tts.py:48-49). */
typedef struct dexb_tv dexb_tv;

/* replaces: TVEncoder.__init__(c_in, c_out, c_out_g, num_layer, c_h, n_emb, commit_w) (ref_encoder.py:110-122).
 * c_h, c_out, c_out_g: multiples of 64, at most 256. */
int dexb_tv_create(int c_in, int c_h, int c_out, int c_out_g, int num_layer, int n_emb, float commit_w, dexb_tv** out);
void dexb_tv_destroy(dexb_tv* h);

/* replaces: load_state_dict for the `tv_encoder.*` tensors; `name` relative to `tv_encoder.` ("in_conv.ln.weight",
 * "vq.embedding", "proj_0.norm_1.gamma", "proj_1.bn.running_var", ...).  vq.ema_count / vq.ema_weight / num_batches_tracked are
 * training bookkeeping and need not be loaded.  Call dexb_tv_finalize_weights once after the last tensor. */
int dexb_tv_load_weight(dexb_tv* h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dexb_tv_finalize_weights(dexb_tv* h, void* stream);

/* replaces: TVEncoder.forward(x, mask) (ref_encoder.py:124-140).
 *   sty_dev (B, c_in, T) fp32 style mel, mask_dev (B, T) in {0,1}
 *   z_before_dev (B, c_out, T) = z_beforeVQ (may be NULL), z_dec_dev (B, c_out_g, T) = z_dec,
 *   vq_loss_dev: one float = commit_w * e_latent_loss (may be NULL), idx_dev (B, T) int32 = the chosen code per frame (may be
 *   NULL; a test / inspection aid, the reference does not return it).
 * Allocation behaviour as dexb_tiv_forward: only the first call of a (B, T) shape allocates. */
int dexb_tv_forward(dexb_tv* h, const float* sty_dev, const float* mask_dev, int B, int T, float* z_before_dev, float* z_dec_dev,
                    float* vq_loss_dev, int32_t* idx_dev, void* stream);
long dexb_tv_last_launch_count(const dexb_tv* h);

/* ---- LF0 encoder and style fusion (last pieces) ------------------------------------------------------
 * replaces: LF0Encoder (DEX-TTS/model/ref_encoder.py:36-56; attached as DeXTTS.lf0_encoder, DEX-TTS/model/tts.py:27,42) in eval
 * mode: in_conv (1 -> c_h) -> bidirectional nn.GRU(c_h, c_h/2, num_layer) over all T frames -> out_conv -> Projection. */
typedef struct dexb_lf0 dexb_lf0;

/* replaces: LF0Encoder.__init__(c_h, c_out, c_out_g, num_layer, c_in=1) (ref_encoder.py:37-44).  c_h must be 192 (the GRU
 * recurrence kernel is instantiated for 96 hidden units per direction) and c_out == c_h. */
int dexb_lf0_create(int c_h, int c_out, int c_out_g, int num_layer, dexb_lf0** out);
void dexb_lf0_destroy(dexb_lf0* h);
/* `name` relative to `lf0_encoder.` ("in_conv.conv.weight", "rnn_layer.weight_hh_l1_reverse", "proj.norm_2.beta", ...). */
int dexb_lf0_load_weight(dexb_lf0* h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dexb_lf0_finalize_weights(dexb_lf0* h, void* stream);
/* replaces: LF0Encoder.forward(lf0, mask) (ref_encoder.py:46-56).  lf0_dev (B, T) fp32, mask_dev (B, T) in {0,1} ->
 * lf0_enc_dev (B, c_out, T), lf0_dec_dev (B, c_out_g, T).  Allocation behaviour as dexb_tiv_forward. */
int dexb_lf0_forward(dexb_lf0* h, const float* lf0_dev, const float* mask_dev, int B, int T, float* lf0_enc_dev, float* lf0_dec_dev,
                     void* stream);
long dexb_lf0_last_launch_count(const dexb_lf0* h);

/* replaces: the style fusion of DeXTTS.forward (DEX-TTS/model/tts.py:45-49), stateless:
 *   sty_enc = sum_t z_before / sum_t sty_mask + sum_t lf0_enc / sum_t lf0_mask                    (B, C)      [optional]
 *   sty     = conv_sty(z_dec + (sum_t lf0_dec / sum_t lf0_mask)[:, :, None])                      (B, N, Ts)  = the loop's `sty`
 * z_before_dev / z_dec_dev (B, C, Ts) and sty_mask_dev (B, Ts) come from dexb_tv_forward, lf0_enc_dev / lf0_dec_dev (B, C, Tl)
 * and lf0_mask_dev (B, Tl) from dexb_lf0_forward; conv_sty_w_dev (N, C, 1), conv_sty_b_dev (N) are DeXTTS.conv_sty's tensors
 * (tts.py:31).  lf0_mean_scratch_dev: (B, C) floats of caller-owned scratch.  sty_enc_dev may be NULL (then z_before_dev,
 * sty_mask_dev and lf0_enc_dev are not read).  fp32 on the CUDA cores; never allocates or synchronises. */
int dexb_style_fuse(const float* z_before_dev, const float* z_dec_dev, const float* sty_mask_dev, int Ts, const float* lf0_enc_dev,
                    const float* lf0_dec_dev, const float* lf0_mask_dev, int Tl, int B, int C, const float* conv_sty_w_dev,
                    const float* conv_sty_b_dev, int N, float* lf0_mean_scratch_dev, float* sty_enc_dev, float* sty_dev, void* stream);

/* replaces: the duration / alignment glue of DeXTTS.forward between the text encoder and the decoder (DEX-TTS/model/tts.py:55-68,
 * GeDEX-TTS/model/tts.py:37-50) over model.utils.sequence_mask / generate_path (DEX-TTS/model/utils.py:6-39).  Two calls around
 * the reference's own host round trip (`int(y_lengths.max())`, tts.py:58), which decides the size of the outputs:
 *
 *   dexb_align_lengths:  w_ceil = ceil(exp(logw) * x_mask) * length_scale;  cum = cumsum_i(w_ceil)   (utils.py:30)
 *                        y_lengths = clamp_min(sum_i w_ceil, 1).long()                               (tts.py:57)
 *     logw_dev, x_mask_dev (B, Tx) fp32 -> cum_dev (B, Tx) fp32 scratch kept for the second call, y_lengths_dev (B) int64, and a
 *     copy in y_lengths_host (B int64, pinned or pageable).  SYNCHRONISES the stream (the one host sync of the path, as upstream).
 *     Additions run in token order without fma contraction (torch.cumsum's CPU order): bit-exact y_lengths whenever the partial
 *     sums are exact in fp32 -- always for the default length_scale = 1 (integers) and for dyadic scales.
 *   host:                Ty = fix_len_compatibility(max_b y_lengths)   (utils.py:13-17; an integer loop, stays in the caller)
 *   dexb_align_expand:   y_mask[b, t] = t < y_lengths[b]                                             (tts.py:62)
 *                        attn[b, i, t] = (cum[b, i-1] <= t < cum[b, i]) * x_mask[b, i] * y_mask[b, t]   (utils.py:26-39, tts.py:63-64)
 *                        mu_y[b, f, t] = sum_i attn[b, i, t] * mu_x[b, f, i]   -- one-hot in i, so a gather (tts.py:67-68)
 *     mu_x_dev (B, n_feats, Tx) -> attn_dev (B, Tx, Ty) or NULL, y_mask_dev (B, Ty), mu_y_dev (B, n_feats, Ty), all fp32.
 *     Never allocates or synchronises.  HBM-bound: one coalesced pass over the outputs. */
int dexb_align_lengths(const float* logw_dev, const float* x_mask_dev, int B, int Tx, float length_scale, float* cum_dev,
                       int64_t* y_lengths_dev, int64_t* y_lengths_host, void* stream);
int dexb_align_expand(const float* cum_dev, const float* x_mask_dev, const int64_t* y_lengths_dev, const float* mu_x_dev, int B,
                      int Tx, int n_feats, int Ty, float* attn_dev, float* y_mask_dev, float* mu_y_dev, void* stream);

/* ---- text encoder ---------------------------------------------------------------------------------
 * replaces: TextEncoder (DEX-TTS/model/text_encoder.py:97-142; attached as DeXTTS.encoder, DEX-TTS/model/tts.py:29,51; GeDEX-TTS/
 * model/text_encoder.py:99-146 / tts.py:24,34 with adaln = 0) in eval mode for n_spks <= 1 and the shipped RetNet settings
 * (use_softmax = True, use_decay = False, GLU feed-forward, pre-RMSNorm: DEX-TTS/config/VCTK/base.yaml:51-61, model/retnet_cfg.py).
 * Tensor names are the module's own state_dict keys relative to `encoder.` ("emb.weight", "prenet.conv_layers.0.weight",
 * "encoder.layers.3.retention.q_proj.weight", "proj_w.norm_1.gamma", ...).  Creation / weight loading / finalisation behave like
 * the dexb_tiv_* calls. */
typedef struct dexb_text dexb_text;
int dexb_text_create(int n_vocab, int n_feats, int n_channels, int filter_channels, int filter_channels_dp, int n_heads, int n_layers,
                     int kernel_size, int adaln, int spk_emb_dim, dexb_text** out);
void dexb_text_destroy(dexb_text* h);
int dexb_text_load_weight(dexb_text* h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dexb_text_finalize_weights(dexb_text* h, void* stream);
/* replaces: TextEncoder.forward(x, x_lengths, sty, spk=None) (text_encoder.py:129-142).  ids_dev (B, Tx) int64 phoneme ids,
 * mask_dev (B, Tx) = sequence_mask(x_lengths) in {0,1}, sty_dev (B, n_channels) style vector (NULL iff adaln = 0), spk_dev (B, spk_emb_dim)
 * speaker embedding (NULL iff spk_emb_dim = 0; n_spks > 1 of GeDEX-TTS/model/text_encoder.py:119-127,141-142: concatenated behind the
 * prenet, everything after it is n_channels + spk_emb_dim wide) ->
 * mu_dev (B, n_feats, Tx), logw_dev (B, 1, Tx).  Allocation behaviour as dexb_tiv_forward. */
int dexb_text_forward(dexb_text* h, const int64_t* ids_dev, const float* mask_dev, const float* sty_dev, const float* spk_dev, int B, int Tx,
                      float* mu_dev, float* logw_dev, void* stream);
long dexb_text_last_launch_count(const dexb_text* h);
/* unit-parity aids (no reference counterpart): n_layers >= 0 makes dexb_text_forward stop after the prenet and that many RetNet
 * layers (mu / logw are then not written), -1 restores the full forward; dexb_text_copy_stream copies the residual stream
 * (B * Tx, n_channels) fp32 rows -- the prenet output for n_layers = 0, RetNetDecoderLayer n - 1's output otherwise. */
int dexb_text_set_layer_limit(dexb_text* h, int n_layers);
int dexb_text_copy_stream(const dexb_text* h, float* rows_dev, void* stream);

/* ---- vocoder: HiFi-GAN v1 generator (the stage behind the loop, mel -> waveform) ---------------------------
 * replaces: hifigan.Generator (DEX-TTS/hifigan/models.py:112-173 over ResBlock :24-109) in the state get_vocoder leaves it in --
 * eval + remove_weight_norm (DEX-TTS/src/utils.py:251-281) -- called as `vocoder(y_dec)` at DEX-TTS/synthesize.py:106.
 * Every Conv1d / ConvTranspose1d runs on the tcgen05 implicit-GEMM engine of the loop; the forward of a (B, T) shape is one CUDA graph. */
typedef struct dexb_voc dexb_voc;
/* replaces: Generator.__init__(h) with h = hifigan/config.json: n_mels 80, upsample_initial_channel 512, upsample_rates (8,8,2,2)
 * (kernel = 2 x rate, as config.json pairs them), resblock_kernel_sizes (3,7,11), resblock_dilation_sizes (1,3,5) shared by the blocks. */
int dexb_voc_create(int n_mels, int initial_channels, const int* upsample_rates, int n_up, const int* resblock_kernels, int n_rk,
                    const int* resblock_dilations, int n_rd, dexb_voc** out);
void dexb_voc_destroy(dexb_voc* h);
/* replaces: load_state_dict + remove_weight_norm.  `name` = the generator's state_dict key after remove_weight_norm
 * ("conv_pre.weight", "ups.2.bias", "resblocks.7.convs2.1.weight", "conv_post.weight", ...); the tensor is copied. */
int dexb_voc_load_weight(dexb_voc* h, const char* name, const float* data_dev, const int64_t* shape, int ndim);
int dexb_voc_finalize_weights(dexb_voc* h, void* stream);
/* replaces: Generator.forward(x) (models.py:157-173): mel_dev (B, n_mels, T) fp32 -> wav_dev (B, 1, T * prod(upsample_rates)) fp32 in
 * [-1, 1].  The first call for a new (B, T) allocates the workspace and captures the graph; later calls only launch it on `stream`. */
int dexb_voc_forward(dexb_voc* h, const float* mel_dev, int B, int T, float* wav_dev, void* stream);
long dexb_voc_last_launch_count(const dexb_voc* h);

/* replaces: monotonic_align.maximum_path(value, mask) (DEX-TTS/model/monotonic_align/__init__.py:8-25 over the Cython kernel
 * core.pyx:9-47; call site DEX-TTS/model/tts.py:108, training).  value_dev, mask_dev (B, Tx, Ty) fp32 (mask in {0,1}, the outer product
 * of the two sequence masks) -> path_dev (B, Tx, Ty) fp32 zeros / ones.  scratch_dev: B * Tx * Ty bytes.  One CTA per utterance;
 * the per-utterance lengths are read from the mask's first column / row as upstream does.  Never allocates or synchronises. */
int dexb_mas_maximum_path(const float* value_dev, const float* mask_dev, int B, int Tx, int Ty, uint8_t* scratch_dev, float* path_dev,
                          void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DEXB200_H_ */
