// C ABI of the library (include/dexb200.h).  No torch types cross this boundary.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "engine.cuh"

namespace dexb {
static thread_local char g_err[1024] = "";
void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dexb

using namespace dexb;

extern "C" {

const char* dexb_last_error(void) { return g_err; }
int dexb_version(void) { return 100; }

int dexb_create(const dexb_config* cfg, dexb_handle** out) {
  DEXB_CHECK(cfg != nullptr && out != nullptr, "dexb_create: null argument");
  int dev = 0, major = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DEXB_CHECK(major == 10, "dexb200 is built for sm_100a only (device %d has compute capability major %d); there is no fallback",
             dev, major);
  DEXB_CHECK(cfg->variant == 0 || cfg->variant == 1, "variant must be 0 (GeDEX-TTS) or 1 (DEX-TTS)");
  DEXB_CHECK(cfg->nsplit == 1 || cfg->nsplit == 3, "nsplit must be 1 or 3");
  DEXB_CHECK(cfg->gemm_engine == 0 || cfg->gemm_engine == 1, "gemm_engine must be 0 or 1");
  DEXB_CHECK(cfg->n_spks <= 1 || (cfg->variant == 0 && cfg->spk_emb_dim >= 1 && cfg->spk_emb_dim <= 1024),
             "n_spks > 1 is a GeDEX-TTS option (variant 0) and needs 1 <= spk_emb_dim <= 1024");
  dexb_handle* h = new dexb_handle();
  h->cfg = *cfg;
  *out = h;
  return 0;
}

void dexb_destroy(dexb_handle* h) {
  if (h == nullptr) return;
  engine_release_plan(h);
  engine_release_weights(h);
  delete h;
}

int dexb_load_weight(dexb_handle* h, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  DEXB_CHECK(h != nullptr && name != nullptr && data_dev != nullptr && shape != nullptr && ndim >= 1 && ndim <= 8,
             "dexb_load_weight: bad argument");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DEXB_CHECK(shape[i] >= 1, "dexb_load_weight(%s): empty dimension", name);
    n *= (size_t)shape[i];
  }
  HostTensor& t = h->w[name];
  if (t.p != nullptr && t.n != n) { cudaFree(t.p); t.p = nullptr; }
  if (t.p == nullptr) DEXB_CUDA_OK(cudaMalloc(&t.p, n * sizeof(float)));
  t.n = n;
  t.shape.assign(shape, shape + ndim);
  DEXB_CUDA_OK(cudaMemcpy(t.p, data_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return 0;
}

int dexb_finalize_weights(dexb_handle* h, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  return engine_finalize(h, (cudaStream_t)stream);
}

int dexb_plan(dexb_handle* h, int B, int T, int Ts, int Tr, int n_steps, const float* sigmas_host, size_t* workspace_bytes) {
  DEXB_CHECK(h != nullptr, "null handle");
  return engine_plan(h, B, T, Ts, Tr, n_steps, sigmas_host, workspace_bytes);
}

int dexb_reverse_diffusion(dexb_handle* h, float* x_inout_dev, const float* mu_dev, const float* mask_dev,
                           const dexb_cond* cond, void* stream) {
  DEXB_CHECK(h != nullptr && x_inout_dev != nullptr && mu_dev != nullptr && mask_dev != nullptr, "null argument");
  return engine_run(h, x_inout_dev, mu_dev, mask_dev, cond, -1, nullptr, (cudaStream_t)stream);
}

int dexb_denoise_once(dexb_handle* h, const float* x_dev, const float* mu_dev, const float* mask_dev, const dexb_cond* cond,
                      int step, float* out_dev, void* stream) {
  DEXB_CHECK(h != nullptr && x_dev != nullptr && mu_dev != nullptr && mask_dev != nullptr && out_dev != nullptr, "null argument");
  DEXB_CHECK(step >= 0, "step must be >= 0");
  return engine_run(h, const_cast<float*>(x_dev), mu_dev, mask_dev, cond, step, out_dev, (cudaStream_t)stream);
}

int dexb_reverse_diffusion_host(dexb_handle* h, float* x_inout_host, const float* mu_host, const float* mask_host,
                                const float* sty_host, const int32_t* sty_len_host, const float* const* ref_skips_host, int Tr,
                                const float* spk_host, void* stream) {
  DEXB_CHECK(h != nullptr && h->planned, "dexb_reverse_diffusion_host: call dexb_plan first");
  DEXB_CHECK(x_inout_host != nullptr && mu_host != nullptr && mask_host != nullptr, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const long n0 = (long)h->B * h->H0 * h->W0;
  const int mid = 2 * h->cfg.dim;
  // staging buffers of the handle double as the device-side copies: upload straight into them
  DEXB_CUDA_OK(cudaMemcpyAsync(h->x, x_inout_host, n0 * 4, cudaMemcpyHostToDevice, st));
  DEXB_CUDA_OK(cudaMemcpyAsync(h->mu, mu_host, n0 * 4, cudaMemcpyHostToDevice, st));
  DEXB_CUDA_OK(cudaMemcpyAsync(h->mask0, mask_host, (long)h->B * h->W0 * 4, cudaMemcpyHostToDevice, st));
  dexb_cond cond;
  memset(&cond, 0, sizeof(cond));
  if (h->cfg.variant == 1) {
    DEXB_CHECK(sty_host != nullptr && sty_len_host != nullptr && ref_skips_host != nullptr, "DEX-TTS needs conditioning");
    DEXB_CHECK(Tr == h->Tr, "Tr %d != planned reference length %d", Tr, h->Tr);
    DEXB_CUDA_OK(cudaMemcpyAsync(h->sty, sty_host, (long)h->B * mid * h->Ts * 4, cudaMemcpyHostToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(h->sty_len, sty_len_host, (long)h->B * 4, cudaMemcpyHostToDevice, st));
    for (int l = 0; l < 6; ++l)
      DEXB_CUDA_OK(cudaMemcpyAsync(h->refs[l], ref_skips_host[l], (long)h->B * mid * Tr * 4, cudaMemcpyHostToDevice, st));
    cond.sty_dev = h->sty; cond.sty_len_dev = h->sty_len; cond.Tr = Tr;
    for (int l = 0; l < 6; ++l) cond.ref_skips_dev[l] = h->refs[l];
  }
  const bool spk = h->cfg.variant == 0 && h->cfg.n_spks > 1;
  if (spk) {
    DEXB_CHECK(spk_host != nullptr, "multi-speaker GeDEX-TTS needs the speaker embedding");
    DEXB_CUDA_OK(cudaMemcpyAsync(h->spk, spk_host, (long)h->B * h->cfg.spk_emb_dim * 4, cudaMemcpyHostToDevice, st));
    cond.spk_dev = h->spk;
  }
  // engine_run's device-to-device staging copies become self-copies (src == dst) and are skipped there
  DEXB_TRY(engine_run(h, h->x, h->mu, h->mask0, (h->cfg.variant == 1 || spk) ? &cond : nullptr, -1, nullptr, st));
  DEXB_CUDA_OK(cudaMemcpyAsync(x_inout_host, h->x, n0 * 4, cudaMemcpyDeviceToHost, st));
  DEXB_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int dexb_debug_tap(dexb_handle* h, const char* name, float* out_dev, int* C, int* H, int* W, void* stream) {
  DEXB_CHECK(h != nullptr && name != nullptr, "null argument");
  return engine_debug_tap(h, name, out_dev, C, H, W, (cudaStream_t)stream);
}

int dexb_profile_step(dexb_handle* h, int step, char* buf, size_t buflen, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  return engine_profile_step(h, step, buf, buflen, (cudaStream_t)stream);
}

long dexb_last_launch_count(const dexb_handle* h) { return h != nullptr ? h->launches : 0; }

int dexb_simt_fallbacks(const dexb_handle* h) {
  if (h == nullptr || !h->planned) return -1;
  int n = 0;
  auto chk = [&](const GemmPlan& g) { if (!g.tc_ok) ++n; };
  const dexb::ResnetW* rs[6] = {&h->d00, &h->d01, &h->d10, &h->d11, &h->u00, &h->u01};
  for (int i = 0; i < 6; ++i) {
    if (i != 0) chk(rs[i]->b1.conv);
    chk(rs[i]->b2.conv);
    if (rs[i]->res_w != nullptr) chk(rs[i]->res);
  }
  const dexb::LinAttW* las[3] = {&h->la0, &h->la1, &h->la2};
  for (int i = 0; i < 3; ++i) { if (!h->fused_la) chk(las[i]->kv); chk(las[i]->apply); }
  chk(h->g_down);
  for (int i = 0; i < 4; ++i) chk(h->g_up[i]);
  if (h->cfg.variant == 1 && !h->fused_tv) { chk(h->g_tvs); chk(h->g_tvo); }
  chk(h->g_pe); chk(h->g_posconv); chk(h->g_final); chk(h->fin.conv);
  for (const auto& k : h->blocks) { chk(k.qkv); chk(k.scores); chk(k.pv); chk(k.proj); chk(k.fc1); chk(k.fc2); }
  return n;
}

int dexb_gemm_test(int engine, int nsplit, const float* a_dev, int nimg, int H, int W, int K, const float* w_dev, int N, int KH,
                   int KW, int offH, int offW, int in_stride, const float* bias_dev, float* out_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_CHECK(in_stride >= 1 && nimg >= 1 && H >= 1 && W >= 1, "gemm_test: bad geometry");
  DEXB_TRY(gemm_global_init());
  const long rows = (long)nimg * H * W;
  const int taps = KH * KW;
  bf16 *as = nullptr, *ws = nullptr;
  DEXB_CUDA_OK(cudaMalloc(&as, rows * 2 * K * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&ws, (long)taps * N * 2 * K * sizeof(bf16)));
  launch_pack_rows(a_dev, as, rows, K, st);
  launch_pack_rows(w_dev, ws, (long)taps * N, K, st);
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.nz = nimg; p.nheads = 1;
  p.H = H; p.W = W;
  p.in_stride = in_stride;
  p.CH = (H + in_stride - 1) / in_stride; p.CW = (W + in_stride - 1) / in_stride;
  p.OH = p.CH; p.OW = p.CW;
  p.out_scale = 1; p.tap_sw = 1;
  p.KH = KH; p.KW = KW; p.offH = offH; p.offW = offW;
  p.K = K; p.N = N;
  p.A = as; p.a_row_stride = 2L * K; p.a_hi = 0; p.a_lo = K;
  p.Bw = ws; p.b_row_stride = 2L * K; p.b_hi = 0; p.b_lo = K; p.b_rows_per_tap = N;
  p.nsplit = nsplit;
  p.epi.alpha = 1.f; p.epi.bias = bias_dev; p.epi.out_s_ncols = 1 << 30;
  p.epi.out_f32 = out_dev; p.epi.out_f32_stride = N;
  p.BW = (p.CW >= 96) ? 128 : (p.CW >= 48 ? 64 : (p.CW >= 24 ? 32 : 16));
  p.BH = 128 / p.BW;
  GemmPlan gp;
  int r = gemm_plan_init(&gp, p, nimg, (long)taps * N, 1);
  if (r == 0 && engine == 0 && !gp.tc_ok) { set_last_error("gemm_test: shape not eligible for the tcgen05 engine"); r = -1; }
  if (r == 0) r = gemm_launch(gp, gp.p, engine, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(as);
  cudaFree(ws);
  if (r != 0) return r;
  DEXB_CHECK(e == cudaSuccess, "gemm_test: kernel failed: %s", cudaGetErrorString(e));
  return 0;
}

// operand fill for the micro-benchmark: pseudo-random bf16 in [-1, 1) (DEXB_BENCH_RANDOM=1; tensor-core power, and with it the SM clock
// under the board's power cap, depends on how many operand bits toggle -- constant operands flatter the kernel)
__global__ void k_bench_fill(bf16* p, long n, unsigned seed) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    unsigned x = (unsigned)i * 2654435761u + seed;
    x ^= x >> 15; x *= 2246822519u; x ^= x >> 13;
    p[i] = __float2bfloat16((float)(x & 0xffff) * (1.f / 32768.f) - 1.f);
  }
}

int dexb_gemm_bench(int nsplit, int nimg, int H, int W, int K, int N, int KH, int KW, int offH, int offW, int in_stride,
                    int out_mode, int dbg, int iters, float* ms_out) {
  DEXB_CHECK(ms_out != nullptr && iters >= 1, "gemm_bench: bad argument");
  DEXB_TRY(gemm_global_init());
  const long rows = (long)nimg * H * W;
  const int taps = KH * KW;
  const int CH = (H + in_stride - 1) / in_stride, CW = (W + in_stride - 1) / in_stride;
  const long orows = (long)nimg * CH * CW;
  bf16 *as = nullptr, *ws = nullptr;
  float* bias = nullptr;
  void* out = nullptr;
  DEXB_CUDA_OK(cudaMalloc(&as, rows * 2 * K * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&ws, (long)taps * N * 2 * K * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&bias, N * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&out, orows * N * 4));
  DEXB_CUDA_OK(cudaMemset(as, 0x3c, rows * 2 * K * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMemset(ws, 0x3c, (long)taps * N * 2 * K * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMemset(bias, 0, N * sizeof(float)));
  {
    const char* e = getenv("DEXB_BENCH_RANDOM");
    if (e != nullptr && atoi(e) != 0) {
      k_bench_fill<<<1024, 256>>>(as, rows * 2 * K, 1u);
      k_bench_fill<<<256, 256>>>(ws, (long)taps * N * 2 * K, 2u);
    }
  }
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.nz = nimg; p.nheads = 1; p.H = H; p.W = W; p.in_stride = in_stride;
  p.CH = CH; p.CW = CW; p.OH = CH; p.OW = CW; p.out_scale = 1; p.tap_sw = 1;
  p.KH = KH; p.KW = KW; p.offH = offH; p.offW = offW; p.K = K; p.N = N;
  p.A = as; p.a_row_stride = 2L * K; p.a_hi = 0; p.a_lo = K;
  p.Bw = ws; p.b_row_stride = 2L * K; p.b_hi = 0; p.b_lo = K; p.b_rows_per_tap = N;
  p.nsplit = nsplit; p.dbg = dbg & 7;
  p.epi.dbg_nostore = (dbg >> 3) & 1;
  if (out_mode & 2) p.epi.act = 1;
  double* gstats = nullptr;
  if (out_mode & 4) {                                       // GroupNorm(8, N) sums in the epilogue, as the U-Net convolutions carry them
    DEXB_CUDA_OK(cudaMalloc(&gstats, (size_t)nimg * 16 * kGnRep * sizeof(double)));
    DEXB_CUDA_OK(cudaMemset(gstats, 0, (size_t)nimg * 16 * kGnRep * sizeof(double)));
    p.epi.gn_stats = gstats; p.epi.gn_gs = N / 8;
  }
  out_mode &= 1;
  p.epi.alpha = 1.f; p.epi.bias = bias; p.epi.out_s_ncols = 1 << 30;
  if (out_mode == 0) { p.epi.out_f32 = (float*)out; p.epi.out_f32_stride = N; }
  else { p.epi.out_s = (bf16*)out; p.epi.out_s_stride = 2L * N; p.epi.out_s_hi = 0; p.epi.out_s_lo = N; }
  int best_bw = 128; long best = -1;
  for (int bw = 128; bw >= 8; bw >>= 1) {
    const long tiles = (long)cdiv(CH, 128 / bw) * cdiv(CW, bw);
    if (bw * in_stride > 256) continue;
    if (best < 0 || tiles < best) { best = tiles; best_bw = bw; }
  }
  p.BW = best_bw; p.BH = 128 / best_bw;
  GemmPlan gp;
  int r = gemm_plan_init(&gp, p, nimg, (long)taps * N, 1);
  if (r == 0 && !gp.tc_ok) { set_last_error("gemm_bench: shape not eligible for the tcgen05 engine"); r = -1; }
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  if (r == 0) r = gemm_launch(gp, gp.p, 0, 0);
  if (r == 0) r = gemm_launch(gp, gp.p, 0, 0);
  cudaEventRecord(e0, 0);
  for (int i = 0; i < iters && r == 0; ++i) r = gemm_launch(gp, gp.p, 0, 0);
  cudaEventRecord(e1, 0);
  cudaError_t e = cudaDeviceSynchronize();
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  *ms_out = ms / iters;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(as); cudaFree(ws); cudaFree(bias); cudaFree(out); cudaFree(gstats);
  if (r != 0) return r;
  DEXB_CHECK(e == cudaSuccess, "gemm_bench: kernel failed: %s", cudaGetErrorString(e));
  return 0;
}

int dexb_attn_test(const float* qkv_dev, int B, int N, int heads, int hid, float* out_dev, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_CHECK(qkv_dev != nullptr && out_dev != nullptr && B >= 1 && N >= 1, "attn_test: bad argument");
  DEXB_CHECK(attn_supported(hid / heads), "attn_test: head dim %d not supported", hid / heads);
  DEXB_TRY(attn_global_init());
  const long M = (long)B * N;
  const int NP = (N + 63) / 64 * 64;
  bf16 *qs = nullptr, *vT = nullptr, *os = nullptr;
  DEXB_CUDA_OK(cudaMalloc(&qs, M * 6 * hid * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&vT, (long)B * hid * 2 * NP * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&os, M * 2 * hid * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMemsetAsync(vT, 0, (long)B * hid * 2 * NP * sizeof(bf16), st));
  launch_pack_rows(qkv_dev, qs, M, 3 * hid, st);
  launch_transpose_v(qs, 6L * hid, 2 * hid, 5 * hid, vT, B, N, NP, hid, hid / heads, st);
  AttnPlan ap;
  const char* vmn = getenv("DEXB_VMN");
  int r = attn_plan_init(&ap, qs, (vmn != nullptr && vmn[0] == '1') ? nullptr : vT, os, B, N, NP, heads, hid);
  float* tail = nullptr;                                 // tail-split partials (tile counts that leave a partial last wave)
  const long tail_floats = attn_tail_scratch_floats(B, N, heads);
  if (r == 0 && tail_floats > 0) {
    if (cudaMalloc(&tail, tail_floats * sizeof(float)) == cudaSuccess) attn_plan_set_tail(&ap, tail);
  }
  if (r == 0) r = attn_launch(ap, st);
  if (r == 0) launch_unpack_rows(os, out_dev, M, hid, st);
  cudaError_t e = cudaStreamSynchronize(st);
  cudaFree(qs); cudaFree(vT); cudaFree(os);
  if (tail != nullptr) cudaFree(tail);
  if (r != 0) return r;
  DEXB_CHECK(e == cudaSuccess, "attn_test: kernel failed: %s", cudaGetErrorString(e));
  return 0;
}

int dexb_stft_mel(const float* wav_dev, int B, int S, const float* window_dev, const float* mel_basis_dev, int n_fft, int hop,
                  int n_mels, float* mel_dev, float* energy_dev, void* stream) {
  DEXB_CHECK(wav_dev != nullptr && window_dev != nullptr && mel_basis_dev != nullptr && mel_dev != nullptr, "null argument");
  return launch_stft_mel(wav_dev, B, S, window_dev, mel_basis_dev, n_fft, hop, n_mels, mel_dev, energy_dev, (cudaStream_t)stream);
}

}  // extern "C"
