// TIV encoder of DEX-TTS (the once-per-utterance stage that produces the six `ref` skip tensors the loop's TIVAdaptor reads):
// TIVEncoder.forward, DEX-TTS/model/ref_encoder.py:83-107 over BasicConv / InstanceNorm1D (DEX-TTS/model/base.py:33-93).
//
//   x  = relu(bn(conv3(ref * mask))) * mask                                   in_conv             (:97)
//   6x { y = (x*mask + conv3(relu(bn(conv3(x*mask))))) * mask ; skips += y ; x = inorm(y) }      (:100-103, :58-68)
//   out = relu(bn(conv3(x * mask))) * mask                                    out_conv            (:104)
//
// The 14 conv1d (k = 3, no bias) run on the tcgen05 implicit-GEMM engine as 1 x 3-tap convolutions over rows [B*T][C]
// (split-bf16 x3, fp32 accumulation -- the same engine and precision as the loop); BatchNorm1d is evaluated with its running
// statistics (eval mode, base.py:45), InstanceNorm1D takes mean / unbiased variance over ALL T frames including padding
// (cal_stats ignores x_lengths, base.py:72-78).  Everything between two convolutions is one small fused kernel.
#include <string.h>

#include <initializer_list>
#include <map>
#include <string>
#include <vector>

#include "../../include/dexb200.h"
#include "enc_graph.cuh"
#include "gemm_host.cuh"

namespace dexb {

struct TivTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  size_t n = 0;
};

struct TivConv {
  bf16* w = nullptr;                 // [3][co][hi(K)|lo(K)], K = ci padded to a multiple of 64
  float *alpha = nullptr, *beta = nullptr;   // eval BatchNorm as y = x * alpha + beta, or null (no norm)
  int ci = 0, co = 0, K = 0;
  GemmPlan plan;
};

}  // namespace dexb

struct dexb_tiv {
  int c_in = 0, c_h = 0, c_out = 0, L = 0;
  std::map<std::string, dexb::TivTensor> w;
  bool finalized = false;
  dexb::TivConv in_conv, out_conv;
  std::vector<dexb::TivConv> conv_a, conv_b;
  // plan (B, T): workspace owned by the handle
  int B = 0, T = 0;
  dexb::bf16 *xs = nullptr, *hs = nullptr;     // split rows [B*T][hi(Kmax)|lo(Kmax)]: block input / hidden activation
  float *acc = nullptr, *xf = nullptr;         // fp32 rows [B*T][c_h]: raw conv output / block input (residual)
  long launches = 0;
  dexb::EncGraph g;                            // CUDA-graph replay of the forward of this plan (enc_graph.cuh)
};

namespace dexb {

static inline int pad64(int k) { return (k + 63) / 64 * 64; }

// ---- weight packing ------------------------------------------------------------------------------------------------------------
// Conv1d weight (co, ci, 3) -> [tap][co][hi(K) | lo(K)], zero beyond ci
__global__ void k_tiv_pack_w(const float* __restrict__ w, bf16* __restrict__ out, int co, int ci, int K) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 3 * co * K) return;
  const int k = i % K, n = (i / K) % co, tap = i / (K * co);
  const float v = k < ci ? w[((long)n * ci + k) * 3 + tap] : 0.f;
  bf16 hi, lo;
  split2(v, hi, lo);
  out[((long)tap * co + n) * 2 * K + k] = hi;
  out[((long)tap * co + n) * 2 * K + K + k] = lo;
}
// eval BatchNorm1d (eps 1e-5, base.py:42): alpha = weight / sqrt(running_var + eps), beta = bias - running_mean * alpha
__global__ void k_tiv_bn_fold(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ rm,
                              const float* __restrict__ rv, float* __restrict__ alpha, float* __restrict__ beta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = g[c] * (1.f / sqrtf(rv[c] + 1e-5f));
  alpha[c] = a;
  beta[c] = b[c] - rm[c] * a;
}

// ---- activations ---------------------------------------------------------------------------------------------------------------
// ref (B, c_in, T) channel-major, mask (B, T) -> split rows of ref * mask (columns >= c_in were zeroed by the caller)
__global__ void k_tiv_in(const float* __restrict__ ref, const float* __restrict__ mask, bf16* __restrict__ xs, int B, int C, int T,
                         int K) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * C * T) return;
  const int t = (int)(i % T), c = (int)((i / T) % C), b = (int)(i / ((long)T * C));
  const float v = ref[i] * mask[(long)b * T + t];
  bf16 hi, lo;
  split2(v, hi, lo);
  bf16* row = xs + ((long)b * T + t) * 2 * K;
  row[c] = hi;
  row[K + c] = lo;
}

// acc rows [rows][C] -> v = relu(acc * alpha + beta) (* mask[row]); written as split rows (os), fp32 rows (of) and / or
// channel-major (B, C, T) (ocm); any of the three may be null
__global__ void k_tiv_bn_relu(const float* __restrict__ acc, const float* __restrict__ alpha, const float* __restrict__ beta,
                              const float* __restrict__ mask, bf16* __restrict__ os, float* __restrict__ of,
                              float* __restrict__ ocm, long rows, int C, int T) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long r = i / C;
  const int c = (int)(i % C);
  float v = fmaxf(fmaf(acc[i], alpha[c], beta[c]), 0.f);
  if (mask != nullptr) v *= mask[r];
  if (os != nullptr) {
    bf16 hi, lo;
    split2(v, hi, lo);
    os[r * 2 * C + c] = hi;
    os[r * 2 * C + C + c] = lo;
  }
  if (of != nullptr) of[i] = v;
  if (ocm != nullptr) ocm[((r / T) * C + c) * T + (r % T)] = v;
}

// One CTA = (utterance b, 32 channels), 256 threads = 32 channels x 8 time lanes.
//   y = (xf + acc) * mask -> skip (B, C, T);  x' = (y - mean_t y) / sqrt(var_t y + 1e-5) (unbiased, all T frames);
//   next block input x' * mask -> xf (fp32 rows) and xs (split rows).
__global__ void __launch_bounds__(256) k_tiv_block_out(const float* __restrict__ acc, float* __restrict__ xf,
                                                       const float* __restrict__ mask, float* __restrict__ skip,
                                                       bf16* __restrict__ xs, int C, int T) {
  pdl_wait();
  __shared__ float red[8][33];
  const int b = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), tl = threadIdx.x >> 5;
  const long row0 = (long)b * T;
  const float* m = mask + row0;
  float s = 0.f;
  for (int t = tl; t < T; t += 8) {
    const long i = (row0 + t) * C + c;
    s += (xf[i] + acc[i]) * m[t];
  }
  red[tl][threadIdx.x & 31] = s;
  __syncthreads();
  float mean = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) mean += red[j][threadIdx.x & 31];
  mean /= (float)T;
  __syncthreads();
  float q = 0.f;
  for (int t = tl; t < T; t += 8) {
    const long i = (row0 + t) * C + c;
    const float d = (xf[i] + acc[i]) * m[t] - mean;
    q = fmaf(d, d, q);
  }
  red[tl][threadIdx.x & 31] = q;
  __syncthreads();
  float var = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) var += red[j][threadIdx.x & 31];
  const float stdv = sqrtf(var / (float)(T - 1) + 1e-5f);
  for (int t = tl; t < T; t += 8) {
    const long i = (row0 + t) * C + c;
    const float y = (xf[i] + acc[i]) * m[t];
    skip[((long)b * C + c) * T + t] = y;
    const float xn = (y - mean) / stdv * m[t];
    xf[i] = xn;
    bf16 hi, lo;
    split2(xn, hi, lo);
    xs[(row0 + t) * 2 * C + c] = hi;
    xs[(row0 + t) * 2 * C + C + c] = lo;
  }
}

// ---- host ----------------------------------------------------------------------------------------------------------------------
static int tiv_get(dexb_tiv* h, const std::string& name, std::initializer_list<int64_t> shape, const float** out) {
  auto it = h->w.find(name);
  DEXB_CHECK(it != h->w.end(), "tiv encoder: weight '%s' was not loaded", name.c_str());
  const std::vector<int64_t> want(shape);
  DEXB_CHECK(it->second.shape == want, "tiv encoder: weight '%s' has the wrong shape", name.c_str());
  *out = it->second.p;
  return 0;
}

static int tiv_pack_conv(dexb_tiv* h, const std::string& prefix, int ci, int co, bool bn, TivConv* c, cudaStream_t st) {
  c->ci = ci; c->co = co; c->K = pad64(ci);
  const float* w = nullptr;
  DEXB_TRY(tiv_get(h, prefix + ".conv.weight", {co, ci, 3}, &w));
  if (c->w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->w, (size_t)3 * co * 2 * c->K * sizeof(bf16)));
  k_tiv_pack_w<<<cdiv(3L * co * c->K, 256), 256, 0, st>>>(w, c->w, co, ci, c->K);
  if (bn) {
    const float *g, *b, *rm, *rv;
    DEXB_TRY(tiv_get(h, prefix + ".bn.weight", {co}, &g));
    DEXB_TRY(tiv_get(h, prefix + ".bn.bias", {co}, &b));
    DEXB_TRY(tiv_get(h, prefix + ".bn.running_mean", {co}, &rm));
    DEXB_TRY(tiv_get(h, prefix + ".bn.running_var", {co}, &rv));
    if (c->alpha == nullptr) {
      DEXB_CUDA_OK(cudaMalloc(&c->alpha, (size_t)co * sizeof(float)));
      DEXB_CUDA_OK(cudaMalloc(&c->beta, (size_t)co * sizeof(float)));
    }
    k_tiv_bn_fold<<<cdiv(co, 128), 128, 0, st>>>(g, b, rm, rv, c->alpha, c->beta, co);
  }
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

static void tiv_free_conv(TivConv* c) {
  cudaFree(c->w); cudaFree(c->alpha); cudaFree(c->beta);
  c->w = nullptr; c->alpha = c->beta = nullptr;
}

static void tiv_release_plan(dexb_tiv* h) {
  enc_graph_release(&h->g);
  cudaFree(h->xs); cudaFree(h->hs); cudaFree(h->acc); cudaFree(h->xf);
  h->xs = h->hs = nullptr;
  h->acc = h->xf = nullptr;
  h->B = h->T = 0;
}

// conv1d(k = 3, padding 1) as a 1 x 3-tap implicit GEMM: A = split rows [B][1][T][2K], output fp32 rows [B*T][co]
static int tiv_plan_conv(dexb_tiv* h, TivConv* c, const bf16* a, float* out) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.nz = h->B; p.nheads = 1;
  p.H = 1; p.W = h->T;
  p.in_stride = 1;
  p.CH = 1; p.CW = h->T; p.OH = 1; p.OW = h->T;
  p.out_scale = 1; p.tap_sw = 1;
  p.KH = 1; p.KW = 3; p.offH = 0; p.offW = -1;
  p.K = c->K; p.N = c->co;
  p.A = a; p.a_row_stride = 2L * c->K; p.a_hi = 0; p.a_lo = c->K;
  p.Bw = c->w; p.b_row_stride = 2L * c->K; p.b_hi = 0; p.b_lo = c->K; p.b_rows_per_tap = c->co;
  p.nsplit = 3;
  p.epi.alpha = 1.f; p.epi.out_s_ncols = 1 << 30;
  p.epi.out_f32 = out; p.epi.out_f32_stride = c->co;
  p.BW = 128; p.BH = 1;                     // one image row per utterance: 1 x 128-frame tiles (frames beyond T are zero-filled by TMA)
  DEXB_TRY(gemm_plan_init(&c->plan, p, h->B, 3L * c->co, 1));
  DEXB_CHECK(c->plan.tc_ok, "tiv encoder: convolution %d -> %d is not eligible for the tcgen05 engine", c->ci, c->co);
  return 0;
}

static int tiv_plan(dexb_tiv* h, int B, int T) {
  if (B == h->B && T == h->T) return 0;
  tiv_release_plan(h);
  DEXB_TRY(gemm_global_init());
  const int Kmax = h->in_conv.K > h->c_h ? h->in_conv.K : h->c_h;
  const long rows = (long)B * T;
  DEXB_CUDA_OK(cudaMalloc(&h->xs, rows * 2 * Kmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->hs, rows * 2 * h->c_h * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->acc, rows * h->c_h * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&h->xf, rows * h->c_h * sizeof(float)));
  h->B = B; h->T = T;
  DEXB_TRY(tiv_plan_conv(h, &h->in_conv, h->xs, h->acc));
  for (int l = 0; l < h->L; ++l) {
    DEXB_TRY(tiv_plan_conv(h, &h->conv_a[l], h->xs, h->acc));
    DEXB_TRY(tiv_plan_conv(h, &h->conv_b[l], h->hs, h->acc));
  }
  DEXB_TRY(tiv_plan_conv(h, &h->out_conv, h->xs, h->acc));
  return 0;
}

}  // namespace dexb

using namespace dexb;

extern "C" {

int dexb_tiv_create(int c_in, int c_h, int c_out, int num_layer, dexb_tiv** out) {
  DEXB_CHECK(out != nullptr, "dexb_tiv_create: null argument");
  int dev = 0, major = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DEXB_CHECK(major == 10, "dexb200 is built for sm_100a only (device %d has compute capability major %d); there is no fallback",
             dev, major);
  DEXB_CHECK(c_in >= 1 && c_in <= 1024 && num_layer >= 1 && num_layer <= DEXB_TIV_MAX_LAYERS,
             "dexb_tiv_create: c_in %d / num_layer %d out of range", c_in, num_layer);
  DEXB_CHECK(c_h >= 64 && c_h % 64 == 0 && c_h <= 1024, "dexb_tiv_create: c_h = %d must be a multiple of 64", c_h);
  DEXB_CHECK(c_out >= 32 && c_out % 32 == 0 && c_out <= 1024, "dexb_tiv_create: c_out = %d must be a multiple of 32", c_out);
  dexb_tiv* h = new dexb_tiv();
  h->c_in = c_in; h->c_h = c_h; h->c_out = c_out; h->L = num_layer;
  h->conv_a.resize(num_layer);
  h->conv_b.resize(num_layer);
  *out = h;
  return 0;
}

void dexb_tiv_destroy(dexb_tiv* h) {
  if (h == nullptr) return;
  tiv_release_plan(h);
  tiv_free_conv(&h->in_conv);
  tiv_free_conv(&h->out_conv);
  for (auto& c : h->conv_a) tiv_free_conv(&c);
  for (auto& c : h->conv_b) tiv_free_conv(&c);
  for (auto& kv : h->w) cudaFree(kv.second.p);
  delete h;
}

int dexb_tiv_load_weight(dexb_tiv* h, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  DEXB_CHECK(h != nullptr && name != nullptr && data_dev != nullptr && shape != nullptr && ndim >= 1 && ndim <= 4,
             "dexb_tiv_load_weight: bad argument");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DEXB_CHECK(shape[i] >= 1, "dexb_tiv_load_weight(%s): empty dimension", name);
    n *= (size_t)shape[i];
  }
  TivTensor& t = h->w[name];
  if (t.p != nullptr && t.n != n) { cudaFree(t.p); t.p = nullptr; }
  if (t.p == nullptr) DEXB_CUDA_OK(cudaMalloc(&t.p, n * sizeof(float)));
  t.n = n;
  t.shape.assign(shape, shape + ndim);
  DEXB_CUDA_OK(cudaMemcpy(t.p, data_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return 0;
}

int dexb_tiv_finalize_weights(dexb_tiv* h, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(tiv_pack_conv(h, "in_conv", h->c_in, h->c_h, true, &h->in_conv, st));
  for (int l = 0; l < h->L; ++l) {
    const std::string p = "conv_blocks." + std::to_string(l) + ".conv_block.";
    DEXB_TRY(tiv_pack_conv(h, p + "0", h->c_h, h->c_h, true, &h->conv_a[l], st));
    DEXB_TRY(tiv_pack_conv(h, p + "1", h->c_h, h->c_h, false, &h->conv_b[l], st));
  }
  DEXB_TRY(tiv_pack_conv(h, "out_conv", h->c_h, h->c_out, true, &h->out_conv, st));
  DEXB_CUDA_OK(cudaStreamSynchronize(st));
  tiv_release_plan(h);                      // plans hold the packed-weight pointers of the previous finalize
  h->finalized = true;
  return 0;
}

}  // extern "C"

static int tiv_enqueue(dexb_tiv* h, const float* ref_dev, const float* mask_dev, int B, int T, float* out_dev, float* const* skips_dev,
                       cudaStream_t st) {
  const long rows = (long)B * T;
  const int C = h->c_h;
  h->launches = 0;
  // in_conv(ref * mask) * mask
  if (h->in_conv.K != h->c_in) DEXB_CUDA_OK(cudaMemsetAsync(h->xs, 0, rows * 2 * h->in_conv.K * sizeof(bf16), st));
  launch_pdl(k_tiv_in, dim3((unsigned)(cdiv(rows * h->c_in, 256))), dim3(256), 0, st, ref_dev, mask_dev, h->xs, B, h->c_in, T, h->in_conv.K);
  DEXB_TRY(gemm_launch(h->in_conv.plan, h->in_conv.plan.p, 0, st));
  launch_pdl(k_tiv_bn_relu, dim3((unsigned)(cdiv(rows * C, 256))), dim3(256), 0, st, h->acc, h->in_conv.alpha, h->in_conv.beta, mask_dev, h->xs, h->xf, nullptr,
                                                     rows, C, T);
  h->launches += 3;
  for (int l = 0; l < h->L; ++l) {
    DEXB_TRY(gemm_launch(h->conv_a[l].plan, h->conv_a[l].plan.p, 0, st));
    launch_pdl(k_tiv_bn_relu, dim3((unsigned)(cdiv(rows * C, 256))), dim3(256), 0, st, h->acc, h->conv_a[l].alpha, h->conv_a[l].beta, nullptr, h->hs, nullptr,
                                                       nullptr, rows, C, T);
    DEXB_TRY(gemm_launch(h->conv_b[l].plan, h->conv_b[l].plan.p, 0, st));
    launch_pdl(k_tiv_block_out, dim3(C / 32, B), dim3(256), 0, st, h->acc, h->xf, mask_dev, skips_dev[l], h->xs, C, T);
    h->launches += 4;
  }
  if (out_dev != nullptr) {                 // `ref` output of TIVEncoder.forward (unused by DeXTTS.forward, tts.py:50)
    DEXB_TRY(gemm_launch(h->out_conv.plan, h->out_conv.plan.p, 0, st));
    launch_pdl(k_tiv_bn_relu, dim3((unsigned)(cdiv(rows * h->c_out, 256))), dim3(256), 0, st, h->acc, h->out_conv.alpha, h->out_conv.beta, mask_dev, nullptr,
                                                              nullptr, out_dev, rows, h->c_out, T);
    h->launches += 2;
  }
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" {

int dexb_tiv_forward(dexb_tiv* h, const float* ref_dev, const float* mask_dev, int B, int T, float* out_dev,
                     float* const* skips_dev, void* stream) {
  DEXB_CHECK(h != nullptr && ref_dev != nullptr && mask_dev != nullptr && skips_dev != nullptr, "dexb_tiv_forward: null argument");
  DEXB_CHECK(h->finalized, "dexb_tiv_forward: call dexb_tiv_finalize_weights first");
  DEXB_CHECK(B >= 1 && T >= 2, "dexb_tiv_forward: B = %d, T = %d (InstanceNorm1D needs at least two frames)", B, T);
  for (int l = 0; l < h->L; ++l) DEXB_CHECK(skips_dev[l] != nullptr, "dexb_tiv_forward: skips_dev[%d] is null", l);
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(tiv_plan(h, B, T));
  if (!enc_graphs_on()) return tiv_enqueue(h, ref_dev, mask_dev, B, T, out_dev, skips_dev, st);
  const size_t rows = (size_t)B * T;
  for (int pass = 0; pass < 2; ++pass) {                      // pass 0 sizes the staging buffer, pass 1 uses it
    EncStage a{pass == 0 ? nullptr : h->g.stage};
    float* g_ref = a.get<float>(rows * h->c_in);
    float* g_mask = a.get<float>(rows);
    float* g_out = a.get<float>(rows * h->c_out);
    float* g_skips[DEXB_TIV_MAX_LAYERS];
    for (int l = 0; l < h->L; ++l) g_skips[l] = a.get<float>(rows * h->c_h);
    if (pass == 0) {
      if (h->g.stage == nullptr) DEXB_CUDA_OK(cudaMalloc(&h->g.stage, a.off + 256));
      continue;
    }
    DEXB_CUDA_OK(cudaMemcpyAsync(g_ref, ref_dev, rows * h->c_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(g_mask, mask_dev, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // out_conv is always part of the graph (one graph per plan, whatever the caller asks for)
    DEXB_TRY(enc_graph_run(&h->g, &h->launches, st, [&](cudaStream_t cs) { return tiv_enqueue(h, g_ref, g_mask, B, T, g_out, g_skips, cs); }));
    for (int l = 0; l < h->L; ++l)
      DEXB_CUDA_OK(cudaMemcpyAsync(skips_dev[l], g_skips[l], rows * h->c_h * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (out_dev != nullptr) DEXB_CUDA_OK(cudaMemcpyAsync(out_dev, g_out, rows * h->c_out * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

long dexb_tiv_last_launch_count(const dexb_tiv* h) { return h != nullptr ? h->launches : 0; }

}  // extern "C"
