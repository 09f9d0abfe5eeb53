// TV encoder, LF0 encoder and style fusion of DEX-TTS: the once-per-utterance stage that builds the loop's `sty`
// (DeXTTS.forward, DEX-TTS/model/tts.py:42-49).  Part 1: the TV encoder; part 2: LF0 encoder + fusion; part 3 (end of file): the
// text encoder (tts.py:51), which is built from the same row layout, GEMM plans and row-kernel pattern.
//
// TVEncoder.forward, DEX-TTS/model/ref_encoder.py:109-140, over BasicConv (model/base.py:33-63), Projection (:8-34),
// VQEmbeddingEMA (:181-235, eval branch) and model.base.LayerNorm (base.py:139-159).
//
//   x = ln(relu(conv3(sty * mask))) * mask                                              in_conv        (:128)
//   6x  x = (x*mask + conv3(ln(relu(conv3(x*mask))))) * mask                            conv_blocks    (:131-133, :70-81)
//   z_beforeVQ = conv3(x * mask) * mask                                                 out_conv       (:134)
//   z = nearest code of z_beforeVQ per frame (fp32 CUDA cores), vq_loss                 vq             (:135, :199-235)
//   z_dec = proj(cln(relu(conv3(cln(relu(conv3(z)))))))  (mask before every conv)       proj_0         (:137-138, :24-34)
//   z_dec = relu(bn(conv3(z_dec * mask))) * mask                                        proj_1         (:139)
//
// Same construction as the TIV encoder (tiv.cu): rows [B*T][C], one image row per utterance, every Conv1d a 1 x k-tap implicit
// GEMM on the tcgen05 engine (split-bf16 x3; conv biases in the GEMM epilogue), and ONE fused row kernel between two
// convolutions (BatchNorm-eval | ReLU | LayerNorm over channels | residual | mask | split-operand store).  The nearest-code
// search stays in fp32 on the CUDA cores: an argmin must not see split-bf16 noise, and it is 0.2 GFLOP per 8 utterances.
#include <string.h>

#include <initializer_list>
#include <map>
#include <string>
#include <vector>

#include "../../include/dexb200.h"
#include "enc_graph.cuh"
#include "gemm_host.cuh"

namespace dexb {

struct TvTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  size_t n = 0;
};

struct TvConv {
  bf16* w = nullptr;                          // [taps][co][hi(K)|lo(K)], K = ci padded to a multiple of 64
  const float* bias = nullptr;                // conv bias (GEMM epilogue) or null
  float *bn_a = nullptr, *bn_b = nullptr;     // eval BatchNorm as y = x * a + b, or null
  const float *ln_g = nullptr, *ln_b = nullptr;
  int ci = 0, co = 0, K = 0, taps = 3;
  GemmPlan plan;
};

// what the fused row kernel does after a convolution
struct TvPost {
  const float* acc;          // [rows][C] raw conv output (bias already added by the GEMM epilogue)
  const float *bn_a, *bn_b;  // [C] or null
  int relu;
  int ln_mode;               // 0 none, 1 nn.LayerNorm (1 / sqrt(var + eps)), 2 model.base.LayerNorm (rsqrt(var + eps))
  float ln_eps;
  const float *ln_g, *ln_b;  // [C]
  const float* resid;        // fp32 rows added after the norm, or null
  const float* mask;         // [rows] or null
  bf16* os;                  // split rows [rows][hi(C)|lo(C)] or null
  float* of;                 // fp32 rows or null
  float* ocm;                // channel-major (B, C, T) or null
  long rows;
  int C, T;
};

// state every encoder handle of this file shares: loaded tensors, the (B, T) plan and its row buffers
struct EncBase {
  std::map<std::string, TvTensor> w;
  bool finalized = false;
  int B = 0, T = 0;
  bf16 *xs = nullptr, *hs = nullptr;           // split rows, up to 2 * max(K) columns
  float *acc = nullptr, *xf = nullptr;         // fp32 rows, up to max(C) columns
  long launches = 0;
  EncGraph g;                                  // CUDA-graph replay of the forward of this plan (enc_graph.cuh)
};

}  // namespace dexb

struct dexb_tv : dexb::EncBase {
  int c_in = 0, c_h = 0, c_out = 0, c_g = 0, L = 0, n_emb = 0;
  float commit_w = 0.25f;
  dexb::TvConv in_conv, out_conv, p_conv1, p_conv2, p_proj, proj1;
  std::vector<dexb::TvConv> conv_a, conv_b;
  const float* codebook = nullptr;             // (n_emb, c_out)
  double* loss_acc = nullptr;                  // [2]: sum of squared code distances over valid frames, number of valid frames
};

namespace dexb {

static inline int tv_pad64(int k) { return (k + 63) / 64 * 64; }
constexpr int kTvMaxC = 256;                   // channels per row the fused row kernel keeps in registers (8 per lane)

// ---- weight packing ------------------------------------------------------------------------------------------------------------
__global__ void k_tv_pack_w(const float* __restrict__ w, bf16* __restrict__ out, int co, int ci, int K, int taps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= taps * co * K) return;
  const int k = i % K, n = (i / K) % co, tap = i / (K * co);
  const float v = k < ci ? w[((long)n * ci + k) * taps + tap] : 0.f;
  bf16 hi, lo;
  split2(v, hi, lo);
  out[((long)tap * co + n) * 2 * K + k] = hi;
  out[((long)tap * co + n) * 2 * K + K + k] = lo;
}
__global__ void k_tv_bn_fold(const float* __restrict__ g, const float* __restrict__ b, const float* __restrict__ rm,
                             const float* __restrict__ rv, float* __restrict__ alpha, float* __restrict__ beta, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float a = g[c] * (1.f / sqrtf(rv[c] + 1e-5f));
  alpha[c] = a;
  beta[c] = b[c] - rm[c] * a;
}

// ---- activations ---------------------------------------------------------------------------------------------------------------
// sty (B, c_in, T) channel-major, mask (B, T) -> split rows of sty * mask (columns >= c_in zeroed by the caller)
__global__ void k_tv_in(const float* __restrict__ x, const float* __restrict__ mask, bf16* __restrict__ xs, int B, int C, int T,
                        int K) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * C * T) return;
  const int t = (int)(i % T), c = (int)((i / T) % C), b = (int)(i / ((long)T * C));
  const float v = x[i] * mask[(long)b * T + t];
  bf16 hi, lo;
  split2(v, hi, lo);
  bf16* row = xs + ((long)b * T + t) * 2 * K;
  row[c] = hi;
  row[K + c] = lo;
}

// one warp per row (frame): everything between two convolutions
__global__ void __launch_bounds__(256) k_tv_post(const TvPost p) {
  pdl_wait();
  const long r = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  const int lane = threadIdx.x & 31;
  const int C = p.C;
  float v[kTvMaxC / 32];
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const int c = lane + 32 * j;
    float x = 0.f;
    if (c < C) {
      x = p.acc[r * C + c];
      if (p.bn_a != nullptr) x = fmaf(x, p.bn_a[c], p.bn_b[c]);
      if (p.relu) x = fmaxf(x, 0.f);
    }
    v[j] = x;
  }
  if (p.ln_mode != 0) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) s += v[j];                 // lanes beyond C hold zeros
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const float d = (lane + 32 * j < C) ? v[j] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float var = warp_sum(q) / (float)C;                         // biased, as both LayerNorm flavours use
    const float rstd = p.ln_mode == 1 ? 1.f / sqrtf(var + p.ln_eps) : rsqrtf(var + p.ln_eps);
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const int c = lane + 32 * j;
      if (c < C) v[j] = (v[j] - mean) * rstd * p.ln_g[c] + p.ln_b[c];
    }
  }
  const float m = p.mask != nullptr ? p.mask[r] : 1.f;
  const long b = r / p.T;
  const int t = (int)(r % p.T);
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const int c = lane + 32 * j;
    if (c >= C) continue;
    float x = v[j];
    if (p.resid != nullptr) x += p.resid[r * C + c];
    x *= m;
    if (p.os != nullptr) {
      bf16 hi, lo;
      split2(x, hi, lo);
      p.os[r * 2 * C + c] = hi;
      p.os[r * 2 * C + C + c] = lo;
    }
    if (p.of != nullptr) p.of[r * C + c] = x;
    if (p.ocm != nullptr) p.ocm[(b * C + c) * p.T + t] = x;
  }
}

// Nearest-code search, VQEmbeddingEMA.forward (ref_encoder.py:199-231), eval mode.  One warp = 4 frames; the codebook streams
// through every warp once (coalesced rows, L2 resident: n_emb * D * 4 B = 393 KB).  Distances are the direct fp32 sums
// sum_d (x_d - e_d)^2 (the reference expands them to |e|^2 + |x|^2 - 2 x.e; both orderings agree far below the code gaps);
// ties resolve to the lowest index like torch.argmin.  Output rows: (x + (e - x)) * mask as the split operand of proj_0.conv_1.
constexpr int kVqRows = 4;
__global__ void __launch_bounds__(256) k_tv_vq(const float* __restrict__ zf, const float* __restrict__ code,
                                               const float* __restrict__ mask, bf16* __restrict__ os, int* __restrict__ idx_out,
                                               double* __restrict__ loss_acc, long rows, int D, int M) {
  pdl_wait();
  const long r0 = (blockIdx.x * 8L + (threadIdx.x >> 5)) * kVqRows;
  if (r0 >= rows) return;
  const int lane = threadIdx.x & 31;
  float x[kVqRows][kTvMaxC / 32];
#pragma unroll
  for (int i = 0; i < kVqRows; ++i)
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const int d = lane + 32 * j;
      x[i][j] = (r0 + i < rows && d < D) ? zf[(r0 + i) * D + d] : 0.f;     // zf is z_beforeVQ * mask already
    }
  float best[kVqRows];
  int bi[kVqRows];
#pragma unroll
  for (int i = 0; i < kVqRows; ++i) { best[i] = INFINITY; bi[i] = 0; }
  for (int m = 0; m < M; ++m) {
    float e[kTvMaxC / 32];
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const int d = lane + 32 * j;
      e[j] = d < D ? code[(long)m * D + d] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kVqRows; ++i) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < kTvMaxC / 32; ++j) {
        const float df = x[i][j] - e[j];
        s = fmaf(df, df, s);
      }
      s = warp_sum(s);
      if (s < best[i]) { best[i] = s; bi[i] = m; }
    }
  }
  double lsum = 0.0, lcnt = 0.0;
#pragma unroll
  for (int i = 0; i < kVqRows; ++i) {
    const long r = r0 + i;
    if (r >= rows) continue;
    const float mk = mask[r];
    if (idx_out != nullptr && lane == 0) idx_out[r] = bi[i];
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const int d = lane + 32 * j;
      if (d >= D) continue;
      const float q = code[(long)bi[i] * D + d];
      const float dl = x[i][j] * mk - q * mk;                           // e_latent_loss term (:223)
      sq = fmaf(dl, dl, sq);
      const float o = (x[i][j] + (q - x[i][j])) * mk;                   // straight-through value, then mask (:227,233)
      bf16 hi, lo;
      split2(o, hi, lo);
      os[r * 2 * D + d] = hi;
      os[r * 2 * D + D + d] = lo;
    }
    sq = warp_sum(sq);
    lsum += (double)sq;
    lcnt += (double)mk;
  }
  if (lane == 0) {
    atomicAdd(&loss_acc[0], lsum);
    atomicAdd(&loss_acc[1], lcnt);
  }
}
__global__ void k_tv_loss(const double* __restrict__ loss_acc, float* __restrict__ out, float commit_w, int D) {
  pdl_wait();
  // commitment_cost * sum((x*m - q*m)^2) / (sum(m) * D)   (:223-224)
  out[0] = commit_w * (float)(loss_acc[0] / (loss_acc[1] * (double)D));
}

// ---- host ----------------------------------------------------------------------------------------------------------------------
static int tv_get(EncBase* h, const std::string& name, std::initializer_list<int64_t> shape, const float** out) {
  auto it = h->w.find(name);
  DEXB_CHECK(it != h->w.end(), "style encoders: weight '%s' was not loaded", name.c_str());
  const std::vector<int64_t> want(shape);
  DEXB_CHECK(it->second.shape == want, "style encoders: weight '%s' has the wrong shape", name.c_str());
  *out = it->second.p;
  return 0;
}

// norm: 0 none, 1 ".ln" (nn.LayerNorm of BasicConv), 2 ".bn" (eval BatchNorm of BasicConv)
static int tv_pack_conv(EncBase* h, const std::string& wname, const std::string& bname, int ci, int co, int taps, TvConv* c,
                        cudaStream_t st) {
  c->ci = ci; c->co = co; c->K = tv_pad64(ci); c->taps = taps;
  const float* w = nullptr;
  DEXB_TRY(tv_get(h, wname, {co, ci, taps}, &w));
  if (c->w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->w, (size_t)taps * co * 2 * c->K * sizeof(bf16)));
  k_tv_pack_w<<<cdiv((long)taps * co * c->K, 256), 256, 0, st>>>(w, c->w, co, ci, c->K, taps);
  c->bias = nullptr;
  if (!bname.empty()) DEXB_TRY(tv_get(h, bname, {co}, &c->bias));
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}
static int tv_basic_conv(EncBase* h, const std::string& p, int ci, int co, int norm, TvConv* c, cudaStream_t st) {
  DEXB_TRY(tv_pack_conv(h, p + ".conv.weight", "", ci, co, 3, c, st));
  if (norm == 1) {
    DEXB_TRY(tv_get(h, p + ".ln.weight", {co}, &c->ln_g));
    DEXB_TRY(tv_get(h, p + ".ln.bias", {co}, &c->ln_b));
  } else if (norm == 2) {
    const float *g, *b, *rm, *rv;
    DEXB_TRY(tv_get(h, p + ".bn.weight", {co}, &g));
    DEXB_TRY(tv_get(h, p + ".bn.bias", {co}, &b));
    DEXB_TRY(tv_get(h, p + ".bn.running_mean", {co}, &rm));
    DEXB_TRY(tv_get(h, p + ".bn.running_var", {co}, &rv));
    if (c->bn_a == nullptr) {
      DEXB_CUDA_OK(cudaMalloc(&c->bn_a, (size_t)co * sizeof(float)));
      DEXB_CUDA_OK(cudaMalloc(&c->bn_b, (size_t)co * sizeof(float)));
    }
    k_tv_bn_fold<<<cdiv(co, 128), 128, 0, st>>>(g, b, rm, rv, c->bn_a, c->bn_b, co);
    DEXB_CUDA_OK(cudaGetLastError());
  }
  return 0;
}

static void tv_free_conv(TvConv* c) {
  cudaFree(c->w); cudaFree(c->bn_a); cudaFree(c->bn_b);
  c->w = nullptr; c->bn_a = c->bn_b = nullptr;
}

static void enc_release_rows(EncBase* h) {
  enc_graph_release(&h->g);
  cudaFree(h->xs); cudaFree(h->hs); cudaFree(h->acc); cudaFree(h->xf);
  h->xs = h->hs = nullptr;
  h->acc = h->xf = nullptr;
  h->B = h->T = 0;
}
static void tv_release_plan(dexb_tv* h) {
  enc_release_rows(h);
  cudaFree(h->loss_acc);
  h->loss_acc = nullptr;
}

// Conv1d(k = taps, padding taps / 2) as a 1 x taps implicit GEMM: A = split rows [B][1][T][2K], output fp32 rows [B*T][co]
static int tv_plan_conv(EncBase* h, TvConv* c, const bf16* a, float* out) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.nz = h->B; p.nheads = 1;
  p.H = 1; p.W = h->T;
  p.in_stride = 1;
  p.CH = 1; p.CW = h->T; p.OH = 1; p.OW = h->T;
  p.out_scale = 1; p.tap_sw = 1;
  p.KH = 1; p.KW = c->taps; p.offH = 0; p.offW = -(c->taps / 2);
  p.K = c->K; p.N = c->co;
  p.A = a; p.a_row_stride = 2L * c->K; p.a_hi = 0; p.a_lo = c->K;
  p.Bw = c->w; p.b_row_stride = 2L * c->K; p.b_hi = 0; p.b_lo = c->K; p.b_rows_per_tap = c->co;
  p.nsplit = 3;
  p.epi.alpha = 1.f; p.epi.out_s_ncols = 1 << 30;
  p.epi.bias = c->bias;
  p.epi.out_f32 = out; p.epi.out_f32_stride = c->co;
  p.BW = 128; p.BH = 1;                     // one image row per utterance: 1 x 128-frame tiles (frames beyond T are zero-filled by TMA)
  DEXB_TRY(gemm_plan_init(&c->plan, p, h->B, (long)c->taps * c->co, 1));
  DEXB_CHECK(c->plan.tc_ok, "style encoders: convolution %d -> %d is not eligible for the tcgen05 engine", c->ci, c->co);
  return 0;
}

static int tv_plan(dexb_tv* h, int B, int T) {
  if (B == h->B && T == h->T) return 0;
  tv_release_plan(h);
  DEXB_TRY(gemm_global_init());
  int Kmax = h->in_conv.K, Cmax = h->c_h;
  if (h->c_h > Kmax) Kmax = h->c_h;
  if (tv_pad64(h->c_out) > Kmax) Kmax = tv_pad64(h->c_out);
  if (tv_pad64(h->c_g) > Kmax) Kmax = tv_pad64(h->c_g);
  if (h->c_out > Cmax) Cmax = h->c_out;
  if (h->c_g > Cmax) Cmax = h->c_g;
  const long rows = (long)B * T;
  DEXB_CUDA_OK(cudaMalloc(&h->xs, rows * 2 * Kmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->hs, rows * 2 * Kmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->acc, rows * Cmax * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&h->xf, rows * Cmax * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&h->loss_acc, 2 * sizeof(double)));
  h->B = B; h->T = T;
  DEXB_TRY(tv_plan_conv(h, &h->in_conv, h->xs, h->acc));
  for (int l = 0; l < h->L; ++l) {
    DEXB_TRY(tv_plan_conv(h, &h->conv_a[l], h->xs, h->acc));
    DEXB_TRY(tv_plan_conv(h, &h->conv_b[l], h->hs, h->acc));
  }
  DEXB_TRY(tv_plan_conv(h, &h->out_conv, h->xs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->p_conv1, h->xs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->p_conv2, h->hs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->p_proj, h->xs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->proj1, h->hs, h->acc));
  return 0;
}

static TvPost tv_post(const EncBase* h, const TvConv& c, int relu, int ln_mode, float ln_eps) {
  TvPost p;
  memset(&p, 0, sizeof(p));
  p.acc = h->acc;
  p.bn_a = c.bn_a; p.bn_b = c.bn_b;
  p.relu = relu;
  p.ln_mode = ln_mode; p.ln_eps = ln_eps; p.ln_g = c.ln_g; p.ln_b = c.ln_b;
  p.rows = (long)h->B * h->T;
  p.C = c.co; p.T = h->T;
  return p;
}
static void tv_launch_post(const TvPost& p, cudaStream_t st) { launch_pdl(k_tv_post, dim3((unsigned)(cdiv(p.rows, 8))), dim3(256), 0, st, p); }

}  // namespace dexb

using namespace dexb;

extern "C" {

int dexb_tv_create(int c_in, int c_h, int c_out, int c_out_g, int num_layer, int n_emb, float commit_w, dexb_tv** out) {
  DEXB_CHECK(out != nullptr, "dexb_tv_create: null argument");
  int dev = 0, major = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DEXB_CHECK(major == 10, "dexb200 is built for sm_100a only (device %d has compute capability major %d); there is no fallback",
             dev, major);
  DEXB_CHECK(c_in >= 1 && c_in <= kTvMaxC && num_layer >= 1 && num_layer <= DEXB_TIV_MAX_LAYERS && n_emb >= 1,
             "dexb_tv_create: c_in %d / num_layer %d / n_emb %d out of range", c_in, num_layer, n_emb);
  DEXB_CHECK(c_h >= 64 && c_h % 64 == 0 && c_h <= kTvMaxC, "dexb_tv_create: c_h = %d must be a multiple of 64 (<= %d)", c_h, kTvMaxC);
  DEXB_CHECK(c_out >= 64 && c_out % 64 == 0 && c_out <= kTvMaxC && c_out_g >= 64 && c_out_g % 64 == 0 && c_out_g <= kTvMaxC,
             "dexb_tv_create: c_out = %d / c_out_g = %d must be multiples of 64 (<= %d)", c_out, c_out_g, kTvMaxC);
  dexb_tv* h = new dexb_tv();
  h->c_in = c_in; h->c_h = c_h; h->c_out = c_out; h->c_g = c_out_g; h->L = num_layer; h->n_emb = n_emb; h->commit_w = commit_w;
  h->conv_a.resize(num_layer);
  h->conv_b.resize(num_layer);
  *out = h;
  return 0;
}

void dexb_tv_destroy(dexb_tv* h) {
  if (h == nullptr) return;
  tv_release_plan(h);
  TvConv* cs[6] = {&h->in_conv, &h->out_conv, &h->p_conv1, &h->p_conv2, &h->p_proj, &h->proj1};
  for (TvConv* c : cs) tv_free_conv(c);
  for (auto& c : h->conv_a) tv_free_conv(&c);
  for (auto& c : h->conv_b) tv_free_conv(&c);
  for (auto& kv : h->w) cudaFree(kv.second.p);
  delete h;
}

int dexb_tv_load_weight(dexb_tv* h, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  DEXB_CHECK(h != nullptr && name != nullptr && data_dev != nullptr && shape != nullptr && ndim >= 1 && ndim <= 4,
             "dexb_tv_load_weight: bad argument");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DEXB_CHECK(shape[i] >= 1, "dexb_tv_load_weight(%s): empty dimension", name);
    n *= (size_t)shape[i];
  }
  TvTensor& t = h->w[name];
  if (t.p != nullptr && t.n != n) { cudaFree(t.p); t.p = nullptr; }
  if (t.p == nullptr) DEXB_CUDA_OK(cudaMalloc(&t.p, n * sizeof(float)));
  t.n = n;
  t.shape.assign(shape, shape + ndim);
  DEXB_CUDA_OK(cudaMemcpy(t.p, data_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return 0;
}

int dexb_tv_finalize_weights(dexb_tv* h, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(tv_basic_conv(h, "in_conv", h->c_in, h->c_h, 1, &h->in_conv, st));
  for (int l = 0; l < h->L; ++l) {
    const std::string p = "conv_blocks." + std::to_string(l) + ".conv_block.";
    DEXB_TRY(tv_basic_conv(h, p + "0", h->c_h, h->c_h, 1, &h->conv_a[l], st));
    DEXB_TRY(tv_basic_conv(h, p + "1", h->c_h, h->c_h, 0, &h->conv_b[l], st));
  }
  DEXB_TRY(tv_basic_conv(h, "out_conv", h->c_h, h->c_out, 0, &h->out_conv, st));
  DEXB_TRY(tv_get(h, "vq.embedding", {h->n_emb, h->c_out}, &h->codebook));
  DEXB_TRY(tv_pack_conv(h, "proj_0.conv_1.weight", "proj_0.conv_1.bias", h->c_out, h->c_g, 3, &h->p_conv1, st));
  DEXB_TRY(tv_get(h, "proj_0.norm_1.gamma", {h->c_g}, &h->p_conv1.ln_g));
  DEXB_TRY(tv_get(h, "proj_0.norm_1.beta", {h->c_g}, &h->p_conv1.ln_b));
  DEXB_TRY(tv_pack_conv(h, "proj_0.conv_2.weight", "proj_0.conv_2.bias", h->c_g, h->c_g, 3, &h->p_conv2, st));
  DEXB_TRY(tv_get(h, "proj_0.norm_2.gamma", {h->c_g}, &h->p_conv2.ln_g));
  DEXB_TRY(tv_get(h, "proj_0.norm_2.beta", {h->c_g}, &h->p_conv2.ln_b));
  DEXB_TRY(tv_pack_conv(h, "proj_0.proj.weight", "proj_0.proj.bias", h->c_g, h->c_g, 1, &h->p_proj, st));
  DEXB_TRY(tv_basic_conv(h, "proj_1", h->c_g, h->c_g, 2, &h->proj1, st));
  DEXB_CUDA_OK(cudaStreamSynchronize(st));
  tv_release_plan(h);                       // plans hold the packed-weight pointers of the previous finalize
  h->finalized = true;
  return 0;
}

static int tv_enqueue(dexb_tv* h, const float* sty_dev, const float* mask_dev, int B, int T, float* z_before_dev, float* z_dec_dev,
                      float* vq_loss_dev, int32_t* idx_dev, cudaStream_t st) {
  const long rows = (long)B * T;
  h->launches = 0;
  // in_conv(sty * mask) * mask
  if (h->in_conv.K != h->c_in) DEXB_CUDA_OK(cudaMemsetAsync(h->xs, 0, rows * 2 * h->in_conv.K * sizeof(bf16), st));
  DEXB_CUDA_OK(cudaMemsetAsync(h->loss_acc, 0, 2 * sizeof(double), st));
  launch_pdl(k_tv_in, dim3((unsigned)(cdiv(rows * h->c_in, 256))), dim3(256), 0, st, sty_dev, mask_dev, h->xs, B, h->c_in, T, h->in_conv.K);
  DEXB_TRY(gemm_launch(h->in_conv.plan, h->in_conv.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->in_conv, 1, 1, 1e-5f);
    p.mask = mask_dev; p.os = h->xs; p.of = h->xf;
    tv_launch_post(p, st);
  }
  h->launches += 3;
  for (int l = 0; l < h->L; ++l) {
    DEXB_TRY(gemm_launch(h->conv_a[l].plan, h->conv_a[l].plan.p, 0, st));
    TvPost pa = tv_post(h, h->conv_a[l], 1, 1, 1e-5f);
    pa.os = h->hs;
    tv_launch_post(pa, st);
    DEXB_TRY(gemm_launch(h->conv_b[l].plan, h->conv_b[l].plan.p, 0, st));
    TvPost pb = tv_post(h, h->conv_b[l], 0, 0, 0.f);
    pb.resid = h->xf; pb.mask = mask_dev; pb.os = h->xs; pb.of = h->xf;      // in place: every element is read and written by one lane
    tv_launch_post(pb, st);
    h->launches += 4;
  }
  // z_beforeVQ = out_conv(x * mask) * mask
  DEXB_TRY(gemm_launch(h->out_conv.plan, h->out_conv.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->out_conv, 0, 0, 0.f);
    p.mask = mask_dev; p.of = h->xf; p.ocm = z_before_dev;
    tv_launch_post(p, st);
  }
  // vector quantisation -> split operand of proj_0.conv_1, loss
  launch_pdl(k_tv_vq, dim3((unsigned)(cdiv(cdiv(rows, kVqRows), 8))), dim3(256), 0, st, h->xf, h->codebook, mask_dev, h->xs, idx_dev, h->loss_acc, rows, h->c_out,
                                                       h->n_emb);
  h->launches += 3;
  if (vq_loss_dev != nullptr) {
    launch_pdl(k_tv_loss, dim3((unsigned)(1)), dim3(1), 0, st, h->loss_acc, vq_loss_dev, h->commit_w, h->c_out);
    h->launches += 1;
  }
  // proj_0: conv_1 -> relu -> norm_1 -> (mask) conv_2 -> relu -> norm_2 -> (mask) proj -> mask
  DEXB_TRY(gemm_launch(h->p_conv1.plan, h->p_conv1.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->p_conv1, 1, 2, 1e-4f);
    p.mask = mask_dev; p.os = h->hs;
    tv_launch_post(p, st);
  }
  DEXB_TRY(gemm_launch(h->p_conv2.plan, h->p_conv2.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->p_conv2, 1, 2, 1e-4f);
    p.mask = mask_dev; p.os = h->xs;
    tv_launch_post(p, st);
  }
  DEXB_TRY(gemm_launch(h->p_proj.plan, h->p_proj.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->p_proj, 0, 0, 0.f);
    p.mask = mask_dev; p.os = h->hs;
    tv_launch_post(p, st);
  }
  // proj_1: relu(bn(conv3(z_dec * mask))) * mask
  DEXB_TRY(gemm_launch(h->proj1.plan, h->proj1.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->proj1, 1, 0, 0.f);
    p.mask = mask_dev; p.ocm = z_dec_dev;
    tv_launch_post(p, st);
  }
  h->launches += 8;
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

int dexb_tv_forward(dexb_tv* h, const float* sty_dev, const float* mask_dev, int B, int T, float* z_before_dev, float* z_dec_dev,
                    float* vq_loss_dev, int32_t* idx_dev, void* stream) {
  DEXB_CHECK(h != nullptr && sty_dev != nullptr && mask_dev != nullptr && z_dec_dev != nullptr, "dexb_tv_forward: null argument");
  DEXB_CHECK(h->finalized, "dexb_tv_forward: call dexb_tv_finalize_weights first");
  DEXB_CHECK(B >= 1 && T >= 1, "dexb_tv_forward: B = %d, T = %d", B, T);
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(tv_plan(h, B, T));
  if (!enc_graphs_on()) return tv_enqueue(h, sty_dev, mask_dev, B, T, z_before_dev, z_dec_dev, vq_loss_dev, idx_dev, st);
  const size_t rows = (size_t)B * T;
  for (int pass = 0; pass < 2; ++pass) {                      // pass 0 sizes the staging buffer, pass 1 uses it
    EncStage a{pass == 0 ? nullptr : h->g.stage};
    float* g_sty = a.get<float>(rows * h->c_in);
    float* g_mask = a.get<float>(rows);
    float* g_zb = a.get<float>(rows * h->c_out);
    float* g_zd = a.get<float>(rows * h->c_g);
    float* g_loss = a.get<float>(1);
    int32_t* g_idx = a.get<int32_t>(rows);
    if (pass == 0) {
      if (h->g.stage == nullptr) DEXB_CUDA_OK(cudaMalloc(&h->g.stage, a.off + 256));
      continue;
    }
    DEXB_CUDA_OK(cudaMemcpyAsync(g_sty, sty_dev, rows * h->c_in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(g_mask, mask_dev, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    // every optional output is always produced into the staging buffers (one graph per plan, whatever the caller asks for)
    DEXB_TRY(enc_graph_run(&h->g, &h->launches, st, [&](cudaStream_t cs) { return tv_enqueue(h, g_sty, g_mask, B, T, g_zb, g_zd, g_loss, g_idx, cs); }));
    if (z_before_dev != nullptr) DEXB_CUDA_OK(cudaMemcpyAsync(z_before_dev, g_zb, rows * h->c_out * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(z_dec_dev, g_zd, rows * h->c_g * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (vq_loss_dev != nullptr) DEXB_CUDA_OK(cudaMemcpyAsync(vq_loss_dev, g_loss, sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (idx_dev != nullptr) DEXB_CUDA_OK(cudaMemcpyAsync(idx_dev, g_idx, rows * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

long dexb_tv_last_launch_count(const dexb_tv* h) { return h != nullptr ? h->launches : 0; }

}  // extern "C"


// ================================================================================================================================
// Part 2: LF0 encoder (LF0Encoder.forward, DEX-TTS/model/ref_encoder.py:36-56, eval) and the style fusion of DeXTTS.forward
// (DEX-TTS/model/tts.py:45-49).
//
//   x = ln(relu(conv3(lf0 * mask))) * mask                       in_conv, 1 -> c_h channels: CUDA cores, fused   (:49)
//   x = BiGRU_{num_layer}(x)  over ALL T frames (no packing)     input projections on the tcgen05 engine (both directions in one
//                                                                GEMM, b_ih in the epilogue), recurrence in fp32 on CUDA cores:
//                                                                one CTA per (utterance, direction), W_hh in registers   (:50)
//   lf0_enc = ln(relu(conv3(x * mask))) * mask                   out_conv                                          (:51)
//   lf0_dec = Projection(lf0_enc)                                proj                                              (:53-54)
// ================================================================================================================================
struct dexb_lf0 : dexb::EncBase {
  int c_h = 0, c_out = 0, c_g = 0, L = 0, H = 0;
  const float *in_w = nullptr, *in_g = nullptr, *in_b = nullptr;     // in_conv (c_h, 1, 3), ln affine
  std::vector<dexb::TvConv> ih;                                      // per layer: [fwd | bwd] input projection (6H x c_h), b_ih
  std::vector<float*> ih_w, ih_b;                                    // their fp32 sources: concatenated weights / biases
  std::vector<const float*> hh_w[2], hh_b[2];                        // per direction, per layer: W_hh (3H, H), b_hh (3H)
  dexb::TvConv out_conv, p_conv1, p_conv2, p_proj;
  float* gi = nullptr;                                               // [B*T][6H] input projections of the current layer
};

namespace dexb {

// in_conv of the LF0 encoder: one warp per frame; 1 input channel, 3 taps, ReLU, nn.LayerNorm(eps 1e-5), mask -> split rows
__global__ void __launch_bounds__(256) k_lf0_in(const float* __restrict__ lf0, const float* __restrict__ mask,
                                                const float* __restrict__ w, const float* __restrict__ g,
                                                const float* __restrict__ bta, bf16* __restrict__ os, long rows, int C, int T) {
  pdl_wait();
  const long r = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  const int t = (int)(r % T);
  const float x0 = t > 0 ? lf0[r - 1] * mask[r - 1] : 0.f;           // zero padding at both ends of the utterance
  const float x1 = lf0[r] * mask[r];
  const float x2 = t + 1 < T ? lf0[r + 1] * mask[r + 1] : 0.f;
  float v[kTvMaxC / 32];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const int c = lane + 32 * j;
    float y = 0.f;
    if (c < C) y = fmaxf(fmaf(w[c * 3 + 2], x2, fmaf(w[c * 3 + 1], x1, w[c * 3] * x0)), 0.f);
    v[j] = y;
    s += y;
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const float d = (lane + 32 * j < C) ? v[j] - mean : 0.f;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.f / sqrtf(warp_sum(q) / (float)C + 1e-5f);
  const float m = mask[r];
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const int c = lane + 32 * j;
    if (c >= C) continue;
    const float y = ((v[j] - mean) * rstd * g[c] + bta[c]) * m;
    bf16 hi, lo;
    split2(y, hi, lo);
    os[r * 2 * C + c] = hi;
    os[r * 2 * C + C + c] = lo;
  }
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

// GRU recurrence of one layer, both directions: grid (B, 2), 3H threads.  Thread j keeps row j of W_hh (gate order r, z, n) in
// registers; h lives in shared memory.  gi = x W_ih^T + b_ih comes from the GEMM ([B*T][6H]: forward gates, then backward gates).
// Output h_t (times omask[t] when given -- the mask the reference applies before out_conv) -> split rows, columns dir*H .. +H.
template <int H>
__global__ void __launch_bounds__(3 * H) k_gru_rec(const float* __restrict__ gi, const float* __restrict__ whh_f,
                                                   const float* __restrict__ bhh_f, const float* __restrict__ whh_b,
                                                   const float* __restrict__ bhh_b, const float* __restrict__ omask,
                                                   bf16* __restrict__ os, int T) {
  pdl_wait();
  __shared__ __align__(16) float h[H];
  __shared__ float gh[3 * H];
  const int b = blockIdx.x, dir = blockIdx.y, j = threadIdx.x;
  const float* whh = dir == 0 ? whh_f : whh_b;
  float wr[H];
#pragma unroll
  for (int k = 0; k < H; ++k) wr[k] = whh[j * H + k];
  const float bj = (dir == 0 ? bhh_f : bhh_b)[j];
  if (j < H) h[j] = 0.f;
  __syncthreads();
  for (int s = 0; s < T; ++s) {
    const int t = dir == 0 ? s : T - 1 - s;
    const long row = (long)b * T + t;
    float gr = 0.f, gz = 0.f, gn = 0.f;
    if (j < H) {                                  // issued before the mat-vec so their latency hides behind it
      const float* g = gi + row * (6 * H) + dir * 3 * H;
      gr = g[j]; gz = g[H + j]; gn = g[2 * H + j];
    }
    float a0 = bj, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
    for (int k = 0; k < H; k += 4) {
      const float4 hv = *reinterpret_cast<const float4*>(&h[k]);
      a0 = fmaf(wr[k], hv.x, a0);
      a1 = fmaf(wr[k + 1], hv.y, a1);
      a2 = fmaf(wr[k + 2], hv.z, a2);
      a3 = fmaf(wr[k + 3], hv.w, a3);
    }
    gh[j] = (a0 + a1) + (a2 + a3);
    __syncthreads();
    if (j < H) {
      const float r = sigmoid_f(gr + gh[j]);
      const float z = sigmoid_f(gz + gh[H + j]);
      const float n = tanhf(gn + r * gh[2 * H + j]);
      const float hn = (1.f - z) * n + z * h[j];
      h[j] = hn;
      const float o = omask != nullptr ? hn * omask[row] : hn;
      bf16 hi, lo;
      split2(o, hi, lo);
      os[row * (4 * H) + dir * H + j] = hi;       // row = [hi(2H) | lo(2H)]
      os[row * (4 * H) + 2 * H + dir * H + j] = lo;
    }
    __syncthreads();
  }
}

// rows [2][R][C] -> [2R][C] concatenation helper for the per-layer [forward | backward] input projection
__global__ void k_cat2(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, long na, long nb) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < na) out[i] = a[i];
  else if (i < na + nb) out[i] = b[i - na];
}

// time mean used by the fusion: out[b][c] (+)= sum_t x[b][c][t] / sum_t mask[b][t]   (x is already masked; tts.py:45,48)
__global__ void __launch_bounds__(256) k_time_mean(const float* __restrict__ x, const float* __restrict__ mask,
                                                   float* __restrict__ out, int C, int T, int accumulate) {
  pdl_wait();
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;      // one warp per (b, c)
  const int b = blockIdx.y;
  if (w >= C) return;
  float s = 0.f, m = 0.f;
  for (int t = lane; t < T; t += 32) {
    s += x[((long)b * C + w) * T + t];
    m += mask[(long)b * T + t];
  }
  s = warp_sum(s);
  m = warp_sum(m);
  if (lane == 0) out[(long)b * C + w] = (accumulate ? out[(long)b * C + w] : 0.f) + s / m;
}

// sty[b][n][t] = bias[n] + sum_k W[n][k] * (z_dec[b][k][t] + v[b][k]),  v = time mean of lf0_dec   (tts.py:48-49; conv_sty is 1x1)
// One CTA = (utterance, 32 frames): the [C][32] input tile (+ v) in shared memory, thread = one frame x N/8 output channels.
__global__ void __launch_bounds__(256) k_conv_sty(const float* __restrict__ z, const float* __restrict__ v,
                                                  const float* __restrict__ w, const float* __restrict__ bias,
                                                  float* __restrict__ out, int C, int N, int T) {
  pdl_wait();
  extern __shared__ float zt[];                  // [C][33]
  const int b = blockIdx.y, t0 = blockIdx.x * 32, tt = threadIdx.x & 31, ng = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < C * 32; i += 256) {
    const int k = i >> 5, x = i & 31;
    zt[k * 33 + x] = (t0 + x < T) ? z[((long)b * C + k) * T + t0 + x] + v[(long)b * C + k] : 0.f;
  }
  __syncthreads();
  if (t0 + tt >= T) return;
  for (int n = ng; n < N; n += 8) {
    const float* wn = w + (long)n * C;
    float a0 = bias[n], a1 = 0.f;
    for (int k = 0; k < C; k += 2) {
      a0 = fmaf(__ldg(wn + k), zt[k * 33 + tt], a0);
      a1 = fmaf(__ldg(wn + k + 1), zt[(k + 1) * 33 + tt], a1);
    }
    out[((long)b * N + n) * T + t0 + tt] = a0 + a1;
  }
}

static void lf0_release_plan(dexb_lf0* h) {
  enc_release_rows(h);
  cudaFree(h->gi);
  h->gi = nullptr;
}

static int lf0_plan(dexb_lf0* h, int B, int T) {
  if (B == h->B && T == h->T) return 0;
  lf0_release_plan(h);
  DEXB_TRY(gemm_global_init());
  int Cmax = h->c_h;
  if (h->c_out > Cmax) Cmax = h->c_out;
  if (h->c_g > Cmax) Cmax = h->c_g;
  const long rows = (long)B * T;
  DEXB_CUDA_OK(cudaMalloc(&h->xs, rows * 2 * Cmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->hs, rows * 2 * Cmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->acc, rows * Cmax * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&h->gi, rows * 6 * h->H * sizeof(float)));
  h->B = B; h->T = T;
  // layer l reads xs (l even) / hs (l odd) and its recurrence writes the other one; out_conv reads what the last layer wrote
  for (int l = 0; l < h->L; ++l) DEXB_TRY(tv_plan_conv(h, &h->ih[l], (l & 1) ? h->hs : h->xs, h->gi));
  bf16* gru_out = (h->L & 1) ? h->hs : h->xs;
  bf16* other = (h->L & 1) ? h->xs : h->hs;
  DEXB_TRY(tv_plan_conv(h, &h->out_conv, gru_out, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->p_conv1, other, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->p_conv2, gru_out, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->p_proj, other, h->acc));
  return 0;
}

}  // namespace dexb

extern "C" {

int dexb_lf0_create(int c_h, int c_out, int c_out_g, int num_layer, dexb_lf0** out) {
  DEXB_CHECK(out != nullptr, "dexb_lf0_create: null argument");
  int dev = 0, major = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DEXB_CHECK(major == 10, "dexb200 is built for sm_100a only (device %d has compute capability major %d); there is no fallback",
             dev, major);
  DEXB_CHECK(c_h == 192 || c_h == 256, "dexb_lf0_create: the GRU recurrence kernel is instantiated for c_h = 192 / 256 (hidden 96 / 128 per "
             "direction: the VCTK and LibriTTS configs), got %d", c_h);
  DEXB_CHECK(c_out >= 64 && c_out % 64 == 0 && c_out <= kTvMaxC && c_out_g >= 64 && c_out_g % 64 == 0 && c_out_g <= kTvMaxC,
             "dexb_lf0_create: c_out = %d / c_out_g = %d must be multiples of 64 (<= %d)", c_out, c_out_g, kTvMaxC);
  DEXB_CHECK(num_layer >= 1 && num_layer <= 8, "dexb_lf0_create: num_layer = %d", num_layer);
  dexb_lf0* h = new dexb_lf0();
  h->c_h = c_h; h->c_out = c_out; h->c_g = c_out_g; h->L = num_layer; h->H = c_h / 2;
  h->ih.resize(num_layer);
  h->ih_w.assign(num_layer, nullptr);
  h->ih_b.assign(num_layer, nullptr);
  for (int d = 0; d < 2; ++d) { h->hh_w[d].assign(num_layer, nullptr); h->hh_b[d].assign(num_layer, nullptr); }
  *out = h;
  return 0;
}

void dexb_lf0_destroy(dexb_lf0* h) {
  if (h == nullptr) return;
  lf0_release_plan(h);
  TvConv* cs[4] = {&h->out_conv, &h->p_conv1, &h->p_conv2, &h->p_proj};
  for (TvConv* c : cs) tv_free_conv(c);
  for (auto& c : h->ih) tv_free_conv(&c);
  for (float* p : h->ih_w) cudaFree(p);
  for (float* p : h->ih_b) cudaFree(p);
  for (auto& kv : h->w) cudaFree(kv.second.p);
  delete h;
}

int dexb_lf0_load_weight(dexb_lf0* h, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  DEXB_CHECK(h != nullptr && name != nullptr && data_dev != nullptr && shape != nullptr && ndim >= 1 && ndim <= 4,
             "dexb_lf0_load_weight: bad argument");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DEXB_CHECK(shape[i] >= 1, "dexb_lf0_load_weight(%s): empty dimension", name);
    n *= (size_t)shape[i];
  }
  TvTensor& t = h->w[name];
  if (t.p != nullptr && t.n != n) { cudaFree(t.p); t.p = nullptr; }
  if (t.p == nullptr) DEXB_CUDA_OK(cudaMalloc(&t.p, n * sizeof(float)));
  t.n = n;
  t.shape.assign(shape, shape + ndim);
  DEXB_CUDA_OK(cudaMemcpy(t.p, data_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return 0;
}

int dexb_lf0_finalize_weights(dexb_lf0* h, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  const int H = h->H, C = h->c_h;
  DEXB_TRY(tv_get(h, "in_conv.conv.weight", {C, 1, 3}, &h->in_w));
  DEXB_TRY(tv_get(h, "in_conv.ln.weight", {C}, &h->in_g));
  DEXB_TRY(tv_get(h, "in_conv.ln.bias", {C}, &h->in_b));
  for (int l = 0; l < h->L; ++l) {
    const std::string p = "rnn_layer.", s = "_l" + std::to_string(l);
    const float *wf, *wb, *bf, *bb;
    DEXB_TRY(tv_get(h, p + "weight_ih" + s, {3 * H, C}, &wf));
    DEXB_TRY(tv_get(h, p + "weight_ih" + s + "_reverse", {3 * H, C}, &wb));
    DEXB_TRY(tv_get(h, p + "bias_ih" + s, {3 * H}, &bf));
    DEXB_TRY(tv_get(h, p + "bias_ih" + s + "_reverse", {3 * H}, &bb));
    DEXB_TRY(tv_get(h, p + "weight_hh" + s, {3 * H, H}, &h->hh_w[0][l]));
    DEXB_TRY(tv_get(h, p + "weight_hh" + s + "_reverse", {3 * H, H}, &h->hh_w[1][l]));
    DEXB_TRY(tv_get(h, p + "bias_hh" + s, {3 * H}, &h->hh_b[0][l]));
    DEXB_TRY(tv_get(h, p + "bias_hh" + s + "_reverse", {3 * H}, &h->hh_b[1][l]));
    if (h->ih_w[l] == nullptr) {
      DEXB_CUDA_OK(cudaMalloc(&h->ih_w[l], (size_t)6 * H * C * sizeof(float)));
      DEXB_CUDA_OK(cudaMalloc(&h->ih_b[l], (size_t)6 * H * sizeof(float)));
    }
    launch_pdl(k_cat2, dim3((unsigned)(cdiv(6L * H * C, 256))), dim3(256), 0, st, wf, wb, h->ih_w[l], 3L * H * C, 3L * H * C);
    launch_pdl(k_cat2, dim3((unsigned)(cdiv(6L * H, 256))), dim3(256), 0, st, bf, bb, h->ih_b[l], 3L * H, 3L * H);
    TvConv& c = h->ih[l];
    c.ci = C; c.co = 6 * H; c.K = tv_pad64(C); c.taps = 1;
    if (c.w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c.w, (size_t)c.co * 2 * c.K * sizeof(bf16)));
    k_tv_pack_w<<<cdiv((long)c.co * c.K, 256), 256, 0, st>>>(h->ih_w[l], c.w, c.co, C, c.K, 1);
    c.bias = h->ih_b[l];
    DEXB_CUDA_OK(cudaGetLastError());
  }
  DEXB_TRY(tv_basic_conv(h, "out_conv", C, h->c_out, 1, &h->out_conv, st));
  DEXB_TRY(tv_pack_conv(h, "proj.conv_1.weight", "proj.conv_1.bias", h->c_out, h->c_g, 3, &h->p_conv1, st));
  DEXB_TRY(tv_get(h, "proj.norm_1.gamma", {h->c_g}, &h->p_conv1.ln_g));
  DEXB_TRY(tv_get(h, "proj.norm_1.beta", {h->c_g}, &h->p_conv1.ln_b));
  DEXB_TRY(tv_pack_conv(h, "proj.conv_2.weight", "proj.conv_2.bias", h->c_g, h->c_g, 3, &h->p_conv2, st));
  DEXB_TRY(tv_get(h, "proj.norm_2.gamma", {h->c_g}, &h->p_conv2.ln_g));
  DEXB_TRY(tv_get(h, "proj.norm_2.beta", {h->c_g}, &h->p_conv2.ln_b));
  DEXB_TRY(tv_pack_conv(h, "proj.proj.weight", "proj.proj.bias", h->c_g, h->c_g, 1, &h->p_proj, st));
  DEXB_CUDA_OK(cudaStreamSynchronize(st));
  lf0_release_plan(h);
  h->finalized = true;
  return 0;
}

static int lf0_enqueue(dexb_lf0* h, const float* lf0_dev, const float* mask_dev, int B, int T, float* lf0_enc_dev, float* lf0_dec_dev,
                       cudaStream_t st) {
  const long rows = (long)B * T;
  h->launches = 0;
  launch_pdl(k_lf0_in, dim3((unsigned)(cdiv(rows, 8))), dim3(256), 0, st, lf0_dev, mask_dev, h->in_w, h->in_g, h->in_b, h->xs, rows, h->c_h, T);
  h->launches += 1;
  for (int l = 0; l < h->L; ++l) {
    bf16* dst = (l & 1) ? h->xs : h->hs;
    DEXB_TRY(gemm_launch(h->ih[l].plan, h->ih[l].plan.p, 0, st));
    if (h->c_h == 192) launch_pdl(k_gru_rec<96>, dim3(B, 2), dim3(3 * 96), 0, st, h->gi, h->hh_w[0][l], h->hh_b[0][l], h->hh_w[1][l], h->hh_b[1][l],
                                                 l == h->L - 1 ? mask_dev : nullptr, dst, T);
    else launch_pdl(k_gru_rec<128>, dim3(B, 2), dim3(3 * 128), 0, st, h->gi, h->hh_w[0][l], h->hh_b[0][l], h->hh_w[1][l], h->hh_b[1][l],
                                                 l == h->L - 1 ? mask_dev : nullptr, dst, T);
    h->launches += 2;
  }
  bf16* gru_out = (h->L & 1) ? h->hs : h->xs;
  bf16* other = (h->L & 1) ? h->xs : h->hs;
  // lf0_enc = out_conv(x * mask) * mask
  DEXB_TRY(gemm_launch(h->out_conv.plan, h->out_conv.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->out_conv, 1, 1, 1e-5f);
    p.mask = mask_dev; p.os = other; p.ocm = lf0_enc_dev;
    tv_launch_post(p, st);
  }
  // lf0_dec = proj(lf0_enc, mask)
  DEXB_TRY(gemm_launch(h->p_conv1.plan, h->p_conv1.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->p_conv1, 1, 2, 1e-4f);
    p.mask = mask_dev; p.os = gru_out;
    tv_launch_post(p, st);
  }
  DEXB_TRY(gemm_launch(h->p_conv2.plan, h->p_conv2.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->p_conv2, 1, 2, 1e-4f);
    p.mask = mask_dev; p.os = other;
    tv_launch_post(p, st);
  }
  DEXB_TRY(gemm_launch(h->p_proj.plan, h->p_proj.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->p_proj, 0, 0, 0.f);
    p.mask = mask_dev; p.ocm = lf0_dec_dev;
    tv_launch_post(p, st);
  }
  h->launches += 8;
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

int dexb_lf0_forward(dexb_lf0* h, const float* lf0_dev, const float* mask_dev, int B, int T, float* lf0_enc_dev, float* lf0_dec_dev,
                     void* stream) {
  DEXB_CHECK(h != nullptr && lf0_dev != nullptr && mask_dev != nullptr && lf0_enc_dev != nullptr && lf0_dec_dev != nullptr,
             "dexb_lf0_forward: null argument");
  DEXB_CHECK(h->finalized, "dexb_lf0_forward: call dexb_lf0_finalize_weights first");
  DEXB_CHECK(B >= 1 && T >= 1, "dexb_lf0_forward: B = %d, T = %d", B, T);
  DEXB_CHECK(h->c_out == h->c_h, "dexb_lf0_forward: out_conv must keep the channel count of the GRU (c_out == c_h)");
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(lf0_plan(h, B, T));
  if (!enc_graphs_on()) return lf0_enqueue(h, lf0_dev, mask_dev, B, T, lf0_enc_dev, lf0_dec_dev, st);
  const size_t rows = (size_t)B * T;
  for (int pass = 0; pass < 2; ++pass) {
    EncStage a{pass == 0 ? nullptr : h->g.stage};
    float* g_lf0 = a.get<float>(rows);
    float* g_mask = a.get<float>(rows);
    float* g_enc = a.get<float>(rows * h->c_out);
    float* g_dec = a.get<float>(rows * h->c_g);
    if (pass == 0) {
      if (h->g.stage == nullptr) DEXB_CUDA_OK(cudaMalloc(&h->g.stage, a.off + 256));
      continue;
    }
    DEXB_CUDA_OK(cudaMemcpyAsync(g_lf0, lf0_dev, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(g_mask, mask_dev, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_TRY(enc_graph_run(&h->g, &h->launches, st, [&](cudaStream_t cs) { return lf0_enqueue(h, g_lf0, g_mask, B, T, g_enc, g_dec, cs); }));
    DEXB_CUDA_OK(cudaMemcpyAsync(lf0_enc_dev, g_enc, rows * h->c_out * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(lf0_dec_dev, g_dec, rows * h->c_g * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

long dexb_lf0_last_launch_count(const dexb_lf0* h) { return h != nullptr ? h->launches : 0; }

int dexb_style_fuse(const float* z_before_dev, const float* z_dec_dev, const float* sty_mask_dev, int Ts, const float* lf0_enc_dev,
                    const float* lf0_dec_dev, const float* lf0_mask_dev, int Tl, int B, int C, const float* conv_sty_w_dev,
                    const float* conv_sty_b_dev, int N, float* lf0_mean_scratch_dev, float* sty_enc_dev, float* sty_dev, void* stream) {
  DEXB_CHECK(z_dec_dev != nullptr && lf0_dec_dev != nullptr && lf0_mask_dev != nullptr && conv_sty_w_dev != nullptr &&
                 conv_sty_b_dev != nullptr && lf0_mean_scratch_dev != nullptr && sty_dev != nullptr,
             "dexb_style_fuse: null argument");
  DEXB_CHECK(B >= 1 && Ts >= 1 && Tl >= 1 && C >= 2 && C % 2 == 0 && C <= 1024 && N >= 1, "dexb_style_fuse: bad shape");
  cudaStream_t st = (cudaStream_t)stream;
  if (sty_enc_dev != nullptr) {                 // text-encoder conditioning (tts.py:45-46): both masked time means added
    DEXB_CHECK(z_before_dev != nullptr && sty_mask_dev != nullptr && lf0_enc_dev != nullptr, "dexb_style_fuse: sty_enc needs its inputs");
    launch_pdl(k_time_mean, dim3(cdiv(C, 8), B), dim3(256), 0, st, z_before_dev, sty_mask_dev, sty_enc_dev, C, Ts, 0);
    launch_pdl(k_time_mean, dim3(cdiv(C, 8), B), dim3(256), 0, st, lf0_enc_dev, lf0_mask_dev, sty_enc_dev, C, Tl, 1);
  }
  launch_pdl(k_time_mean, dim3(cdiv(C, 8), B), dim3(256), 0, st, lf0_dec_dev, lf0_mask_dev, lf0_mean_scratch_dev, C, Tl, 0);
  launch_pdl(k_conv_sty, dim3(cdiv(Ts, 32), B), dim3(256), (size_t)C * 33 * sizeof(float), st, z_dec_dev, lf0_mean_scratch_dev, conv_sty_w_dev,
                                                                                 conv_sty_b_dev, sty_dev, C, N, Ts);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"

// =====================================================================================================================
// Part 3: text encoder -- TextEncoder.forward, DEX-TTS/model/text_encoder.py:129-142
// (GeDEX-TTS/model/text_encoder.py:132-146 is the same code without the style input), eval mode, n_spks <= 1:
//
//   x  = emb[ids] * sqrt(C)                                                              (:130)
//   x  = (x + proj(3 x relu(cln(conv5(x * mask))))) * mask                               prenet, ConvReluNorm (:56-64)
//   8x RetNetDecoderLayer (retention.py:458-514) with use_softmax, no decay -- softmax attention in RetNet clothing:
//        h = h + out_proj(swish(g) * rms_head(softmax(rope(q) rope(k d^-0.5)^T | pair mask, fill -1e4) v)),  q/k/v/g = W rms(h) w
//        h = adaln_1(h, sty)                                                             (DEX-TTS only; base.py:180-194)
//        h = h + fc2(gelu(fc1 rms(h) w) * gate rms(h) w);  h = adaln_2(h, sty)           GLU (retention.py:371-381)
//   x  = rms(h) w * mask                                                                 (retnet.py:162, text_encoder.py:137)
//   mu = proj_m(x) * mask;  logw = proj(cln(relu(conv3(cln(relu(conv3(x*mask))) * mask))) * mask) * mask     (:138-141, :84-95)
//
// Rows [B*Tx][C] like the other encoders; every Linear / Conv1d is a 1 x k-tap implicit GEMM on the tcgen05 engine (split-bf16 x3,
// the plans of part 1), with one fp32 row kernel between two GEMMs: k_txt_row (residual | channel LayerNorm + ReLU | AdaLN |
// RMSNorm | mask | split-operand store, one warp per token).  Attention over <= a few hundred tokens with head dim 96 is 25 MFLOP
// per utterance and layer: it stays in fp32 on the CUDA cores (k_txt_attn: one warp per (utterance, head, query), online softmax,
// the per-head RMS norm and the swish gate fused behind it) -- the stage is bound by its ~120 launch latencies, not by a roofline.
// =====================================================================================================================

struct TxtLayer {
  dexb::TvConv qkvg, o, ffg, fc2;                   // q | k | v | g and fc1 | gate share their operand: one GEMM each (N = 4 C, 2 Fc)
  const float *rln = nullptr, *fln = nullptr;        // retention_layer_norm.weight, final_layer_norm.weight
};

struct dexb_text : dexb::EncBase {
  int n_vocab = 0, n_feats = 0, C = 0, Fc = 0, Fd = 0, heads = 0, L = 0, ksz = 3, adaln = 1;
  int C0 = 0, spk_dim = 0;                           // C0 = embedding / prenet width; C = C0 + spk_dim behind the prenet (n_spks > 1)
  const float *emb = nullptr, *angle = nullptr, *out_ln = nullptr, *dpw = nullptr, *dpb = nullptr;
  dexb::TvConv pre[3], pre_proj, proj_m, dp1, dp2;
  std::vector<TxtLayer> layers;
  float *adaW = nullptr, *adaB = nullptr;            // packed [L][4][C][C] / [L][4][C]: adaln_1.W_scale, .W_bias, adaln_2.W_scale, .W_bias
  // (B, Tx) plan: residual stream, prenet input, q / k / v / g rows, the two GLU branches, AdaLN scale / bias [L][4][B][C]
  float *hf = nullptr, *x0f = nullptr, *qf = nullptr, *kf = nullptr, *vf = nullptr, *gf = nullptr, *f1 = nullptr, *f2 = nullptr,
        *ada = nullptr;
  int layer_limit = -1;                              // unit-parity aid: >= 0 stops the forward after that many RetNet layers
};

namespace dexb {

constexpr int kTxtMaxDpl = 4;                       // head dim / 32 the attention kernel keeps per lane (96 -> 3)

struct TxtRow {
  const float* in;           // [rows][C]
  const float* resid;        // [rows][C] added first (RetNetDecoderLayer.residual_connection with alpha = 1), or null
  const float *ln_g, *ln_b;  // channel LayerNorm of model.text_encoder.LayerNorm (eps 1e-4, rsqrt), or null
  int relu;                  // ReLU after that LayerNorm (ConvReluNorm: conv -> norm -> relu)
  const float *ada_scale, *ada_bias;   // [B][C]: AdaptiveLayerNorm (x - mean) / sqrt(var + 1e-5) * scale + bias, or null
  float* out_f;              // fp32 rows of the value at this point (the residual stream), or null
  const float* rms_w;        // then RMSNorm (eps 1e-6) * weight, or null
  const float* mask;         // [rows] or null
  bf16* os;                  // split rows [rows][hi(C)|lo(C)] or null
  float* of2;                // fp32 rows of the final value or null
  long rows;
  int C, T;
};

// ids (B*Tx) -> x = emb[id] * sqrt(C) (text_encoder.py:130): fp32 rows (the prenet's residual input) and the split rows of x * mask.
// Ids outside [0, n_vocab) are clamped (upstream raises an IndexError on the host; a device kernel cannot).
__global__ void k_txt_embed(const long long* __restrict__ ids, const float* __restrict__ emb, const float* __restrict__ mask,
                            float* __restrict__ x0, bf16* __restrict__ xs, long rows, int C, int n_vocab, float scale) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rows * C) return;
  const long r = i / C;
  const int c = (int)(i % C);
  long long id = ids[r];
  id = id < 0 ? 0 : (id >= n_vocab ? n_vocab - 1 : id);
  const float v = emb[id * C + c] * scale;
  x0[i] = v;
  bf16 hi, lo;
  split2(v * mask[r], hi, lo);
  xs[r * 2 * C + c] = hi;
  xs[r * 2 * C + C + c] = lo;
}

// n_spks > 1 (text_encoder.py:135-136): rows [C0] of the prenet output + the utterance's speaker embedding (NOT masked: upstream
// repeats it over all Tx positions) -> the residual stream rows [C0 + S]
__global__ void k_txt_cat_spk(const float* __restrict__ x, const float* __restrict__ spk, float* __restrict__ out, long rows, int C0, int S,
                              int T) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int C = C0 + S;
  if (i >= rows * C) return;
  const long r = i / C;
  const int c = (int)(i % C);
  out[i] = c < C0 ? x[r * C0 + c] : spk[(r / T) * S + (c - C0)];
}

// one warp per token: everything between two GEMMs of the text encoder
__global__ void __launch_bounds__(256) k_txt_row(const TxtRow p) {
  pdl_wait();
  const long r = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (r >= p.rows) return;
  const int lane = threadIdx.x & 31;
  const int C = p.C;
  float v[kTvMaxC / 32];
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const int c = lane + 32 * j;
    float x = 0.f;
    if (c < C) {
      x = p.in[r * C + c];
      if (p.resid != nullptr) x += p.resid[r * C + c];
    }
    v[j] = x;                                                          // lanes beyond C hold zeros throughout
  }
  if (p.ln_g != nullptr || p.ada_scale != nullptr) {
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) s += v[j];
    const float mean = warp_sum(s) / (float)C;
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const float d = (lane + 32 * j < C) ? v[j] - mean : 0.f;
      q = fmaf(d, d, q);
    }
    const float var = warp_sum(q) / (float)C;                          // biased, both norms
    if (p.ln_g != nullptr) {
      const float rstd = rsqrtf(var + 1e-4f);
#pragma unroll
      for (int j = 0; j < kTvMaxC / 32; ++j) {
        const int c = lane + 32 * j;
        if (c < C) {
          const float y = (v[j] - mean) * rstd * p.ln_g[c] + p.ln_b[c];
          v[j] = p.relu ? fmaxf(y, 0.f) : y;
        }
      }
    } else {
      const float sd = sqrtf(var + 1e-5f);
      const long b = r / p.T;
#pragma unroll
      for (int j = 0; j < kTvMaxC / 32; ++j) {
        const int c = lane + 32 * j;
        if (c < C) v[j] = (v[j] - mean) / sd * p.ada_scale[b * C + c] + p.ada_bias[b * C + c];
      }
    }
  }
  if (p.out_f != nullptr) {
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const int c = lane + 32 * j;
      if (c < C) p.out_f[r * C + c] = v[j];
    }
  }
  if (p.rms_w != nullptr) {
    float q = 0.f;
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) q = fmaf(v[j], v[j], q);
    const float rr = rsqrtf(warp_sum(q) / (float)C + 1e-6f);
#pragma unroll
    for (int j = 0; j < kTvMaxC / 32; ++j) {
      const int c = lane + 32 * j;
      if (c < C) v[j] = v[j] * rr * p.rms_w[c];
    }
  }
  const float m = p.mask != nullptr ? p.mask[r] : 1.f;
#pragma unroll
  for (int j = 0; j < kTvMaxC / 32; ++j) {
    const int c = lane + 32 * j;
    if (c >= C) continue;
    const float x = v[j] * m;
    if (p.os != nullptr) {
      bf16 hi, lo;
      split2(x, hi, lo);
      p.os[r * 2 * C + c] = hi;
      p.os[r * 2 * C + C + c] = lo;
    }
    if (p.of2 != nullptr) p.of2[r * C + c] = x;
  }
}

// AdaLN scale / bias of every layer for this batch: out[m][b][c] = W[m][c][:] . sty[b][:] + bias[m][c]  (base.py:189-190); one warp each
__global__ void __launch_bounds__(256) k_txt_ada(const float* __restrict__ W, const float* __restrict__ bias, const float* __restrict__ sty,
                                                 float* __restrict__ out, int M, int B, int C) {
  pdl_wait();
  const long wid = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (wid >= (long)M * B * C) return;
  const int lane = threadIdx.x & 31;
  const int c = (int)(wid % C), b = (int)((wid / C) % B);
  const long m = wid / ((long)C * B);
  const float* w = W + (m * C + c) * C;
  const float* s = sty + (long)b * C;
  float part = 0.f;
  for (int k = lane; k < C; k += 32) part = fmaf(w[k], s[k], part);
  part = warp_sum(part);
  if (lane == 0) out[wid] = part + bias[m * C + c];
}

// q, k rows [rows][C] in place: k *= d^-0.5 (retention.py:281), then theta_shift on both (retention.py:28-37) with
// sin / cos(t * angle[i]), angle repeated per pair (retention.py:75-76, 140-142).  One thread per (token, channel pair).
// q / k are column blocks of the merged projection output: row stride ld (= 4 C)
__global__ void k_txt_rope(float* __restrict__ q, float* __restrict__ k, const float* __restrict__ angle, long rows, int C, int d, int T,
                           float scaling, int ld) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int half = C / 2;
  if (i >= rows * half) return;
  const long r = i / half;
  const int c0 = 2 * (int)(i % half);
  const int t = (int)(r % T);
  const float ph = __fmul_rn((float)t, angle[c0 % d]);
  const float sn = sinf(ph), cs = cosf(ph);
  float* qp = q + r * ld + c0;
  float* kp = k + r * ld + c0;
  const float q0 = qp[0], q1 = qp[1];
  qp[0] = q0 * cs + (-q1) * sn;
  qp[1] = q1 * cs + q0 * sn;
  const float k0 = kp[0] * scaling, k1 = kp[1] * scaling;
  kp[0] = k0 * cs + (-k1) * sn;
  kp[1] = k1 * cs + k0 * sn;
}

// MultiScaleRetention.parallel_retention with use_softmax (retention.py:223-250) + group_norm + gate (:290-292).  One warp per
// (utterance, head, query); lane l holds dims l, l + 32, ... of the head.  Scores of (query, key) pairs with a padded member are
// -1e4 as upstream (a padded query therefore averages v over ALL keys).  Output: split rows of swish(g) * rms_head(softmax(s) v).
__global__ void __launch_bounds__(256) k_txt_attn(const float* __restrict__ q, const float* __restrict__ k, const float* __restrict__ v,
                                                  const float* __restrict__ g, const float* __restrict__ mask, bf16* __restrict__ os,
                                                  int B, int T, int C, int heads, int ld) {
  pdl_wait();
  const long wid = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (wid >= (long)B * heads * T) return;
  const int lane = threadIdx.x & 31;
  const int t = (int)(wid % T), hd = (int)((wid / T) % heads), b = (int)(wid / ((long)T * heads));
  const int d = C / heads, dpl = d / 32;
  const long r = (long)b * T + t;
  const float qm = mask[r];
  float qv[kTxtMaxDpl], acc[kTxtMaxDpl];
#pragma unroll
  for (int i = 0; i < kTxtMaxDpl; ++i) {
    qv[i] = i < dpl ? q[r * ld + hd * d + lane + 32 * i] : 0.f;
    acc[i] = 0.f;
  }
  float mx = -INFINITY, l = 0.f;
  const float* kb = k + (long)b * T * ld + hd * d + lane;
  const float* vb = v + (long)b * T * ld + hd * d + lane;
  const float* mb = mask + (long)b * T;
  for (int j = 0; j < T; ++j) {
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < kTxtMaxDpl; ++i)
      if (i < dpl) part = fmaf(qv[i], kb[(long)j * ld + 32 * i], part);
    float s = warp_sum(part);
    if (qm == 0.f || mb[j] == 0.f) s = -1e4f;
    const float mn = fmaxf(mx, s);
    const float corr = expf(mx - mn);                                  // first key: exp(-inf) = 0
    const float pj = expf(s - mn);
    l = l * corr + pj;
#pragma unroll
    for (int i = 0; i < kTxtMaxDpl; ++i)
      if (i < dpl) acc[i] = acc[i] * corr + pj * vb[(long)j * ld + 32 * i];
    mx = mn;
  }
  const float inv = 1.f / l;
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < kTxtMaxDpl; ++i) {
    acc[i] *= inv;                                                     // slots beyond dpl stay zero
    ss = fmaf(acc[i], acc[i], ss);
  }
  const float rr = rsqrtf(warp_sum(ss) / (float)d + 1e-6f);            // group_norm: RMSNorm(head_dim), no affine
#pragma unroll
  for (int i = 0; i < kTxtMaxDpl; ++i) {
    if (i >= dpl) continue;
    const int c = hd * d + lane + 32 * i;
    const float gv = g[r * ld + c];
    const float o = gv / (1.f + expf(-gv)) * (acc[i] * rr);            // swish gate
    bf16 hi, lo;
    split2(o, hi, lo);
    os[r * 2 * C + c] = hi;
    os[r * 2 * C + C + c] = lo;
  }
}

// GLU.forward (retention.py:371-381): gelu(fc1 x) * gate x (exact gelu) -> split rows [rows][hi(F)|lo(F)], the operand of fc2
__global__ void k_txt_glu(const float* __restrict__ a, const float* __restrict__ gate, bf16* __restrict__ os, long rows, int F, int ld) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rows * F) return;
  const long r = i / F;
  const int c = (int)(i % F);
  const float x = a[r * ld + c];                            // fc1 | gate are the two column halves of one GEMM output (ld = 2 F)
  const float y = x * 0.5f * (1.f + erff(x * 0.70710678118654752f)) * gate[r * ld + c];
  bf16 hi, lo;
  split2(y, hi, lo);
  os[r * 2 * F + c] = hi;
  os[r * 2 * F + F + c] = lo;
}

// DurationPredictor.proj (text_encoder.py:94-95): Conv1d(Fd, 1, 1) of the masked rows, * mask; one warp per token
__global__ void __launch_bounds__(256) k_txt_dp_out(const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
                                                    const float* __restrict__ mask, float* __restrict__ logw, long rows, int C) {
  pdl_wait();
  const long r = blockIdx.x * 8L + (threadIdx.x >> 5);
  if (r >= rows) return;
  const int lane = threadIdx.x & 31;
  float part = 0.f;
  for (int c = lane; c < C; c += 32) part = fmaf(x[r * C + c], w[c], part);
  part = warp_sum(part);
  if (lane == 0) logw[r] = (part + bias[0]) * mask[r];
}

// Linear (co, ci) or Conv1d (co, ci, taps) weight -> packed split operand (same layout as tv_pack_conv)
// Several Linear weights with the same input stacked along the output dimension: rows [r0, r0 + co_i) of one packed operand
static int txt_pack_stacked(EncBase* h, const std::vector<std::string>& wnames, int ci, int co_each, TvConv* c, cudaStream_t st) {
  const int n = (int)wnames.size();
  c->ci = ci; c->co = n * co_each; c->K = tv_pad64(ci); c->taps = 1; c->bias = nullptr;
  if (c->w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->w, (size_t)c->co * 2 * c->K * sizeof(bf16)));
  for (int i = 0; i < n; ++i) {
    const float* w = nullptr;
    DEXB_TRY(tv_get(h, wnames[i], {co_each, ci}, &w));
    k_tv_pack_w<<<cdiv((long)co_each * c->K, 256), 256, 0, st>>>(w, c->w + (size_t)i * co_each * 2 * c->K, co_each, ci, c->K, 1);
  }
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

static int txt_pack(EncBase* h, const std::string& wname, const std::string& bname, int ci, int co, int taps, bool linear, TvConv* c,
                    cudaStream_t st) {
  c->ci = ci; c->co = co; c->K = tv_pad64(ci); c->taps = taps;
  const float* w = nullptr;
  if (linear) DEXB_TRY(tv_get(h, wname, {co, ci}, &w));
  else DEXB_TRY(tv_get(h, wname, {co, ci, taps}, &w));
  if (c->w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->w, (size_t)taps * co * 2 * c->K * sizeof(bf16)));
  k_tv_pack_w<<<cdiv((long)taps * co * c->K, 256), 256, 0, st>>>(w, c->w, co, ci, c->K, taps);
  c->bias = nullptr;
  if (!bname.empty()) DEXB_TRY(tv_get(h, bname, {co}, &c->bias));
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

static void txt_release_plan(dexb_text* h) {
  enc_release_rows(h);
  float** bufs[] = {&h->hf, &h->x0f, &h->qf, &h->f1, &h->ada};
  for (float** b : bufs) { cudaFree(*b); *b = nullptr; }
  h->kf = h->vf = h->gf = h->f2 = nullptr;             // column views of qf / f1
}

static int txt_plan_build(dexb_text* h, int B, int T) {
  DEXB_TRY(gemm_global_init());
  const int C = h->C;
  int Kmax = C, Cmax = C;
  if (h->Fc > Kmax) Kmax = h->Fc;
  if (h->Fd > Kmax) Kmax = h->Fd;
  if (h->Fd > Cmax) Cmax = h->Fd;
  if (h->n_feats > Cmax) Cmax = h->n_feats;
  const long rows = (long)B * T;
  DEXB_CUDA_OK(cudaMalloc(&h->xs, rows * 2 * Kmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->hs, rows * 2 * Kmax * sizeof(bf16)));
  DEXB_CUDA_OK(cudaMalloc(&h->acc, rows * Cmax * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&h->xf, rows * Cmax * sizeof(float)));
  float** cbufs[] = {&h->hf, &h->x0f};
  for (float** b : cbufs) DEXB_CUDA_OK(cudaMalloc(b, rows * C * sizeof(float)));
  DEXB_CUDA_OK(cudaMalloc(&h->qf, rows * 4 * C * sizeof(float)));      // [rows][q | k | v | g]
  h->kf = h->qf + C; h->vf = h->qf + 2 * C; h->gf = h->qf + 3 * C;
  DEXB_CUDA_OK(cudaMalloc(&h->f1, rows * 2 * h->Fc * sizeof(float))); // [rows][fc1 | gate]
  h->f2 = h->f1 + h->Fc;
  DEXB_CUDA_OK(cudaMalloc(&h->ada, (size_t)h->L * 4 * B * C * sizeof(float)));
  h->B = B; h->T = T;
  for (int i = 0; i < 3; ++i) DEXB_TRY(tv_plan_conv(h, &h->pre[i], h->xs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->pre_proj, h->xs, h->acc));
  for (auto& ly : h->layers) {
    DEXB_TRY(tv_plan_conv(h, &ly.qkvg, h->xs, h->qf));
    DEXB_TRY(tv_plan_conv(h, &ly.o, h->hs, h->acc));
    DEXB_TRY(tv_plan_conv(h, &ly.ffg, h->xs, h->f1));
    DEXB_TRY(tv_plan_conv(h, &ly.fc2, h->hs, h->acc));
  }
  DEXB_TRY(tv_plan_conv(h, &h->proj_m, h->xs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->dp1, h->xs, h->acc));
  DEXB_TRY(tv_plan_conv(h, &h->dp2, h->hs, h->acc));
  return 0;
}
static int txt_plan(dexb_text* h, int B, int T) {
  if (B == h->B && T == h->T) return 0;
  txt_release_plan(h);
  const int rc = txt_plan_build(h, B, T);
  if (rc != 0) txt_release_plan(h);          // never keep a half-built plan: the next call starts from scratch
  return rc;
}

static TxtRow txt_row(const dexb_text* h, const float* in, int C = 0) {
  TxtRow p;
  memset(&p, 0, sizeof(p));
  p.in = in;
  p.rows = (long)h->B * h->T;
  p.C = C > 0 ? C : h->C; p.T = h->T;
  return p;
}
static void txt_launch_row(const TxtRow& p, cudaStream_t st) { launch_pdl(k_txt_row, dim3((unsigned)(cdiv(p.rows, 8))), dim3(256), 0, st, p); }

}  // namespace dexb

extern "C" {

int dexb_text_create(int n_vocab, int n_feats, int n_channels, int filter_channels, int filter_channels_dp, int n_heads, int n_layers,
                     int kernel_size, int adaln, int spk_emb_dim, dexb_text** out) {
  DEXB_CHECK(out != nullptr, "dexb_text_create: null argument");
  int dev = 0, major = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DEXB_CHECK(major == 10, "dexb200 is built for sm_100a only (device %d has compute capability major %d); there is no fallback",
             dev, major);
  DEXB_CHECK(n_vocab >= 1 && n_feats >= 1 && n_feats <= kTvMaxC && n_layers >= 1 && n_layers <= 64 && (kernel_size == 1 || kernel_size == 3 || kernel_size == 5),
             "dexb_text_create: n_vocab %d / n_feats %d / n_layers %d / kernel_size %d out of range", n_vocab, n_feats, n_layers, kernel_size);
  DEXB_CHECK(n_channels >= 64 && n_channels % 64 == 0 && n_channels <= kTvMaxC && filter_channels_dp >= 64 && filter_channels_dp % 64 == 0 &&
                 filter_channels_dp <= kTvMaxC && filter_channels >= 64 && filter_channels % 64 == 0,
             "dexb_text_create: n_channels %d / filter_channels_dp %d must be multiples of 64 (<= %d), filter_channels %d a multiple of 64",
             n_channels, filter_channels_dp, kTvMaxC, filter_channels);
  DEXB_CHECK(spk_emb_dim >= 0 && spk_emb_dim % 64 == 0 && n_channels + spk_emb_dim <= kTvMaxC,
             "dexb_text_create: spk_emb_dim %d must be a multiple of 64 with n_channels + spk_emb_dim <= %d", spk_emb_dim, kTvMaxC);
  DEXB_CHECK(spk_emb_dim == 0 || !adaln, "dexb_text_create: the speaker channel exists for GeDEX-TTS only (DeXTTS.forward passes spk=None "
             "to its encoder, DEX-TTS/model/tts.py:52)");
  const int Cw = n_channels + spk_emb_dim;
  DEXB_CHECK(n_heads >= 1 && Cw % n_heads == 0 && (Cw / n_heads) % 32 == 0 && Cw / n_heads <= 32 * kTxtMaxDpl,
             "dexb_text_create: head dim %d / %d must be a multiple of 32 (<= %d)", Cw, n_heads, 32 * kTxtMaxDpl);
  dexb_text* h = new dexb_text();
  h->n_vocab = n_vocab; h->n_feats = n_feats; h->C = Cw; h->C0 = n_channels; h->spk_dim = spk_emb_dim;
  h->Fc = filter_channels; h->Fd = filter_channels_dp;
  h->heads = n_heads; h->L = n_layers; h->ksz = kernel_size; h->adaln = adaln ? 1 : 0;
  h->layers.resize(n_layers);
  *out = h;
  return 0;
}

void dexb_text_destroy(dexb_text* h) {
  if (h == nullptr) return;
  txt_release_plan(h);
  TvConv* cs[7] = {&h->pre[0], &h->pre[1], &h->pre[2], &h->pre_proj, &h->proj_m, &h->dp1, &h->dp2};
  for (TvConv* c : cs) tv_free_conv(c);
  for (auto& ly : h->layers) {
    TvConv* ls[4] = {&ly.qkvg, &ly.o, &ly.ffg, &ly.fc2};
    for (TvConv* c : ls) tv_free_conv(c);
  }
  cudaFree(h->adaW); cudaFree(h->adaB);
  for (auto& kv : h->w) cudaFree(kv.second.p);
  delete h;
}

int dexb_text_load_weight(dexb_text* h, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  DEXB_CHECK(h != nullptr && name != nullptr && data_dev != nullptr && shape != nullptr && ndim >= 1 && ndim <= 4,
             "dexb_text_load_weight: bad argument");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DEXB_CHECK(shape[i] >= 1, "dexb_text_load_weight(%s): empty dimension", name);
    n *= (size_t)shape[i];
  }
  TvTensor& t = h->w[name];
  if (t.p != nullptr && t.n != n) { cudaFree(t.p); t.p = nullptr; }
  if (t.p == nullptr) DEXB_CUDA_OK(cudaMalloc(&t.p, n * sizeof(float)));
  t.n = n;
  t.shape.assign(shape, shape + ndim);
  DEXB_CUDA_OK(cudaMemcpy(t.p, data_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return 0;
}

int dexb_text_finalize_weights(dexb_text* h, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  const int C = h->C, C0 = h->C0;
  DEXB_TRY(tv_get(h, "emb.weight", {h->n_vocab, C0}, &h->emb));
  for (int i = 0; i < 3; ++i) {
    const std::string s = std::to_string(i);
    DEXB_TRY(txt_pack(h, "prenet.conv_layers." + s + ".weight", "prenet.conv_layers." + s + ".bias", C0, C0, 5, false, &h->pre[i], st));
    DEXB_TRY(tv_get(h, "prenet.norm_layers." + s + ".gamma", {C0}, &h->pre[i].ln_g));
    DEXB_TRY(tv_get(h, "prenet.norm_layers." + s + ".beta", {C0}, &h->pre[i].ln_b));
  }
  DEXB_TRY(txt_pack(h, "prenet.proj.weight", "prenet.proj.bias", C0, C0, 1, false, &h->pre_proj, st));
  if (h->adaln) {
    if (h->adaW == nullptr) {
      DEXB_CUDA_OK(cudaMalloc(&h->adaW, (size_t)h->L * 4 * C * C * sizeof(float)));
      DEXB_CUDA_OK(cudaMalloc(&h->adaB, (size_t)h->L * 4 * C * sizeof(float)));
    }
  }
  for (int l = 0; l < h->L; ++l) {
    const std::string p = "encoder.layers." + std::to_string(l) + ".";
    TxtLayer& ly = h->layers[l];
    DEXB_TRY(txt_pack_stacked(h, {p + "retention.q_proj.weight", p + "retention.k_proj.weight", p + "retention.v_proj.weight",
                                  p + "retention.g_proj.weight"}, C, C, &ly.qkvg, st));
    DEXB_TRY(txt_pack(h, p + "retention.out_proj.weight", "", C, C, 1, true, &ly.o, st));
    DEXB_TRY(txt_pack_stacked(h, {p + "ffn.fc1.weight", p + "ffn.gate.weight"}, C, h->Fc, &ly.ffg, st));
    DEXB_TRY(txt_pack(h, p + "ffn.fc2.weight", "", h->Fc, C, 1, true, &ly.fc2, st));
    DEXB_TRY(tv_get(h, p + "retention_layer_norm.weight", {C}, &ly.rln));
    DEXB_TRY(tv_get(h, p + "final_layer_norm.weight", {C}, &ly.fln));
    if (h->adaln) {
      const char* names[4] = {"adaln_1.W_scale", "adaln_1.W_bias", "adaln_2.W_scale", "adaln_2.W_bias"};
      for (int m = 0; m < 4; ++m) {
        const float *w = nullptr, *b = nullptr;
        DEXB_TRY(tv_get(h, p + names[m] + ".weight", {C, C}, &w));
        DEXB_TRY(tv_get(h, p + names[m] + ".bias", {C}, &b));
        DEXB_CUDA_OK(cudaMemcpyAsync(h->adaW + ((size_t)l * 4 + m) * C * C, w, (size_t)C * C * sizeof(float), cudaMemcpyDeviceToDevice, st));
        DEXB_CUDA_OK(cudaMemcpyAsync(h->adaB + ((size_t)l * 4 + m) * C, b, (size_t)C * sizeof(float), cudaMemcpyDeviceToDevice, st));
      }
    }
  }
  DEXB_TRY(tv_get(h, "encoder.layer_norm.weight", {C}, &h->out_ln));
  DEXB_TRY(tv_get(h, "encoder.retnet_rel_pos.angle", {C / h->heads}, &h->angle));
  DEXB_TRY(txt_pack(h, "proj_m.weight", "proj_m.bias", C, h->n_feats, 1, false, &h->proj_m, st));
  DEXB_TRY(txt_pack(h, "proj_w.conv_1.weight", "proj_w.conv_1.bias", C, h->Fd, h->ksz, false, &h->dp1, st));
  DEXB_TRY(tv_get(h, "proj_w.norm_1.gamma", {h->Fd}, &h->dp1.ln_g));
  DEXB_TRY(tv_get(h, "proj_w.norm_1.beta", {h->Fd}, &h->dp1.ln_b));
  DEXB_TRY(txt_pack(h, "proj_w.conv_2.weight", "proj_w.conv_2.bias", h->Fd, h->Fd, h->ksz, false, &h->dp2, st));
  DEXB_TRY(tv_get(h, "proj_w.norm_2.gamma", {h->Fd}, &h->dp2.ln_g));
  DEXB_TRY(tv_get(h, "proj_w.norm_2.beta", {h->Fd}, &h->dp2.ln_b));
  DEXB_TRY(tv_get(h, "proj_w.proj.weight", {1, h->Fd, 1}, &h->dpw));
  DEXB_TRY(tv_get(h, "proj_w.proj.bias", {1}, &h->dpb));
  DEXB_CUDA_OK(cudaStreamSynchronize(st));
  txt_release_plan(h);                      // plans hold the packed-weight pointers of the previous finalize
  h->finalized = true;
  return 0;
}

// every launch of one forward on `st`; all pointers are device pointers that stay valid until the work has run
static int text_enqueue(dexb_text* h, const int64_t* ids_dev, const float* mask_dev, const float* sty_dev, const float* spk_dev, int B, int Tx,
                        float* mu_dev, float* logw_dev, cudaStream_t st) {
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
  const int C = h->C, C0 = h->C0, T = Tx;
  const long rows = (long)B * T;
  h->launches = 0;
  if (h->adaln) {
    launch_pdl(k_txt_ada, dim3((unsigned)(cdiv((long)h->L * 4 * B * C, 8))), dim3(256), 0, st, h->adaW, h->adaB, sty_dev, h->ada, h->L * 4, B, C);
    h->launches += 1;
  }
  launch_pdl(k_txt_embed, dim3((unsigned)(cdiv(rows * C0, 256))), dim3(256), 0, st, reinterpret_cast<const long long*>(ids_dev), h->emb, mask_dev, h->x0f, h->xs, rows, C0,
                                                   h->n_vocab, (float)sqrt((double)C0));
  h->launches += 1;
  // prenet: 3 x (conv5 -> channel LayerNorm -> ReLU), input masked before every conv; then (x + proj(.)) * mask
  for (int i = 0; i < 3; ++i) {
    DEXB_TRY(gemm_launch(h->pre[i].plan, h->pre[i].plan.p, 0, st));
    TxtRow p = txt_row(h, h->acc, C0);
    p.ln_g = h->pre[i].ln_g; p.ln_b = h->pre[i].ln_b; p.relu = 1;
    p.mask = mask_dev; p.os = h->xs;
    txt_launch_row(p, st);
    h->launches += 2;
  }
  DEXB_TRY(gemm_launch(h->pre_proj.plan, h->pre_proj.plan.p, 0, st));
  {
    TxtRow p = txt_row(h, h->acc, C0);
    p.resid = h->x0f; p.mask = mask_dev;
    p.of2 = h->spk_dim > 0 ? h->qf : h->hf;                            // (x + proj(x)) * mask: the residual stream (its first C0 channels)
    txt_launch_row(p, st);
    if (h->spk_dim > 0) {
      launch_pdl(k_txt_cat_spk, dim3((unsigned)(cdiv(rows * C, 256))), dim3(256), 0, st, h->qf, spk_dev, h->hf, rows, C0, h->spk_dim, T);
      h->launches += 1;
    }
    TxtRow n = txt_row(h, h->hf);
    n.rms_w = h->layers[0].rln; n.os = h->xs;                          // operand of layer 0's q / k / v / g projections
    txt_launch_row(n, st);
    h->launches += 3;
  }
  const int d = C / h->heads;
  for (int l = 0; l < h->L; ++l) {
    if (h->layer_limit >= 0 && l >= h->layer_limit) break;
    TxtLayer& ly = h->layers[l];
    DEXB_TRY(gemm_launch(ly.qkvg.plan, ly.qkvg.plan.p, 0, st));          // one C -> 4 C GEMM for q | k | v | g
    launch_pdl(k_txt_rope, dim3((unsigned)(cdiv(rows * (C / 2), 256))), dim3(256), 0, st, h->qf, h->kf, h->angle, rows, C, d, T, 1.f / sqrtf((float)d), 4 * C);
    launch_pdl(k_txt_attn, dim3((unsigned)(cdiv((long)B * h->heads * T, 8))), dim3(256), 0, st, h->qf, h->kf, h->vf, h->gf, mask_dev, h->hs, B, T, C, h->heads, 4 * C);
    DEXB_TRY(gemm_launch(ly.o.plan, ly.o.plan.p, 0, st));
    {
      TxtRow p = txt_row(h, h->acc);                                   // h = adaln_1(h + out_proj(.)); operand = rms(h) * w
      p.resid = h->hf;
      if (h->adaln) {
        p.ada_scale = h->ada + ((size_t)l * 4 + 0) * B * C;
        p.ada_bias = h->ada + ((size_t)l * 4 + 1) * B * C;
      }
      p.out_f = h->hf; p.rms_w = ly.fln; p.os = h->xs;
      txt_launch_row(p, st);
    }
    DEXB_TRY(gemm_launch(ly.ffg.plan, ly.ffg.plan.p, 0, st));            // one C -> 2 Fc GEMM for fc1 | gate
    launch_pdl(k_txt_glu, dim3((unsigned)(cdiv(rows * h->Fc, 256))), dim3(256), 0, st, h->f1, h->f2, h->hs, rows, h->Fc, 2 * h->Fc);
    DEXB_TRY(gemm_launch(ly.fc2.plan, ly.fc2.plan.p, 0, st));
    {
      const bool last = l == h->L - 1;
      TxtRow p = txt_row(h, h->acc);                                   // h = adaln_2(h + ffn(.)); operand of the next layer / the heads
      p.resid = h->hf;
      if (h->adaln) {
        p.ada_scale = h->ada + ((size_t)l * 4 + 2) * B * C;
        p.ada_bias = h->ada + ((size_t)l * 4 + 3) * B * C;
      }
      p.out_f = h->hf;
      p.rms_w = last ? h->out_ln : h->layers[l + 1].rln;
      if (last) p.mask = mask_dev;                                     // x = layer_norm(h) * x_mask (text_encoder.py:137)
      p.os = h->xs;
      txt_launch_row(p, st);
    }
    h->launches += 9;              // q|k|v|g projection, rope, attention, out_proj, row, fc1|gate, glu, fc2, row
  }
  if (h->layer_limit >= 0) {                 // unit parity: the residual stream is read back with dexb_text_copy_stream
    DEXB_CUDA_OK(cudaGetLastError());
    return 0;
  }
  // mu = proj_m(x) * mask -> (B, n_feats, Tx)
  DEXB_TRY(gemm_launch(h->proj_m.plan, h->proj_m.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->proj_m, 0, 0, 0.f);
    p.mask = mask_dev; p.ocm = mu_dev;
    tv_launch_post(p, st);
  }
  // logw = DurationPredictor(x, mask): conv -> relu -> norm -> (mask) conv -> relu -> norm -> (mask) proj -> mask
  DEXB_TRY(gemm_launch(h->dp1.plan, h->dp1.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->dp1, 1, 2, 1e-4f);
    p.mask = mask_dev; p.os = h->hs;
    tv_launch_post(p, st);
  }
  DEXB_TRY(gemm_launch(h->dp2.plan, h->dp2.plan.p, 0, st));
  {
    TvPost p = tv_post(h, h->dp2, 1, 2, 1e-4f);
    p.mask = mask_dev; p.of = h->xf;
    tv_launch_post(p, st);
  }
  launch_pdl(k_txt_dp_out, dim3((unsigned)(cdiv(rows, 8))), dim3(256), 0, st, h->xf, h->dpw, h->dpb, mask_dev, logw_dev, rows, h->Fd);
  h->launches += 7;
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

int dexb_text_forward(dexb_text* h, const int64_t* ids_dev, const float* mask_dev, const float* sty_dev, const float* spk_dev, int B, int Tx,
                      float* mu_dev, float* logw_dev, void* stream) {
  DEXB_CHECK(h != nullptr && ids_dev != nullptr && mask_dev != nullptr && mu_dev != nullptr && logw_dev != nullptr,
             "dexb_text_forward: null argument");
  DEXB_CHECK(h->finalized, "dexb_text_forward: call dexb_text_finalize_weights first");
  DEXB_CHECK(B >= 1 && Tx >= 1, "dexb_text_forward: B = %d, Tx = %d", B, Tx);
  DEXB_CHECK((sty_dev != nullptr) == (h->adaln != 0), "dexb_text_forward: the style vector is %s for this encoder",
             h->adaln ? "required (DEX-TTS: AdaptiveLayerNorm)" : "not taken (GeDEX-TTS)");
  DEXB_CHECK((spk_dev != nullptr) == (h->spk_dim > 0), "dexb_text_forward: the speaker embedding is %s for this encoder",
             h->spk_dim > 0 ? "required (n_spks > 1)" : "not taken (n_spks <= 1)");
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(txt_plan(h, B, Tx));
  if (!enc_graphs_on() || h->layer_limit >= 0)
    return text_enqueue(h, ids_dev, mask_dev, sty_dev, spk_dev, B, Tx, mu_dev, logw_dev, st);
  // graph replay: inputs / outputs go through fixed staging buffers of the plan
  const size_t rows = (size_t)B * Tx;
  for (int pass = 0; pass < 2; ++pass) {
    EncStage a{pass == 0 ? nullptr : h->g.stage};
    int64_t* g_ids = a.get<int64_t>(rows);
    float* g_mask = a.get<float>(rows);
    float* g_sty = a.get<float>((size_t)B * h->C0);
    float* g_spk = a.get<float>((size_t)B * (h->spk_dim > 0 ? h->spk_dim : 1));
    float* g_mu = a.get<float>(rows * h->n_feats);
    float* g_logw = a.get<float>(rows);
    if (pass == 0) {
      if (h->g.stage == nullptr) DEXB_CUDA_OK(cudaMalloc(&h->g.stage, a.off + 256));
      continue;
    }
    DEXB_CUDA_OK(cudaMemcpyAsync(g_ids, ids_dev, rows * sizeof(int64_t), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(g_mask, mask_dev, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (sty_dev != nullptr) DEXB_CUDA_OK(cudaMemcpyAsync(g_sty, sty_dev, (size_t)B * h->C0 * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (spk_dev != nullptr) DEXB_CUDA_OK(cudaMemcpyAsync(g_spk, spk_dev, (size_t)B * h->spk_dim * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_TRY(enc_graph_run(&h->g, &h->launches, st, [&](cudaStream_t cs) {
      return text_enqueue(h, g_ids, g_mask, sty_dev != nullptr ? g_sty : nullptr, spk_dev != nullptr ? g_spk : nullptr, B, Tx, g_mu, g_logw, cs);
    }));
    DEXB_CUDA_OK(cudaMemcpyAsync(mu_dev, g_mu, rows * h->n_feats * sizeof(float), cudaMemcpyDeviceToDevice, st));
    DEXB_CUDA_OK(cudaMemcpyAsync(logw_dev, g_logw, rows * sizeof(float), cudaMemcpyDeviceToDevice, st));
  }
  return 0;
}

long dexb_text_last_launch_count(const dexb_text* h) { return h != nullptr ? h->launches : 0; }

int dexb_text_set_layer_limit(dexb_text* h, int n_layers) {
  DEXB_CHECK(h != nullptr && n_layers >= -1 && n_layers <= h->L, "dexb_text_set_layer_limit: bad argument");
  h->layer_limit = n_layers;
  return 0;
}

int dexb_text_copy_stream(const dexb_text* h, float* rows_dev, void* stream) {
  DEXB_CHECK(h != nullptr && rows_dev != nullptr && h->hf != nullptr, "dexb_text_copy_stream: no forward has run on this handle");
  DEXB_CUDA_OK(cudaMemcpyAsync(rows_dev, h->hf, (size_t)h->B * h->T * h->C * sizeof(float), cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
  return 0;
}

}  // extern "C"
