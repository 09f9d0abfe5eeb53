// HiFi-GAN v1 generator -- the vocoder the reference synthesises with (mel -> waveform):
// Generator.forward, DEX-TTS/hifigan/models.py:157-173, ResBlock.forward :96-103, in the state get_vocoder leaves it in
// (eval + remove_weight_norm, DEX-TTS/src/utils.py:251-281; sizes from DEX-TTS/hifigan/config.json).
//
//   x = conv_pre(mel)                                   Conv1d(80 -> 512, k 7)
//   4x { x = ConvTranspose1d(leaky_relu(x, 0.1))        stride u in (8, 8, 2, 2), kernel 2u, channels halve
//        x = (ResBlock_3(x) + ResBlock_7(x) + ResBlock_11(x)) / 3 }
//   ResBlock_k: 3x { x = x + conv_k(leaky_relu(conv_k,dil d(leaky_relu(x)))) }, d in (1, 3, 5)
//   wav = tanh(conv_post(leaky_relu(x, 0.01)))          Conv1d(32 -> 1, k 7)
//
// Layout: rows [B * T_i][channels] (time-major, one image row per utterance), so every Conv1d is a 1 x k-tap implicit GEMM on the
// tcgen05 engine of the loop (gemm.cuh: split-bf16 x3, fp32 accumulation in TMEM, TMA zero fill = the convolution padding, the
// dilation is the tap step `tap_sw`).  No elementwise kernel runs inside a ResBlock: the producing GEMM's epilogue writes the
// residual stream x as fp32 rows AND leaky_relu(x) as the split-bf16 operand of the next convolution (EpiParams::s_lrelu), and the
// second convolution of a pair adds the residual in its epilogue (in place on x).
// ConvTranspose1d(k = 2u, stride u, padding u/2) is ONE 3-tap GEMM with N = u * C_out: output time t = q u + r (phase r) only sees
// the inputs q - 1, q (r < u/2) or q, q + 1 (r >= u/2), so the weight of (tap, phase) is w[:, :, r + u/2 - (tap - 1) u] where that
// index is inside [0, 2u) and zero elsewhere (1.5x the MACs of the two live taps, but one launch, and the N = u * C_out columns of
// an input row ARE the u output rows, contiguous).  The split operand of the next stage is written per output time step through
// the grouped column map of the epilogue (out_s_gshift / out_s_gpitch).
// The whole forward of a (B, T) shape is captured in a CUDA graph at first use (79 GEMMs + 6 small kernels).
#include <stdlib.h>
#include <string.h>

#include <initializer_list>
#include <map>
#include <string>
#include <vector>

#include "../../include/dexb200.h"
#include "gemm_host.cuh"

namespace dexb {

struct VocTensor {
  float* p = nullptr;
  std::vector<int64_t> shape;
  size_t n = 0;
};

struct VocConv {
  bf16* w = nullptr;                 // [tap][N][hi(K)|lo(K)], K = input channels padded to a multiple of 64
  float* bias = nullptr;             // [N]
  int ci = 0, K = 0, N = 0, taps = 0, dil = 1, off = 0;
  GemmPlan plan;
};

constexpr int kVocStages = 4, kVocRes = 3, kVocDil = 3;

}  // namespace dexb

struct dexb_voc {
  int n_mels = 80, ch0 = 512;
  int rates[dexb::kVocStages] = {8, 8, 2, 2};
  int rk[dexb::kVocRes] = {3, 7, 11};
  int rd[dexb::kVocDil] = {1, 3, 5};
  std::map<std::string, dexb::VocTensor> w;
  bool finalized = false;
  dexb::VocConv pre, ups[dexb::kVocStages], c1[dexb::kVocStages][dexb::kVocRes][dexb::kVocDil],
      c2[dexb::kVocStages][dexb::kVocRes][dexb::kVocDil];
  float* post_w = nullptr;           // [7][32]
  float* post_b = nullptr;
  // plan
  int B = 0, T = 0;
  void* ws = nullptr;
  size_t ws_bytes = 0;
  dexb::bf16* mel_s = nullptr;       // [B*T][hi(128)|lo(128)]
  dexb::bf16* x0_s = nullptr;        // leaky_relu(conv_pre) [B*T][hi(512)|lo(512)]
  float* xup[dexb::kVocStages] = {};             // ConvTranspose output, fp32 rows
  dexb::bf16* xup_s[dexb::kVocStages] = {};      // leaky_relu of it, split rows
  float* xr[dexb::kVocStages][dexb::kVocRes] = {};   // residual stream of each ResBlock
  dexb::bf16* xr_s[dexb::kVocStages] = {};       // leaky_relu(residual stream), split rows (one buffer: the ResBlocks run in turn)
  dexb::bf16* h_s[dexb::kVocStages] = {};        // leaky_relu(conv1 output)
  dexb::bf16* nxt_s[dexb::kVocStages] = {};      // leaky_relu(stage average): operand of the next ConvTranspose
  float* yfin = nullptr;             // leaky_relu(last average, 0.01), fp32 rows [B*T*256][32]
  float* mel_in = nullptr;           // staged input / output (fixed addresses for the graph)
  float* wav_out = nullptr;
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t cap_stream = nullptr;
  long launches = 0;
};

namespace dexb {

static inline int vpad64(int k) { return (k + 63) / 64 * 64; }
static inline int ilog2(int v) { int s = 0; while ((1 << s) < v) ++s; return s; }

// ---- weight packing ------------------------------------------------------------------------------------------------------------
// Conv1d weight (co, ci, k) -> [tap][co][hi(K) | lo(K)], zero beyond ci
__global__ void k_voc_pack_conv(const float* __restrict__ w, bf16* __restrict__ out, int co, int ci, int k, int K) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)k * co * K) return;
  const int kk = (int)(i % K), n = (int)((i / K) % co), tap = (int)(i / ((long)K * co));
  const float v = kk < ci ? w[((long)n * ci + kk) * k + tap] : 0.f;
  bf16 hi, lo;
  split2(v, hi, lo);
  out[((long)tap * co + n) * 2 * K + kk] = hi;
  out[((long)tap * co + n) * 2 * K + K + kk] = lo;
}
// ConvTranspose1d weight (ci, co, 2u) -> 3 taps (input q - 1, q, q + 1) x N = u * co rows (n = r * co + c, phase r):
// W[tap][n][k] = w[k][c][r + u/2 - (tap - 1) u] where that kernel index lies in [0, 2u), else 0
__global__ void k_voc_pack_convT(const float* __restrict__ w, bf16* __restrict__ out, int ci, int co, int u, int K) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const int N = u * co;
  if (i >= 3L * N * K) return;
  const int kk = (int)(i % K), n = (int)((i / K) % N), tap = (int)(i / ((long)K * N));
  const int r = n / co, c = n % co;
  const int j = r + u / 2 - (tap - 1) * u;
  const float v = (kk < ci && j >= 0 && j < 2 * u) ? w[((long)kk * co + c) * (2 * u) + j] : 0.f;
  bf16 hi, lo;
  split2(v, hi, lo);
  out[((long)tap * N + n) * 2 * K + kk] = hi;
  out[((long)tap * N + n) * 2 * K + K + kk] = lo;
}
__global__ void k_voc_tile_bias(const float* __restrict__ b, float* __restrict__ out, int co, int u) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < co * u) out[i] = b[i % co];
}
// conv_post weight (1, C, 7) -> [7][C]
__global__ void k_voc_pack_post(const float* __restrict__ w, float* __restrict__ out, int C) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 7 * C) out[i] = w[(i % C) * 7 + i / C];
}

// ---- activations ---------------------------------------------------------------------------------------------------------------
// mel (B, C, T) channel-major -> split rows [B*T][hi(K)|lo(K)] (columns >= C stay zero: the buffer is cleared once per plan)
__global__ void k_voc_in(const float* __restrict__ mel, bf16* __restrict__ xs, int B, int C, int T, int K) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * C * T) return;
  const int t = (int)(i % T), c = (int)((i / T) % C), b = (int)(i / ((long)T * C));
  bf16 hi, lo;
  split2(mel[i], hi, lo);
  bf16* row = xs + ((long)b * T + t) * 2 * K;
  row[c] = hi;
  row[K + c] = lo;
}
// x = (a + b + c) / 3 (models.py:163-168, in that order), y = leaky_relu(x, slope) -> split rows [rows][hi(K)|lo(K)] and / or fp32 rows
__global__ void k_voc_avg3(const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ c, bf16* __restrict__ os,
                           float* __restrict__ of, long rows, int C, int K, float slope) {
  pdl_wait();
  const long i = (blockIdx.x * (long)blockDim.x + threadIdx.x) * 4;
  if (i >= rows * C) return;
  const float4 va = *reinterpret_cast<const float4*>(a + i), vb = *reinterpret_cast<const float4*>(b + i),
               vc = *reinterpret_cast<const float4*>(c + i);
  float v[4] = {((va.x + vb.x) + vc.x) / 3.f, ((va.y + vb.y) + vc.y) / 3.f, ((va.z + vb.z) + vc.z) / 3.f, ((va.w + vb.w) + vc.w) / 3.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) v[j] = v[j] > 0.f ? v[j] : v[j] * slope;
  if (of != nullptr) *reinterpret_cast<float4*>(of + i) = make_float4(v[0], v[1], v[2], v[3]);
  if (os != nullptr) {
    const long r = i / C;
    const int col = (int)(i % C);
    bf16 h[4], l[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) split2(v[j], h[j], l[j]);
    bf16* row = os + r * 2 * K;
    *reinterpret_cast<uint2*>(row + col) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(row + K + col) = *reinterpret_cast<const uint2*>(l);
  }
}
// wav[b][t] = tanh(bias + sum_{j < 7, c < C} w[j][c] * y[b][t + j - 3][c]), zero padding at both ends of an utterance.
// A block stages its 256 + 6 rows of y in shared memory with coalesced loads (row stride C + 1: conflict-free for the row-per-thread
// reads); one thread per row reading its seven rows straight from global memory was 32 sectors per load instruction (235 us for 134 MB).
template <int C>
__global__ void __launch_bounds__(256) k_voc_post(const float* __restrict__ y, const float* __restrict__ w, const float* __restrict__ bias,
                                                  float* __restrict__ wav, int B, int L) {
  pdl_wait();
  __shared__ float ws[7 * C];
  __shared__ float ys[(256 + 6) * (C + 1)];
  for (int i = threadIdx.x; i < 7 * C; i += 256) ws[i] = w[i];
  const long i0 = blockIdx.x * 256L;                         // first output of the block (blocks never straddle utterances: L % 256 == 0
  const int b = (int)(i0 / L), t0 = (int)(i0 % L);           //  is checked by the launcher; otherwise rows are clamped per element below)
  for (int k = threadIdx.x; k < (256 + 6) * (C / 4); k += 256) {
    const int r = k / (C / 4), c4 = k % (C / 4);
    const int tt = t0 + r - 3;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (tt >= 0 && tt < L) q = *reinterpret_cast<const float4*>(y + ((long)b * L + tt) * C + c4 * 4);
    float* d = ys + r * (C + 1) + c4 * 4;
    d[0] = q.x; d[1] = q.y; d[2] = q.z; d[3] = q.w;
  }
  __syncthreads();
  const int t = t0 + threadIdx.x;
  if (t >= L) return;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 7; ++j) {
    const float* yr = ys + (threadIdx.x + j) * (C + 1);
#pragma unroll
    for (int c = 0; c < C; ++c) acc = fmaf(yr[c], ws[j * C + c], acc);
  }
  wav[i0 + threadIdx.x] = tanhf(acc + bias[0]);
}

// ---- host ----------------------------------------------------------------------------------------------------------------------
static int voc_get(dexb_voc* h, const std::string& name, std::initializer_list<int64_t> shape, const float** out) {
  auto it = h->w.find(name);
  DEXB_CHECK(it != h->w.end(), "vocoder: weight '%s' was not loaded", name.c_str());
  const std::vector<int64_t> want(shape);
  DEXB_CHECK(it->second.shape == want, "vocoder: weight '%s' has the wrong shape", name.c_str());
  *out = it->second.p;
  return 0;
}

static int voc_pack_conv(dexb_voc* h, const std::string& prefix, int ci, int co, int k, int dil, VocConv* c, cudaStream_t st) {
  const float *w = nullptr, *b = nullptr;
  DEXB_TRY(voc_get(h, prefix + ".weight", {co, ci, k}, &w));
  DEXB_TRY(voc_get(h, prefix + ".bias", {co}, &b));
  c->ci = ci; c->K = vpad64(ci); c->N = co; c->taps = k; c->dil = dil; c->off = -(dil * (k - 1)) / 2;
  if (c->w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->w, (size_t)k * co * 2 * c->K * sizeof(bf16)));
  if (c->bias == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->bias, (size_t)co * sizeof(float)));
  k_voc_pack_conv<<<cdiv((long)k * co * c->K, 256), 256, 0, st>>>(w, c->w, co, ci, k, c->K);
  DEXB_CUDA_OK(cudaMemcpyAsync(c->bias, b, (size_t)co * sizeof(float), cudaMemcpyDeviceToDevice, st));
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

static int voc_pack_up(dexb_voc* h, int i, int ci, int co, int u, VocConv* c, cudaStream_t st) {
  const float *w = nullptr, *b = nullptr;
  const std::string prefix = "ups." + std::to_string(i);
  DEXB_TRY(voc_get(h, prefix + ".weight", {ci, co, 2 * u}, &w));
  DEXB_TRY(voc_get(h, prefix + ".bias", {co}, &b));
  c->ci = ci; c->K = vpad64(ci); c->N = u * co; c->taps = 3; c->dil = 1; c->off = -1;
  if (c->w == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->w, (size_t)3 * c->N * 2 * c->K * sizeof(bf16)));
  if (c->bias == nullptr) DEXB_CUDA_OK(cudaMalloc(&c->bias, (size_t)c->N * sizeof(float)));
  k_voc_pack_convT<<<cdiv(3L * c->N * c->K, 256), 256, 0, st>>>(w, c->w, ci, co, u, c->K);
  k_voc_tile_bias<<<cdiv(c->N, 256), 256, 0, st>>>(b, c->bias, co, u);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

static void voc_free_conv(VocConv* c) {
  cudaFree(c->w); cudaFree(c->bias);
  c->w = nullptr; c->bias = nullptr;
}

static void voc_release_plan(dexb_voc* h) {
  if (h->graph_exec != nullptr) { cudaGraphExecDestroy(h->graph_exec); h->graph_exec = nullptr; }
  if (h->graph != nullptr) { cudaGraphDestroy(h->graph); h->graph = nullptr; }
  if (h->ws != nullptr) { cudaFree(h->ws); h->ws = nullptr; }
  h->B = h->T = 0;
}

// Conv1d / phase-stacked ConvTranspose1d as a 1 x taps implicit GEMM over rows [B][1][L][2K]
static int voc_plan_gemm(dexb_voc* h, VocConv* c, const bf16* a, int L, float* out_f, const float* resid, bf16* out_s, long s_stride,
                         int s_lo, int gshift, int gpitch, float slope) {
  GemmParams p;
  memset(&p, 0, sizeof(p));
  p.nz = h->B; p.nheads = 1;
  p.H = 1; p.W = L;
  p.in_stride = 1;
  p.CH = 1; p.CW = L; p.OH = 1; p.OW = L;
  p.out_scale = 1; p.tap_sw = c->dil;
  p.KH = 1; p.KW = c->taps; p.offH = 0; p.offW = c->off;
  p.K = c->K; p.N = c->N;
  p.A = a; p.a_row_stride = 2L * c->K; p.a_hi = 0; p.a_lo = c->K;
  p.Bw = c->w; p.b_row_stride = 2L * c->K; p.b_hi = 0; p.b_lo = c->K; p.b_rows_per_tap = c->N;
  p.nsplit = 3;
  p.epi.alpha = 1.f; p.epi.out_s_ncols = 1 << 30;
  p.epi.bias = c->bias;
  p.epi.out_f32 = out_f; p.epi.out_f32_stride = c->N;
  p.epi.resid_f32 = resid; p.epi.resid_f32_stride = c->N;
  p.epi.out_s = out_s; p.epi.out_s_stride = s_stride; p.epi.out_s_hi = 0; p.epi.out_s_lo = s_lo;
  p.epi.out_s_gshift = gshift; p.epi.out_s_gpitch = gpitch;
  p.epi.s_lrelu = slope;
  p.BW = 128; p.BH = 1;
  DEXB_TRY(gemm_plan_init(&c->plan, p, h->B, (long)c->taps * c->N, 1));
  DEXB_CHECK(c->plan.tc_ok, "vocoder: convolution %d -> %d (k %d) is not eligible for the tcgen05 engine", c->ci, c->N, c->taps);
  return 0;
}

struct VocArena {
  char* base = nullptr;
  size_t off = 0;
  template <class T> T* get(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base != nullptr ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};

static void voc_layout(dexb_voc* h, VocArena& ar, int B, int T) {
  const long rows0 = (long)B * T;
  h->mel_in = ar.get<float>((size_t)rows0 * h->n_mels);
  h->mel_s = ar.get<bf16>((size_t)rows0 * 2 * vpad64(h->n_mels));
  h->x0_s = ar.get<bf16>((size_t)rows0 * 2 * h->ch0);
  long L = T;
  int ch = h->ch0;
  for (int i = 0; i < kVocStages; ++i) {
    L *= h->rates[i]; ch /= 2;
    const size_t rows = (size_t)B * L, K = vpad64(ch);
    h->xup[i] = ar.get<float>(rows * ch);
    h->xup_s[i] = ar.get<bf16>(rows * 2 * K);
    for (int r = 0; r < kVocRes; ++r) h->xr[i][r] = ar.get<float>(rows * ch);
    h->xr_s[i] = ar.get<bf16>(rows * 2 * K);
    h->h_s[i] = ar.get<bf16>(rows * 2 * K);
    h->nxt_s[i] = (i + 1 < kVocStages) ? ar.get<bf16>(rows * 2 * K) : nullptr;
  }
  h->yfin = ar.get<float>((size_t)B * L * ch);
  h->wav_out = ar.get<float>((size_t)B * L);
}

static int voc_plan(dexb_voc* h, int B, int T) {
  if (B == h->B && T == h->T) return 0;
  voc_release_plan(h);
  DEXB_TRY(gemm_global_init());
  VocArena m;
  voc_layout(h, m, B, T);
  h->ws_bytes = m.off + 256;
  DEXB_CUDA_OK(cudaMalloc(&h->ws, h->ws_bytes));
  DEXB_CUDA_OK(cudaMemset(h->ws, 0, h->ws_bytes));        // the zero padding of K (80 -> 128 mel bins, 32 -> 64 channels) is never written again
  VocArena ar; ar.base = reinterpret_cast<char*>(h->ws);
  voc_layout(h, ar, B, T);
  h->B = B; h->T = T;
  // conv_pre: split(leaky_relu(.)) only
  DEXB_TRY(voc_plan_gemm(h, &h->pre, h->mel_s, T, nullptr, nullptr, h->x0_s, 2L * h->ch0, h->ch0, 0, 0, 0.1f));
  long L = T;
  int ch = h->ch0;
  const bf16* in_s = h->x0_s;
  for (int i = 0; i < kVocStages; ++i) {
    const int u = h->rates[i], co = ch / 2, K = vpad64(co);
    // ConvTranspose: fp32 rows [B*L][u*co] == [B*L*u][co]; split(leaky_relu) per output time step
    DEXB_TRY(voc_plan_gemm(h, &h->ups[i], in_s, (int)L, h->xup[i], nullptr, h->xup_s[i], (long)u * 2 * K, K, ilog2(co), 2 * K, 0.1f));
    L *= u; ch = co;
    for (int r = 0; r < kVocRes; ++r)
      for (int d = 0; d < kVocDil; ++d) {
        const bf16* a1 = (d == 0) ? h->xup_s[i] : h->xr_s[i];
        const float* res = (d == 0) ? h->xup[i] : h->xr[i][r];
        DEXB_TRY(voc_plan_gemm(h, &h->c1[i][r][d], a1, (int)L, nullptr, nullptr, h->h_s[i], 2L * K, K, 0, 0, 0.1f));
        // x = x + conv2(.): fp32 residual stream in place, split(leaky_relu(x)) for the next pair (not needed after the last one)
        DEXB_TRY(voc_plan_gemm(h, &h->c2[i][r][d], h->h_s[i], (int)L, h->xr[i][r], res, (d + 1 < kVocDil) ? h->xr_s[i] : nullptr,
                               2L * K, K, 0, 0, (d + 1 < kVocDil) ? 0.1f : 0.f));
      }
    in_s = h->nxt_s[i];
  }
  return 0;
}

static int voc_enqueue(dexb_voc* h, cudaStream_t st) {
  const int B = h->B, T = h->T;
  h->launches = 0;
  launch_pdl(k_voc_in, dim3((unsigned)(cdiv((long)B * h->n_mels * T, 256))), dim3(256), 0, st, h->mel_in, h->mel_s, B, h->n_mels, T, vpad64(h->n_mels));
  DEXB_TRY(gemm_launch(h->pre.plan, h->pre.plan.p, 0, st));
  h->launches += 2;
  long L = T;
  int ch = h->ch0;
  for (int i = 0; i < kVocStages; ++i) {
    DEXB_TRY(gemm_launch(h->ups[i].plan, h->ups[i].plan.p, 0, st));
    ++h->launches;
    L *= h->rates[i]; ch /= 2;
    for (int r = 0; r < kVocRes; ++r)
      for (int d = 0; d < kVocDil; ++d) {
        DEXB_TRY(gemm_launch(h->c1[i][r][d].plan, h->c1[i][r][d].plan.p, 0, st));
        DEXB_TRY(gemm_launch(h->c2[i][r][d].plan, h->c2[i][r][d].plan.p, 0, st));
        h->launches += 2;
      }
    const long rows = (long)B * L;
    const bool last = i + 1 == kVocStages;
    launch_pdl(k_voc_avg3, dim3((unsigned)(cdiv(rows * ch / 4, 256))), dim3(256), 0, st, h->xr[i][0], h->xr[i][1], h->xr[i][2], last ? nullptr : h->nxt_s[i],
                                                         last ? h->yfin : nullptr, rows, ch, vpad64(ch), last ? 0.01f : 0.1f);
    ++h->launches;
  }
  DEXB_CHECK(L % 256 == 0, "vocoder: %ld output samples per utterance are not a multiple of the hop size 256", (long)L);
  launch_pdl(k_voc_post<32>, dim3((unsigned)(cdiv((long)B * L, 256))), dim3(256), 0, st, h->yfin, h->post_w, h->post_b, h->wav_out, B, (int)L);
  ++h->launches;
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dexb

using namespace dexb;

extern "C" {

int dexb_voc_create(int n_mels, int initial_channels, const int* upsample_rates, int n_up, const int* resblock_kernels, int n_rk,
                    const int* resblock_dilations, int n_rd, dexb_voc** out) {
  DEXB_CHECK(out != nullptr && upsample_rates != nullptr && resblock_kernels != nullptr && resblock_dilations != nullptr,
             "dexb_voc_create: null argument");
  int dev = 0, major = 0;
  DEXB_CUDA_OK(cudaGetDevice(&dev));
  DEXB_CUDA_OK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  DEXB_CHECK(major == 10, "dexb200 is built for sm_100a only (device %d has compute capability major %d); there is no fallback",
             dev, major);
  DEXB_CHECK(n_up == kVocStages && n_rk == kVocRes && n_rd == kVocDil,
             "dexb_voc_create: the HiFi-GAN v1 layout is instantiated (4 upsample stages, 3 ResBlocks x 3 dilations), got %d / %d / %d",
             n_up, n_rk, n_rd);
  DEXB_CHECK(n_mels >= 1 && n_mels <= 128, "dexb_voc_create: n_mels = %d out of range", n_mels);
  DEXB_CHECK(initial_channels == 512 || initial_channels == 1024, "dexb_voc_create: upsample_initial_channel must be 512 or 1024 (got %d)",
             initial_channels);
  dexb_voc* h = new dexb_voc();
  h->n_mels = n_mels; h->ch0 = initial_channels;
  int ch = initial_channels;
  for (int i = 0; i < kVocStages; ++i) {
    const int u = upsample_rates[i];
    ch /= 2;
    if (!(u == 2 || u == 4 || u == 8 || u == 16) || (u * ch) % 32 != 0) {
      delete h;
      DEXB_CHECK(false, "dexb_voc_create: upsample rate %d of stage %d is not supported (even power of two <= 16)", u, i);
    }
    h->rates[i] = u;
  }
  if (initial_channels != 512) { delete h; DEXB_CHECK(false, "dexb_voc_create: conv_post is instantiated for 32 final channels (initial 512)"); }
  for (int i = 0; i < kVocRes; ++i) {
    if (resblock_kernels[i] % 2 != 1 || resblock_kernels[i] > 15) { delete h; DEXB_CHECK(false, "dexb_voc_create: ResBlock kernel %d", resblock_kernels[i]); }
    h->rk[i] = resblock_kernels[i];
  }
  for (int i = 0; i < kVocDil; ++i) {
    if (resblock_dilations[i] < 1 || resblock_dilations[i] > 8) { delete h; DEXB_CHECK(false, "dexb_voc_create: dilation %d", resblock_dilations[i]); }
    h->rd[i] = resblock_dilations[i];
  }
  *out = h;
  return 0;
}

void dexb_voc_destroy(dexb_voc* h) {
  if (h == nullptr) return;
  voc_release_plan(h);
  if (h->cap_stream != nullptr) cudaStreamDestroy(h->cap_stream);
  voc_free_conv(&h->pre);
  for (int i = 0; i < kVocStages; ++i) {
    voc_free_conv(&h->ups[i]);
    for (int r = 0; r < kVocRes; ++r)
      for (int d = 0; d < kVocDil; ++d) { voc_free_conv(&h->c1[i][r][d]); voc_free_conv(&h->c2[i][r][d]); }
  }
  cudaFree(h->post_w); cudaFree(h->post_b);
  for (auto& kv : h->w) cudaFree(kv.second.p);
  delete h;
}

int dexb_voc_load_weight(dexb_voc* h, const char* name, const float* data_dev, const int64_t* shape, int ndim) {
  DEXB_CHECK(h != nullptr && name != nullptr && data_dev != nullptr && shape != nullptr && ndim >= 1 && ndim <= 3,
             "dexb_voc_load_weight: bad argument");
  size_t n = 1;
  for (int i = 0; i < ndim; ++i) {
    DEXB_CHECK(shape[i] >= 1, "dexb_voc_load_weight(%s): empty dimension", name);
    n *= (size_t)shape[i];
  }
  VocTensor& t = h->w[name];
  if (t.p != nullptr && t.n != n) { cudaFree(t.p); t.p = nullptr; }
  if (t.p == nullptr) DEXB_CUDA_OK(cudaMalloc(&t.p, n * sizeof(float)));
  t.n = n;
  t.shape.assign(shape, shape + ndim);
  DEXB_CUDA_OK(cudaMemcpy(t.p, data_dev, n * sizeof(float), cudaMemcpyDeviceToDevice));
  h->finalized = false;
  return 0;
}

int dexb_voc_finalize_weights(dexb_voc* h, void* stream) {
  DEXB_CHECK(h != nullptr, "null handle");
  cudaStream_t st = (cudaStream_t)stream;
  voc_release_plan(h);                      // plans and the captured graph hold the packed-weight pointers of the previous finalize
  DEXB_TRY(voc_pack_conv(h, "conv_pre", h->n_mels, h->ch0, 7, 1, &h->pre, st));
  int ch = h->ch0;
  for (int i = 0; i < kVocStages; ++i) {
    DEXB_TRY(voc_pack_up(h, i, ch, ch / 2, h->rates[i], &h->ups[i], st));
    ch /= 2;
    for (int r = 0; r < kVocRes; ++r)
      for (int d = 0; d < kVocDil; ++d) {
        const std::string p = "resblocks." + std::to_string(i * kVocRes + r);
        DEXB_TRY(voc_pack_conv(h, p + ".convs1." + std::to_string(d), ch, ch, h->rk[r], h->rd[d], &h->c1[i][r][d], st));
        DEXB_TRY(voc_pack_conv(h, p + ".convs2." + std::to_string(d), ch, ch, h->rk[r], 1, &h->c2[i][r][d], st));
      }
  }
  const float *pw = nullptr, *pb = nullptr;
  DEXB_TRY(voc_get(h, "conv_post.weight", {1, ch, 7}, &pw));
  DEXB_TRY(voc_get(h, "conv_post.bias", {1}, &pb));
  DEXB_CHECK(ch == 32, "vocoder: conv_post is instantiated for 32 input channels (got %d)", ch);
  if (h->post_w == nullptr) DEXB_CUDA_OK(cudaMalloc(&h->post_w, (size_t)7 * ch * sizeof(float)));
  if (h->post_b == nullptr) DEXB_CUDA_OK(cudaMalloc(&h->post_b, sizeof(float)));
  k_voc_pack_post<<<1, 256, 0, st>>>(pw, h->post_w, ch);
  DEXB_CUDA_OK(cudaMemcpyAsync(h->post_b, pb, sizeof(float), cudaMemcpyDeviceToDevice, st));
  DEXB_CUDA_OK(cudaGetLastError());
  DEXB_CUDA_OK(cudaStreamSynchronize(st));
  h->finalized = true;
  return 0;
}

int dexb_voc_forward(dexb_voc* h, const float* mel_dev, int B, int T, float* wav_dev, void* stream) {
  DEXB_CHECK(h != nullptr && mel_dev != nullptr && wav_dev != nullptr, "dexb_voc_forward: null argument");
  DEXB_CHECK(h->finalized, "dexb_voc_forward: call dexb_voc_finalize_weights first");
  DEXB_CHECK(B >= 1 && T >= 1, "dexb_voc_forward: B = %d, T = %d", B, T);
  cudaStream_t st = (cudaStream_t)stream;
  DEXB_TRY(voc_plan(h, B, T));
  long L = T;
  for (int i = 0; i < kVocStages; ++i) L *= h->rates[i];
  DEXB_CUDA_OK(cudaMemcpyAsync(h->mel_in, mel_dev, (size_t)B * h->n_mels * T * sizeof(float), cudaMemcpyDeviceToDevice, st));
  const char* ng = getenv("DEXB_NO_GRAPH");
  if (ng != nullptr && ng[0] == '1') {
    DEXB_TRY(voc_enqueue(h, st));
  } else {
    if (h->graph_exec == nullptr) {
      if (h->cap_stream == nullptr) DEXB_CUDA_OK(cudaStreamCreateWithFlags(&h->cap_stream, cudaStreamNonBlocking));
      DEXB_CUDA_OK(cudaStreamBeginCapture(h->cap_stream, cudaStreamCaptureModeThreadLocal));
      const int r = voc_enqueue(h, h->cap_stream);
      cudaGraph_t g = nullptr;
      const cudaError_t e = cudaStreamEndCapture(h->cap_stream, &g);
      if (r != 0) { if (g != nullptr) cudaGraphDestroy(g); return r; }
      DEXB_CHECK(e == cudaSuccess && g != nullptr, "vocoder: graph capture failed: %s", cudaGetErrorString(e));
      h->graph = g;
      DEXB_CUDA_OK(cudaGraphInstantiate(&h->graph_exec, g, 0));
    }
    DEXB_CUDA_OK(cudaGraphLaunch(h->graph_exec, st));
  }
  DEXB_CUDA_OK(cudaMemcpyAsync(wav_dev, h->wav_out, (size_t)B * L * sizeof(float), cudaMemcpyDeviceToDevice, st));
  return 0;
}

long dexb_voc_last_launch_count(const dexb_voc* h) { return h != nullptr ? h->launches : 0; }

}  // extern "C"
