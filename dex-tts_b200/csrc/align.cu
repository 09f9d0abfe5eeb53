// Duration / alignment glue between the text encoder and the decoder (the part that holds the host
// round trip): DeXTTS.forward, DEX-TTS/model/tts.py:55-68 (GeDEX-TTS/model/tts.py:37-50) over model.utils.sequence_mask /
// fix_len_compatibility / generate_path (DEX-TTS/model/utils.py:6-39).
//
//   w_ceil    = ceil(exp(logw) * x_mask) * length_scale                                   (tts.py:55-56)
//   y_lengths = clamp_min(sum(w_ceil), 1).long()                                          (:57)      -> host: Ty = fix_len(max)
//   attn      = generate_path(w_ceil, x_mask (x) y_mask)   (hard monotonic alignment)     (:62-64)
//   mu_y      = attn^T mu_x                                                               (:67)
//
// The alignment is one-hot per output frame, so `attn^T mu_x` is a gather (bit-exact: a matmul row with one non-zero term) and
// generate_path's two sequence masks are a search of each frame index in the cumulative durations.  HBM-bound byte work: the
// cumulative sums are sequential fp32 adds in torch.cumsum's order (one thread per utterance, Tx is a few hundred), everything
// else is one coalesced pass over the outputs.
#include <stdint.h>

#include "../../include/dexb200.h"
#include "common.cuh"

namespace dexb {

// one thread per utterance walks the tokens in order (same fp32 addition order as torch.cumsum on the CPU)
__global__ void k_align_len(const float* __restrict__ logw, const float* __restrict__ x_mask, float length_scale,
                            float* __restrict__ cum, long long* __restrict__ y_len, int B, int Tx) {
  pdl_wait();
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float c = 0.f;
  for (int i = 0; i < Tx; ++i) {
    const float w = __fmul_rn(expf(logw[(long)b * Tx + i]), x_mask[(long)b * Tx + i]);
    c = __fadd_rn(c, __fmul_rn(ceilf(w), length_scale));     // separate mul and add (no fma contraction), as the reference's two ops
    cum[(long)b * Tx + i] = c;
  }
  y_len[b] = (long long)fmaxf(c, 1.f);                       // clamp_min(sum, 1).long(): truncation
}

// one thread per (utterance, output frame, group of ALIGN_FG features): token i with cum[i-1] <= t < cum[i].  The feature groups
// (blockIdx.y) put B*Ty/256 * F/8 CTAs on the machine (160 at B = 8, Ty = 512, F = 80: one wave of the 148 SMs) instead of 16;
// each repeats the 9-step search in the L1-resident cum row, group 0 also writes the alignment column and the frame mask.
constexpr int ALIGN_FG = 8;
__global__ void k_align_expand(const float* __restrict__ cum, const float* __restrict__ x_mask, const long long* __restrict__ y_len,
                               const float* __restrict__ mu_x, float* __restrict__ attn, float* __restrict__ y_mask,
                               float* __restrict__ mu_y, int B, int Tx, int F, int Ty) {
  pdl_wait();
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)B * Ty) return;
  const int b = (int)(idx / Ty), t = (int)(idx % Ty);
  const float ym = (long long)t < y_len[b] ? 1.f : 0.f;
  // first i with t < cum[i] (cum is non-decreasing): binary search
  const float* cb = cum + (long)b * Tx;
  const float ft = (float)t;
  int lo = 0, hi = Tx;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (ft < cb[mid]) hi = mid; else lo = mid + 1;
  }
  const float a = lo < Tx ? x_mask[(long)b * Tx + lo] * ym : 0.f;      // path * (x_mask (x) y_mask)
  if (blockIdx.y == 0) {
    y_mask[idx] = ym;
    if (attn != nullptr && lo < Tx) attn[((long)b * Tx + lo) * Ty + t] = a;   // the rest of the column was zeroed by the caller
  }
  const int f0 = blockIdx.y * ALIGN_FG, f1 = min(F, f0 + ALIGN_FG);
  for (int f = f0; f < f1; ++f)                                         // stores coalesced over t; the mu_x column is a broadcast
    mu_y[((long)b * F + f) * Ty + t] = lo < Tx ? a * mu_x[((long)b * F + f) * Tx + lo] : 0.f;
}

}  // namespace dexb

using namespace dexb;

extern "C" {

int dexb_align_lengths(const float* logw_dev, const float* x_mask_dev, int B, int Tx, float length_scale, float* cum_dev,
                       int64_t* y_lengths_dev, int64_t* y_lengths_host, void* stream) {
  DEXB_CHECK(logw_dev != nullptr && x_mask_dev != nullptr && cum_dev != nullptr && y_lengths_dev != nullptr && y_lengths_host != nullptr,
             "dexb_align_lengths: null argument");
  DEXB_CHECK(B >= 1 && Tx >= 1 && length_scale > 0.f, "dexb_align_lengths: B = %d, Tx = %d, length_scale = %g", B, Tx, length_scale);
  cudaStream_t st = (cudaStream_t)stream;
  static_assert(sizeof(long long) == sizeof(int64_t), "int64_t layout");
  launch_pdl(k_align_len, dim3((unsigned)(cdiv(B, 32))), dim3(32), 0, st, logw_dev, x_mask_dev, length_scale, cum_dev, reinterpret_cast<long long*>(y_lengths_dev), B, Tx);
  DEXB_CUDA_OK(cudaGetLastError());
  DEXB_CUDA_OK(cudaMemcpyAsync(y_lengths_host, y_lengths_dev, (size_t)B * sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  DEXB_CUDA_OK(cudaStreamSynchronize(st));          // the reference's own host round trip: int(y_lengths.max()), tts.py:58
  return 0;
}

int dexb_align_expand(const float* cum_dev, const float* x_mask_dev, const int64_t* y_lengths_dev, const float* mu_x_dev, int B,
                      int Tx, int n_feats, int Ty, float* attn_dev, float* y_mask_dev, float* mu_y_dev, void* stream) {
  DEXB_CHECK(cum_dev != nullptr && x_mask_dev != nullptr && y_lengths_dev != nullptr && mu_x_dev != nullptr && y_mask_dev != nullptr &&
                 mu_y_dev != nullptr, "dexb_align_expand: null argument");
  DEXB_CHECK(B >= 1 && Tx >= 1 && n_feats >= 1 && Ty >= 1 && cdiv(n_feats, ALIGN_FG) <= 65535,
             "dexb_align_expand: B = %d, Tx = %d, n_feats = %d, Ty = %d", B, Tx, n_feats, Ty);
  cudaStream_t st = (cudaStream_t)stream;
  if (attn_dev != nullptr) DEXB_CUDA_OK(cudaMemsetAsync(attn_dev, 0, (size_t)B * Tx * Ty * sizeof(float), st));
  launch_pdl(k_align_expand, dim3(cdiv((long)B * Ty, 256), cdiv(n_feats, ALIGN_FG)), dim3(256), 0, st, cum_dev, x_mask_dev, reinterpret_cast<const long long*>(y_lengths_dev), mu_x_dev,
                                                          attn_dev, y_mask_dev, mu_y_dev, B, Tx, n_feats, Ty);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
