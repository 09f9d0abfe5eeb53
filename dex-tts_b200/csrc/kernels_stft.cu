// STFT -> magnitude -> mel filterbank -> log, one pass over the audio (bandwidth-bound; no tensor cores).
// Reference semantics: DEX-TTS/audio/stft.py:52-81 (reflect pad n_fft/2, Hann-windowed DFT as conv1d, stride = hop),
// :159-178 (mel_basis @ magnitude, log(clamp(., 1e-5))), DEX-TTS/audio/audio_processing.py:85-91.
//
// One CTA (256 threads) produces kFramesPerCta consecutive frames of one utterance: the overlapping audio segment is
// staged once in shared memory with coalesced loads, every frame is transformed by a 1024-point radix-2 FFT in shared
// memory (the reference's dense 1026x1024 DFT matmul is 2.1 MFLOP/frame; the FFT is 0.05), magnitudes are folded
// through the mel matrix by warps and the log-mel columns are written out per mel row.
#include "kernels.cuh"

namespace dexb {

constexpr int kFftN = 1024;
constexpr int kFftLog = 10;
constexpr int kFramesPerCta = 4;

__device__ __forceinline__ int reflect_idx(int i, int S) {
  if (i < 0) i = -i;
  if (i >= S) i = 2 * (S - 1) - i;
  return min(max(i, 0), S - 1);        // only the unused tail of the last CTA's segment is ever clamped
}

__global__ void __launch_bounds__(256) k_stft_mel(const float* __restrict__ wav, int S, int hop,
                                                  const float* __restrict__ window, const float* __restrict__ mel_basis,
                                                  int n_mels, int n_frames, float* __restrict__ mel) {
  __shared__ float seg[kFftN + (kFramesPerCta - 1) * 256];
  __shared__ float2 buf[kFftN];
  __shared__ float2 tw[kFftN / 2];
  __shared__ float mag[kFftN / 2 + 1];
  __shared__ float win[kFftN];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerCta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* w = wav + (long)b * S;
  const int seg_len = kFftN + (kFramesPerCta - 1) * hop;
  const int start = f0 * hop - kFftN / 2;
  for (int i = tid; i < seg_len; i += 256) seg[i] = w[reflect_idx(start + i, S)];
  for (int i = tid; i < kFftN; i += 256) win[i] = window[i];
  for (int k = tid; k < kFftN / 2; k += 256) {
    float s, c;
    sincospif((float)(2 * k) / (float)kFftN, &s, &c);
    tw[k] = make_float2(c, -s);
  }
  __syncthreads();
  const int nf = min(kFramesPerCta, n_frames - f0);
  for (int f = 0; f < nf; ++f) {
    // windowed frame in bit-reversed order
    for (int i = tid; i < kFftN; i += 256) {
      const int r = __brev((unsigned)i) >> (32 - kFftLog);
      buf[r] = make_float2(seg[f * hop + i] * win[i], 0.f);
    }
    __syncthreads();
#pragma unroll 1
    for (int s = 1; s <= kFftLog; ++s) {
      const int half = 1 << (s - 1);
      const int tstep = kFftN >> s;
#pragma unroll
      for (int jj = 0; jj < 2; ++jj) {
        const int j = tid + jj * 256;
        const int pos = j & (half - 1);
        const int i0 = ((j >> (s - 1)) << s) + pos;
        const int i1 = i0 + half;
        const float2 t = tw[pos * tstep];
        const float2 a = buf[i0], c = buf[i1];
        const float2 wc = make_float2(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
        buf[i0] = make_float2(a.x + wc.x, a.y + wc.y);
        buf[i1] = make_float2(a.x - wc.x, a.y - wc.y);
      }
      __syncthreads();
    }
    for (int k = tid; k <= kFftN / 2; k += 256) mag[k] = sqrtf(buf[k].x * buf[k].x + buf[k].y * buf[k].y);
    __syncthreads();
    for (int m = warp; m < n_mels; m += 8) {
      const float* row = mel_basis + (long)m * (kFftN / 2 + 1);
      float acc = 0.f;
      for (int k = lane; k <= kFftN / 2; k += 32) acc = fmaf(row[k], mag[k], acc);
      acc = warp_sum(acc);
      if (lane == 0) mel[((long)b * n_mels + m) * n_frames + f0 + f] = logf(fmaxf(acc, 1e-5f));
    }
    __syncthreads();
  }
}

int launch_stft_mel(const float* wav, int B, int S, const float* window, const float* mel_basis, int n_fft, int hop,
                    int n_mels, float* mel, cudaStream_t st) {
  DEXB_CHECK(n_fft == kFftN && hop == 256, "stft_mel is instantiated for n_fft 1024 / hop 256 (got %d / %d)", n_fft, hop);
  DEXB_CHECK(S > n_fft / 2, "stft_mel: reflect padding needs more than n_fft/2 samples (got %d)", S);
  DEXB_CHECK(B >= 1 && n_mels >= 1, "stft_mel: empty batch");
  const int n_frames = S / hop + 1;
  dim3 grid(cdiv(n_frames, kFramesPerCta), B);
  k_stft_mel<<<grid, 256, 0, st>>>(wav, S, hop, window, mel_basis, n_mels, n_frames, mel);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dexb
