// STFT -> magnitude -> mel filterbank -> log (and the per-frame spectral energy), one pass over the audio.
// Reference semantics: DEX-TTS/audio/stft.py:52-81 (reflect pad n_fft/2, Hann-windowed DFT as conv1d, stride = hop),
// :159-178 (mel_basis @ magnitude, log(clamp(., 1e-5)), energy = ||magnitude||_2), DEX-TTS/audio/audio_processing.py:85-91.
//
// Bound: HBM on paper (4 S bytes in, 4 (n_mels + 1) (1 + S / hop) bytes out per utterance = 349 KB for 3 s of audio -- 11 MB for a
// batch of 32, two microseconds of HBM time), in practice the latency of the FFT passes: the kernel is sized so that the whole
// batch is resident at once (one CTA per 8 frames, ~46 KB of shared memory, 4 CTAs per SM).
//
// One CTA (256 threads) produces 8 consecutive frames of one utterance:
//   1. the overlapping audio segment (1024 + 7 * 256 samples) is staged once with coalesced, 128-bit loads where the row alignment
//      allows (reflect padding resolved while staging);
//   2. the 8 real frames are transformed as 4 complex 1024-point FFTs (frame 2j in the real part, 2j+1 in the imaginary part; the
//      two spectra are separated with X1[k] = (Z[k] + conj Z[N-k]) / 2, X2[k] = (Z[k] - conj Z[N-k]) / 2i), all four advancing
//      through the 10 radix-2 passes together -- one block barrier per pass for 8 frames.  The reference's dense 1026 x 1024 DFT
//      convolution is 2.1 MFLOP per frame, the packed FFT 0.03;
//   3. the mel matrix of the reference (librosa, Slaney) is triangular: row m is non-zero on one contiguous run of ~13 of the 513
//      bins.  The runs are found once per CTA (ballot scan, the matrix stays in L2) and the product is taken over the runs only:
//      thread = (mel row, frame), 41 k -> ~1.1 k multiply-adds per frame;
//   4. log-mel is written as 32 B row segments (8 frames of one mel row), the energy by one warp per frame.
#include "kernels.cuh"

namespace dexb {

constexpr int kFftN = 1024;
constexpr int kFftLog = 10;
constexpr int kFramesPerCta = 8;
constexpr int kHop = 256;
constexpr int kSegLen = kFftN + (kFramesPerCta - 1) * kHop;      // 2816
constexpr int kBins = kFftN / 2 + 1;
constexpr int kMaxMels = 128;

__device__ __forceinline__ int reflect_idx(int i, int S) {
  if (i < 0) i = -i;
  if (i >= S) i = 2 * (S - 1) - i;
  return min(max(i, 0), S - 1);        // only the unused tail of the last CTA's segment is ever clamped
}

__global__ void __launch_bounds__(256, 4) k_stft_mel(const float* __restrict__ wav, int S, const float* __restrict__ window,
                                                     const float* __restrict__ mel_basis, int n_mels, int n_frames,
                                                     float* __restrict__ mel, float* __restrict__ energy) {
  __shared__ __align__(16) float seg[kSegLen];
  __shared__ float2 buf[kFramesPerCta / 2][kFftN];               // 4 complex FFTs; afterwards reused as mag[8][513]
  __shared__ float2 tw[kFftN / 2];
  __shared__ short run_lo[kMaxMels], run_hi[kMaxMels];
  const int b = blockIdx.y;
  const int f0 = blockIdx.x * kFramesPerCta;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* w = wav + (long)b * S;
  const int start = f0 * kHop - kFftN / 2;
  // ---- 1. stage the segment
  const bool vec = ((reinterpret_cast<uintptr_t>(w + start) & 15) == 0) && start >= 0 && start + kSegLen <= S;
  if (vec) {
    const float4* src = reinterpret_cast<const float4*>(w + start);
    for (int i = tid; i < kSegLen / 4; i += 256) reinterpret_cast<float4*>(seg)[i] = __ldg(src + i);
  } else {
    for (int i = tid; i < kSegLen; i += 256) seg[i] = __ldg(w + reflect_idx(start + i, S));
  }
  for (int k = tid; k < kFftN / 2; k += 256) {
    float s, c;
    sincospif((float)(2 * k) / (float)kFftN, &s, &c);
    tw[k] = make_float2(c, -s);
  }
  // ---- 3a. non-zero run of every mel row (independent of the audio: overlaps the staging loads).  All 17 loads of a row are
  // issued before the first ballot -- with one dependent load per ballot the scan alone took ~100 k cycles per CTA on a cold L2.
  for (int m = warp; m < n_mels; m += 8) {
    const float* row = mel_basis + (long)m * kBins;
    constexpr int kChunks = (kBins + 31) / 32;                   // 17
    float v[kChunks];
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const int k = c * 32 + lane;
      v[c] = (k < kBins) ? __ldg(row + k) : 0.f;
    }
    int lo = kBins, hi = 0;
#pragma unroll
    for (int c = 0; c < kChunks; ++c) {
      const unsigned nz = __ballot_sync(0xffffffffu, v[c] != 0.f);
      if (nz != 0u) {
        lo = min(lo, c * 32 + __ffs(nz) - 1);
        hi = max(hi, c * 32 + 32 - __clz(nz));
      }
    }
    if (lane == 0) { run_lo[m] = (short)min(lo, hi); run_hi[m] = (short)hi; }
  }
  __syncthreads();
  // ---- 2. windowed frames, two per complex FFT, in bit-reversed order
  for (int i = tid; i < kFftN; i += 256) {
    const int r = __brev((unsigned)i) >> (32 - kFftLog);
    const float wi = __ldg(window + i);
#pragma unroll
    for (int j = 0; j < kFramesPerCta / 2; ++j)
      buf[j][r] = make_float2(seg[(2 * j) * kHop + i] * wi, seg[(2 * j + 1) * kHop + i] * wi);
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
#pragma unroll
      for (int q = 0; q < kFramesPerCta / 2; ++q) {
        const float2 a = buf[q][i0], c = buf[q][i1];
        const float2 wc = make_float2(c.x * t.x - c.y * t.y, c.x * t.y + c.y * t.x);
        buf[q][i0] = make_float2(a.x + wc.x, a.y + wc.y);
        buf[q][i1] = make_float2(a.x - wc.x, a.y - wc.y);
      }
    }
    __syncthreads();
  }
  // separate the two real spectra of every FFT and take magnitudes: registers first (the magnitudes overwrite the spectra)
  float m1[kFramesPerCta / 2][3], m2[kFramesPerCta / 2][3];
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int k = tid + it * 256;
    if (k < kBins) {
      const int kn = (kFftN - k) & (kFftN - 1);
#pragma unroll
      for (int q = 0; q < kFramesPerCta / 2; ++q) {
        const float2 z = buf[q][k], zn = buf[q][kn];
        const float ar = 0.5f * (z.x + zn.x), ai = 0.5f * (z.y - zn.y);      // X1 = (Z[k] + conj Z[N-k]) / 2
        const float br = 0.5f * (z.y + zn.y), bi = 0.5f * (zn.x - z.x);      // X2 = (Z[k] - conj Z[N-k]) / 2i
        m1[q][it] = sqrtf(ar * ar + ai * ai);
        m2[q][it] = sqrtf(br * br + bi * bi);
      }
    }
  }
  __syncthreads();
  float* mag = reinterpret_cast<float*>(&buf[0][0]);             // [8][kBins + 3] (row stride 516 floats)
  constexpr int kMagLd = kBins + 3;
#pragma unroll
  for (int it = 0; it < 3; ++it) {
    const int k = tid + it * 256;
    if (k < kBins) {
#pragma unroll
      for (int q = 0; q < kFramesPerCta / 2; ++q) {
        mag[(2 * q) * kMagLd + k] = m1[q][it];
        mag[(2 * q + 1) * kMagLd + k] = m2[q][it];
      }
    }
  }
  __syncthreads();
  // ---- 3b / 4. mel rows over their runs; thread = (mel row, frame), frame fastest: 32 B store segments
  const int nf = min(kFramesPerCta, n_frames - f0);
  for (int o = tid; o < n_mels * kFramesPerCta; o += 256) {
    const int m = o / kFramesPerCta, f = o % kFramesPerCta;
    if (f >= nf) continue;
    const float* row = mel_basis + (long)m * kBins;
    const float* mg = mag + f * kMagLd;
    float acc = 0.f;
    const int k1 = run_hi[m];
    int k = run_lo[m];
    for (; k + 4 <= k1; k += 4) {                                // four independent loads in flight; same summation order
      const float w0 = __ldg(row + k), w1 = __ldg(row + k + 1), w2 = __ldg(row + k + 2), w3 = __ldg(row + k + 3);
      acc = fmaf(w0, mg[k], acc); acc = fmaf(w1, mg[k + 1], acc); acc = fmaf(w2, mg[k + 2], acc); acc = fmaf(w3, mg[k + 3], acc);
    }
    for (; k < k1; ++k) acc = fmaf(__ldg(row + k), mg[k], acc);
    mel[((long)b * n_mels + m) * n_frames + f0 + f] = logf(fmaxf(acc, 1e-5f));
  }
  if (energy != nullptr && warp < nf) {                          // torch.norm(magnitudes, dim=1), stft.py:176
    const float* mg = mag + warp * kMagLd;
    float acc = 0.f;
    for (int k = lane; k < kBins; k += 32) acc = fmaf(mg[k], mg[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) energy[(long)b * n_frames + f0 + warp] = sqrtf(acc);
  }
}

int launch_stft_mel(const float* wav, int B, int S, const float* window, const float* mel_basis, int n_fft, int hop,
                    int n_mels, float* mel, float* energy, cudaStream_t st) {
  DEXB_CHECK(n_fft == kFftN && hop == kHop, "stft_mel is instantiated for n_fft 1024 / hop 256 (got %d / %d)", n_fft, hop);
  DEXB_CHECK(S > n_fft / 2, "stft_mel: reflect padding needs more than n_fft/2 samples (got %d)", S);
  DEXB_CHECK(B >= 1 && n_mels >= 1 && n_mels <= kMaxMels, "stft_mel: need 1 <= n_mels <= %d and a non-empty batch", kMaxMels);
  const int n_frames = S / hop + 1;
  dim3 grid(cdiv(n_frames, kFramesPerCta), B);
  k_stft_mel<<<grid, 256, 0, st>>>(wav, S, window, mel_basis, n_mels, n_frames, mel, energy);
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // namespace dexb
