// Monotonic Alignment Search on the GPU (training side) -- replaces the reference's only native component,
// the Cython kernel maximum_path_c / maximum_path_each (DEX-TTS/model/monotonic_align/core.pyx:9-47) behind
// monotonic_align.maximum_path (DEX-TTS/model/monotonic_align/__init__.py:8-25; call site DEX-TTS/model/tts.py:108):
//
//   v = value * mask;  t_x = sum_x mask[:, x, 0];  t_y = sum_y mask[:, 0, y]
//   forward, column by column:  v[x, y] += max(x == y ? -1e9 : v[x, y-1],  x == 0 ? (y == 0 ? 0 : -1e9) : v[x-1, y-1])   on the band
//                               max(0, t_x + y - t_y) <= x < min(t_x, y + 1)
//   backtrack from (t_x - 1, t_y - 1):  path[index, y] = 1;  index -= 1  if index != 0 and (index == y or v[index, y-1] < v[index-1, y-1])
//
// One CTA per utterance (upstream: one OpenMP thread per utterance), threads over the tokens x, one barrier per frame y: the previous
// column lives in shared memory (every value a column reads lies inside the previous column's band), and the backtrack's comparison
// is evaluated in the forward pass on the very same two floats and kept as one byte per (y, x) -- so the walk back (thread 0) is one
// dependent load per frame instead of two, and bit-identical to upstream's.  Integer / fp32 add-compare work, latency-bound: Ty
// barriers of a 32 ... 1024-thread CTA (about 1 us each) plus Ty dependent loads; the inputs are staged 32 frames at a time (below).
// Inputs with t_y < t_x (more tokens than frames) are undefined upstream (out-of-bounds reads); here they give a path that is
// monotonic but not meaningful, and never an invalid access.
#include <stdint.h>

#include "../../include/dexb200.h"
#include "common.cuh"

namespace dexb {

constexpr float kMasNeg = -1e9f;

// YB frames of value * mask are staged through shared memory per round: the DP walks the frames (y) with the tokens (x) across the
// lanes, but both inputs are (Tx, Ty) row-major, i.e. a lane-per-token read touches one 32 B sector per lane and frame.  A warp
// therefore loads YB consecutive frames of one token row (one or four full sectors), rows round-robin over the warps, into a padded
// tile [band row][YB + 1] that the DP then reads conflict-free.
template <int YB>
__global__ void k_mas(const float* __restrict__ value, const float* __restrict__ mask, unsigned char* __restrict__ dec,
                      float* __restrict__ path, int Tx, int Ty) {
  extern __shared__ float s_col[];                    // [2][Tx]: DP values of the previous / current column, then the tile [Tx][YB + 1]
  __shared__ int s_tx, s_ty;
  const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
  const float* vb = value + (long)b * Tx * Ty;
  const float* mb = mask + (long)b * Tx * Ty;
  unsigned char* db = dec + (long)b * Tx * Ty;        // [Ty][Tx]
  if (tid == 0) { s_tx = 0; s_ty = 0; }
  __syncthreads();
  int cx = 0, cy = 0;
  for (int x = tid; x < Tx; x += nt) cx += mb[(long)x * Ty] != 0.f;
  for (int y = tid; y < Ty; y += nt) cy += mb[y] != 0.f;
  if (cx) atomicAdd(&s_tx, cx);
  if (cy) atomicAdd(&s_ty, cy);
  __syncthreads();
  const int t_x = s_tx, t_y = s_ty;
  if (t_x <= 0 || t_y <= 0) return;                   // empty utterance: all-zero path (upstream would index path[-1])
  float* prev = s_col;
  float* cur = s_col + Tx;
  float* tile = s_col + 2 * Tx;
  const int lane = tid & 31, warp = tid >> 5, nw = nt >> 5;
  for (int y0 = 0; y0 < t_y; y0 += YB) {
    const int y1 = min(t_y, y0 + YB);
    const int lo0 = max(0, t_x + y0 - t_y), hi1 = min(t_x, y1);      // union of the bands of frames y0 .. y1 - 1
    for (int x = lo0 + warp; x < hi1; x += nw)
      for (int yy = lane; yy < YB; yy += 32)
        if (y0 + yy < y1) tile[(x - lo0) * (YB + 1) + yy] = vb[(long)x * Ty + y0 + yy] * mb[(long)x * Ty + y0 + yy];
    __syncthreads();
    for (int y = y0; y < y1; ++y) {
      const int lo = max(0, t_x + y - t_y), hi = min(t_x, y + 1);
      for (int x = lo + tid; x < hi; x += nt) {
        const float raw = tile[(x - lo0) * (YB + 1) + (y - y0)];
        const float v_cur = (x == y) ? kMasNeg : prev[x];
        const float v_prev = (x == 0) ? (y == 0 ? 0.f : kMasNeg) : prev[x - 1];
        cur[x] = fmaxf(v_cur, v_prev) + raw;
        db[(long)y * Tx + x] = (x == y || v_cur < v_prev) ? 1 : 0;      // upstream's backtrack test for (index = x, frame y)
      }
      __syncthreads();
      float* t = prev; prev = cur; cur = t;
    }
  }
  if (tid == 0) {
    int index = t_x - 1;
    float* pb = path + (long)b * Tx * Ty;
    for (int y = t_y - 1; y >= 0; --y) {
      pb[(long)index * Ty + y] = 1.f;
      if (index != 0 && y > 0 && db[(long)y * Tx + index]) --index;
    }
  }
}

}  // namespace dexb

using namespace dexb;

extern "C" {

int dexb_mas_maximum_path(const float* value_dev, const float* mask_dev, int B, int Tx, int Ty, uint8_t* scratch_dev, float* path_dev,
                          void* stream) {
  DEXB_CHECK(value_dev != nullptr && mask_dev != nullptr && scratch_dev != nullptr && path_dev != nullptr,
             "dexb_mas_maximum_path: null argument");
  DEXB_CHECK(B >= 1 && Tx >= 1 && Ty >= 1 && Tx <= 4096, "dexb_mas_maximum_path: B = %d, Tx = %d (<= 4096), Ty = %d", B, Tx, Ty);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t n = (size_t)B * Tx * Ty;
  DEXB_CUDA_OK(cudaMemsetAsync(path_dev, 0, n * sizeof(float), st));
  DEXB_CUDA_OK(cudaMemsetAsync(scratch_dev, 0, n, st));
  int threads = (Tx + 31) / 32 * 32;
  if (threads > 1024) threads = 1024;
  // 32 frames per staging round while the tile fits the 227 KB carve-out (Tx <= 1536), 8 beyond
  if (Tx <= 1536) {
    const size_t smem = (2 + 33) * (size_t)Tx * sizeof(float);
    DEXB_CUDA_OK(cudaFuncSetAttribute(k_mas<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mas<32><<<B, threads, smem, st>>>(value_dev, mask_dev, scratch_dev, path_dev, Tx, Ty);
  } else {
    const size_t smem = (2 + 9) * (size_t)Tx * sizeof(float);
    DEXB_CUDA_OK(cudaFuncSetAttribute(k_mas<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_mas<8><<<B, threads, smem, st>>>(value_dev, mask_dev, scratch_dev, path_dev, Tx, Ty);
  }
  DEXB_CUDA_OK(cudaGetLastError());
  return 0;
}

}  // extern "C"
