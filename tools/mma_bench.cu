// Micro-benchmark of tcgen05.mma on B200 (tuning aid, not part of the library):
//   1. cycles per MMA for M = 128, K = 16 bf16, N in {16..256}, A from shared memory (SS) or tensor memory (TS);
//   2. correctness of a K-major SWIZZLE_128B A descriptor whose start address is shifted by whole 128 B rows
//      (not 1024 B aligned) -- the "halo" trick a 3x3 convolution needs to reuse one staged row slab for its dx taps.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dex-tts_b200/csrc tools/mma_bench.cu -o gpurun_out/mma_bench
#include <cuda_bf16.h>
#include <stdio.h>

#include "ptx.cuh"

using namespace dexb;

__global__ void __launch_bounds__(128, 1) k_time(int n, int ts_mode, int reps, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 64 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, n);
    const uint32_t a = ptx::smem_u32(smem), b = a + 16384;
    long long best = 1ll << 60;
    for (int r = 0; r < 5; ++r) {
      __syncwarp();
      long long t0 = 0, t1 = 0;
      if (ptx::elect_one()) {
        t0 = clock64();
        for (int i = 0; i < reps; ++i) {
          const uint32_t ko = (i & 3) * 32;
          if (ts_mode) ptx::mma_bf16_ts(tmem, tmem + 256 + (i & 3) * 8, ptx::make_desc_k128(b + ko), idesc, i ? 1u : 0u);
          else ptx::mma_bf16_ss(tmem, ptx::make_desc_k128(a + ko), ptx::make_desc_k128(b + ko), idesc, i ? 1u : 0u);
        }
        ptx::mma_commit(&bar);
      }
      __syncwarp();
      ptx::mbar_wait(&bar, r & 1);
      t1 = clock64();
      long long mx = t0;
      for (int o = 16; o > 0; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, mx, o); mx = v > mx ? v : mx; }
      if (t1 - mx < best) best = t1 - mx;
    }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = best;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

// The conv / linear kernels' exact per-stage pattern: 4 x { MMA(A_hi, [B_hi|B_lo], N = 2*bn) ; MMA(A_lo, B_hi, N = bn) } then one
// tcgen05.commit, `stages` times, alternating between two TMEM accumulators every 9 stages.  mode 1 adds nothing else; it shows
// whether the tensor pipe itself sustains N/2 + 43 cycles per MMA on this mix.
__global__ void __launch_bounds__(128, 1) k_pattern(int bn, int stages, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[4];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { for (int i = 0; i < 4; ++i) ptx::mbar_init(&bar[i], 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    const uint32_t idesc = ptx::make_idesc_bf16(128, bn), idesc2 = ptx::make_idesc_bf16(128, 2 * bn);
    long long t0 = 0, t1 = 0;
    __syncwarp();
    if (ptx::elect_one()) {
      t0 = clock64();
      for (int st = 0; st < stages; ++st) {
        const uint32_t base = ptx::smem_u32(smem) + (st % 3) * 65536;
        const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768;
        const uint32_t tacc = tmem + ((st / 9) & 1) * 256;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ko = kk * 32;
          ptx::mma_bf16_ss(tacc, ptx::make_desc_k128(a_hi + ko), ptx::make_desc_k128(b_hi + ko), idesc2, (st % 9 || kk) ? 1u : 0u);
          ptx::mma_bf16_ss(tacc, ptx::make_desc_k128(a_lo + ko), ptx::make_desc_k128(b_hi + ko), idesc, 1u);
        }
        ptx::mma_commit(&bar[1 + (st % 3)]);
      }
      ptx::mma_commit(&bar[0]);
    }
    __syncwarp();
    ptx::mbar_wait(&bar[0], 0);                              // only the final commit arrives on bar[0]
    t1 = clock64();
    long long mx = t0;
    for (int o = 16; o > 0; o >>= 1) { long long v = __shfl_xor_sync(0xffffffffu, mx, o); mx = v > mx ? v : mx; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - mx;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

// The same MMA pattern behind a producer / consumer mbarrier ring of depth D (no TMA: the producer thread only waits for `empty`
// and arrives on `full`): what does the barrier round trip (tcgen05.commit -> empty -> producer -> full -> issuer) cost?
__global__ void __launch_bounds__(128, 1) k_ring(int bn, int stages, int depth, int flags, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8], done;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&done, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 1) {                                           // producer
    if (ptx::elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int st = 0; st < stages && !(flags & 2); ++st) {
        if (!(flags & 4)) ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_arrive(&full[s]);
        if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 0) {                                    // MMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, bn), idesc2 = ptx::make_idesc_bf16(128, 2 * bn);
      uint32_t s = 0, ph = 0;
      for (int st = 0; st < stages; ++st) {
        if (flags & 8) {                                     // mbarrier.test_wait (no suspend) instead of try_wait
          uint32_t ok = 0;
          while (!ok) {
            asm volatile("{\n\t.reg .pred P;\n\tmbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
                         : "=r"(ok) : "r"(ptx::smem_u32(&full[s])), "r"(ph) : "memory");
          }
        } else if (!(flags & 2)) ptx::mbar_wait(&full[s], ph);
        if (!(flags & 1)) ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem) + (st % 3) * 65536;
        const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768;
        const uint32_t tacc = tmem + ((st / 9) & 1) * 256;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ko = kk * 32;
          ptx::mma_bf16_ss(tacc, ptx::make_desc_k128(a_hi + ko), ptx::make_desc_k128(b_hi + ko), idesc2, (st % 9 || kk) ? 1u : 0u);
          ptx::mma_bf16_ss(tacc, ptx::make_desc_k128(a_lo + ko), ptx::make_desc_k128(b_hi + ko), idesc, 1u);
        }
        if (!(flags & 4)) ptx::mma_commit(&empty[s]);
        if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
      }
      ptx::mma_commit(&done);
    }
    __syncwarp();
    ptx::mbar_wait(&done, 0);
    long long t1 = clock64();
    if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

// Two MMA-issuing warps, each with its own accumulator and its own mbarrier ring (a third / fourth warp flips the barriers):
// does the tensor pipe interleave the two streams so that one issuer's barrier handshake hides behind the other's MMAs?
__global__ void __launch_bounds__(128, 1) k_ring2(int bn, int stages, int depth, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[2][8], empty[2][8], done[2];
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int w = 0; w < 2; ++w) {
      for (int i = 0; i < 8; ++i) { ptx::mbar_init(&full[w][i], 1); ptx::mbar_init(&empty[w][i], 1); }
      ptx::mbar_init(&done[w], 1);
    }
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp >= 2) {                                           // producers: warp 2 feeds issuer 0, warp 3 issuer 1
    const int w = warp - 2;
    if (ptx::elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int st = 0; st < stages; ++st) {
        ptx::mbar_wait(&empty[w][s], ph ^ 1);
        ptx::mbar_arrive(&full[w][s]);
        if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
      }
    }
  } else {                                                   // issuers
    const int w = warp;
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, bn), idesc2 = ptx::make_idesc_bf16(128, 2 * bn);
      uint32_t s = 0, ph = 0;
      for (int st = 0; st < stages; ++st) {
        ptx::mbar_wait(&full[w][s], ph);
        ptx::tc_fence_after();
        const uint32_t base = ptx::smem_u32(smem) + ((2 * st + w) % 3) * 65536;
        const uint32_t a_hi = base, a_lo = base + 16384, b_hi = base + 32768;
        const uint32_t tacc = tmem + w * 256;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ko = kk * 32;
          ptx::mma_bf16_ss(tacc, ptx::make_desc_k128(a_hi + ko), ptx::make_desc_k128(b_hi + ko), idesc2, (st % 9 || kk) ? 1u : 0u);
          ptx::mma_bf16_ss(tacc, ptx::make_desc_k128(a_lo + ko), ptx::make_desc_k128(b_hi + ko), idesc, 1u);
        }
        ptx::mma_commit(&empty[w][s]);
        if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
      }
      ptx::mma_commit(&done[w]);
    }
    __syncwarp();
    ptx::mbar_wait(&done[w], 0);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0 && blockIdx.x == 0) out[w] = t1 - t0;
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<512>(tmem);
}

// A[r][k] = (k == 0) ? r + 1 : 0 for r in [0, 144), written with the TMA 128B swizzle; B[n][k] = (n == 0 && k == 0).
// D = A_shift B^T  -> D[m][0] must be m + shift + 1.
__global__ void __launch_bounds__(128, 1) k_shift(int shift, int use_base_offset, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  __nv_bfloat16* A = reinterpret_cast<__nv_bfloat16*>(smem);
  __nv_bfloat16* Bm = reinterpret_cast<__nv_bfloat16*>(smem + 32768);
  for (int r = threadIdx.x; r < 144; r += 128) {
    // element (r, k = 0): 16-byte chunk 0 of row r lands at chunk (0 ^ (r & 7))
    A[(r * 128 + ((0 ^ (r & 7)) << 4)) / 2] = __float2bfloat16((float)(r + 1));
  }
  if (threadIdx.x == 0) Bm[0] = __float2bfloat16(1.f);
  if (threadIdx.x == 0) { ptx::mbar_init(&bar, 1); ptx::fence_barrier_init(); }
  if (threadIdx.x < 32) ptx::tmem_alloc<32>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  if (threadIdx.x < 32) {
    if (ptx::elect_one()) {
      const uint32_t a = ptx::smem_u32(smem) + shift * 128;
      uint64_t da = ptx::make_desc_k128(a);
      if (use_base_offset) da |= (uint64_t)((a >> 7) & 7) << 49;
      ptx::mma_bf16_ss(tmem, da, ptx::make_desc_k128(ptx::smem_u32(smem + 32768)), ptx::make_idesc_bf16(128, 16), 0u);
      ptx::mma_commit(&bar);
    }
    __syncwarp();
  }
  ptx::mbar_wait(&bar, 0);
  ptx::tc_fence_after();
  float v[32];
  {
    uint32_t* q = reinterpret_cast<uint32_t*>(v);
    const uint32_t ta = tmem + ((uint32_t)((threadIdx.x >> 5) * 32) << 16);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(q[0]), "=r"(q[1]), "=r"(q[2]), "=r"(q[3]), "=r"(q[4]), "=r"(q[5]), "=r"(q[6]), "=r"(q[7])
                 : "r"(ta)
                 : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  out[threadIdx.x] = v[0];
  ptx::tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) ptx::tmem_dealloc<32>(tmem);
}

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k_time, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
  cudaFuncSetAttribute(k_shift, cudaFuncAttributeMaxDynamicSharedMemorySize, 50 * 1024);
  const int reps = 512;
  const int ns[] = {16, 32, 64, 96, 128, 160, 192, 256};
  for (int grid : {1, 148}) {
    for (int mode = 0; mode < 2; ++mode) {
      for (int n : ns) {
        k_time<<<grid, 128, 66 * 1024>>>(n, mode, reps, d);
        long long c = 0;
        cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
        printf("grid %3d  %s  M=128 N=%3d K=16: %7.1f cycles/MMA  (ideal %5.1f)\n", grid, mode ? "TS" : "SS", n, (double)c / reps,
               128.0 * n / 256.0);
      }
    }
  }
  cudaFuncSetAttribute(k_pattern, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  for (int bn : {64, 128}) {
    const int stages = 9 * 32;
    k_pattern<<<148, 128, 194 * 1024>>>(bn, stages, d);
    long long c = 0;
    cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("pattern error %s\n", cudaGetErrorString(e)); return 1; }
    printf("kernel pattern  BLOCK_N=%3d: %7.1f cycles per stage (4 x [N=%d + N=%d]); model 4 x (%d/2+43 + %d/2+43) = %d\n", bn,
           (double)c / stages, 2 * bn, bn, 2 * bn, bn, 4 * (bn + 43 + bn / 2 + 43));
  }
  cudaFuncSetAttribute(k_ring, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  for (int cfgi = 0; cfgi < 9; ++cfgi) {
    const int depths[9] = {1, 2, 4, 8, 4, 4, 4, 4, 4}, flagv[9] = {0, 0, 0, 0, 1, 2, 8, 9, 9};
    const int depth = depths[cfgi], flags = flagv[cfgi];
    const int stages = 9 * 32;
    k_ring<<<148, 128, 194 * 1024>>>(64, stages, depth, flags, d);
    long long c = 0;
    cudaError_t e = cudaMemcpy(&c, d, 8, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("ring error %s\n", cudaGetErrorString(e)); return 1; }
    printf("mbarrier ring depth %d flags %d (1 = no fence, 2 = no full wait, 8 = test_wait), BLOCK_N=64 stage (8 MMAs, 473 "
           "cycles of tensor time): %7.1f cycles per stage\n", depth, flags, (double)c / stages);
  }
  cudaFuncSetAttribute(k_ring2, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  {
    long long* d2;
    cudaMalloc(&d2, 16);
    for (int bn : {64, 128}) {
      const int stages = 9 * 16;
      k_ring2<<<148, 128, 194 * 1024>>>(bn, stages, 2, d2);
      long long c[2] = {0, 0};
      cudaError_t e = cudaMemcpy(c, d2, 16, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("ring2 error %s\n", cudaGetErrorString(e)); return 1; }
      printf("two issuers x mbarrier ring depth 2, BLOCK_N=%d: %7.1f / %7.1f cycles per stage PAIR  (single issuer with ring: %d per stage)\n",
             bn, (double)c[0] / stages, (double)c[1] / stages, bn == 64 ? 726 : 1020);
    }
  }
  float* o;
  cudaMalloc(&o, 128 * 4);
  for (int bo = 0; bo < 2; ++bo)
    for (int shift = 0; shift < 10; ++shift) {
      k_shift<<<1, 128, 50 * 1024>>>(shift, bo, o);
      float h[128];
      cudaError_t e = cudaMemcpy(h, o, sizeof(h), cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("shift %d: error %s\n", shift, cudaGetErrorString(e)); return 1; }
      int bad = 0;
      for (int m = 0; m < 128; ++m) bad += (h[m] != (float)(m + shift + 1));
      printf("row-shifted A descriptor: shift %d rows, base_offset field %s: %s (D[0..3][0] = %g %g %g %g, D[127][0] = %g)\n", shift,
             bo ? "set" : "0", bad ? "MISMATCH" : "ok", h[0], h[1], h[2], h[3], h[127]);
    }
  return 0;
}
