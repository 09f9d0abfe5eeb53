// Micro-benchmark + semantics check of 2-CTA tcgen05 MMAs (cta_group::2) for the split-bf16 stage mix of the GEMM engine (tuning aid, not
// part of the library).  Questions (profiles/r02_issue_bench.md): (1) which rows of B does each CTA of the pair supply, and where does D
// land; (2) does a stage of 4 x [N = 2 BN, N = BN] MMAs retire faster when every CTA reads only half of B from its shared memory?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dex-tts_b200/csrc tools/cta2_bench.cu -o tools/_bin/cta2_bench
#include <cuda_bf16.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "ptx.cuh"

using namespace dexb;
using bf16 = __nv_bfloat16;

namespace p2 {
__device__ __forceinline__ uint32_t cta_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc2(uint32_t* dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(ptx::smem_u32(dst_smem)), "r"(NCOLS) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc2(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(NCOLS) : "memory");
}
__device__ __forceinline__ void mma2_bf16_ss(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void mma2_commit(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(ptx::smem_u32(bar)),
               "h"(mask)
               : "memory");
}
}  // namespace p2

constexpr uint64_t kDescBase = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);

// element (row, k) of a K-major SWIZZLE_128B tile with 64 bf16 per row
__host__ __device__ inline int sw128(int row, int k) { return row * 128 + (((k >> 3) ^ (row & 7)) << 4) + (k & 7) * 2; }

// smem per CTA and slot (64 KB): A_hi 16 KB | A_lo 16 KB | B 32 KB (up to 256 rows)
// mode 0: cta_group::1, every CTA on its own.  mode 1: cta_group::2, the even CTA of a pair issues M = 256 MMAs.
// check: one stage (K = 64) on operands from global memory, D written back (fp32 [cta][128][2 BN]).
template <int mode>
__global__ void __launch_bounds__(128, 1) k_cta2(int bn, int stages, const bf16* __restrict__ a_g, const bf16* __restrict__ b_g,
                                                 float* __restrict__ d_g, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t done;
  __shared__ uint32_t slot;
  const uint32_t rank = mode ? p2::cta_rank() : 0;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  __syncthreads();
  if (a_g != nullptr) {
    // A: this CTA's 128 rows (hi at slot + 0, lo at slot + 16 KB), [cta][2][128][64]; B rows: [cta][rows][64] already in the order the
    // CTA keeps them
    const bf16* a = a_g + (long)blockIdx.x * 2 * 128 * 64;
    for (int i = threadIdx.x; i < 2 * 128 * 64; i += 128) {
      const int part = i / (128 * 64), r = (i / 64) % 128, k = i % 64;
      *reinterpret_cast<bf16*>(smem + part * 16384 + sw128(r, k)) = a[i];
    }
    const bf16* b = b_g + (long)blockIdx.x * 256 * 64;
    for (int i = threadIdx.x; i < 256 * 64; i += 128) *reinterpret_cast<bf16*>(smem + 32768 + sw128(i / 64, i % 64)) = b[i];
  }
  if (threadIdx.x == 0) {
    ptx::mbar_init(&done, 1);
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) {
    if (mode) p2::tmem_alloc2<512>(&slot); else ptx::tmem_alloc<512>(&slot);
  }
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  if (mode) p2::cluster_sync();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  long long t0 = clock64();
  if (warp == 0 && rank == 0) {
    if (ptx::elect_one()) {
      const int M = mode ? 256 : 128;
      const uint32_t idesc = ptx::make_idesc_bf16(M, bn), idesc2 = ptx::make_idesc_bf16(M, 2 * bn);
      for (int st = 0; st < stages; ++st) {
        const uint32_t base = ptx::smem_u32(smem) + (st % 3) * 65536;
        const uint32_t tacc = tmem + ((st / 9) & 1) * 256;
        const bool first = (st % 9) == 0;
        const uint32_t a_hi = base >> 4, a_lo = (base + 16384) >> 4, b_hi = (base + 32768) >> 4;
        // mode 1: the second MMA's B half of this CTA sits behind its half of the stacked tile (rows bn .. bn + bn/2)
        const uint32_t b_2nd = mode ? (base + 32768 + bn * 128) >> 4 : b_hi;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t ko = kk * 2;
          if (mode) {
            p2::mma2_bf16_ss(tacc, kDescBase + (a_hi + ko), kDescBase + (b_hi + ko), idesc2, (!first || kk) ? 1u : 0u);
            p2::mma2_bf16_ss(tacc, kDescBase + (a_lo + ko), kDescBase + (b_2nd + ko), idesc, 1u);
          } else {
            ptx::mma_bf16_ss(tacc, kDescBase + (a_hi + ko), kDescBase + (b_hi + ko), idesc2, (!first || kk) ? 1u : 0u);
            ptx::mma_bf16_ss(tacc, kDescBase + (a_lo + ko), kDescBase + (b_hi + ko), idesc, 1u);
          }
        }
      }
      if (mode) p2::mma2_commit(&done, 3); else ptx::mma_commit(&done);
    }
    __syncwarp();
  }
  ptx::mbar_wait(&done, 0);
  long long t1 = clock64();
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = t1 - t0;
  ptx::tc_fence_after();
  if (d_g != nullptr) {
    float* d = d_g + (long)blockIdx.x * 128 * 256;
    for (int c0 = 0; c0 < 2 * bn; c0 += 32) {
      float v[32];
      ptx::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
      for (int i = 0; i < 32; ++i) d[(long)threadIdx.x * 256 + c0 + i] = v[i];
    }
  }
  ptx::tc_fence_before();
  __syncthreads();
  if (mode) p2::cluster_sync();
  if (threadIdx.x < 32) {
    if (mode) p2::tmem_dealloc2<512>(tmem); else ptx::tmem_dealloc<512>(tmem);
  }
}

static void launch(int ctas, int bn, int stages, int mode, const bf16* a, const bf16* b, float* d, long long* out) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(128);
  cfg.dynamicSmemBytes = 194 * 1024;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = mode ? 2 : 1;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = mode ? 1 : 0;
  cudaError_t e = mode ? cudaLaunchKernelEx(&cfg, k_cta2<1>, bn, stages, a, b, d, out) : cudaLaunchKernelEx(&cfg, k_cta2<0>, bn, stages, a, b, d, out);
  if (e != cudaSuccess) { printf("launch error %s (ctas %d mode %d)\n", cudaGetErrorString(e), ctas, mode); exit(1); }
}

int main() {
  long long* dcyc;
  cudaMalloc(&dcyc, 8);
  cudaFuncSetAttribute(k_cta2<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  cudaFuncSetAttribute(k_cta2<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  // ---- semantics: one pair, one stage, small integers
  for (int bn : {64, 128}) {
    const int N2 = 2 * bn;
    std::vector<bf16> a(2 * 2 * 128 * 64), b(2 * 256 * 64);
    std::vector<float> af(a.size()), bfull_hi(bn * 64), bfull_lo(bn * 64);
    srand(7 + bn);
    for (size_t i = 0; i < a.size(); ++i) { af[i] = (float)(rand() % 5 - 2); a[i] = __float2bfloat16(af[i]); }
    for (int i = 0; i < bn * 64; ++i) { bfull_hi[i] = (float)(rand() % 5 - 2); bfull_lo[i] = (float)(rand() % 7 - 3); }
    // hypothesis: an N-wide 2-CTA MMA takes rows [r N/2, (r+1) N/2) of B from CTA r.  Stacked tile [B_hi; B_lo] (N = 2 bn): CTA 0 holds
    // B_hi, CTA 1 holds B_lo.  Second MMA (A_lo x B_hi, N = bn): CTA 0 uses B_hi[0 : bn/2] and CTA 1 B_hi[bn/2 : bn] -- both placed
    // behind the stacked half (rows bn ...) so that one descriptor serves both CTAs.
    std::vector<float> bf(b.size(), 0.f);
    for (int r = 0; r < bn; ++r)
      for (int k = 0; k < 64; ++k) {
        bf[(0 * 256 + r) * 64 + k] = bfull_hi[r * 64 + k];
        bf[(1 * 256 + r) * 64 + k] = bfull_lo[r * 64 + k];
      }
    for (int r = 0; r < bn / 2; ++r)
      for (int k = 0; k < 64; ++k) {
        bf[(0 * 256 + bn + r) * 64 + k] = bfull_hi[r * 64 + k];
        bf[(1 * 256 + bn + r) * 64 + k] = bfull_hi[(bn / 2 + r) * 64 + k];
      }
    for (size_t i = 0; i < b.size(); ++i) b[i] = __float2bfloat16(bf[i]);
    bf16 *da, *db;
    float* dd;
    cudaMalloc(&da, a.size() * 2); cudaMalloc(&db, b.size() * 2); cudaMalloc(&dd, 2 * 128 * 256 * 4);
    cudaMemcpy(da, a.data(), a.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(db, b.data(), b.size() * 2, cudaMemcpyHostToDevice);
    cudaMemset(dd, 0, 2 * 128 * 256 * 4);
    launch(2, bn, 1, 1, da, db, dd, dcyc);
    std::vector<float> d(2 * 128 * 256);
    cudaError_t e = cudaMemcpy(d.data(), dd, d.size() * 4, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { printf("check error %s\n", cudaGetErrorString(e)); return 1; }
    double worst = 0.;
    for (int c = 0; c < 2; ++c)
      for (int m = 0; m < 128; ++m)
        for (int n = 0; n < N2; ++n) {
          double ref = 0.;
          const float* ahi = &af[((c * 2 + 0) * 128 + m) * 64];
          const float* alo = &af[((c * 2 + 1) * 128 + m) * 64];
          for (int k = 0; k < 64; ++k) {
            if (n < bn) ref += (double)ahi[k] * bfull_hi[n * 64 + k] + (double)alo[k] * bfull_hi[n * 64 + k];
            else ref += (double)ahi[k] * bfull_lo[(n - bn) * 64 + k];
          }
          const double err = fabs(ref - d[((long)c * 128 + m) * 256 + n]);
          if (err > worst) worst = err;
        }
    printf("semantics BLOCK_N=%3d: max |D - ref| = %g %s\n", bn, worst, worst == 0. ? "(hypothesis holds)" : "(MISMATCH)");
    cudaFree(da); cudaFree(db); cudaFree(dd);
  }
  // ---- timing
  {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148); cfg.blockDim = dim3(128); cfg.dynamicSmemBytes = 194 * 1024;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_cta2<1>, &cfg);
    printf("cudaOccupancyMaxActiveClusters(cluster 2, 194 KB): %d (%s)\n", nclusters, cudaGetErrorString(e));
  }
  for (int bn : {64, 128}) {
    for (int mode = 0; mode < 2; ++mode) {
      const int stages = 9 * 32;
      launch(148, bn, stages, mode, nullptr, nullptr, nullptr, dcyc);
      long long cyc = 0;
      cudaError_t e = cudaMemcpy(&cyc, dcyc, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s (mode %d)\n", cudaGetErrorString(e), mode); return 1; }
      printf("BLOCK_N=%3d  cta_group::%d  %7.1f cycles per stage of 8 MMAs (128 rows per SM)\n", bn, mode + 1, (double)cyc / stages);
    }
  }
  return 0;
}
