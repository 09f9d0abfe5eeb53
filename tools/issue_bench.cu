// Micro-benchmark of the single-thread MMA issue loop behind an mbarrier ring (tuning aid, not part of the library).
// Question (profiles/r01_ncu_gemm_epilogue.md): the kernel's stage of 8 MMAs retires in 473 cycles without barriers but ~726 behind a
// full/empty ring -- which part of the handshake costs, and can it be hidden by (1) waiting for the NEXT stage's `full` barrier in
// the middle of the current stage's MMAs, (2) larger stages (one barrier round per G k-chunks), (3) a watcher warp that turns the
// mbarrier completion into a plain shared-memory flag?
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I dex-tts_b200/csrc tools/issue_bench.cu -o tools/_bin/issue_bench
#include <cuda_bf16.h>
#include <stdio.h>

#include "ptx.cuh"

using namespace dexb;

__device__ __forceinline__ void mbar_wait_relaxed(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0, spins = 0;
  while (!ok) {
    asm volatile("{\n\t.reg .pred P;\n\tmbarrier.try_wait.parity.relaxed.cta.shared::cta.b64 P, [%1], %2;\n\tselp.b32 %0, 1, 0, P;\n\t}\n"
                 : "=r"(ok) : "r"(ptx::smem_u32(bar)), "r"(parity) : "memory");
    if (++spins > (1u << 26)) __trap();
  }
}

// descriptors as the kernel forms them: base + (addr >> 4)
template <int K0, int K1>
__device__ __forceinline__ void issue_mmas(uint32_t tacc, uint32_t base, uint32_t idesc, uint32_t idesc2, bool first) {
  constexpr uint64_t kDescBase = ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)2 << 61);
  const uint32_t a_hi = base >> 4, a_lo = (base + 16384) >> 4, b_hi = (base + 32768) >> 4;
#pragma unroll
  for (int kk = K0; kk < K1; ++kk) {
    const uint32_t ko = kk * 2;
    ptx::mma_bf16_ss(tacc, kDescBase + (a_hi + ko), kDescBase + (b_hi + ko), idesc2, (!first || kk) ? 1u : 0u);
    ptx::mma_bf16_ss(tacc, kDescBase + (a_lo + ko), kDescBase + (b_hi + ko), idesc, 1u);
  }
}

// mode 0 baseline | 1 early wait after `arg` k-steps (of 4) | 2 grouped: one barrier round per `arg` stages | 3 watcher warp + flag
// mode 4 no full wait | 6 no barriers at all | 7 = mode 2 with the wait for the NEXT round placed before this round's commit
// mode 8 = mode 2 with mbarrier.try_wait.relaxed.cta | 9 = mode 7 + relaxed
__global__ void __launch_bounds__(128, 1) k_issue(int bn, int stages, int depth, int mode, int arg, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t full[8], empty[8], done;
  __shared__ uint32_t slot;
  __shared__ volatile int flag;
  for (int i = threadIdx.x; i < 192 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 8; ++i) { ptx::mbar_init(&full[i], 1); ptx::mbar_init(&empty[i], 1); }
    ptx::mbar_init(&done, 1);
    flag = 0;
    ptx::fence_barrier_init();
  }
  if (threadIdx.x < 32) ptx::tmem_alloc<512>(&slot);
  ptx::fence_proxy_async_smem();
  ptx::tc_fence_before();
  __syncthreads();
  ptx::tc_fence_after();
  const uint32_t tmem = slot;
  const int warp = threadIdx.x >> 5;
  const int G = (mode == 2 || mode >= 7) ? arg : 1;                       // stages per barrier round
  const int rounds = stages / G;
  long long t0 = clock64();
  if (warp == 1) {                                           // producer: flips barriers only
    if (ptx::elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int r = 0; r < rounds && mode != 6; ++r) {
        ptx::mbar_wait(&empty[s], ph ^ 1);
        ptx::mbar_arrive(&full[s]);
        if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 2 && mode == 3) {                       // watcher: mbarrier completion -> sequence number in shared memory
    if (ptx::elect_one()) {
      uint32_t s = 0, ph = 0;
      for (int r = 0; r < rounds; ++r) {
        ptx::mbar_wait(&full[s], ph);
        flag = r + 1;
        if (++s == (uint32_t)depth) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 0) {                                    // MMA issuer
    if (ptx::elect_one()) {
      const uint32_t idesc = ptx::make_idesc_bf16(128, bn), idesc2 = ptx::make_idesc_bf16(128, 2 * bn);
      uint32_t s = 0, ph = 0;
      bool prewaited = false;
      int st = 0;
      for (int r = 0; r < rounds; ++r) {
        if (mode == 3) { while (flag < r + 1) {} }
        else if (mode != 4 && mode != 6 && !prewaited) { if (mode >= 8) mbar_wait_relaxed(&full[s], ph); else ptx::mbar_wait(&full[s], ph); }
        prewaited = false;
        ptx::tc_fence_after();
        uint32_t sn = s + 1, phn = ph;
        if (sn == (uint32_t)depth) { sn = 0; phn ^= 1; }
        for (int g = 0; g < G; ++g, ++st) {
          const uint32_t base = ptx::smem_u32(smem) + (st % 3) * 65536;
          const uint32_t tacc = tmem + ((st / 9) & 1) * 256;
          const bool first = (st % 9) == 0;
          if (mode == 1) {
            if (arg == 1) issue_mmas<0, 1>(tacc, base, idesc, idesc2, first);
            else if (arg == 2) issue_mmas<0, 2>(tacc, base, idesc, idesc2, first);
            else issue_mmas<0, 3>(tacc, base, idesc, idesc2, first);
            if (r + 1 < rounds) { ptx::mbar_wait(&full[sn], phn); prewaited = true; }
            if (arg == 1) issue_mmas<1, 4>(tacc, base, idesc, idesc2, first);
            else if (arg == 2) issue_mmas<2, 4>(tacc, base, idesc, idesc2, first);
            else issue_mmas<3, 4>(tacc, base, idesc, idesc2, first);
          } else {
            issue_mmas<0, 4>(tacc, base, idesc, idesc2, first);
          }
        }
        if ((mode == 7 || mode == 9) && r + 1 < rounds) {
          if (mode == 9) mbar_wait_relaxed(&full[sn], phn); else ptx::mbar_wait(&full[sn], phn);
          prewaited = true;
        }
        if (mode != 6) ptx::mma_commit(&empty[s]);
        s = sn; ph = phn;
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

int main() {
  long long* d;
  cudaMalloc(&d, 8);
  cudaFuncSetAttribute(k_issue, cudaFuncAttributeMaxDynamicSharedMemorySize, 194 * 1024);
  struct Cfg { int mode, arg; const char* what; };
  const Cfg cfgs[] = {
      {0, 0, "baseline: wait full, 8 MMAs, commit empty"},
      {4, 0, "no wait on full"},
      {6, 0, "no barriers at all (MMAs only)"},
      {1, 1, "early wait for stage s+1 after 2 of 8 MMAs"},
      {1, 2, "early wait for stage s+1 after 4 of 8 MMAs"},
      {1, 3, "early wait for stage s+1 after 6 of 8 MMAs"},
      {2, 2, "one barrier round per 2 stages (16 MMAs)"},
      {2, 3, "one barrier round per 3 stages (24 MMAs)"},
      {2, 9, "one barrier round per 9 stages (72 MMAs)"},
      {3, 0, "watcher warp + shared-memory flag instead of the issuer's mbarrier wait"},
      {7, 1, "wait for the next stage BEFORE this stage's commit"},
      {7, 3, "3 stages per round, wait for the next round before the commit"},
      {8, 1, "try_wait.relaxed.cta"},
      {8, 3, "3 stages per round, try_wait.relaxed.cta"},
      {9, 1, "relaxed wait for the next stage before the commit"},
      {9, 3, "3 stages per round, relaxed wait before the commit"},
  };
  for (int bn : {64, 128}) {
    for (const Cfg& c : cfgs) {
      const int stages = 9 * 32;
      k_issue<<<148, 128, 194 * 1024>>>(bn, stages, 4, c.mode, c.arg, d);
      long long cyc = 0;
      cudaError_t e = cudaMemcpy(&cyc, d, 8, cudaMemcpyDeviceToHost);
      if (e != cudaSuccess) { printf("error %s (mode %d)\n", cudaGetErrorString(e), c.mode); return 1; }
      printf("BLOCK_N=%3d  %-75s %7.1f cycles per stage of 8 MMAs\n", bn, c.what, (double)cyc / stages);
    }
  }
  return 0;
}
