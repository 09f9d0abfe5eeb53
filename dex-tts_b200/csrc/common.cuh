// Shared device helpers: split-bf16 storage, accurate activations, warp reductions, error plumbing.
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

namespace dexb {

typedef __nv_bfloat16 bf16;

// ---- error plumbing (host) -------------------------------------------------------------------
void set_last_error(const char* fmt, ...);
#define DEXB_CUDA_OK(expr)                                                                        \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      ::dexb::set_last_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return -2;                                                                                  \
    }                                                                                             \
  } while (0)
#define DEXB_CHECK(cond, ...)                  \
  do {                                         \
    if (!(cond)) {                             \
      ::dexb::set_last_error(__VA_ARGS__);     \
      return -1;                               \
    }                                          \
  } while (0)
#define DEXB_TRY(expr)        \
  do {                        \
    int _r = (expr);          \
    if (_r != 0) return _r;   \
  } while (0)

// ---- split-bf16 ("S" tensors) ------------------------------------------------------------------
// A value v is stored as hi = bf16(v), lo = bf16(v - hi): hi + lo carries ~16 mantissa bits, and a GEMM on
// (Ah*Bh + Ah*Bl + Al*Bh) with fp32 accumulation reproduces the fp32 product to ~2^-17 relative.
// An S tensor row holds [ ... hi(C) ... | ... lo(C) ... ]; hi and lo column offsets are explicit so views
// into channel-concatenated buffers work.
__device__ __forceinline__ void split2(float v, bf16& hi, bf16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}
__device__ __forceinline__ float join2(bf16 hi, bf16 lo) { return __bfloat162float(hi) + __bfloat162float(lo); }

// store 8 consecutive values as 8 hi (16 B) + 8 lo (16 B)
__device__ __forceinline__ void store_split8(bf16* hi_ptr, bf16* lo_ptr, const float* v) {
  __align__(16) bf16 h[8];
  __align__(16) bf16 l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) split2(v[i], h[i], l[i]);
  *reinterpret_cast<uint4*>(hi_ptr) = *reinterpret_cast<const uint4*>(h);
  *reinterpret_cast<uint4*>(lo_ptr) = *reinterpret_cast<const uint4*>(l);
}
__device__ __forceinline__ void load_split8(const bf16* hi_ptr, const bf16* lo_ptr, float* v) {
  uint4 hq = *reinterpret_cast<const uint4*>(hi_ptr);
  uint4 lq = *reinterpret_cast<const uint4*>(lo_ptr);
  const bf16* h = reinterpret_cast<const bf16*>(&hq);
  const bf16* l = reinterpret_cast<const bf16*>(&lq);
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = join2(h[i], l[i]);
}

// ---- 256-bit global accesses (sm_100: STG.E.ENL2.256 / LDG.E.ENL2.256): the GEMM epilogue is bound by the number of
// store instructions a warp has in flight, so 32 B per lane halves its cost.  Pointers must be 32 B aligned.
__device__ __forceinline__ void st256_f32(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld256_f32(const float* p, float* v) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p)
               : "memory");
}
// 16 consecutive values -> 16 hi (32 B) + 16 lo (32 B)
__device__ __forceinline__ void store_split16(bf16* hi_ptr, bf16* lo_ptr, const float* v) {
  uint32_t h[8], l[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - __uint_as_float(hb << 16), v[2 * i + 1] - __uint_as_float(hb & 0xffff0000u));
    h[i] = hb;
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(hi_ptr), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]),
               "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7])
               : "memory");
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(lo_ptr), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]),
               "r"(l[4]), "r"(l[5]), "r"(l[6]), "r"(l[7])
               : "memory");
}

// ---- activations (accurate; never compile this project with --use_fast_math) --------------------
// Mish(x) = x * tanh(softplus(x)); with w = e^x, tanh(log(1+w)) = w(w+2) / (w(w+2) + 2): one exp, one divide,
// no cancellation.  torch's softplus switches to identity above 20, where tanh(x) == 1 in fp32 anyway.
__device__ __forceinline__ float mish_f(float x) {
  if (x > 20.f) return x;
  float w = expf(x);
  float n = w * (w + 2.f);
  return x * (n / (n + 2.f));
}
__device__ __forceinline__ float silu_f(float x) { return x / (1.f + expf(-x)); }
__device__ __forceinline__ float gelu_f(float x) { return 0.5f * x * (1.f + erff(x * 0.70710678118654752440f)); }
// Exact-GELU for the GEMM epilogues: erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7; measured 4.7e-7 absolute on the GELU
// over [-8, 8] in fp32, two orders below the split-bf16 noise floor): one MUFU.RCP, one MUFU.EX2 and 8 FMAs, branch-free --
// erff costs ~3x the instructions and the fc1 epilogue is instruction-bound.
__device__ __forceinline__ float gelu_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752440f;
  float t;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, z, 1.f)));
  float p = 1.061405429f;
  p = fmaf(p, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  p *= t;
  float e;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(z * z * -1.4426950408889634f));
  const float erf_abs = fmaf(-p, e, 1.f);
  return 0.5f * x * (1.f + copysignf(erf_abs, x));
}

// ---- warp reductions ---------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

static inline int cdiv(long a, long b) { return (int)((a + b - 1) / b); }

// ---- programmatic dependent launch ---------------------------------------------------------------
// The trajectory is one CUDA graph of ~4 600 kernels; each boundary costs a grid drain + launch.  Kernels launched through
// launch_pdl carry the programmatic-stream-serialization attribute (captured as a programmatic edge): their blocks may start while
// the previous kernel drains, run their prologue (barrier init, TMEM allocation, descriptor prefetch) and then block in pdl_wait()
// until the previous grid has completed and its writes are visible.  EVERY kernel launched this way must call pdl_wait() before it
// touches global memory.  DEXB_PDL=0 launches them as ordinary kernels (pdl_wait is then a no-op).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// (An early `griddepcontrol.launch_dependents` at kernel entry was measured slower: 174.1 vs 170.0 ms per trajectory -- the waiting
// blocks of the next kernel take resources from the running one -- so dependents launch when the blocks of the primary exit.)
bool pdl_enabled();
template <class... KArgs, class... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace dexb
