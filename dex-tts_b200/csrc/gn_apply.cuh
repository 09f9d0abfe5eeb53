// GroupNorm-apply + Mish + mask (+ time bias | + residual) -> split-bf16 operand of the next convolution
// (Block / ResnetBlock, DEX-TTS/model/diffusion.py:44-74).  The item code is shared by two callers:
//   * k_gn_apply (kernels_unet.cu): a stand-alone pass, used behind the CUDA-core first convolution and as the fallback;
//   * the epilogue warps of the persistent tcgen05 convolution kernel (gemm.cuh, GNF = true): every CTA applies its share of
//     image i while the tensor pipe works on image i + 2 -- the raw fp32 image is read back from L2 (it was written one or two
//     images ago), so the stand-alone pass and its HBM read disappear.
#pragma once
#include "common.cuh"
#include "gemm_host.cuh"

namespace dexb {

// per-step scalars of the EDM sampler / preconditioner (DEX-TTS/model/edm.py:88-98,185-203)
struct StepScalars {
  float sigma, sigma_next, c_skip, c_out, c_in, c_noise;
};

struct SView {            // a split-bf16 tensor view: element (row, c) hi at p[row*stride + hi + c], lo at ... + lo
  bf16* p;
  long stride;
  int hi, lo;
};

struct GnApplyArgs {
  const float* raw; int C; int G;          // F[M][C]
  const double* stats;                      // [B][kGnRep][G][2]
  const float* gamma; const float* beta;
  int B, P, W;                              // P pixels per image, W image width (mask column = pixel % W)
  const float* mask; long mask_stride;      // [B][W]
  const float* tbias;                       // [C] added after Mish*mask (then masked again), or null
  SView resid_s;                            // identity residual (already masked) or p == null
  const float* resid_f; long resid_f_stride;   // fp32 residual (res_conv output incl. bias), masked on the fly
  // residual computed from the network input (first ResnetBlock, res_conv 1x1 on 2 channels)
  const float* rin_w; const float* rin_b;   // [C][2], [C] or null
  const float* x; const float* mu; const StepScalars* tab; int step;
  const float* spk_s; int H;                // third input channel spk_s[b][h] of the multi-speaker GeDEX-TTS (rin_w is [C][3] then)
  SView out;
  int reverse;                              // stand-alone pass: walk the images / chunks backwards (L2 reuse, see k_gn_apply)
};

// GroupNorm-apply fused into the convolution that produces `a.raw` (gemm.cuh, GNF kernels)
struct GnFuse {
  GnApplyArgs a;
  unsigned* done;                           // [B] tiles of image b whose raw rows and statistics are globally visible (zeroed per step)
  int enabled;
  int lag;                                  // images between the tile front and the in-loop apply (tuning: DEXB_GN_LAG, default 2)
  int mode;                                 // tuning aid (DEXB_GN_MODE): 0 = normal, 1 = apply everything after the last tile, 2 = never apply (wrong results)
};

#ifdef __CUDACC__
// packed fp32 pairs (FFMA2 / FMUL2 / FADD2 on sm_100): half the issue slots of the scalar forms for the element-wise chain below
__device__ __forceinline__ uint64_t gp_pack(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void gp_unpack(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t gp_fma(uint64_t a, uint64_t b, uint64_t c) { uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ uint64_t gp_add(uint64_t a, uint64_t b) { uint64_t r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
__device__ __forceinline__ uint64_t gp_mul(uint64_t a, uint64_t b) { uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r; }
// Mish of a pair: the same operations as mish_fast, two lanes per instruction except the clamp and the two MUFU evaluations
__device__ __forceinline__ uint64_t mish_fast2(uint64_t y2) {
  float y0, y1, w0, w1, r0, r1, d0, d1;
  gp_unpack(y2, y0, y1);
  const uint64_t e2 = gp_mul(gp_pack(fminf(y0, 20.f), fminf(y1, 20.f)), gp_pack(1.4426950408889634f, 1.4426950408889634f));
  float e0, e1;
  gp_unpack(e2, e0, e1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w0) : "f"(e0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w1) : "f"(e1));
  const uint64_t two2 = gp_pack(2.f, 2.f);
  const uint64_t w2 = gp_pack(w0, w1);
  const uint64_t n2 = gp_mul(w2, gp_add(w2, two2));
  gp_unpack(gp_add(n2, two2), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return gp_mul(y2, gp_mul(n2, gp_pack(r0, r1)));
}

// Mish, branch-free on the raw MUFU approximations: w = 2^(min(x, 20) log2 e), n = w (w + 2), mish = x n / (n + 2) -- 9 instructions.
// (__expf / __fdividef carry denormal range checks and the x > 20 early-out was a divergent branch: 19 instructions and two
//  BSSY / BSYNC pairs per element; the stand-alone pass issued 49 instructions per element and was issue-bound, not HBM-bound.)
// Above 20 the clamp gives n / (n + 2) == 1 in fp32, i.e. x itself, which is what torch's softplus threshold does.
__device__ __forceinline__ float mish_fast(float x) {
  float w, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(w) : "f"(fminf(x, 20.f) * 1.4426950408889634f));
  const float n = w * (w + 2.f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(n + 2.f));
  return x * (n * r);
}

// mean / rstd of one (image, group) straight from the double sums, per thread: two broadcast loads and a handful of DP
// instructions -- no shared memory, no block barrier (a per-block prologue with __syncthreads was the top stall reason of the
// GroupNorm-apply kernels: 2-2.8 stalled warps per issue slot, profiles/r01_ncu_small_kernels.md).
// L2 = true: the sums were accumulated by other CTAs of the SAME kernel -> read them from L2 (ld.global.cg), never from L1.
template <bool L2 = false>
__device__ __forceinline__ void gn_thread_stats(const double* __restrict__ stats, int G, double inv_n, int b, int g, float& mean,
                                                float& rstd) {
  const double2* sp = reinterpret_cast<const double2*>(stats + ((long)b * kGnRep * G + g) * 2);
  double2 s = make_double2(0., 0.);
#pragma unroll
  for (int r = 0; r < kGnRep; ++r) {                         // the writers spread their atomics over kGnRep replicas (gemm_host.cuh)
    const double2 t = L2 ? __ldcg(sp + (long)r * G) : sp[(long)r * G];
    s.x += t.x; s.y += t.y;
  }
  const double mean_d = s.x * inv_n;
  double var_d = s.y * inv_n - mean_d * mean_d;
  if (var_d < 0.) var_d = 0.;
  mean = (float)mean_d;
  rstd = 1.f / sqrtf((float)(var_d + 1e-5));                // sums and the variance in double, only the root in fp32
}

struct GnItem {
  float4 r0, r1;                                             // 8 raw channels
  uint4 q0, q1;                                              // residual: (hi, lo) of an S row or two float4 of an F row
};

// Residual variant of a launch, compiled in (KIND >= 0) or looked up per item (KIND = kGnAny: the fused path, one kernel for all):
constexpr int kGnAny = -1, kGnPlain = 0, kGnResS = 1, kGnResF = 2, kGnRin = 3;
__host__ __device__ inline int gn_kind_of(const GnApplyArgs& a) {
  return a.resid_s.p != nullptr ? kGnResS : (a.resid_f != nullptr ? kGnResF : (a.rin_w != nullptr ? kGnRin : kGnPlain));
}

// all global loads of one item (pixel `pix` of the whole batch, channels c0 .. c0 + 7); L2: see gn_thread_stats
template <bool L2, int KIND = kGnAny>
__device__ __forceinline__ void gn_item_load(const GnApplyArgs& a, long pix, int c0, GnItem& it) {
  const float4* rp = reinterpret_cast<const float4*>(a.raw + pix * a.C + c0);
  if (L2) { it.r0 = __ldcg(rp); it.r1 = __ldcg(rp + 1); }
  else { it.r0 = rp[0]; it.r1 = rp[1]; }
  it.q0 = make_uint4(0, 0, 0, 0); it.q1 = make_uint4(0, 0, 0, 0);
  if (KIND == kGnResS || (KIND == kGnAny && a.resid_s.p != nullptr)) {
    const bf16* q = a.resid_s.p + pix * a.resid_s.stride + c0;
    it.q0 = *reinterpret_cast<const uint4*>(q + a.resid_s.hi);
    it.q1 = *reinterpret_cast<const uint4*>(q + a.resid_s.lo);
  } else if (KIND == kGnResF || (KIND == kGnAny && a.resid_f != nullptr)) {
    const float* q = a.resid_f + pix * a.resid_f_stride + c0;
    it.q0 = *reinterpret_cast<const uint4*>(q);
    it.q1 = *reinterpret_cast<const uint4*>(q + 4);
  }
}

// normalise, Mish, mask, (+ time bias), + residual, split store.  p = pixel inside image b; ga[] = rstd * gamma.
// w = image column of the pixel (p % W), which the caller tracks incrementally.
template <int KIND = kGnAny>
__device__ __forceinline__ void gn_item_finish(const GnApplyArgs& a, int b, unsigned p, int w, int c0, const GnItem& it, float mean,
                                               const float (&ga)[8], const float (&be)[8], const float (&tb)[8]) {
  const long pix = (long)b * a.P + p;
  const float m = a.mask[(long)b * a.mask_stride + w];
  float v[8] = {it.r0.x, it.r0.y, it.r0.z, it.r0.w, it.r1.x, it.r1.y, it.r1.z, it.r1.w};
  float res[8];
  if (KIND == kGnResS || (KIND == kGnAny && a.resid_s.p != nullptr)) {
    // (hi, lo) pairs of bf16 -> fp32: a bf16 is the upper half of a float
    const uint32_t hq[4] = {it.q0.x, it.q0.y, it.q0.z, it.q0.w}, lq[4] = {it.q1.x, it.q1.y, it.q1.z, it.q1.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      res[2 * i] = __uint_as_float(hq[i] << 16) + __uint_as_float(lq[i] << 16);
      res[2 * i + 1] = __uint_as_float(hq[i] & 0xffff0000u) + __uint_as_float(lq[i] & 0xffff0000u);
    }
  } else if (KIND == kGnResF || (KIND == kGnAny && a.resid_f != nullptr)) {
    const float* f0 = reinterpret_cast<const float*>(&it.q0);
    const float* f1 = reinterpret_cast<const float*>(&it.q1);
#pragma unroll
    for (int i = 0; i < 4; ++i) { res[i] = f0[i] * m; res[4 + i] = f1[i] * m; }
  } else if (KIND == kGnRin || (KIND == kGnAny && a.rin_w != nullptr)) {
    // res_conv(x * mask) of the first ResnetBlock: 1x1 conv on stack[mu, c_in*x]
    const float in0 = a.mu[pix] * m, in1 = (a.tab[a.step].c_in * a.x[pix]) * m;
    if (a.spk_s == nullptr) {
#pragma unroll
      for (int i = 0; i < 8; ++i) res[i] = (a.rin_b[c0 + i] + a.rin_w[(c0 + i) * 2] * in0 + a.rin_w[(c0 + i) * 2 + 1] * in1) * m;
    } else {
      const float in2 = a.spk_s[b * a.H + (int)(p / (unsigned)a.W)] * m;      // speaker channel: constant along time
#pragma unroll
      for (int i = 0; i < 8; ++i)
        res[i] = (a.rin_b[c0 + i] + a.rin_w[(c0 + i) * 3] * in0 + a.rin_w[(c0 + i) * 3 + 1] * in1 + a.rin_w[(c0 + i) * 3 + 2] * in2) * m;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) res[i] = 0.f;
  }
  {
    // ga[] holds rstd * gamma (gn_thread_scale).  m is 0 or 1, so both products below are exact and the two fused multiply-adds
    // round exactly like (mish * m + tb) * m + res.  Pairs of channels share every instruction but the clamp and the MUFUs.
    const uint64_t nm2 = gp_pack(-mean, -mean), m2 = gp_pack(m, m);
#pragma unroll
    for (int i = 0; i < 8; i += 2) {
      const uint64_t y2 = mish_fast2(gp_fma(gp_add(gp_pack(v[i], v[i + 1]), nm2), gp_pack(ga[i], ga[i + 1]), gp_pack(be[i], be[i + 1])));
      gp_unpack(gp_fma(gp_fma(y2, m2, gp_pack(tb[i], tb[i + 1])), m2, gp_pack(res[i], res[i + 1])), v[i], v[i + 1]);
    }
  }
  bf16* op = a.out.p + pix * a.out.stride + c0;
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {                              // packed conversions: one F2FP per pair (store_split16's scheme)
    const __nv_bfloat162 h2 = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    const uint32_t hb = *reinterpret_cast<const uint32_t*>(&h2);
    const __nv_bfloat162 l2 = __floats2bfloat162_rn(v[2 * i] - __uint_as_float(hb << 16), v[2 * i + 1] - __uint_as_float(hb & 0xffff0000u));
    h[i] = hb;
    l[i] = *reinterpret_cast<const uint32_t*>(&l2);
  }
  *reinterpret_cast<uint4*>(op + a.out.hi) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(op + a.out.lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

// rstd folded into gamma once per thread: the per-element normalisation is one subtraction and one fused multiply-add
__device__ __forceinline__ void gn_thread_scale(float rstd, float (&ga)[8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) ga[i] *= rstd;
}

// gamma / beta / time bias of the 8 channels a thread keeps for all its items
__device__ __forceinline__ void gn_thread_affine(const GnApplyArgs& a, int c0, float (&ga)[8], float (&be)[8], float (&tb)[8]) {
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(a.gamma + c0 + 4));
  const float4 b0v = __ldg(reinterpret_cast<const float4*>(a.beta + c0)), b1v = __ldg(reinterpret_cast<const float4*>(a.beta + c0 + 4));
  ga[0] = g0.x; ga[1] = g0.y; ga[2] = g0.z; ga[3] = g0.w; ga[4] = g1.x; ga[5] = g1.y; ga[6] = g1.z; ga[7] = g1.w;
  be[0] = b0v.x; be[1] = b0v.y; be[2] = b0v.z; be[3] = b0v.w; be[4] = b1v.x; be[5] = b1v.y; be[6] = b1v.z; be[7] = b1v.w;
#pragma unroll
  for (int i = 0; i < 8; ++i) tb[i] = 0.f;
  if (a.tbias != nullptr) {
    const float4 t0 = __ldg(reinterpret_cast<const float4*>(a.tbias + c0)), t1 = __ldg(reinterpret_cast<const float4*>(a.tbias + c0 + 4));
    tb[0] = t0.x; tb[1] = t0.y; tb[2] = t0.z; tb[3] = t0.w; tb[4] = t1.x; tb[5] = t1.y; tb[6] = t1.z; tb[7] = t1.w;
  }
}

// Fused path (gemm.cuh, GNF): a thread of the convolution's epilogue warps owns the items gi = first + tid + j * nthr (j = 0 .. J-1,
// gi < last) of every image, where [first, last) is its CTA's share; nthr % (C / 8) == 0, so a thread keeps one channel octet.
// gn_apply_items processes n item rounds starting at gi, four at a time: all four items' loads are in flight together (one L2 round
// trip per group of four).
__device__ __forceinline__ void gn_apply_items(const GnApplyArgs& a, int b, unsigned gi, unsigned last, unsigned nthr, int n) {
  const int cpt = a.C >> 3;                                  // threads per pixel: 8 or 16
  const int cshift = 31 - __clz(cpt);
  const int gs = a.C / a.G;
  const int c0 = (int)(gi & (unsigned)(cpt - 1)) * 8;
  float ga[8], be[8], tb[8];
  gn_thread_affine(a, c0, ga, be, tb);
  float mean, rstd;
  gn_thread_stats<true>(a.stats, a.G, 1.0 / ((double)a.P * gs), b, c0 / gs, mean, rstd);
  gn_thread_scale(rstd, ga);
  const long img_row0 = (long)b * a.P;
#pragma unroll 1
  for (int j = 0; j < n; j += 4, gi += 4u * nthr) {
    GnItem it[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned g = gi + (unsigned)k * nthr;
      if (j + k < n && g < last) gn_item_load<true>(a, img_row0 + (g >> cshift), c0, it[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const unsigned g = gi + (unsigned)k * nthr;
      if (j + k < n && g < last) gn_item_finish(a, b, g >> cshift, (int)((g >> cshift) % (unsigned)a.W), c0, it[k], mean, ga, be, tb);
    }
  }
}
#endif  // __CUDACC__

}  // namespace dexb
