// U-Net side kernels: first conv, GroupNorm-apply/Mish fusions, EDM update, LinearAttention reductions.
// Reference semantics: DEX-TTS/model/diffusion.py:44-105,190-236 and DEX-TTS/model/edm.py:88-98,185-203.
#include "kernels.cuh"

namespace dexb {

// ------------------------------------------------------------------------------------------------
__global__ void k_fill_zero(uint4* p, size_t n16) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
}
void launch_fill_zero(void* p, size_t bytes, cudaStream_t st) {
  const size_t n16 = bytes / 16;                       // callers keep scratch regions 16 B granular
  int blocks = (int)((n16 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  k_fill_zero<<<blocks, 256, 0, st>>>(reinterpret_cast<uint4*>(p), n16);
}

__global__ void k_scale(float* x, long n, float s) {
  long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i < n) x[i] *= s;
}
void launch_scale(float* x, long n, float s, cudaStream_t st) { k_scale<<<cdiv(n, 256), 256, 0, st>>>(x, n, s); }

__global__ void k_mask_down(const float* mask, float* mask1, int B, int T, int W1) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B * W1) { int b = i / W1, w = i % W1; mask1[i] = mask[(long)b * T + 2 * w]; }
}
void launch_mask_down(const float* mask, float* mask1, int B, int T, int W1, cudaStream_t st) {
  k_mask_down<<<cdiv((long)B * W1, 256), 256, 0, st>>>(mask, mask1, B, T, W1);
}

// ------------------------------------------------------------------------------------------------
// conv_in: one thread per pixel computes all C(=64) outputs of the 2->C 3x3 conv (K = 18: CUDA cores, the
// contraction is too short for tensor cores).  Input = stack[mu, c_in * x] * mask  (diffusion.py:198,52; edm.py:96).
// ------------------------------------------------------------------------------------------------
template <int C>
__global__ void __launch_bounds__(128) k_conv_in(const float* __restrict__ x, const float* __restrict__ mu,
                                                 const float* __restrict__ mask, const StepScalars* __restrict__ tab,
                                                 int step, const float* __restrict__ w, const float* __restrict__ bias,
                                                 float* __restrict__ raw, double* __restrict__ stats, int B, int H,
                                                 int W) {
  __shared__ float ws[C * 18];
  __shared__ float bs[C];
  __shared__ float red[4][C / 8][2];
  for (int i = threadIdx.x; i < C * 18; i += 128) ws[i] = w[i];
  for (int i = threadIdx.x; i < C; i += 128) bs[i] = bias[i];
  __syncthreads();
  const int wt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int wcol = wt * 128 + threadIdx.x;
  const bool valid = wcol < W;
  const float c_in = tab[step].c_in;
  float in[18];
#pragma unroll
  for (int dy = 0; dy < 3; ++dy)
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int hh = h + dy - 1, ww = wcol + dx - 1;
      float a = 0.f, c = 0.f;
      if (valid && hh >= 0 && hh < H && ww >= 0 && ww < W) {
        const float m = mask[(long)b * W + ww];
        const long idx = ((long)b * H + hh) * W + ww;
        a = mu[idx] * m;
        c = (c_in * x[idx]) * m;
      }
      in[dy * 3 + dx] = a;
      in[9 + dy * 3 + dx] = c;
    }
  float* orow = raw + (((long)b * H + h) * W + wcol) * C;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll 1
  for (int g = 0; g < C / 8; ++g) {
    float o[8];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int co = g * 8 + j;
      float acc = bs[co];
#pragma unroll
      for (int k = 0; k < 18; ++k) acc = fmaf(ws[co * 18 + k], in[k], acc);
      o[j] = acc;
      if (valid) { s += acc; ss += acc * acc; }
    }
    if (valid) {
      *reinterpret_cast<float4*>(orow + g * 8) = make_float4(o[0], o[1], o[2], o[3]);
      *reinterpret_cast<float4*>(orow + g * 8 + 4) = make_float4(o[4], o[5], o[6], o[7]);
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    if (lane == 0) { red[warp][g][0] = s; red[warp][g][1] = ss; }
  }
  __syncthreads();
  if (threadIdx.x < C / 8) {
    const int g = threadIdx.x;               // channels-per-group == 8 when C == 64 (GroupNorm(8, 64))
    double s = 0., ss = 0.;
    for (int wq = 0; wq < 4; ++wq) { s += red[wq][g][0]; ss += red[wq][g][1]; }
    atomicAdd(&stats[((long)b * (C / 8) + g) * 2], s);
    atomicAdd(&stats[((long)b * (C / 8) + g) * 2 + 1], ss);
  }
}

void launch_conv_in(const float* x, const float* mu, const float* mask, const StepScalars* tab, int step,
                    const float* w, const float* bias, float* raw, double* stats, int B, int H, int W, int C,
                    cudaStream_t st) {
  dim3 grid(cdiv(W, 128), H, B);
  // stats layout is [B][8 groups][2]; with C == 64 a group is 8 channels, with C == 128 it is 16 (two 8-chunks):
  // the C==128 instantiation folds pairs of chunks on the host side by giving G = C/8 "half groups" -- not needed
  // for the shipped configs (dim 64), so only C == 64 is instantiated.
  if (C == 64) k_conv_in<64><<<grid, 128, 0, st>>>(x, mu, mask, tab, step, w, bias, raw, stats, B, H, W);
}

// ------------------------------------------------------------------------------------------------
// GroupNorm-apply + Mish + mask (+ time bias | + residual) -> split-bf16.  8 channels per thread.
// ------------------------------------------------------------------------------------------------
// mean / rstd of the (at most two) images a 256-thread block touches, computed once per block from the double sums
__device__ __forceinline__ void gn_block_stats(const double* __restrict__ stats, int B, int G, double n, int b0,
                                               float (*s_mean)[8], float (*s_rstd)[8]) {
  if (threadIdx.x < 2 * G) {
    const int bi = threadIdx.x / G, g = threadIdx.x % G;
    const int b = b0 + bi;
    if (b < B) {
      const double s = stats[((long)b * G + g) * 2], ss = stats[((long)b * G + g) * 2 + 1];
      const double mean_d = s / n;
      double var_d = ss / n - mean_d * mean_d;
      if (var_d < 0.) var_d = 0.;
      s_mean[bi][g] = (float)mean_d;
      s_rstd[bi][g] = (float)(1.0 / sqrt(var_d + 1e-5));
    }
  }
  __syncthreads();
}
// Mish with fast intrinsics (ex2.approx / approximate divide: ~1e-6 relative, far inside the split-bf16 noise floor)
__device__ __forceinline__ float mish_fast(float x) {
  if (x > 20.f) return x;
  const float w = __expf(x);
  const float n = w * (w + 2.f);
  return x * __fdividef(n, n + 2.f);
}

__global__ void __launch_bounds__(256) k_gn_apply(const GnApplyArgs a) {
  __shared__ float s_mean[2][8], s_rstd[2][8];
  const int cpt = a.C / 8;                                   // threads per pixel
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)a.B * a.P * cpt;
  const int gs = a.C / a.G;
  const int b0 = (int)((blockIdx.x * (long)blockDim.x / cpt) / a.P);
  gn_block_stats(a.stats, a.B, a.G, (double)a.P * gs, b0, s_mean, s_rstd);
  if (gid >= total) return;
  const int c0 = (int)(gid % cpt) * 8;
  const long pix = gid / cpt;                                // global pixel row
  const int b = (int)(pix / a.P);
  const int w = (int)((pix % a.P) % a.W);
  const int g = c0 / gs;
  const float mean = s_mean[b - b0][g], rstd = s_rstd[b - b0][g];
  const float m = a.mask[(long)b * a.mask_stride + w];
  const float* rp = a.raw + pix * a.C + c0;
  const float4 r0 = *reinterpret_cast<const float4*>(rp);
  const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
  float v[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
  float res[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) res[i] = 0.f;
  if (a.resid_s.p != nullptr) {
    const bf16* q = a.resid_s.p + pix * a.resid_s.stride + c0;
    load_split8(q + a.resid_s.hi, q + a.resid_s.lo, res);
  } else if (a.resid_f != nullptr) {
    const float* q = a.resid_f + pix * a.resid_f_stride + c0;
    const float4 q0 = *reinterpret_cast<const float4*>(q);
    const float4 q1 = *reinterpret_cast<const float4*>(q + 4);
    res[0] = q0.x * m; res[1] = q0.y * m; res[2] = q0.z * m; res[3] = q0.w * m;
    res[4] = q1.x * m; res[5] = q1.y * m; res[6] = q1.z * m; res[7] = q1.w * m;
  } else if (a.rin_w != nullptr) {
    // res_conv(x * mask) of the first ResnetBlock: 1x1 conv on stack[mu, c_in*x]
    const float in0 = a.mu[pix] * m, in1 = (a.tab[a.step].c_in * a.x[pix]) * m;
#pragma unroll
    for (int i = 0; i < 8; ++i) res[i] = (a.rin_b[c0 + i] + a.rin_w[(c0 + i) * 2] * in0 + a.rin_w[(c0 + i) * 2 + 1] * in1) * m;
  }
  float ga[8], be[8], tb[8];
  {
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
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float y = (v[i] - mean) * rstd * ga[i] + be[i];
    y = mish_fast(y) * m;
    y = (y + tb[i]) * m;                                     // tb == 0 without a time bias: (y*m)*m == y*m for m in {0,1}
    v[i] = y + res[i];
  }
  bf16* op = a.out.p + pix * a.out.stride + c0;
  store_split8(op + a.out.hi, op + a.out.lo, v);
}
void launch_gn_apply(const GnApplyArgs& a, cudaStream_t st) {
  const long total = (long)a.B * a.P * (a.C / 8);
  k_gn_apply<<<cdiv(total, 256), 256, 0, st>>>(a);
}

// ------------------------------------------------------------------------------------------------
// final: GN + Mish + mask -> 1x1 conv (C -> 1) + bias -> mask = F_x;  D = c_skip*x + c_out*F_x;
// d = x/sigma - D/sigma;  x <- x + (sigma_next - sigma) * d        (edm.py:97,197,203)
// 8 lanes per pixel (C == 64).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_gn_final(const float* __restrict__ raw, int C, int G,
                                                  const double* __restrict__ stats, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, const float* __restrict__ fc_w,
                                                  const float* __restrict__ fc_b, const float* __restrict__ mask,
                                                  float* __restrict__ x, float* __restrict__ den_out,
                                                  const StepScalars* __restrict__ tab, int step, int B, int P, int W) {
  const int cpt = C / 8;                                      // == 8
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * P * cpt;
  const bool active = gid < total;
  const long pix = active ? gid / cpt : 0;
  const int c0 = (int)(gid % cpt) * 8;
  const int b = (int)(pix / P);
  const int w = (int)((pix % P) % W);
  __shared__ float s_mean[2][8], s_rstd[2][8];
  const int b0 = (int)((blockIdx.x * (long)blockDim.x / cpt) / P);
  gn_block_stats(stats, B, G, (double)P * (C / G), b0, s_mean, s_rstd);
  float part = 0.f;
  float m = 0.f;
  if (active) {
    const int gs = C / G;
    const int g = c0 / gs;
    const float mean = s_mean[b - b0][g], rstd = s_rstd[b - b0][g];
    m = mask[(long)b * W + w];
    const float* rp = raw + pix * C + c0;
    const float4 r0 = *reinterpret_cast<const float4*>(rp);
    const float4 r1 = *reinterpret_cast<const float4*>(rp + 4);
    const float v[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      const float y = mish_fast((v[i] - mean) * rstd * gamma[c] + beta[c]) * m;
      part = fmaf(fc_w[c], y * m, part);
    }
  }
  // reduce over the 8 lanes of a pixel (cpt == 8, aligned groups of lanes)
  part += __shfl_xor_sync(0xffffffffu, part, 1);
  part += __shfl_xor_sync(0xffffffffu, part, 2);
  part += __shfl_xor_sync(0xffffffffu, part, 4);
  if (active && c0 == 0) {
    const StepScalars sc = tab[step];
    const float fx = (part + fc_b[0]) * m;
    const float xv = x[pix];
    const float den = sc.c_skip * xv + sc.c_out * fx;
    const float inv = 1.f / sc.sigma;
    const float d = inv * xv - inv * den;
    if (den_out != nullptr) den_out[pix] = den;
    else x[pix] = xv + (sc.sigma_next - sc.sigma) * d;
  }
}
void launch_gn_final(const float* raw, int C, int G, const double* stats, const float* gamma, const float* beta,
                     const float* fc_w, const float* fc_b, const float* mask, float* x, float* den_out,
                     const StepScalars* tab, int step, int B, int H, int W, cudaStream_t st) {
  const long total = (long)B * H * W * (C / 8);
  k_gn_final<<<cdiv(total, 256), 256, 0, st>>>(raw, C, G, stats, gamma, beta, fc_w, fc_b, mask, x, den_out, tab, step, B,
                                                H * W, W);
}

// ------------------------------------------------------------------------------------------------
// LinearAttention (diffusion.py:82-95): softmax of k over ALL pixels, context = softmax(k) v^T (32x32 per head).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_ord(float f) {
  unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_ord(unsigned u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// column max of k: kv rows are [k(128) | v(128)]; block = 128 threads (one per k column) over a chunk of pixels
__global__ void __launch_bounds__(128) k_la_colmax(const float* __restrict__ kv, unsigned* __restrict__ kmax, int P,
                                                   int chunk) {
  const int b = blockIdx.y;
  const long p0 = (long)blockIdx.x * chunk;
  long p1 = p0 + chunk;
  if (p1 > P) p1 = P;
  float m = -INFINITY;
  const float* base = kv + ((long)b * P) * 256 + threadIdx.x;
  for (long p = p0; p < p1; ++p) m = fmaxf(m, base[p * 256]);
  atomicMax(&kmax[b * 128 + threadIdx.x], enc_ord(m));
}
void launch_la_colmax(const float* kv, unsigned* kmax_enc, int B, int P, cudaStream_t st) {
  const int chunk = 256;
  dim3 grid(cdiv(P, chunk), B);
  k_la_colmax<<<grid, 128, 0, st>>>(kv, kmax_enc, P, chunk);
}

// ctx[b][h][d][e] += sum_n exp(k[n][h*32+d] - max) * v[n][h*32+e];  ssum[b][h*32+d] += sum_n exp(..)
// One block (256 threads) = one (b, pixel chunk), all 4 heads: the full 1 KiB kv row of every pixel is read once,
// coalesced; thread (h, dq, eq) keeps a 4x4 register block of the 32x32 context of head h (16 FMA per 2 LDS.128).
constexpr int kLaTile = 32;
constexpr int kLaPartial = 4096 + 128;       // one block's partial: ctx[4][32][32] | ssum[128]
__global__ void __launch_bounds__(256) k_la_ctx(const float* __restrict__ kv, const unsigned* __restrict__ kmax,
                                                float* __restrict__ part, int P, int chunk) {
  __shared__ __align__(16) float ps[kLaTile][128];
  __shared__ __align__(16) float vs[kLaTile][128];
  const int b = blockIdx.y;
  const long p0 = (long)blockIdx.x * chunk;
  long p1 = p0 + chunk;
  if (p1 > P) p1 = P;
  const int tid = threadIdx.x;
  const int h = tid >> 6, dq = (tid & 63) >> 3, eq = tid & 7;
  const int d0 = h * 32 + dq * 4, e0 = h * 32 + eq * 4;
  // loader: thread handles float4 column group lc4 (0..63: 0..31 = k, 32..63 = v) of rows lr, lr+4, ...
  const int lc4 = tid & 63, lr = tid >> 6;
  const float kLog2e = 1.4426950408889634f;
  float4 kmx = make_float4(0.f, 0.f, 0.f, 0.f);           // column maxima, pre-multiplied by log2(e)
  if (lc4 < 32) {
    const unsigned* km = kmax + b * 128 + lc4 * 4;
    kmx = make_float4(dec_ord(km[0]) * kLog2e, dec_ord(km[1]) * kLog2e, dec_ord(km[2]) * kLog2e, dec_ord(km[3]) * kLog2e);
  }
  float acc[4][4];
  float4 psum = make_float4(0.f, 0.f, 0.f, 0.f);          // loader threads: running sum of exp(k - max) of their 4 columns
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (long t0 = p0; t0 < p1; t0 += kLaTile) {
    __syncthreads();
#pragma unroll
    for (int rr = 0; rr < kLaTile / 4; ++rr) {
      const int r = lr + rr * 4;
      const long pp = t0 + r;
      float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
      if (pp < p1) {
        x = *reinterpret_cast<const float4*>(kv + ((long)b * P + pp) * 256 + lc4 * 4);
        if (lc4 < 32) {                                   // exp(k - max) = exp2(k log2e - max log2e): one FFMA + one MUFU
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x.x) : "f"(fmaf(x.x, kLog2e, -kmx.x)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x.y) : "f"(fmaf(x.y, kLog2e, -kmx.y)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x.z) : "f"(fmaf(x.z, kLog2e, -kmx.z)));
          asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(x.w) : "f"(fmaf(x.w, kLog2e, -kmx.w)));
          psum.x += x.x; psum.y += x.y; psum.z += x.z; psum.w += x.w;
        }
      }
      if (lc4 < 32) *reinterpret_cast<float4*>(&ps[r][lc4 * 4]) = x;
      else *reinterpret_cast<float4*>(&vs[r][(lc4 - 32) * 4]) = x;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < kLaTile; ++r) {
      const float4 p4 = *reinterpret_cast<const float4*>(&ps[r][d0]);
      const float4 v4 = *reinterpret_cast<const float4*>(&vs[r][e0]);
      const float pp[4] = {p4.x, p4.y, p4.z, p4.w};
      const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(pp[i], vv[j], acc[i][j]);
    }
  }
  // per-block partial sums, reduced in a fixed order by k_la_reduce: deterministic (fp32 atomics made the whole
  // trajectory vary by ~5e-5 from run to run) and no contention on the 4224 addresses of an image
  float* pb = part + ((long)b * gridDim.x + blockIdx.x) * kLaPartial;
#pragma unroll
  for (int i = 0; i < 4; ++i)
    *reinterpret_cast<float4*>(pb + ((h * 32 + dq * 4 + i) * 32 + eq * 4)) = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
  // column sums: the 4 loader rows (lr = 0..3) of a column group combine through shared memory in a fixed order
  __syncthreads();
  if (lc4 < 32) *reinterpret_cast<float4*>(&ps[lr][lc4 * 4]) = psum;
  __syncthreads();
  if (tid < 128) pb[4096 + tid] = ((ps[0][tid] + ps[1][tid]) + ps[2][tid]) + ps[3][tid];
}
// ctx[b][4096] | ssum[b][128]  =  sum over the blocks of an image, in block order
__global__ void __launch_bounds__(256) k_la_reduce(const float* __restrict__ part, float* __restrict__ ctx,
                                                   float* __restrict__ ssum, int B, int nblk) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * kLaPartial) return;
  const int b = i / kLaPartial, k = i % kLaPartial;
  const float* p = part + (long)b * nblk * kLaPartial + k;
  float a = 0.f;
  for (int j = 0; j < nblk; ++j) a += p[(long)j * kLaPartial];
  if (k < 4096) ctx[(long)b * 4096 + k] = a;
  else ssum[b * 128 + (k - 4096)] = a;
}
int la_ctx_blocks(int B, int P) {
  int chunk = 512;
  while (chunk > 128 && (long)cdiv(P, chunk) * B < 4 * 148) chunk >>= 1;
  return cdiv(P, chunk);
}
void launch_la_ctx(const float* kv, const unsigned* kmax_enc, float* part, float* ctx, float* ssum, int B, int P,
                   cudaStream_t st) {
  const int nblk = la_ctx_blocks(B, P);
  const int chunk = cdiv(cdiv(P, nblk), kLaTile) * kLaTile;
  dim3 grid(nblk, B);
  k_la_ctx<<<grid, 256, 0, st>>>(kv, kmax_enc, part, P, chunk);
  k_la_reduce<<<cdiv((long)B * kLaPartial, 256), 256, 0, st>>>(part, ctx, ssum, B, nblk);
}

// Merge the split-KV partials of the tensor-core context kernel (attn.cu, out_mode 2), in split order:
//   M[d] = max_s m_s[d];  ctx[b][h][d][e] = sum_s O_s[d][h*32+e] exp(m_s[d] - M[d]);  ssum[b][d] = sum_s l_s[d] exp(m_s[d] - M[d])
// (only the 4 diagonal 32x32 head blocks of the 128x128 product are context).  One thread per output.
__global__ void __launch_bounds__(256) k_la_combine(const float* __restrict__ part_o, const float* __restrict__ part_l,
                                                    const float* __restrict__ part_m, float* __restrict__ ctx,
                                                    float* __restrict__ ssum, int B, int S) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int per = 4096 + 128;
  if (i >= B * per) return;
  const int b = i / per, k = i % per;
  int d, col;
  if (k < 4096) { const int h = k >> 10, dl = (k >> 5) & 31, el = k & 31; d = h * 32 + dl; col = h * 32 + el; }
  else { d = k - 4096; col = -1; }
  float M = -INFINITY;
  for (int s = 0; s < S; ++s) M = fmaxf(M, part_m[((long)b * S + s) * 128 + d]);
  float acc = 0.f;
  for (int s = 0; s < S; ++s) {
    const long ps = (long)b * S + s;
    const float w = expf(part_m[ps * 128 + d] - M);          // exp(-inf) = 0 for an empty split
    const float v = (col >= 0) ? part_o[(ps * 128 + d) * 128 + col] : part_l[ps * 128 + d];
    acc = fmaf(v, w, acc);
  }
  if (col >= 0) ctx[(long)b * 4096 + k] = acc;
  else ssum[b * 128 + d] = acc;
}
void launch_la_combine(const float* part_o, const float* part_l, const float* part_m, float* ctx, float* ssum, int B, int S,
                       cudaStream_t st) {
  k_la_combine<<<cdiv((long)B * (4096 + 128), 256), 256, 0, st>>>(part_o, part_l, part_m, ctx, ssum, B, S);
}

// W_eff[b][co][ci] = delta(co,ci) + g * sum_{h,e} Wout[co][h*32+e] * sum_d (ctx[b][h][d][e]/ssum[b][h][d]) * Wq[h*32+d][ci]
// (q is linear in x, so  x + g*to_out(ctx^T q) == W_eff x + g*b_out : the whole attention read-out is one
//  per-sample CxC matrix.)   Two small fully parallel kernels: m1 = ctxn^T Wq, then W_eff = I + g Wout m1.
__global__ void __launch_bounds__(256) k_la_m1(const float* __restrict__ ctx, const float* __restrict__ ssum,
                                               const float* __restrict__ wq, float* __restrict__ m1, int B, int C) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)B * 128 * C) return;
  const int ci = (int)(idx % C);
  const int he = (int)((idx / C) % 128);
  const int b = (int)(idx / ((long)C * 128));
  const int h = he >> 5, e = he & 31;
  float acc = 0.f;
#pragma unroll 8
  for (int d = 0; d < 32; ++d) {
    const float cn = ctx[(((long)b * 4 + h) * 32 + d) * 32 + e] / ssum[b * 128 + h * 32 + d];
    acc = fmaf(cn, wq[(h * 32 + d) * C + ci], acc);
  }
  m1[idx] = acc;
}
__global__ void __launch_bounds__(256) k_la_weff(const float* __restrict__ m1, const float* __restrict__ wout,
                                                 const float* __restrict__ bout, const float* __restrict__ g,
                                                 bf16* __restrict__ weff, float* __restrict__ beff, int B, int C) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (idx >= (long)B * C * C) return;
  const int ci = (int)(idx % C);
  const int co = (int)((idx / C) % C);
  const int b = (int)(idx / ((long)C * C));
  const float gg = g[0];
  const float* mp = m1 + (long)b * 128 * C + ci;
  const float* wp = wout + (long)co * 128;
  float acc = 0.f;
#pragma unroll 8
  for (int he = 0; he < 128; ++he) acc = fmaf(wp[he], mp[(long)he * C], acc);
  const float v = gg * acc + (co == ci ? 1.f : 0.f);
  bf16* row = weff + ((long)b * C + co) * (2 * C);
  split2(v, row[ci], row[C + ci]);
  if (ci == 0) beff[b * C + co] = gg * bout[co];
}
int kernels_global_init() { return 0; }
void launch_la_weff(const float* ctx, const float* ssum, const float* wq, const float* wout, const float* bout,
                    const float* g, float* m1, bf16* weff, float* beff, int B, int C, cudaStream_t st) {
  k_la_m1<<<cdiv((long)B * 128 * C, 256), 256, 0, st>>>(ctx, ssum, wq, m1, B, C);
  k_la_weff<<<cdiv((long)B * C * C, 256), 256, 0, st>>>(m1, wout, bout, g, weff, beff, B, C);
}

// ------------------------------------------------------------------------------------------------
// per-(image, channel) sum / sumsq (InstanceNorm2D statistics, base.py:95-103)
// block = 256 threads = 32 pixel-lanes x (C/8 <= 16 ... ) ; simple: thread handles 8 channels of a pixel stream
// ------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_chan_stats(const bf16* __restrict__ xs, long s_stride, int hi, int lo,
                                                    const float* __restrict__ xf, long f_stride,
                                                    double* __restrict__ stats, int P, int C, int chunk) {
  __shared__ float red[256][17];
  const int cpt = C / 8;                                      // 16 for C == 128, 8 for C == 64
  const int rows_per_iter = 256 / cpt;
  const int cg = threadIdx.x % cpt, pr = threadIdx.x / cpt;
  const int b = blockIdx.y;
  const long p0 = (long)blockIdx.x * chunk;
  long p1 = p0 + chunk;
  if (p1 > P) p1 = P;
  float s[8], ss[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = 0.f; ss[i] = 0.f; }
  for (long p = p0 + pr; p < p1; p += rows_per_iter) {
    float v[8];
    const long row = (long)b * P + p;
    if (SPLIT) {
      const bf16* q = xs + row * s_stride + cg * 8;
      load_split8(q + hi, q + lo, v);
    } else {
      const float* q = xf + row * f_stride + cg * 8;
      const float4 a0 = *reinterpret_cast<const float4*>(q);
      const float4 a1 = *reinterpret_cast<const float4*>(q + 4);
      v[0] = a0.x; v[1] = a0.y; v[2] = a0.z; v[3] = a0.w; v[4] = a1.x; v[5] = a1.y; v[6] = a1.z; v[7] = a1.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { s[i] += v[i]; ss[i] = fmaf(v[i], v[i], ss[i]); }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[threadIdx.x][i] = s[i]; red[threadIdx.x][8 + i] = ss[i]; }
  __syncthreads();
  // thread t < C sums channel t over the pixel-rows of the block
  if (threadIdx.x < C) {
    const int c = threadIdx.x;
    const int g = c / 8, i = c % 8;
    double a = 0., q = 0.;
    for (int r = 0; r < rows_per_iter; ++r) { a += red[r * cpt + g][i]; q += red[r * cpt + g][8 + i]; }
    atomicAdd(&stats[((long)b * C + c) * 2], a);
    atomicAdd(&stats[((long)b * C + c) * 2 + 1], q);
  }
}
void launch_chan_stats_s(SView x, double* stats, int B, int P, int C, cudaStream_t st) {
  const int chunk = 128;                                 // 80 -> 640 blocks at C2: the kernel was latency-bound
  dim3 grid(cdiv(P, chunk), B);
  k_chan_stats<true><<<grid, 256, 0, st>>>(x.p, x.stride, x.hi, x.lo, nullptr, 0, stats, P, C, chunk);
}
void launch_chan_stats_f(const float* x, long stride, double* stats, int B, int P, int C, cudaStream_t st) {
  const int chunk = 128;                                 // 80 -> 640 blocks at C2: the kernel was latency-bound
  dim3 grid(cdiv(P, chunk), B);
  k_chan_stats<false><<<grid, 256, 0, st>>>(nullptr, 0, 0, 0, x, stride, stats, P, C, chunk);
}

}  // namespace dexb
