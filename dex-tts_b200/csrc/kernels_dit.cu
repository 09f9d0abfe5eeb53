// Bottleneck-side kernels: TV/TIV adaptor glue, DiT patch embed front, LayerNorm+modulate, softmaxes, unpatchify,
// the tiny per-step table builders and the one-off weight packers.
// Reference semantics: DEX-TTS/model/ref_encoder.py:142-179,239-273, DEX-TTS/model/dit.py:31-90,219-326,434-519.
#include "kernels.cuh"

namespace dexb {

// ------------------------------------------------------------------------------------------------
// TV adaptor fold (per step).  With q = W_q ((x - mean)/std)  (InstanceNorm2D, unbiased variance, eps 1e-5):
//   S[pixel][j] = (q . k_j)/sqrt(C) = sum_c KQ[j][c] x[c] + sb[j],   KQ[j][c] = KW[j][c]/std[c],
//   sb[j] = -sum_c KQ[j][c] mean[c],   KW[j][c] = sum_o K[j][o] W_q[o][c] / sqrt(C)   (step-invariant for j >= 1).
// One warp per (b, j).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tv_fold(const float* __restrict__ kw, const float* __restrict__ kw0,
                                                 const double* __restrict__ stats, int P, bf16* __restrict__ kq,
                                                 float* __restrict__ sbias, int B, int NK, int NKR, int C) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * NK) return;
  const int b = warp / NK, j = warp % NK;
  const float* src = (j == 0) ? kw0 : kw + ((long)b * (NK - 1) + (j - 1)) * C;   // kw holds the NK-1 style rows
  bf16* dst = kq + ((long)b * NKR + j) * (2 * C);
  float sb = 0.f;
  for (int c = lane; c < C; c += 32) {
    const double s = stats[((long)b * C + c) * 2], ss = stats[((long)b * C + c) * 2 + 1];
    const double mean = s / P;
    double var = (ss - s * mean) / (double)(P - 1);          // unbiased (torch.var default)
    if (var < 0.) var = 0.;
    const float istd = (float)(1.0 / sqrt(var + 1e-5));
    const float v = src[c] * istd;
    split2(v, dst[c], dst[C + c]);
    sb -= v * (float)mean;
  }
  sb = warp_sum(sb);
  if (lane == 0) sbias[(long)b * NKR + j] = sb;
}
void launch_tv_fold(const float* kw, const float* kw0, const double* stats, int P, bf16* kq, float* sbias, int B,
                    int NK, int NKR, int C, cudaStream_t st) {
  launch_pdl(k_tv_fold, dim3((unsigned)(cdiv((long)B * NK * 32, 256))), dim3(256), 0, st, kw, kw0, stats, P, kq, sbias, B, NK, NKR, C);
}

// column 0 (the time token) of VL^T changes every step:  vlt[b][c][0] = vl0[c]
__global__ void k_tv_vl0(const float* __restrict__ vl0, bf16* __restrict__ vlt, int B, int C, int KP) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * C) return;
  const int c = i % C;
  bf16* row = vlt + (long)i * (2 * KP);
  split2(vl0[c], row[0], row[KP]);
}
void launch_tv_vl0(const float* vl0, bf16* vlt, int B, int C, int KP, cudaStream_t st) {
  launch_pdl(k_tv_vl0, dim3((unsigned)(cdiv((long)B * C, 128))), dim3(128), 0, st, vl0, vlt, B, C, KP);
}

// masked softmax over the NK = Ts+1 style tokens (key 0 = time token, always visible; masked keys get -1e4,
// ref_encoder.py:171-172).  One warp per pixel row; writes split P with zero padding up to KP.
__global__ void __launch_bounds__(256) k_tv_softmax(const float* __restrict__ scores, long sstride,
                                                    const int* __restrict__ sty_len, bf16* __restrict__ P_, long rows,
                                                    int Ppix, int NK, int KP) {
  pdl_wait();
  const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = (int)(row / Ppix);
  const int vis = sty_len[b] + 1;
  const float* sp = scores + row * sstride;
  // three passes over the (L2-resident) score row: any number of style tokens (the fused attention kernel takes NK <= 512; this
  // path serves longer reference utterances -- synthesize.py:96-99 passes the whole reference mel as `sty`)
  float mx = -INFINITY;
  for (int j = lane; j < NK; j += 32) mx = fmaxf(mx, (j < vis) ? sp[j] : -1e4f);
  mx = warp_max(mx);
  float sum = 0.f;
  for (int j = lane; j < NK; j += 32) sum += expf(((j < vis) ? sp[j] : -1e4f) - mx);
  sum = warp_sum(sum);
  const float inv = 1.f / sum;
  bf16* op = P_ + row * (2 * (long)KP);
  for (int j = lane; j < KP; j += 32) {
    const float e = (j < NK) ? expf(((j < vis) ? sp[j] : -1e4f) - mx) * inv : 0.f;
    split2(e, op[j], op[KP + j]);
  }
}
void launch_tv_softmax(const float* scores, long sstride, const int* sty_len, bf16* P_, int B, int Ppix, int NK, int KP,
                       cudaStream_t st) {
  const long rows = (long)B * Ppix;
  launch_pdl(k_tv_softmax, dim3((unsigned)(cdiv(rows * 32, 256))), dim3(256), 0, st, scores, sstride, sty_len, P_, rows, Ppix, NK, KP);
}

// ------------------------------------------------------------------------------------------------
// DiT front: (TIV AdaIN affine) -> F.pad to a multiple of PATCH size with zeros -> depthwise conv p x p, stride s,
// padding p//2 -> SiLU -> S tokens.  dit.py:434-441,50-52.   One thread = 8 channels of one token.
// ------------------------------------------------------------------------------------------------
// AdaIN of the TIV adaptor as a per-(b, c) affine map  y = a x + d:  a = sc / std,  d = sh - mean * a
// (InstanceNorm2D: unbiased variance over all pixels, eps 1e-5; base.py:95-109, ref_encoder.py:264-273)
__global__ void k_tiv_affine(const double* __restrict__ stats, const float* __restrict__ sc, const float* __restrict__ sh,
                             float* __restrict__ a_out, float* __restrict__ d_out, int n, int P) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double sm = stats[(long)i * 2], ss = stats[(long)i * 2 + 1];
  const double mean = sm / P;
  double var = (ss - sm * mean) / (double)(P - 1);
  if (var < 0.) var = 0.;
  const float istd = (float)(1.0 / sqrt(var + 1e-5));
  const float a = istd * sc[i];
  a_out[i] = a;
  d_out[i] = sh[i] - (float)mean * a;
}
void launch_tiv_affine(const double* stats, const float* sc, const float* sh, float* a_out, float* d_out, int B, int C, int P,
                       cudaStream_t st) {
  launch_pdl(k_tiv_affine, dim3((unsigned)(cdiv((long)B * C, 128))), dim3(128), 0, st, stats, sc, sh, a_out, d_out, B * C, P);
}

// PS = compile-time patch size (3: DEX-TTS, 7: GeDEX-TTS; 0 = run-time p): the taps of a row are loaded together (out-of-range taps
// as predicated zero weights) -- as a rolled loop with `continue` every tap was its own dependent round trip (49 of them at p = 7).
// The depthwise weights are staged transposed ([tap][channel]) in shared memory: two 16 B reads per tap instead of 8 scalar loads.
template <bool SPLIT_IN, int PS>
__global__ void __launch_bounds__(256) k_dw_patch(const float* __restrict__ xin, SView sin,
                                                  const double* __restrict__ stats, const float* __restrict__ tiv_scale,
                                                  const float* __restrict__ tiv_shift, int use_tiv,
                                                  const float* __restrict__ dw_w, const float* __restrict__ dw_b,
                                                  SView out, int B, int H, int W, int C, int p_rt, int s, int Fq, int Wq) {
  pdl_wait();
  extern __shared__ __align__(16) float dw_s[];               // [p * p][C]
  const int p = PS > 0 ? PS : p_rt;
  for (int i = threadIdx.x; i < C * p * p; i += 256) dw_s[(i % (p * p)) * C + i / (p * p)] = dw_w[i];
  __syncthreads();
  const int cpt = C / 8;
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * Fq * Wq * cpt;
  if (gid >= total) return;
  const int c0 = (int)(gid % cpt) * 8;
  const long tok = gid / cpt;
  const int wq = (int)(tok % Wq), hq = (int)((tok / Wq) % Fq), b = (int)(tok / ((long)Wq * Fq));
  float a[8], d[8], acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = 1.f; d[i] = 0.f;
    acc[i] = dw_b[c0 + i];
  }
  if (use_tiv) {                                               // a, d precomputed per (b, c) by k_tiv_affine
    const float4 a0 = *reinterpret_cast<const float4*>(tiv_scale + (long)b * C + c0), a1 = *reinterpret_cast<const float4*>(tiv_scale + (long)b * C + c0 + 4);
    const float4 d0 = *reinterpret_cast<const float4*>(tiv_shift + (long)b * C + c0), d1 = *reinterpret_cast<const float4*>(tiv_shift + (long)b * C + c0 + 4);
    a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
    d[0] = d0.x; d[1] = d0.y; d[2] = d0.z; d[3] = d0.w; d[4] = d1.x; d[5] = d1.y; d[6] = d1.z; d[7] = d1.w;
  }
  const int pad = p / 2;
  constexpr int KXU = PS > 0 ? PS : 1;                         // taps of a row in flight together
#pragma unroll 1
  for (int ky = 0; ky < p; ++ky) {
    const int y = hq * s - pad + ky;
    if (y < 0 || y >= H) continue;                             // (uniform enough: whole rows of the padding zone)
    for (int kx0 = 0; kx0 < p; kx0 += KXU) {
      float v[KXU][8];
      bool ok[KXU];
#pragma unroll
      for (int u = 0; u < KXU; ++u) {
        const int x = wq * s - pad + kx0 + u;
        ok[u] = x >= 0 && x < W;                               // conv zero padding and the F.pad zone are both 0
        const long row = ((long)b * H + y) * W + (ok[u] ? x : 0);
        if (SPLIT_IN) {
          const bf16* q = sin.p + row * sin.stride + c0;
          load_split8(q + sin.hi, q + sin.lo, v[u]);
        } else {
          const float* q = xin + row * C + c0;
          const float4 r0 = *reinterpret_cast<const float4*>(q);
          const float4 r1 = *reinterpret_cast<const float4*>(q + 4);
          v[u][0] = r0.x; v[u][1] = r0.y; v[u][2] = r0.z; v[u][3] = r0.w; v[u][4] = r1.x; v[u][5] = r1.y; v[u][6] = r1.z; v[u][7] = r1.w;
        }
      }
#pragma unroll
      for (int u = 0; u < KXU; ++u) {
        const float* wp = dw_s + (ky * p + kx0 + u) * C + c0;
        const float4 w0 = *reinterpret_cast<const float4*>(wp), w1 = *reinterpret_cast<const float4*>(wp + 4);
        const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
        if (ok[u]) {
#pragma unroll
          for (int i = 0; i < 8; ++i) acc[i] = fmaf(wv[i], fmaf(a[i], v[u][i], d[i]), acc[i]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = silu_f(acc[i]);
  bf16* op = out.p + tok * out.stride + c0;
  store_split8(op + out.hi, op + out.lo, acc);
}
template <bool SPLIT_IN, class... Args>
static void launch_dw_patch_p(long total, int C, int p, cudaStream_t st, Args... args) {
  const size_t smem = (size_t)C * p * p * sizeof(float);      // 4.6 KB (p = 3, C = 128) ... 25 KB (p = 7); < 48 KB checked by the engine's config
  const dim3 grid((unsigned)cdiv(total, 256));
  if (p == 3) launch_pdl(k_dw_patch<SPLIT_IN, 3>, grid, dim3(256), smem, st, args...);
  else if (p == 7) launch_pdl(k_dw_patch<SPLIT_IN, 7>, grid, dim3(256), smem, st, args...);
  else launch_pdl(k_dw_patch<SPLIT_IN, 0>, grid, dim3(256), smem, st, args...);
}
void launch_dw_patch(const float* tv, const double* stats, const float* tiv_scale, const float* tiv_shift, int use_tiv,
                     const float* dw_w, const float* dw_b, SView out, int B, int H, int W, int C, int p, int s, int Fq,
                     int Wq, cudaStream_t st) {
  const long total = (long)B * Fq * Wq * (C / 8);
  SView none = {nullptr, 0, 0, 0};
  launch_dw_patch_p<false>(total, C, p, st, tv, none, stats, tiv_scale, tiv_shift, use_tiv, dw_w, dw_b, out, B, H, W, C, p, s, Fq, Wq);
}
void launch_dw_patch_s(SView in, const float* dw_w, const float* dw_b, SView out, int B, int H, int W, int C, int p,
                       int s, int Fq, int Wq, cudaStream_t st) {
  const long total = (long)B * Fq * Wq * (C / 8);
  launch_dw_patch_p<true>(total, C, p, st, (const float*)nullptr, in, (const double*)nullptr, (const float*)nullptr, (const float*)nullptr, 0, dw_w, dw_b,
                          out, B, H, W, C, p, s, Fq, Wq);
}

// ------------------------------------------------------------------------------------------------
// LayerNorm(eps 1e-6, no affine) + adaLN modulate -> S.  One warp per token, D = 256 (8 per lane) or 384 (12).
// ------------------------------------------------------------------------------------------------
template <int PER_LANE>
__device__ __forceinline__ void ln_mod_row(float (&v)[PER_LANE], const float* shift, const float* scale, bf16* op,
                                           int hi, int lo, int lane, int D) {
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) s += v[i];
  const float mean = warp_sum(s) / D;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < PER_LANE; ++i) { const float t = v[i] - mean; q = fmaf(t, t, q); }
  const float rstd = rsqrtf(warp_sum(q) / D + 1e-6f);
#pragma unroll
  for (int g = 0; g < PER_LANE; g += 4) {
    const int c = (g / 4) * 128 + lane * 4;                   // lane owns 4 consecutive channels per 128-chunk
    __align__(8) bf16 h[4];
    __align__(8) bf16 l[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float y = (v[g + i] - mean) * rstd * (1.f + scale[c + i]) + shift[c + i];
      split2(y, h[i], l[i]);
    }
    *reinterpret_cast<uint2*>(op + hi + c) = *reinterpret_cast<const uint2*>(h);
    *reinterpret_cast<uint2*>(op + lo + c) = *reinterpret_cast<const uint2*>(l);
  }
}

// Two tokens per warp: the pass is latency-bound (one load round trip, two dependent warp reductions, one store per token and ~2 waves of
// warps per SM), so a warp carries two independent chains.
template <int PER_LANE>
__global__ void __launch_bounds__(256) k_ln_mod(const float* __restrict__ x, const float* __restrict__ shift,
                                                const float* __restrict__ scale, SView out, long M, int D) {
  pdl_wait();
  const long row = ((blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5) * 2;
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const bool two = row + 1 < M;
  float v[PER_LANE], u[PER_LANE];
  const float* xp = x + row * D;
#pragma unroll
  for (int g = 0; g < PER_LANE; g += 4) {
    const float4 t = *reinterpret_cast<const float4*>(xp + (g / 4) * 128 + lane * 4);
    v[g] = t.x; v[g + 1] = t.y; v[g + 2] = t.z; v[g + 3] = t.w;
    const float4 t2 = two ? *reinterpret_cast<const float4*>(xp + D + (g / 4) * 128 + lane * 4) : make_float4(0.f, 0.f, 0.f, 0.f);
    u[g] = t2.x; u[g + 1] = t2.y; u[g + 2] = t2.z; u[g + 3] = t2.w;
  }
  ln_mod_row<PER_LANE>(v, shift, scale, out.p + row * out.stride, out.hi, out.lo, lane, D);
  if (two) ln_mod_row<PER_LANE>(u, shift, scale, out.p + (row + 1) * out.stride, out.hi, out.lo, lane, D);
}
void launch_ln_mod(const float* x, const float* shift, const float* scale, SView out, long M, int D, cudaStream_t st) {
  if (D == 256) launch_pdl(k_ln_mod<8>, dim3(cdiv(cdiv(M, 2) * 32, 256)), dim3(256), 0, st, x, shift, scale, out, M, D);
  else if (D == 384) launch_pdl(k_ln_mod<12>, dim3(cdiv(cdiv(M, 2) * 32, 256)), dim3(256), 0, st, x, shift, scale, out, M, D);
}

// pe[b][w][c] = mean over the frequency rows of pg[b][h][w][c] (dit.py:445), summed in a fixed order.  One thread = 4 channels.
__global__ void __launch_bounds__(256) k_freq_mean(const float* __restrict__ pg, float* __restrict__ pe, int B, int Fq, int Wq,
                                                   int D) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * Wq * (D / 4);
  if (i >= total) return;
  const int c = (int)(i % (D / 4)) * 4;
  const int w = (int)((i / (D / 4)) % Wq);
  const int b = (int)(i / ((long)(D / 4) * Wq));
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int h = 0; h < Fq; ++h) {
    const float4 q = *reinterpret_cast<const float4*>(pg + (((long)b * Fq + h) * Wq + w) * D + c);
    a.x += q.x; a.y += q.y; a.z += q.z; a.w += q.w;
  }
  const float inv = 1.f / (float)Fq;
  *reinterpret_cast<float4*>(pe + ((long)b * Wq + w) * D + c) = make_float4(a.x * inv, a.y * inv, a.z * inv, a.w * inv);
}
void launch_freq_mean(const float* pg, float* pe, int B, int Fq, int Wq, int D, cudaStream_t st) {
  launch_pdl(k_freq_mean, dim3((unsigned)(cdiv((long)B * Wq * (D / 4), 256))), dim3(256), 0, st, pg, pe, B, Fq, Wq, D);
}

// x = xe + pe[b][w] + fpos[h]  (dit.py:444-447), stored fp32 (residual stream) and LN+modulated for block 0
template <int PER_LANE>
__global__ void __launch_bounds__(256) k_tok_assemble(const float* __restrict__ xe, const float* __restrict__ pe,
                                                      const float* __restrict__ fpos, float* __restrict__ x,
                                                      const float* __restrict__ shift, const float* __restrict__ scale,
                                                      SView out, int B, int Fq, int Wq, int D) {
  pdl_wait();
  const long row = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const long M = (long)B * Fq * Wq;
  if (row >= M) return;
  const int wq = (int)(row % Wq), hq = (int)((row / Wq) % Fq), b = (int)(row / ((long)Wq * Fq));
  float v[PER_LANE];
#pragma unroll
  for (int g = 0; g < PER_LANE; g += 4) {
    const int c = (g / 4) * 128 + lane * 4;
    const float4 a = *reinterpret_cast<const float4*>(xe + row * D + c);
    const float4 p = *reinterpret_cast<const float4*>(pe + ((long)b * Wq + wq) * D + c);
    const float4 f = *reinterpret_cast<const float4*>(fpos + (long)hq * D + c);
    v[g] = (a.x + p.x) + f.x; v[g + 1] = (a.y + p.y) + f.y; v[g + 2] = (a.z + p.z) + f.z; v[g + 3] = (a.w + p.w) + f.w;
    *reinterpret_cast<float4*>(x + row * D + c) = make_float4(v[g], v[g + 1], v[g + 2], v[g + 3]);
  }
  ln_mod_row<PER_LANE>(v, shift, scale, out.p + row * out.stride, out.hi, out.lo, lane, D);
}
void launch_tok_assemble(const float* xe, const float* pe, const float* fpos, float* x, const float* shift,
                         const float* scale, SView out, int B, int Fq, int Wq, int D, cudaStream_t st) {
  const long M = (long)B * Fq * Wq;
  if (D == 256) launch_pdl(k_tok_assemble<8>, dim3((unsigned)(cdiv(M * 32, 256))), dim3(256), 0, st, xe, pe, fpos, x, shift, scale, out, B, Fq, Wq, D);
  else if (D == 384) launch_pdl(k_tok_assemble<12>, dim3((unsigned)(cdiv(M * 32, 256))), dim3(256), 0, st, xe, pe, fpos, x, shift, scale, out, B, Fq, Wq, D);
}

// ------------------------------------------------------------------------------------------------
// attention softmax: one block per score row (N keys), row staged in shared memory.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_attn_softmax(const float* __restrict__ scores, long NS, bf16* __restrict__ P_,
                                                      long NP, int N) {
  pdl_wait();
  extern __shared__ float rowbuf[];
  __shared__ float red[8];
  const long row = blockIdx.x;
  const float* sp = scores + row * NS;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float mx = -INFINITY;
  for (int j = tid; j < N; j += 256) { const float v = sp[j]; rowbuf[j] = v; mx = fmaxf(mx, v); }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  mx = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) mx = fmaxf(mx, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int j = tid; j < N; j += 256) { const float e = expf(rowbuf[j] - mx); rowbuf[j] = e; sum += e; }
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.f / sum;
  bf16* op = P_ + row * (2 * NP);
  for (long j = tid; j < NP; j += 256) {
    const float v = (j < N) ? rowbuf[j] * inv : 0.f;
    split2(v, op[j], op[NP + j]);
  }
}
void launch_attn_softmax(const float* scores, long NS, bf16* P_, long NP, long rows, int N, cudaStream_t st) {
  launch_pdl(k_attn_softmax, dim3((unsigned)((unsigned)rows)), dim3(256), (size_t)N * sizeof(float), st, scores, NS, P_, NP, N);
}

// ------------------------------------------------------------------------------------------------
// V of the attention: qkv rows hold v token-major; the P.V GEMM wants it as a K-major B operand, i.e. transposed
// vT[b][head][d][hi(NP)|lo(NP)].  32x32 bf16 tiles through shared memory, coalesced on both sides.
// grid = (ceil(N/32), D/32, B*2 [hi|lo]), block = (32, 8).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_transpose_v(const bf16* __restrict__ qkv, long row_stride, int v_hi, int v_lo,
                                                     bf16* __restrict__ vT, int N, long NP, int D, int hd) {
  pdl_wait();
  __shared__ bf16 tile[32][34];
  const int b = blockIdx.z >> 1, part = blockIdx.z & 1;
  const int t0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int src_col = (part ? v_lo : v_hi) + c0;
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int t = t0 + r;
    tile[r][threadIdx.x] = (t < N) ? qkv[((long)b * N + t) * row_stride + src_col + threadIdx.x] : __float2bfloat16_rn(0.f);
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += 8) {
    const int c = c0 + r;                                   // channel = head * hd + d
    const int t = t0 + threadIdx.x;
    if (t < N) vT[((long)b * D + c) * (2 * NP) + (part ? NP : 0) + t] = tile[threadIdx.x][r];
  }
}
void launch_transpose_v(const bf16* qkv, long row_stride, int v_hi, int v_lo, bf16* vT, int B, int N, long NP, int D, int hd,
                        cudaStream_t st) {
  dim3 grid(cdiv(N, 32), D / 32, B * 2), block(32, 8);
  launch_pdl(k_transpose_v, grid, block, 0, st, qkv, row_stride, v_hi, v_lo, vT, N, NP, D, hd);
}

// ------------------------------------------------------------------------------------------------
// unpatchify 'B (h w) (p1 p2 C) -> B C (h p1) (w p2)', crop to W, mask (dit.py:452-457,516-517) -> S (NHWC)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_unpatchify(const float* __restrict__ y, SView out,
                                                    const float* __restrict__ mask1, int B, int Fq, int Wq, int s,
                                                    int C, int H, int W) {
  pdl_wait();
  const int cpt = C / 8;
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * H * W * cpt;
  if (gid >= total) return;
  const int c0 = (int)(gid % cpt) * 8;
  const long pix = gid / cpt;
  const int w = (int)(pix % W), h = (int)((pix / W) % H), b = (int)(pix / ((long)W * H));
  const int hq = h / s, p1 = h % s, wq = w / s, p2 = w % s;
  const float m = mask1[(long)b * W + w];
  float v[8];
  if (hq < Fq && wq < Wq) {
    const float* q = y + (((long)b * Fq + hq) * Wq + wq) * ((long)s * s * C) + (p1 * s + p2) * C + c0;
    const float4 r0 = *reinterpret_cast<const float4*>(q);
    const float4 r1 = *reinterpret_cast<const float4*>(q + 4);
    v[0] = r0.x * m; v[1] = r0.y * m; v[2] = r0.z * m; v[3] = r0.w * m;
    v[4] = r1.x * m; v[5] = r1.y * m; v[6] = r1.z * m; v[7] = r1.w * m;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  bf16* op = out.p + pix * out.stride + c0;
  store_split8(op + out.hi, op + out.lo, v);
}
void launch_unpatchify(const float* y, SView out, const float* mask1, int B, int Fq, int Wq, int s, int C, int H, int W,
                       cudaStream_t st) {
  const long total = (long)B * H * W * (C / 8);
  launch_pdl(k_unpatchify, dim3((unsigned)(cdiv(total, 256))), dim3(256), 0, st, y, out, mask1, B, Fq, Wq, s, C, H, W);
}

// ------------------------------------------------------------------------------------------------
// tiny linears for the per-step tables: one warp per output element
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ float act_f(float x, int a) { return a == 1 ? mish_f(x) : (a == 2 ? silu_f(x) : x); }

__global__ void __launch_bounds__(256) k_small_linear(const float* __restrict__ in, long in_stride,
                                                      const float* __restrict__ w, const float* __restrict__ b,
                                                      float* __restrict__ out, long out_stride, int R, int N, int K,
                                                      int act_in, int act_out) {
  pdl_wait();
  const long gw = (blockIdx.x * (long)blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long)R * N) return;
  const int r = (int)(gw / N), n = (int)(gw % N);
  float acc = 0.f;
  for (int k = lane; k < K; k += 32) acc = fmaf(act_f(in[r * in_stride + k], act_in), w[(long)n * K + k], acc);
  acc = warp_sum(acc);
  if (lane == 0) out[r * out_stride + n] = act_f(acc + (b != nullptr ? b[n] : 0.f), act_out);
}
void launch_small_linear(const float* in, long in_stride, const float* w, const float* b, float* out, long out_stride,
                         int R, int N, int K, int act_in, int act_out, cudaStream_t st) {
  launch_pdl(k_small_linear, dim3((unsigned)(cdiv((long)R * N * 32, 256))), dim3(256), 0, st, in, in_stride, w, b, out, out_stride, R, N, K, act_in, act_out);
}

// mode 0: SinusoidalPosEmb (diffusion.py:113-120): arg = scale * t * exp(-k * ln(1e4)/(half-1)), out = [sin | cos]
// mode 1: timestep_embedding (dit.py:233-251):      arg = t * exp(-ln(1e4) * k / half),        out = [cos | sin]
__global__ void k_time_embed(const StepScalars* __restrict__ tab, int steps, float* __restrict__ out, int dim,
                             float scale, int mode, float c0) {
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int half = dim / 2;
  if (i >= steps * half) return;
  const int s = i / half, k = i % half;
  const float t = tab[s].c_noise;
  if (mode == 0) {
    const float f = expf((float)k * -c0);              // c0 = (float)(ln(1e4) / (half - 1)), rounded on the host
    const float arg = (scale * t) * f;
    out[s * dim + k] = sinf(arg);
    out[s * dim + half + k] = cosf(arg);
  } else {
    const float f = expf((c0 * (float)k) / (float)half);   // c0 = (float)(-ln(1e4))
    const float arg = t * f;
    out[s * dim + k] = cosf(arg);
    out[s * dim + half + k] = sinf(arg);
  }
}
void launch_time_embed(const StepScalars* tab, int steps, float* out, int dim, float scale, int mode, cudaStream_t st) {
  const int half = dim / 2;
  const float c0 = (mode == 0) ? (float)(log(10000.0) / (double)(half - 1)) : (float)(-log(10000.0));
  launch_pdl(k_time_embed, dim3((unsigned)(cdiv((long)steps * half, 128))), dim3(128), 0, st, tab, steps, out, dim, scale, mode, c0);
}

// ------------------------------------------------------------------------------------------------
// weight packers (run once per load)
// ------------------------------------------------------------------------------------------------
__global__ void k_pack_split(const float* __restrict__ w, long ld, bf16* __restrict__ out, long out_stride, int lo_off,
                             int N, int K) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)N * K) return;
  const int n = (int)(i / K), k = (int)(i % K);
  bf16* row = out + (long)n * out_stride;
  split2(w[(long)n * ld + k], row[k], row[lo_off + k]);
}
void launch_pack_split(const float* w, long ld, bf16* out, long out_stride, int lo_off, int N, int K, cudaStream_t st) {
  k_pack_split<<<cdiv((long)N * K, 256), 256, 0, st>>>(w, ld, out, out_stride, lo_off, N, K);
}

__global__ void k_pack_conv(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Ci, int KH, int KW) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)Co * Ci * KH * KW;
  if (i >= total) return;
  int t = (int)i;
  const int kx = t % KW; t /= KW;
  const int ky = t % KH; t /= KH;
  const int ci = t % Ci; t /= Ci;
  const int co = t;
  const int tap = ky * KW + kx;
  bf16* row = out + ((long)tap * Co + co) * (2 * Ci);
  split2(w[i], row[ci], row[Ci + ci]);
}
void launch_pack_conv(const float* w, bf16* out, int Co, int Ci, int KH, int KW, cudaStream_t st) {
  k_pack_conv<<<cdiv((long)Co * Ci * KH * KW, 256), 256, 0, st>>>(w, out, Co, Ci, KH, KW);
}

}  // namespace dexb
