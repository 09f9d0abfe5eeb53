// Once-per-call / once-per-load helper kernels: reference statistics for the TIV adaptor, self-attention pooling
// tables, layout transposes, pair packing for the positional conv, ConvTranspose / pos-conv weight packers.
// Reference semantics: DEX-TTS/model/diffusion.py:177-188, DEX-TTS/model/base.py:72-78,
// DEX-TTS/model/ref_encoder.py:239-273, DEX-TTS/model/dit.py:75-90.
#include "kernels.cuh"

#include <stdlib.h>

namespace dexb {

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DEXB_PDL");
    v = (e != nullptr && e[0] == '0') ? 0 : 1;
  }
  return v != 0;
}


// ------------------------------------------------------------------------------------------------
// InstanceNorm1D.cal_stats over the time axis (lengths ignored, unbiased variance, std = sqrt(var + 1e-5)).
// ref (B, C, Tr) -> mean/std [B][L][C] at layer slot l.  One warp per (b, c).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_ref_stats(const float* __restrict__ ref, float* __restrict__ mean,
                                                   float* __restrict__ stdv, int B, int C, int Tr, int L, int l) {
  pdl_wait();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B * C) return;
  const int b = warp / C, c = warp % C;
  const float* p = ref + (long)warp * Tr;
  double s = 0.;
  for (int t = lane; t < Tr; t += 32) s += (double)p[t];
  s = warp_sum_d(s);
  const double m = s / Tr;
  double q = 0.;
  for (int t = lane; t < Tr; t += 32) { const double d = (double)p[t] - m; q += d * d; }
  q = warp_sum_d(q);
  if (lane == 0) {
    const double var = q / (double)(Tr - 1);
    mean[((long)b * L + l) * C + c] = (float)m;
    stdv[((long)b * L + l) * C + c] = sqrtf((float)var + 1e-5f);
  }
}
void launch_ref_stats(const float* ref, float* mean, float* stdv, int B, int C, int Tr, int L, int l, cudaStream_t st) {
  launch_pdl(k_ref_stats, dim3((unsigned)(cdiv((long)B * C * 32, 256))), dim3(256), 0, st, ref, mean, stdv, B, C, Tr, L, l);
}

// ------------------------------------------------------------------------------------------------
// SelfAttentionPooling over [time token ; L reference rows] for every (step, b):
//   a_l = softmax_l(z_l . W + bias),  out[step][b][c] = sum_l a_l z_l[c]      (ref_encoder.py:246-253)
// One block of 128 threads per (step, b); C <= 512.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_tiv_sap(const float* __restrict__ t_tok, const float* __restrict__ rows,
                                                 const float* __restrict__ W, const float* __restrict__ bias,
                                                 float* __restrict__ out, int B, int C, int L) {
  pdl_wait();
  __shared__ float logit[16];
  __shared__ float red[4];
  const int step = blockIdx.x / B, b = blockIdx.x % B;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int l = 0; l <= L; ++l) {
    const float* z = (l == 0) ? t_tok + (long)step * C : rows + ((long)b * L + (l - 1)) * C;
    float acc = 0.f;
    for (int c = tid; c < C; c += 128) acc = fmaf(z[c], W[c], acc);
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (tid == 0) logit[l] = ((red[0] + red[1]) + (red[2] + red[3])) + bias[0];
    __syncthreads();
  }
  float mx = -INFINITY;
  for (int l = 0; l <= L; ++l) mx = fmaxf(mx, logit[l]);
  float den = 0.f;
  for (int l = 0; l <= L; ++l) den += expf(logit[l] - mx);
  for (int c = tid; c < C; c += 128) {
    float acc = 0.f;
    for (int l = 0; l <= L; ++l) {
      const float* z = (l == 0) ? t_tok + (long)step * C : rows + ((long)b * L + (l - 1)) * C;
      acc = fmaf(expf(logit[l] - mx) / den, z[c], acc);
    }
    out[((long)step * B + b) * C + c] = acc;
  }
}
void launch_tiv_sap(const float* t_tok, const float* rows, const float* W, const float* bias, float* out, int steps,
                    int B, int C, int L, cudaStream_t st) {
  launch_pdl(k_tiv_sap, dim3((unsigned)(steps * B)), dim3(128), 0, st, t_tok, rows, W, bias, out, B, C, L);
}

// out[c][r] = in[r][c] * scale
__global__ void k_transpose_scale(const float* __restrict__ in, float* __restrict__ out, int R, int Cc, float scale) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)R * Cc) return;
  const int c = (int)(i / R), r = (int)(i % R);
  out[i] = in[(long)r * Cc + c] * scale;
}
void launch_transpose_scale(const float* in, float* out, int R, int Cc, float scale, cudaStream_t st) {
  launch_pdl(k_transpose_scale, dim3((unsigned)(cdiv((long)R * Cc, 256))), dim3(256), 0, st, in, out, R, Cc, scale);
}

// (B, C, T) -> (B, T, C)
__global__ void k_bct_to_btc(const float* __restrict__ in, float* __restrict__ out, int B, int C, int T) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * C * T) return;
  const int c = (int)(i % C);
  const int t = (int)((i / C) % T);
  const int b = (int)(i / ((long)C * T));
  out[i] = in[((long)b * C + c) * T + t];
}
void launch_bct_to_btc(const float* in, float* out, int B, int C, int T, cudaStream_t st) {
  launch_pdl(k_bct_to_btc, dim3((unsigned)(cdiv((long)B * C * T, 256))), dim3(256), 0, st, in, out, B, C, T);
}

// VL rows of the style tokens (j >= 1) -> transposed split operand vlt[b][c][hi(KP)|lo(KP)] at column j
__global__ void k_tv_vlt_pack(const float* __restrict__ vl, bf16* __restrict__ vlt, int B, int Ts, int C, int KP) {
  pdl_wait();
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * Ts * C) return;
  const int j = (int)(i % Ts);
  const int c = (int)((i / Ts) % C);
  const int b = (int)(i / ((long)Ts * C));
  bf16* row = vlt + ((long)b * C + c) * (2 * KP);
  split2(vl[((long)b * Ts + j) * C + c], row[j + 1], row[KP + j + 1]);
}
void launch_tv_vlt_pack(const float* vl, bf16* vlt, int B, int Ts, int C, int KP, cudaStream_t st) {
  launch_pdl(k_tv_vlt_pack, dim3((unsigned)(cdiv((long)B * Ts * C, 256))), dim3(256), 0, st, vl, vlt, B, Ts, C, KP);
}

// ------------------------------------------------------------------------------------------------
// pair packing for the grouped 16x16 positional conv (K = 32 or 48 per group and tap is too short for a 128 B swizzle
// span, so PF x-adjacent taps are fused into one K = PF * Cg chunk: PF = 2 for Cg = 32 (K = 64), 4 for Cg = 48 (K = 192)):
//   pairs[b][y][xx][g][f*Cg + ci] = e[b][y][xx - (PF - 1) + f][g*Cg + ci],  xx in [0, Wq + PF - 1), zero outside the image.
// Row layout: [hi(G*PF*Cg) | lo(G*PF*Cg)].  One thread per (pixel xx, fused tap f, 8 channels).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pair_pack(const float* __restrict__ e, bf16* __restrict__ pairs, int B, int Fq,
                                                   int Wq, int D, int Cg, int PF) {
  pdl_wait();
  const int cpt = D / 8;
  const int Wp = Wq + PF - 1;
  const long gid = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)B * Fq * Wp * PF * cpt;
  if (gid >= total) return;
  const int c0 = (int)(gid % cpt) * 8;
  long t = gid / cpt;
  const int half = (int)(t % PF); t /= PF;
  const int xx = (int)(t % Wp); t /= Wp;
  const int y = (int)(t % Fq);
  const int b = (int)(t / Fq);
  const int x = xx - (PF - 1) + half;
  float v[8];
  if (x >= 0 && x < Wq) {
    const float* q = e + ((((long)b * Fq + y) * Wq) + x) * D + c0;
    const float4 r0 = *reinterpret_cast<const float4*>(q);
    const float4 r1 = *reinterpret_cast<const float4*>(q + 4);
    v[0] = r0.x; v[1] = r0.y; v[2] = r0.z; v[3] = r0.w; v[4] = r1.x; v[5] = r1.y; v[6] = r1.z; v[7] = r1.w;
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = 0.f;
  }
  const int g = c0 / Cg, ci = c0 % Cg;
  bf16* row = pairs + ((((long)b * Fq + y) * Wp) + xx) * (2L * PF * D);
  const int col = g * PF * Cg + half * Cg + ci;
  store_split8(row + col, row + PF * D + col, v);
}
void launch_pair_pack(const float* e, bf16* pairs, int B, int Fq, int Wq, int D, int Cg, int PF, cudaStream_t st) {
  const long total = (long)B * Fq * (Wq + PF - 1) * PF * (D / 8);
  launch_pdl(k_pair_pack, dim3((unsigned)(cdiv(total, 256))), dim3(256), 0, st, e, pairs, B, Fq, Wq, D, Cg, PF);
}

// pos-conv weight [Co][Cg][KP][KP] -> [tap = ky*(KP/PF) + kx/PF][Co][hi(PF Cg)|lo(PF Cg)], k = (kx % PF)*Cg + ci
__global__ void k_pack_posconv(const float* __restrict__ w, bf16* __restrict__ out, int Co, int Cg, int KP, int PF) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)Co * Cg * KP * KP;
  if (i >= total) return;
  long t = i;
  const int kx = (int)(t % KP); t /= KP;
  const int ky = (int)(t % KP); t /= KP;
  const int ci = (int)(t % Cg);
  const int co = (int)(t / Cg);
  const int tap = ky * (KP / PF) + kx / PF;
  const int k = (kx % PF) * Cg + ci;
  bf16* row = out + ((long)tap * Co + co) * (2L * PF * Cg);
  split2(w[i], row[k], row[PF * Cg + k]);
}
void launch_pack_posconv(const float* w, bf16* out, int Co, int Cg, int KP, int PF, cudaStream_t st) {
  k_pack_posconv<<<cdiv((long)Co * Cg * KP * KP, 256), 256, 0, st>>>(w, out, Co, Cg, KP, PF);
}

// ConvTranspose2d(4x4, stride 2, pad 1) weight [Ci][Co][4][4] -> per output-parity phase (ry, rx) a 2x2-tap conv:
//   out[2m+ry][2n+rx] = sum_{ty,tx} in[m + ty + offH(ry)][n + tx + offW(rx)] . Wp[phase][ty*2+tx]
//   with off(r) = r - 1 and kernel index k(r, t) = 3 - 2t (r = 0) | 2 - 2t (r = 1).
// Packed as [phase = ry*2+rx][tap][Co][hi(Ci)|lo(Ci)].
__global__ void k_pack_convT(const float* __restrict__ w, bf16* __restrict__ out, int Ci, int Co) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = 16L * Ci * Co;
  if (i >= total) return;
  long t = i;
  const int ci = (int)(t % Ci); t /= Ci;
  const int co = (int)(t % Co); t /= Co;
  const int tx = (int)(t % 2); t /= 2;
  const int ty = (int)(t % 2); t /= 2;
  const int rx = (int)(t % 2);
  const int ry = (int)(t / 2);
  const int ky = (ry == 0 ? 3 : 2) - 2 * ty, kx = (rx == 0 ? 3 : 2) - 2 * tx;
  const float v = w[(((long)ci * Co + co) * 4 + ky) * 4 + kx];
  bf16* row = out + ((((long)(ry * 2 + rx) * 4) + (ty * 2 + tx)) * Co + co) * (2L * Ci);
  split2(v, row[ci], row[Ci + ci]);
}
void launch_pack_convT(const float* w, bf16* out, int Ci, int Co, cudaStream_t st) {
  k_pack_convT<<<cdiv(16L * Ci * Co, 256), 256, 0, st>>>(w, out, Ci, Co);
}

// fp32 [rows][K] -> split rows (generic activation packer used by the unit-test entry point)
void launch_pack_rows(const float* in, bf16* out, long rows, int K, cudaStream_t st) {
  launch_pack_split(in, K, out, 2L * K, K, (int)rows, K, st);
}

// split rows [hi(K) | lo(K)] -> fp32 [rows][K] (unit-test helper)
__global__ void k_unpack_rows(const bf16* __restrict__ in, float* __restrict__ out, long rows, int K) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const long r = i / K;
  const int k = (int)(i % K);
  out[i] = join2(in[r * 2 * K + k], in[r * 2 * K + K + k]);
}
void launch_unpack_rows(const bf16* in, float* out, long rows, int K, cudaStream_t st) {
  k_unpack_rows<<<cdiv(rows * K, 256), 256, 0, st>>>(in, out, rows, K);
}


// ------------------------------------------------------------------------------------------------
// Debug tap (test aid, dexb_debug_tap): an internal NHWC activation (S view or F rows) -> fp32 NCHW like the reference's tensors.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_tap_nchw(const bf16* __restrict__ sp, long s_stride, int hi, int lo,
                                                  const float* __restrict__ fp, long f_stride, float* __restrict__ out, int B,
                                                  int C, long P) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  if (i >= (long)B * C * P) return;
  const long p = i % P;
  const int c = (int)((i / P) % C), b = (int)(i / (P * C));
  const long row = (long)b * P + p;
  out[i] = (sp != nullptr) ? join2(sp[row * s_stride + hi + c], sp[row * s_stride + lo + c]) : fp[row * f_stride + c];
}
void launch_tap_nchw(const bf16* sp, long s_stride, int hi, int lo, const float* fp, long f_stride, float* out, int B, int C, long P,
                     cudaStream_t st) {
  k_tap_nchw<<<cdiv((long)B * C * P, 256), 256, 0, st>>>(sp, s_stride, hi, lo, fp, f_stride, out, B, C, P);
}

}  // namespace dexb
