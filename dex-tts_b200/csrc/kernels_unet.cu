// U-Net side kernels: first conv, GroupNorm-apply/Mish fusions, EDM update, LinearAttention reductions.
// Reference semantics: DEX-TTS/model/diffusion.py:44-105,190-236 and DEX-TTS/model/edm.py:88-98,185-203.
#include "kernels.cuh"

namespace dexb {

// ------------------------------------------------------------------------------------------------
__global__ void k_fill_zero(uint4* p, size_t n16) {
  pdl_wait();
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n16; i += stride) p[i] = make_uint4(0, 0, 0, 0);
}
void launch_fill_zero(void* p, size_t bytes, cudaStream_t st) {
  const size_t n16 = bytes / 16;                       // callers keep scratch regions 16 B granular
  int blocks = (int)((n16 + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  launch_pdl(k_fill_zero, dim3(blocks), dim3(256), 0, st, reinterpret_cast<uint4*>(p), n16);
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
// Thread = 8 output channels (one GroupNorm group at C == 64) x 4 x-adjacent pixels: the 144 weights of its channel octet
// come from shared memory as 36 LDS.128 for 576 FMAs (the one-pixel-per-thread version issued one LDS per FMA and was
// shared-memory bound: 57 us for 0.75 GFLOP).  8 consecutive lanes write the 256 B row of a pixel.
// A block walks `rows` consecutive rows of its 128-pixel (C = 64) column strip with a rolling three-row window of the inputs: the
// weights are staged once, every new row costs one row of loads instead of three, and the GroupNorm sums leave the block once.  (One
// row per block was bound by the block's own latency chain -- weight staging, loads, reduction, atomics: 2560 blocks in 5.8 waves,
// 60 us for an 84 MB write.)
template <int C, int CIN>
__global__ void __launch_bounds__(256) k_conv_in(const float* __restrict__ x, const float* __restrict__ mu,
                                                 const float* __restrict__ spk_s,
                                                 const float* __restrict__ mask, const StepScalars* __restrict__ tab,
                                                 int step, const float* __restrict__ w, const float* __restrict__ bias,
                                                 float* __restrict__ raw, double* __restrict__ stats, int B, int H,
                                                 int W, int rows) {
  static_assert(C == 64 || C == 128, "GroupNorm(8, C): one or two channel octets per group");
  pdl_wait();
  constexpr int KT = 9 * CIN;                                // taps x input channels: [mu | c_in x | speaker channel]
  constexpr int OCTS = C / 8, QUADS = 256 / OCTS;            // a block covers QUADS * 4 pixels of one row: 128 (C = 64) or 64 (C = 128)
  __shared__ __align__(16) float wsT[KT][C];                 // [k][co]
  __shared__ double red[8][OCTS][2];
  for (int i = threadIdx.x; i < C * KT; i += 256) wsT[i % KT][i / KT] = w[i];
  __syncthreads();
  const int h0 = blockIdx.y * rows, b = blockIdx.z;
  const int oct = threadIdx.x % OCTS, quad = threadIdx.x / OCTS;
  const int c0 = oct * 8;
  const int x0 = blockIdx.x * (QUADS * 4) + quad * 4;
  const float c_in = tab[step].c_in;
  float mk[6];                                               // mask at columns x0-1 .. x0+4 (0 outside the image)
#pragma unroll
  for (int col = 0; col < 6; ++col) {
    const int ww = x0 + col - 1;
    mk[col] = (ww >= 0 && ww < W) ? mask[(long)b * W + ww] : 0.f;
  }
  float va[3][6], vc[3][6], vs[3][6];                        // mu*mask, c_in*x*mask (and spk*mask) at rows h-1..h+1, columns x0-1..x0+4
  auto load_row = [&](int hh, float (&ra)[6], float (&rc)[6], float (&rs)[6]) {
    const bool rok = hh >= 0 && hh < H;
    const float sp = (CIN == 3 && rok) ? spk_s[b * H + hh] : 0.f;
#pragma unroll
    for (int col = 0; col < 6; ++col) {
      const int ww = x0 + col - 1;
      float a = 0.f, c = 0.f;
      if (rok && ww >= 0 && ww < W) {
        const long idx = ((long)b * H + hh) * W + ww;
        a = mu[idx] * mk[col];
        c = (c_in * x[idx]) * mk[col];
      }
      ra[col] = a; rc[col] = c; rs[col] = sp * mk[col];
    }
  };
  load_row(h0 - 1, va[0], vc[0], vs[0]);
  load_row(h0, va[1], vc[1], vs[1]);
  float bs[8];
  {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + c0)), b1 = __ldg(reinterpret_cast<const float4*>(bias + c0 + 4));
    bs[0] = b0.x; bs[1] = b0.y; bs[2] = b0.z; bs[3] = b0.w; bs[4] = b1.x; bs[5] = b1.y; bs[6] = b1.z; bs[7] = b1.w;
  }
  double s = 0., ss = 0.;                                    // per-row fp32 sums (32 values) are added up in double
#pragma unroll 1
  for (int r = 0; r < rows; ++r) {
    const int h = h0 + r;
    if (h >= H) break;
    float sr = 0.f, ssr = 0.f;
    load_row(h + 1, va[2], vc[2], vs[2]);
    float acc[4][8];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[p][i] = bs[i];
#pragma unroll
    for (int k = 0; k < KT; ++k) {                           // accumulation order: mu taps, x taps(, speaker taps)
      const float4 w0 = *reinterpret_cast<const float4*>(&wsT[k][c0]);
      const float4 w1 = *reinterpret_cast<const float4*>(&wsT[k][c0 + 4]);
      const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
      const int dy = (k % 9) / 3, dx = k % 3;
#pragma unroll
      for (int p = 0; p < 4; ++p) {
        const float iv = (k < 9) ? va[dy][p + dx] : ((k < 18) ? vc[dy][p + dx] : vs[dy][p + dx]);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[p][i] = fmaf(wv[i], iv, acc[p][i]);
      }
    }
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      const int wcol = x0 + p;
      if (wcol < W) {
        float* orow = raw + (((long)b * H + h) * W + wcol) * C + c0;
        *reinterpret_cast<float4*>(orow) = make_float4(acc[p][0], acc[p][1], acc[p][2], acc[p][3]);
        *reinterpret_cast<float4*>(orow + 4) = make_float4(acc[p][4], acc[p][5], acc[p][6], acc[p][7]);
#pragma unroll
        for (int i = 0; i < 8; ++i) { sr += acc[p][i]; ssr += acc[p][i] * acc[p][i]; }
      }
    }
    s += (double)sr; ss += (double)ssr;
#pragma unroll
    for (int col = 0; col < 6; ++col) {                      // roll the window down one row
      va[0][col] = va[1][col]; va[1][col] = va[2][col];
      vc[0][col] = vc[1][col]; vc[1][col] = vc[2][col];
      vs[0][col] = vs[1][col]; vs[1][col] = vs[2][col];
    }
  }
  if (OCTS == 8) { s += __shfl_xor_sync(0xffffffffu, s, 8);  ss += __shfl_xor_sync(0xffffffffu, ss, 8); }   // lanes of one octet
  s += __shfl_xor_sync(0xffffffffu, s, 16); ss += __shfl_xor_sync(0xffffffffu, ss, 16);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane < OCTS) { red[warp][lane][0] = s; red[warp][lane][1] = ss; }
  __syncthreads();
  if (threadIdx.x < 8) {
    const int g = threadIdx.x;               // GroupNorm(8, C): group g = OCTS / 8 consecutive channel octets
    double ds = 0., dss = 0.;
    for (int wq = 0; wq < 8; ++wq)
      for (int o = 0; o < OCTS / 8; ++o) { ds += red[wq][g * (OCTS / 8) + o][0]; dss += red[wq][g * (OCTS / 8) + o][1]; }
    double* dst = stats + (((long)b * kGnRep + ((blockIdx.x + blockIdx.y) & (kGnRep - 1))) * 8 + g) * 2;
    atomicAdd(dst, ds);
    atomicAdd(dst + 1, dss);
  }
}

void launch_conv_in(const float* x, const float* mu, const float* spk_s, const float* mask, const StepScalars* tab, int step,
                    const float* w, const float* bias, float* raw, double* stats, int B, int H, int W, int C,
                    cudaStream_t st) {
  // stats layout is [B][kGnRep][8 groups][2]; C = decoder.dim in {64, 128} (engine_finalize rejects other widths)
  // rows per block: the fewest that let the whole grid be resident at once (2 blocks of 128 registers per SM) -- one wave, no tail
  int sms = 148, dev = 0;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int strips = cdiv(W, C == 64 ? 128 : 64);
  int rows = 1;
  while (rows < H && (long)cdiv(H, rows) * strips * B > 2L * sms) ++rows;
  dim3 grid(strips, cdiv(H, rows), B);
  if (C == 64) {
    if (spk_s == nullptr) launch_pdl(k_conv_in<64, 2>, grid, dim3(256), 0, st, x, mu, nullptr, mask, tab, step, w, bias, raw, stats, B, H, W, rows);
    else launch_pdl(k_conv_in<64, 3>, grid, dim3(256), 0, st, x, mu, spk_s, mask, tab, step, w, bias, raw, stats, B, H, W, rows);
  } else if (C == 128) {
    if (spk_s == nullptr) launch_pdl(k_conv_in<128, 2>, grid, dim3(256), 0, st, x, mu, nullptr, mask, tab, step, w, bias, raw, stats, B, H, W, rows);
    else launch_pdl(k_conv_in<128, 3>, grid, dim3(256), 0, st, x, mu, spk_s, mask, tab, step, w, bias, raw, stats, B, H, W, rows);
  }
}

// ------------------------------------------------------------------------------------------------
// GroupNorm-apply + Mish + mask (+ time bias | + residual) -> split-bf16.  8 channels per thread.
// ------------------------------------------------------------------------------------------------
// The item code (loads, statistics, Mish, residual variants, split store) lives in gn_apply.cuh: it is shared with the epilogue
// warps of the tcgen05 convolution kernel, which apply an image in-kernel (gemm.cuh, GNF).
//
// grid = (chunks of ITEMS * 256 eight-channel groups, image): all index math is 32-bit with shifts (C/8 is a power of two) --
// the first version derived (image, pixel, column) from a flat 64-bit index with four 64-bit divisions per item and was bound
// by that integer code (2.7 TB/s), not by memory.  Every thread has ITEMS independent 32 B loads (+ residual) in flight;
// 256 % (C/8) == 0, so a thread keeps the same channel group c0 for all its items (gamma / beta / time bias live in registers).
// A block covers ROUNDS x ITEMS x 256 consecutive groups: the per-thread prologue (affine coefficients, statistics in double) was a
// third of the instructions of a two-item thread, and the kernel is issue-bound (ncu: 25.4 M warp instructions, 61 % SM busy at
// 40 % of the HBM peak), so it is amortised over up to eight items; the residual variant is a template parameter and the image column
// of a pixel (mask index) is advanced by the fixed pixel step of a round instead of a modulo per item.
template <int ITEMS, int KIND>
__global__ void __launch_bounds__(256, 3) k_gn_apply(const GnApplyArgs a, const int rounds) {
  pdl_wait();
  // Images and chunks are walked BACKWARDS: the convolution that produced `raw` wrote image B-1 last, so the tail of the tensor is
  // what the 126 MB L2 still holds; and this pass then leaves image 0 hottest for the next convolution, which starts there.
  const int b = a.reverse ? (int)gridDim.y - 1 - (int)blockIdx.y : (int)blockIdx.y;
  const int cpt = a.C >> 3;                                  // threads per pixel: 8 or 16
  const int cshift = 31 - __clz(cpt);
  const unsigned ngroups = (unsigned)a.P << cshift;          // eight-channel groups per image
  const unsigned base = (a.reverse ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x) * (unsigned)(256 * ITEMS) * (unsigned)rounds;
  const int gs = a.C / a.G;
  const int c0 = (int)(threadIdx.x & (cpt - 1)) * 8;
  float ga[8], be[8], tb[8];
  gn_thread_affine(a, c0, ga, be, tb);
  const long img_row0 = (long)b * a.P;                       // first pixel row of this image
  float mean, rstd;
  gn_thread_stats(a.stats, a.G, 1.0 / ((double)a.P * gs), b, c0 / gs, mean, rstd);
  gn_thread_scale(rstd, ga);
  const int pstep = 256 >> cshift;                           // pixels between consecutive items of a thread
  unsigned gi = base + threadIdx.x;
  int w = (int)((gi >> cshift) % (unsigned)a.W);             // column of the thread's first pixel; then += pstep (mod W) per item
#pragma unroll 1
  for (int r = 0; r < rounds; ++r) {
    GnItem it[ITEMS];
#pragma unroll
    for (int j = 0; j < ITEMS; ++j) {
      const unsigned g = gi + j * 256;
      gn_item_load<false, KIND>(a, img_row0 + ((g < ngroups) ? (g >> cshift) : 0u), c0, it[j]);
    }
#pragma unroll
    for (int j = 0; j < ITEMS; ++j, gi += 256) {
      if (gi < ngroups) gn_item_finish<KIND>(a, b, gi >> cshift, w, c0, it[j], mean, ga, be, tb);
      w += pstep;
      while (w >= a.W) w -= a.W;
    }
  }
}
template <int ITEMS, int KIND>
static void launch_gn_apply_k(const GnApplyArgs& a, long ngroups, cudaStream_t st) {
  // rounds: as many as keep >= ~3 full waves of blocks (148 SMs x 3 resident blocks), at most 4
  const long blocks1 = cdiv(ngroups, ITEMS * 256) * a.B;
  int rounds = (int)(blocks1 / 1280);
  rounds = rounds < 1 ? 1 : (rounds > 4 ? 4 : rounds);
  dim3 grid(cdiv(ngroups, (long)ITEMS * 256 * rounds), a.B);
  launch_pdl(k_gn_apply<ITEMS, KIND>, grid, dim3(256), 0, st, a, rounds);
}
template <int ITEMS>
static void launch_gn_apply_i(const GnApplyArgs& a, long ngroups, cudaStream_t st) {
  switch (gn_kind_of(a)) {
    case kGnResS: launch_gn_apply_k<ITEMS, kGnResS>(a, ngroups, st); break;
    case kGnResF: launch_gn_apply_k<ITEMS, kGnResF>(a, ngroups, st); break;
    case kGnRin: launch_gn_apply_k<ITEMS, kGnRin>(a, ngroups, st); break;
    default: launch_gn_apply_k<ITEMS, kGnPlain>(a, ngroups, st); break;
  }
}
void launch_gn_apply(const GnApplyArgs& a, cudaStream_t st) {
  const long ngroups = (long)a.P * (a.C / 8);
  if (ngroups >= 8 * 256) launch_gn_apply_i<2>(a, ngroups, st);
  else launch_gn_apply_i<1>(a, ngroups, st);
}

// ------------------------------------------------------------------------------------------------
// final: GN + Mish + mask -> 1x1 conv (C -> 1) + bias -> mask = F_x;  D = c_skip*x + c_out*F_x;
// d = x/sigma - D/sigma;  x <- x + (sigma_next - sigma) * d        (edm.py:97,197,203)
// 8 lanes per pixel (C == 64).
// ------------------------------------------------------------------------------------------------
template <int ITEMS>
__global__ void __launch_bounds__(256, 3) k_gn_final(const float* __restrict__ raw, int C, int G,
                                                  const double* __restrict__ stats, const float* __restrict__ gamma,
                                                  const float* __restrict__ beta, const float* __restrict__ fc_w,
                                                  const float* __restrict__ fc_b, const float* __restrict__ mask,
                                                  float* __restrict__ x, float* __restrict__ den_out,
                                                  const StepScalars* __restrict__ tab, int step, int B, int P, int W) {
  // grid = (chunks of ITEMS * 256 eight-channel groups, image); C / 8 = 8 or 16 lanes per pixel (32-bit index math, see k_gn_apply)
  pdl_wait();
  const int b = blockIdx.y;
  const int cpt = C >> 3;
  const int cshift = 31 - __clz(cpt);
  const unsigned ngroups = (unsigned)P << cshift;
  const unsigned base = blockIdx.x * (unsigned)(256 * ITEMS);
  const int c0 = (int)(threadIdx.x & (cpt - 1)) * 8;
  const int g = c0 / (C / G);
  float ga[8], be[8], fw[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { ga[i] = gamma[c0 + i]; be[i] = beta[c0 + i]; fw[i] = fc_w[c0 + i]; }
  const long img_row0 = (long)b * P;
  float mean, rstd;                                          // (before the item loads: the replica sums need registers of their own)
  gn_thread_stats(stats, G, 1.0 / ((double)P * (C / G)), b, g, mean, rstd);
  gn_thread_scale(rstd, ga);
  float4 r0[ITEMS], r1[ITEMS];
  float xin[ITEMS];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const unsigned gi = base + j * 256 + threadIdx.x;
    const float* rp = raw + (img_row0 + ((gi < ngroups) ? (gi >> cshift) : 0u)) * C + c0;
    r0[j] = *reinterpret_cast<const float4*>(rp);
    r1[j] = *reinterpret_cast<const float4*>(rp + 4);
    xin[j] = (gi < ngroups && c0 == 0) ? x[img_row0 + (gi >> cshift)] : 0.f;
  }
  const StepScalars sc = tab[step];
  const float fcb = fc_b[0];
#pragma unroll
  for (int j = 0; j < ITEMS; ++j) {
    const unsigned gi = base + j * 256 + threadIdx.x;
    const bool active = gi < ngroups;
    const unsigned p = active ? (gi >> cshift) : 0u;
    const long pix = img_row0 + p;
    float part = 0.f;
    float m = 0.f;
    if (active) {
      m = mask[(long)b * W + (int)(p % (unsigned)W)];
      const float v[8] = {r0[j].x, r0[j].y, r0[j].z, r0[j].w, r1[j].x, r1[j].y, r1[j].z, r1[j].w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float y = mish_fast(fmaf(v[i] - mean, ga[i], be[i])) * m;
        part = fmaf(fw[i], y * m, part);
      }
    }
    // reduce over the 8 / 16 lanes of a pixel (aligned groups of lanes)
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (cpt == 16) part += __shfl_xor_sync(0xffffffffu, part, 8);
    if (active && c0 == 0) {
      const float fx = (part + fcb) * m;
      const float xv = xin[j];
      const float den = sc.c_skip * xv + sc.c_out * fx;
      const float inv = 1.f / sc.sigma;
      const float d = inv * xv - inv * den;
      if (den_out != nullptr) den_out[pix] = den;
      else x[pix] = xv + (sc.sigma_next - sc.sigma) * d;
    }
  }
}
void launch_gn_final(const float* raw, int C, int G, const double* stats, const float* gamma, const float* beta,
                     const float* fc_w, const float* fc_b, const float* mask, float* x, float* den_out,
                     const StepScalars* tab, int step, int B, int H, int W, cudaStream_t st) {
  const long ngroups = (long)H * W * (C / 8);               // C in {64, 128} (engine_finalize)
  if (ngroups >= 8 * 256) {
    dim3 grid(cdiv(ngroups, 2 * 256), B);
    launch_pdl(k_gn_final<2>, grid, dim3(256), 0, st, raw, C, G, stats, gamma, beta, fc_w, fc_b, mask, x, den_out, tab, step, B, H * W, W);
  } else {
    dim3 grid(cdiv(ngroups, 256), B);
    launch_pdl(k_gn_final<1>, grid, dim3(256), 0, st, raw, C, G, stats, gamma, beta, fc_w, fc_b, mask, x, den_out, tab, step, B, H * W, W);
  }
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
  pdl_wait();
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
  launch_pdl(k_la_colmax, grid, dim3(128), 0, st, kv, kmax_enc, P, chunk);
}

// ctx[b][h][d][e] += sum_n exp(k[n][h*32+d] - max) * v[n][h*32+e];  ssum[b][h*32+d] += sum_n exp(..)
// One block (256 threads) = one (b, pixel chunk), all 4 heads: the full 1 KiB kv row of every pixel is read once,
// coalesced; thread (h, dq, eq) keeps a 4x4 register block of the 32x32 context of head h (16 FMA per 2 LDS.128).
constexpr int kLaTile = 32;
constexpr int kLaPartial = 4096 + 128;       // one block's partial: ctx[4][32][32] | ssum[128]
__global__ void __launch_bounds__(256) k_la_ctx(const float* __restrict__ kv, const unsigned* __restrict__ kmax,
                                                float* __restrict__ part, int P, int chunk) {
  pdl_wait();
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
  pdl_wait();
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
  launch_pdl(k_la_ctx, grid, dim3(256), 0, st, kv, kmax_enc, part, P, chunk);
  launch_pdl(k_la_reduce, dim3((unsigned)(cdiv((long)B * kLaPartial, 256))), dim3(256), 0, st, part, ctx, ssum, B, nblk);
}

// Merge the split-KV partials of the tensor-core context kernel (attn.cu, out_mode 2) in split order, then finish the
// reassociated context:  the kernel produced G = softmax(k)^T x per key range (128 x C, un-normalised), so
//   M[d] = max_s m_s[d];  G[d][c] = sum_s G_s[d][c] exp(m_s[d] - M[d]);  ssum[b][d] = sum_s l_s[d] exp(m_s[d] - M[d]);
//   ctx[b][h][d][e] = sum_c G[h*32+d][c] W_v[h*32+e][c]          (= softmax(k)^T v restricted to the 4 diagonal head blocks).
// One block of 256 threads per (head, group of 8 rows, image): four thread groups walk the key ranges s = g, g + 4, ... with their own
// running maximum, then merge in the fixed order g = 0 .. 3 (deterministic); thread = (group, row d, a contiguous eighth of the C columns).
// (A single utterance is cut into up to 148 key ranges: one group of 64 threads walked them in a chain of dependent L2 round trips --
// 28 us per call at C1, 7.6 % of its step.)
template <int C>
__global__ void __launch_bounds__(256) k_la_combine(const float* __restrict__ part_o, const float* __restrict__ part_l,
                                                     const float* __restrict__ part_m, const float* __restrict__ wv,
                                                     float* __restrict__ ctx, float* __restrict__ ssum, int S) {
  pdl_wait();
  __shared__ float Gh[8][C + 1];
  __shared__ float Wvs[32][C + 1];
  __shared__ float sM[4][8], sL[4][8];
  __shared__ float sAcc[3][8][C];                             // partial sums of groups 1 .. 3
  const int h = blockIdx.x >> 2, rg = blockIdx.x & 3, b = blockIdx.y, tid = threadIdx.x;
  const int grp = tid >> 6, t = tid & 63;
  const int dl = t >> 3, seg = t & 7;
  const int d = rg * 8 + dl;
  constexpr int CS = C / 8;                                  // columns per thread: 8 or 16
  const int row = h * 32 + d;
  {
    const float4* wsrc = reinterpret_cast<const float4*>(wv + (long)h * 32 * C);         // 32 x C contiguous floats, requested first
#pragma unroll
    for (int j = 0; j < 32 * C / 4 / 256; ++j) {
      const int i4 = tid + j * 256;
      const float4 q = __ldg(wsrc + i4);
      const int r_ = (i4 * 4) / C, c_ = (i4 * 4) % C;
      Wvs[r_][c_] = q.x; Wvs[r_][c_ + 1] = q.y; Wvs[r_][c_ + 2] = q.z; Wvs[r_][c_ + 3] = q.w;
    }
  }
  float M = -INFINITY;
#pragma unroll 8
  for (int s = grp; s < S; s += 4) M = fmaxf(M, part_m[((long)b * S + s) * 128 + row]);
  float acc[CS];
#pragma unroll
  for (int j = 0; j < CS; ++j) acc[j] = 0.f;
  float l = 0.f;
  if (M > -INFINITY) {                                       // (a group whose ranges are all empty contributes nothing)
#pragma unroll 4
    for (int s = grp; s < S; s += 4) {
      const long pi = ((long)b * S + s) * 128 + row;
      const float w = expf(part_m[pi] - M);                  // exp(-inf) = 0 for an empty split
      l = fmaf(part_l[pi], w, l);
      const float4* po = reinterpret_cast<const float4*>(part_o + pi * 128 + seg * CS);
#pragma unroll
      for (int j = 0; j < CS / 4; ++j) {
        const float4 v = po[j];
        acc[4 * j] = fmaf(v.x, w, acc[4 * j]); acc[4 * j + 1] = fmaf(v.y, w, acc[4 * j + 1]);
        acc[4 * j + 2] = fmaf(v.z, w, acc[4 * j + 2]); acc[4 * j + 3] = fmaf(v.w, w, acc[4 * j + 3]);
      }
    }
  }
  if (seg == 0) sM[grp][dl] = M;
  __syncthreads();
  const float Mt = fmaxf(fmaxf(sM[0][dl], sM[1][dl]), fmaxf(sM[2][dl], sM[3][dl]));
  const float f = (M > -INFINITY) ? expf(M - Mt) : 0.f;      // rescale this group's sums to the common maximum
#pragma unroll
  for (int j = 0; j < CS; ++j) acc[j] *= f;
  l *= f;
  if (grp > 0) {
#pragma unroll
    for (int j = 0; j < CS; ++j) sAcc[grp - 1][dl][seg * CS + j] = acc[j];
  }
  if (seg == 0) sL[grp][dl] = l;
  __syncthreads();
  if (grp == 0) {
#pragma unroll
    for (int j = 0; j < CS; ++j)
      Gh[dl][seg * CS + j] = ((acc[j] + sAcc[0][dl][seg * CS + j]) + sAcc[1][dl][seg * CS + j]) + sAcc[2][dl][seg * CS + j];
    if (seg == 0) ssum[b * 128 + row] = ((sL[0][dl] + sL[1][dl]) + sL[2][dl]) + sL[3][dl];
  }
  __syncthreads();
  {
    // 8 rows x 32 columns of the context block: one output per thread
    const int dd = tid >> 5, e = tid & 31;
    float a = 0.f;
#pragma unroll 8
    for (int c = 0; c < C; ++c) a = fmaf(Gh[dd][c], Wvs[e][c], a);
    ctx[(((long)b * 4 + h) * 32 + rg * 8 + dd) * 32 + e] = a;
  }
}
void launch_la_combine(const float* part_o, const float* part_l, const float* part_m, const float* wv, float* ctx, float* ssum,
                       int B, int S, int C, cudaStream_t st) {
  dim3 grid(16, B);
  if (C == 64) launch_pdl(k_la_combine<64>, grid, dim3(256), 0, st, part_o, part_l, part_m, wv, ctx, ssum, S);
  else launch_pdl(k_la_combine<128>, grid, dim3(256), 0, st, part_o, part_l, part_m, wv, ctx, ssum, S);
}

// W_eff[b][co][ci] = delta(co,ci) + g * sum_{h,e} Wout[co][h*32+e] * sum_d (ctx[b][h][d][e]/ssum[b][h][d]) * Wq[h*32+d][ci]
// (q is linear in x, so  x + g*to_out(ctx^T q) == W_eff x + g*b_out : the whole attention read-out is one
//  per-sample CxC matrix.)
// One block = (sample b, 8 output channels): T[r][h,d] = sum_e W_out[co][h,e] ctxn[b][h][d][e] through shared memory, then
// W_eff[co][ci] = delta + g * sum_{h,d} T[r][h,d] W_q[h,d][ci].  (The first version materialised ctxn^T W_q with two dependent
// launches of 128-step serial loops: 31 us per LinearAttention for 6 MFLOP.)
__global__ void __launch_bounds__(256) k_la_weff(const float* __restrict__ ctx, const float* __restrict__ ssum,
                                                 const float* __restrict__ wq, const float* __restrict__ wout,
                                                 const float* __restrict__ bout, const float* __restrict__ g,
                                                 bf16* __restrict__ weff, float* __restrict__ beff, int C) {
  pdl_wait();
  __shared__ float cn[4][32][33];
  __shared__ float Ts[8][128];
  const int b = blockIdx.y, co0 = blockIdx.x * 8, tid = threadIdx.x;
  // (all 16 loads of a thread in flight at once: as a rolled loop this was 16 dependent round trips, 40 % of the kernel's samples)
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const int i = tid + j * 256;
    const int h = i >> 10, d = (i >> 5) & 31, e = i & 31;
    cn[h][d][e] = ctx[(long)b * 4096 + i] / ssum[b * 128 + h * 32 + d];
  }
  __syncthreads();
  for (int o = tid; o < 1024; o += 256) {
    const int r = o >> 7, hd = o & 127, h = hd >> 5, d = hd & 31;
    const float* wp = wout + (long)(co0 + r) * 128 + h * 32;
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 32; ++e) acc = fmaf(__ldg(wp + e), cn[h][d][e], acc);
    Ts[r][hd] = acc;
  }
  __syncthreads();
  // thread = input channel ci and rows r0, r0 + 256/C, ...: one W_q column serves all its outputs; 32 independent loads in flight
  const float gg = g[0];
  const int ci = tid % C, r0 = tid / C, rstep = 256 / C;     // C == 64: rows r0, r0+4;  C == 128: rows r0, r0+2, r0+4, r0+6
  float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  const int nr = 8 / rstep;                                  // 2, 4 or 8 (C = 256: every thread owns one input channel and all 8 rows)
#pragma unroll 4
  for (int h0 = 0; h0 < 128; h0 += 32) {
    float wv[32];
#pragma unroll
    for (int k = 0; k < 32; ++k) wv[k] = __ldg(wq + (long)(h0 + k) * C + ci);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (i < nr) {
#pragma unroll
        for (int k = 0; k < 32; ++k) acc[i] = fmaf(Ts[r0 + i * rstep][h0 + k], wv[k], acc[i]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    if (i >= nr) break;
    const int co = co0 + r0 + i * rstep;
    const float v = gg * acc[i] + (co == ci ? 1.f : 0.f);
    bf16* row = weff + ((long)b * C + co) * (2 * C);
    split2(v, row[ci], row[C + ci]);
    if (ci == 0) beff[b * C + co] = gg * bout[co];
  }
}
int kernels_global_init() { return 0; }
void launch_la_weff(const float* ctx, const float* ssum, const float* wq, const float* wout, const float* bout,
                    const float* g, bf16* weff, float* beff, int B, int C, cudaStream_t st) {
  dim3 grid(C / 8, B);
  launch_pdl(k_la_weff, grid, dim3(256), 0, st, ctx, ssum, wq, wout, bout, g, weff, beff, C);
}

// ------------------------------------------------------------------------------------------------
// per-(image, channel) sum / sumsq (InstanceNorm2D statistics, base.py:95-103)
// block = 256 threads = 32 pixel-lanes x (C/8 <= 16 ... ) ; simple: thread handles 8 channels of a pixel stream
// ------------------------------------------------------------------------------------------------
template <bool SPLIT>
__global__ void __launch_bounds__(256) k_chan_stats(const bf16* __restrict__ xs, long s_stride, int hi, int lo,
                                                    const float* __restrict__ xf, long f_stride,
                                                    double* __restrict__ stats, int P, int C, int chunk) {
  pdl_wait();
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
  launch_pdl(k_chan_stats<true>, grid, dim3(256), 0, st, x.p, x.stride, x.hi, x.lo, nullptr, 0, stats, P, C, chunk);
}
void launch_chan_stats_f(const float* x, long stride, double* stats, int B, int P, int C, cudaStream_t st) {
  const int chunk = 128;                                 // 80 -> 640 blocks at C2: the kernel was latency-bound
  dim3 grid(cdiv(P, chunk), B);
  launch_pdl(k_chan_stats<false>, grid, dim3(256), 0, st, nullptr, 0, 0, 0, x, stride, stats, P, C, chunk);
}

}  // namespace dexb
