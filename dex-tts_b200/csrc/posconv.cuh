// Weight-stationary grouped positional convolution of the DiT patch embedding (posconv.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dexb {

struct PosConvParams {
  int B, Fq, Wq, D, G, KS;
  int XTn;                   // 72-pixel output tiles per grid row
  int total_tiles;           // B * Fq * G * XTn
  const bf16* w;             // [g][ky][kx/4][(kx%4)*32 + n][hi(32 ci) | lo(32 ci)]
  const float* bias;         // [D]
  float* out;                // GELU(conv + bias), fp32 [b][y][x][D]
};

struct PosConvPlan {
  CUtensorMap tmIn;          // input [b*G + g][y][x][hi(32) | lo(32)] bf16
  CUtensorMap tmW;           // weights [(g, ky, kx/4, kx%4, n)][hi(32 ci) | lo(32 ci)] bf16
  PosConvParams p;
  int grid;
};

int posconv_global_init();
bool posconv_supported(int hidden, int groups, int ks);
int posconv_plan_init(PosConvPlan* pp, const bf16* pin, const bf16* pw, const float* bias, float* out, int B, int Fq, int Wq,
                      int D, int G, int KS);
int posconv_launch(const PosConvPlan& pp, cudaStream_t st);
double posconv_flop(const PosConvPlan& pp);
void launch_posconv_pack_in(const float* xe, bf16* pin, int B, int Fq, int Wq, int D, int G, cudaStream_t st);
void launch_posconv_pack_w(const float* w, bf16* pw, int G, int KS, cudaStream_t st);

}  // namespace dexb
