// Fused attention kernel of the DiT blocks (attn.cu).
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace dexb {

struct AttnParams {
  int N, NP;                 // tokens per sample, padded token count of the V^T rows
  int nheads, hid;           // heads, hidden size (qkv rows are [hi(3*hid) | lo(3*hid)], q | k | v, head-major inside)
  int nt;                    // key tiles of 64
  const bf16* qkv;           // the qkv rows themselves (Q is staged to tensor memory by the softmax warps)
  float scale_log2e;         // hd^-0.5 * log2(e)
  bf16* out;                 // split rows [hi(hid) | lo(hid)], head h at column h*hd
  long out_stride;
  int out_hi, out_lo;
};

struct AttnPlan {
  CUtensorMap tmQ, tmK, tmV;
  AttnParams p;
  int B;
};

int attn_global_init();
bool attn_supported(int hd);
int attn_plan_init(AttnPlan* ap, const bf16* qkv, const bf16* vT, bf16* out, int B, int N, int NP, int heads, int hid);
int attn_launch(const AttnPlan& ap, cudaStream_t st);
double attn_flop(const AttnPlan& ap);

}  // namespace dexb
