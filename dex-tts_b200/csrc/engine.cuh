// Reverse-diffusion engine: owns the packed weights, the per-plan workspace and the launch sequence of one
// EDM/Euler trajectory (DEX-TTS/model/edm.py:183-209 over DEX-TTS/model/diffusion.py:190-236).
#pragma once
#include <map>
#include <string>
#include <vector>

#include "../../include/dexb200.h"
#include "attn.cuh"
#include "gemm_host.cuh"
#include "kernels.cuh"
#include "posconv.cuh"

namespace dexb {

struct HostTensor {
  float* p = nullptr;                 // device copy, fp32, contiguous
  std::vector<int64_t> shape;
  size_t n = 0;
};

// bump allocator with a measuring pass: layout code runs twice (measure -> cudaMalloc -> assign)
struct Arena {
  char* base = nullptr;
  size_t off = 0;
  template <class T>
  T* get(size_t count) {
    off = (off + 1023) & ~(size_t)1023;           // 1 KiB granularity (TMA bases need 16 B; keep it generous)
    T* r = reinterpret_cast<T*>(base + off);      // base == nullptr while measuring: pointers are offsets, never dereferenced
    off += count * sizeof(T);
    return r;
  }
};

struct BlockW {                       // Block = conv3x3 + GroupNorm(8) (+ Mish)
  bf16* w = nullptr;                  // [9][co][hi(ci)|lo(ci)]
  const float *bias = nullptr, *gamma = nullptr, *beta = nullptr;
  int ci = 0, co = 0;
  GemmPlan conv;                      // raw = conv(in) + bias, GN partial sums
  int slot = 0;                       // GroupNorm statistics slot
};

struct ResnetW {
  BlockW b1, b2;
  const float *mlp_w = nullptr, *mlp_b = nullptr;
  float* tbias = nullptr;             // [steps][co]
  bf16* res_w = nullptr;              // [co][hi(ci)|lo(ci)] or null (identity / 2-channel special case)
  const float *res_b = nullptr, *rin_w = nullptr;
  GemmPlan res;
  int ci = 0, co = 0;
};

struct LinAttW {
  bf16* kv_w = nullptr;               // [256][hi(C)|lo(C)]  (rows 128..383 of to_qkv)
  const float *wq = nullptr, *wv = nullptr, *wout = nullptr, *bout = nullptr, *g = nullptr;
  int C = 0;
  bool fused = false;                 // context on the tensor-core kernel (C <= 128) or the colmax / ctx CUDA-core kernels
  GemmPlan kv, apply;
  AttnPlan ctx_plan;
  int splits = 1, PP = 0;
  float *part_o = nullptr, *part_l = nullptr, *part_m = nullptr;
  unsigned* kmax = nullptr;           // [B][128]
  float *ctx = nullptr, *ssum = nullptr, *beff = nullptr, *part = nullptr;
  bf16* weff = nullptr;               // [B][C][hi(C)|lo(C)]
};

struct DitBlockW {
  bf16 *qkv_w = nullptr, *proj_w = nullptr, *fc1_w = nullptr, *fc2_w = nullptr;
  const float *qkv_b = nullptr, *proj_b = nullptr, *fc1_b = nullptr, *fc2_b = nullptr;
  const float *ada_w = nullptr, *ada_b = nullptr;
  GemmPlan qkv, scores, pv, proj, fc1, fc2;
};

}  // namespace dexb

struct dexb_handle {
  dexb_config cfg;
  std::map<std::string, dexb::HostTensor> w;
  bool finalized = false;
  bool planned = false;
  char* packed_base = nullptr;
  char* ws_base = nullptr;
  size_t ws_bytes = 0;
  long launches = 0;

  // ---- packed / referenced weights ----
  dexb::ResnetW d00, d01, d10, d11, u00, u01;
  dexb::LinAttW la0, la1, la2;
  dexb::BlockW fin;
  dexb::bf16 *down_w = nullptr, *up_w = nullptr;      // conv 3x3 s2 [9][co][..], convT [4][4][co][..]
  const float *down_b = nullptr, *up_b = nullptr;
  const float *conv_in_w = nullptr, *conv_in_b = nullptr;
  int cin = 2;                                         // input channels of the first conv: [mu, x] (+ speaker channel)
  const float *spk_w0 = nullptr, *spk_b0 = nullptr, *spk_w2 = nullptr, *spk_b2 = nullptr;   // spk_mlp (GeDEX-TTS, n_spks > 1)
  float *spk = nullptr, *spk_hid = nullptr, *spk_s = nullptr;   // staged embedding (B, E), hidden (B, 4E), channel values (B, n_feats)
  const float *fc_w = nullptr, *fc_b = nullptr;
  // TV / TIV adaptors
  float *wqT_s = nullptr;                              // [C][C]  W_q^T / sqrt(C)
  const float *tv_wk = nullptr, *tv_wv = nullptr, *tv_wl = nullptr;
  const float *sap_m_w = nullptr, *sap_m_b = nullptr, *sap_s_w = nullptr, *sap_s_b = nullptr;
  // DiT
  const float *dw_w = nullptr, *dw_b = nullptr, *pe_b = nullptr, *posconv_b = nullptr, *final_b = nullptr;
  dexb::bf16 *pe_w = nullptr, *posconv_w = nullptr, *final_w = nullptr;
  float* fpos = nullptr;                               // [Fq][hidden]
  std::vector<dexb::DitBlockW> blocks;

  // ---- plan ----
  int B = 0, T = 0, Ts = 0, steps = 0;
  int H0 = 0, W0 = 0, H1 = 0, W1 = 0, Fq = 0, Wq = 0, Ntok = 0, NP = 0, NK = 0, KP = 0;
  std::vector<dexb::StepScalars> tab_host;
  dexb::StepScalars* tab = nullptr;
  // inputs staged in the workspace (so a captured graph sees fixed addresses)
  float *x = nullptr, *mu = nullptr, *mask0 = nullptr, *mask1 = nullptr;
  float *sty = nullptr, *refs[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  int* sty_len = nullptr;
  int Tr = 0;
  // tables
  float *t_init = nullptr, *t_hid = nullptr, *t_unet = nullptr, *t_adap = nullptr, *t_sty = nullptr, *t_hid2 = nullptr;
  float *temb = nullptr, *tc_hid = nullptr, *tc = nullptr, *mod = nullptr, *fmod = nullptr;
  float *k0 = nullptr, *kw0 = nullptr, *v0 = nullptr, *vl0 = nullptr;
  float *ref_mean = nullptr, *ref_std = nullptr, *tiv_shift = nullptr, *tiv_scale = nullptr;
  float *styT = nullptr, *kmat = nullptr, *kw = nullptr, *vmat = nullptr, *vl = nullptr;
  // activations
  float *raw0 = nullptr, *raw1 = nullptr, *resid1 = nullptr, *kv = nullptr, *tvout = nullptr, *tvscores = nullptr;
  dexb::bf16 *A0 = nullptr, *B0 = nullptr, *C0 = nullptr, *D1 = nullptr, *A1 = nullptr, *B1 = nullptr, *C1 = nullptr,
             *cat = nullptr, *tvP = nullptr, *kq = nullptr, *vlt = nullptr;
  float* sbias = nullptr;
  dexb::bf16 *tokS = nullptr, *pairs = nullptr, *hS = nullptr, *qk = nullptr, *vT = nullptr, *P = nullptr,
             *attnS = nullptr, *h2S = nullptr;
  float* attn_tail = nullptr;               // partials of the tail-split DiT attention tiles (attn.cuh)
  float *pg = nullptr, *tiv_a = nullptr, *tiv_d = nullptr;
  float *xe = nullptr, *pe = nullptr, *xtok = nullptr, *scores = nullptr, *ytok = nullptr;
  // per-step zeroed region
  char* zero_base = nullptr;
  size_t zero_bytes = 0;
  double* gn_stats = nullptr;          // [slots][B][kGnRep][8][2]
  int gn_reverse = 1;                  // stand-alone GroupNorm-apply walks the images backwards (DEXB_GN_REVERSE)
  int gn_lag = 2, gn_mode = 0;         // tuning aids of the fused GroupNorm-apply (DEXB_GN_LAG / DEXB_GN_MODE)
  unsigned* gn_done = nullptr;         // [slots][B] tiles of an image whose raw rows + sums are published (fused GroupNorm-apply)
  double* cstats = nullptr;            // [2][B][C1][2]
  int n_slots = 0;
  // plans of the non-block GEMMs
  dexb::GemmPlan g_down, g_up[4], g_tvs, g_tvo, g_pe, g_posconv, g_final;
  dexb::AttnPlan attn, attn_tv;
  dexb::PosConvPlan pc_plan;
  bool fused_attn = false, fused_tv = false, fused_la = false, ws_posconv = false;
  dexb::bf16 *pc_w = nullptr, *pc_in = nullptr;      // weight-stationary pos-conv: packed weights / packed input rows
  // CUDA graph of one whole trajectory
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t graph_exec = nullptr;
  bool use_graph = true;
  cudaStream_t cap_stream = nullptr;
  // per-launch profiling (dexb_profile_step): events around every launch of one un-graphed step
  struct ProfRec { std::string tag; cudaEvent_t a, b; double flop; };
  bool prof = false;
  std::vector<ProfRec> prof_recs;
  long graph_launches = 0;
};

namespace dexb {
int engine_finalize(dexb_handle* h, cudaStream_t st);
int engine_plan(dexb_handle* h, int B, int T, int Ts, int Tr, int n_steps, const float* sigmas_host, size_t* ws_bytes);
int engine_run(dexb_handle* h, float* x_inout, const float* mu, const float* mask, const dexb_cond* cond,
               int only_step, float* den_out, cudaStream_t st);
void engine_release_plan(dexb_handle* h);
int engine_profile_step(dexb_handle* h, int step, char* buf, size_t buflen, cudaStream_t st);
void engine_release_weights(dexb_handle* h);
int engine_debug_tap(dexb_handle* h, const char* name, float* out, int* C, int* H, int* W, cudaStream_t st);
}  // namespace dexb
