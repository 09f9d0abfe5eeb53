// CUDA-graph replay for the once-per-utterance stages (TIV / TV / LF0 / text encoders): they are bound by their 13 ... 122 launch
// latencies, not by a roofline, so the forward of a (B, T) plan is captured at first use and replayed with one cudaGraphLaunch.
// A graph records pointers: the call's inputs / outputs are staged through fixed buffers of the handle (EncStage), a few small
// device-to-device copies around the launch.  DEXB_NO_GRAPH=1 keeps plain launches (sanitizer runs, per-kernel profiling).
#pragma once
#include <stdlib.h>

#include "common.cuh"

namespace dexb {

struct EncGraph {
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  cudaStream_t cap_stream = nullptr;
  char* stage = nullptr;               // fixed-address copies of the inputs / outputs of a forward
  long launches = 0;                   // kernels captured
};

static inline void enc_graph_release(EncGraph* g) {
  if (g->exec != nullptr) { cudaGraphExecDestroy(g->exec); g->exec = nullptr; }
  if (g->graph != nullptr) { cudaGraphDestroy(g->graph); g->graph = nullptr; }
  if (g->cap_stream != nullptr) { cudaStreamDestroy(g->cap_stream); g->cap_stream = nullptr; }
  cudaFree(g->stage); g->stage = nullptr;
}
static inline bool enc_graphs_on() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("DEXB_NO_GRAPH"); v = (e != nullptr && e[0] == '1') ? 0 : 1; }
  return v != 0;
}
// bump allocator over the staging buffer (base == nullptr: measuring pass)
struct EncStage {
  char* base;
  size_t off = 0;
  template <class T> T* get(size_t n) {
    off = (off + 255) & ~(size_t)255;
    T* p = base != nullptr ? reinterpret_cast<T*>(base + off) : nullptr;
    off += n * sizeof(T);
    return p;
  }
};
// Replay the captured forward on `st`; capture it first by running `enqueue(stream)` -- which must only launch work that reads /
// writes buffers owned by the handle -- on a private stream.  `launches` is the handle's launch counter (set by enqueue).
template <class F>
static inline int enc_graph_run(EncGraph* g, long* launches, cudaStream_t st, F enqueue) {
  if (g->exec == nullptr) {
    if (g->cap_stream == nullptr) DEXB_CUDA_OK(cudaStreamCreateWithFlags(&g->cap_stream, cudaStreamNonBlocking));
    DEXB_CUDA_OK(cudaStreamBeginCapture(g->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int r = enqueue(g->cap_stream);
    cudaGraph_t gr = nullptr;
    const cudaError_t e = cudaStreamEndCapture(g->cap_stream, &gr);
    if (r != 0) { if (gr != nullptr) cudaGraphDestroy(gr); return r; }
    DEXB_CHECK(e == cudaSuccess && gr != nullptr, "encoder graph capture failed: %s", cudaGetErrorString(e));
    g->graph = gr;
    DEXB_CUDA_OK(cudaGraphInstantiate(&g->exec, gr, 0));
    g->launches = *launches;
  }
  *launches = g->launches;
  DEXB_CUDA_OK(cudaGraphLaunch(g->exec, st));
  return 0;
}

}  // namespace dexb
