// CUDA-graph helpers for the Krylov iteration bodies (internal).
//
// A Krylov iteration here is a fixed sequence of launches whose scalars live in device memory, so
// it is captured once per solver handle and replayed; the host only polls one double per
// iteration.  Capture needs a non-legacy stream: work arriving on the legacy default stream is
// forked onto a private non-blocking stream and joined back with events.
#pragma once
#include "sfem_common.cuh"
#include "sfem_internal.h"

#include <vector>

namespace sfem {

bool profiling_active();      // sfem_vector.cu: per-launch profiler on -> graphs are bypassed
bool graphs_enabled();        // env SFEM_GRAPHS (default on) and not profiling

// Registry epoch: bumped whenever something a captured launch sequence depends on -- but that is not visible in the
// pointers baked into it -- changes: a sliced-ELL / staged plan (un)registered, a halo pattern attached or detached,
// the communicator (de)activated, a multigrid handle created or destroyed.  Every graph cache stores the epoch of
// its capture and re-captures on mismatch, so a graph can never be replayed over freed or re-purposed buffers that
// happen to sit at the same addresses (graph_epoch / graph_epoch_bump: sfem_internal.h, counter in sfem_vector.cu).
struct GraphExec {
  cudaGraphExec_t exec = nullptr;
  long long nodes = 0;        // kernel launches recorded in the graph (for sfem_launch_count)
  void reset() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
    nodes = 0;
  }
};

// Captures body() (which enqueues work on st and returns an SFEM code) into g.
template <class F>
int graph_capture(cudaStream_t st, GraphExec& g, F&& body) {
  g.reset();
  SFEM_CUDA(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  const long long before = g_launches.load();
  const int rc = body();
  const long long after = g_launches.load();
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(st, &graph);
  g_launches.fetch_sub(after - before);          // nothing ran yet
  if (rc != SFEM_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) {
    set_error(std::string("cudaStreamEndCapture: ") + cudaGetErrorString(e));
    return SFEM_ERR_CUDA;
  }
  const cudaError_t e2 = cudaGraphInstantiate(&g.exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess) {
    g.exec = nullptr;
    set_error(std::string("cudaGraphInstantiate: ") + cudaGetErrorString(e2));
    return SFEM_ERR_CUDA;
  }
  g.nodes = after - before;
  return SFEM_OK;
}

// Everything a captured Krylov iteration bakes in: the pointers, the sizes, the smoother degree, the communicator
// size and the registry epoch.
struct GraphKey {
  const void* a[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  long long nnz = 0;
  int n = 0, m = 0, degree = 0, nranks = 0;
  unsigned long long epoch = 0;
  bool operator==(const GraphKey& o) const {
    for (int i = 0; i < 8; ++i)
      if (a[i] != o.a[i]) return false;
    return nnz == o.nnz && n == o.n && m == o.m && degree == o.degree && nranks == o.nranks && epoch == o.epoch;
  }
};
struct GraphCache {
  GraphKey key;
  std::vector<GraphExec> g;
  void invalidate(size_t n) {
    for (auto& e : g) e.reset();
    g.assign(n, GraphExec());
  }
  ~GraphCache() { for (auto& e : g) e.reset(); }
};

inline int graph_launch(GraphExec& g, cudaStream_t st) {
  SFEM_CUDA(cudaGraphLaunch(g.exec, st));
  g_launches.fetch_add(g.nodes, std::memory_order_relaxed);
  return SFEM_OK;
}

// Private work stream with fork/join against the caller's stream.
struct WorkStream {
  cudaStream_t s = nullptr;
  cudaEvent_t ev = nullptr;
  int init() {
    if (s) return SFEM_OK;
    SFEM_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    SFEM_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    return SFEM_OK;
  }
  void destroy() {
    if (ev) cudaEventDestroy(ev);
    if (s) cudaStreamDestroy(s);
    s = nullptr; ev = nullptr;
  }
  int fork(cudaStream_t user) {
    SFEM_TRY(init());
    SFEM_CUDA(cudaEventRecord(ev, user));
    SFEM_CUDA(cudaStreamWaitEvent(s, ev, 0));
    return SFEM_OK;
  }
  int join(cudaStream_t user) {
    SFEM_CUDA(cudaEventRecord(ev, s));
    SFEM_CUDA(cudaStreamWaitEvent(user, ev, 0));
    return SFEM_OK;
  }
};

}  // namespace sfem
