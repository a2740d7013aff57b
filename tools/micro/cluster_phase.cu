// Micro-benchmark: cost of one PHASE of a fused multi-phase kernel inside a thread-block cluster (sfem_mg_tail.cu):
//   (a) cluster.sync() alone, (b) + a store, (c) + one ld.global.cg gather of a value another CTA wrote + store,
//   (d) + a chain of three dependent L2 loads (rowptr -> cols -> gather) + store,
// for clusters of 8 and 16 CTAs x 1024 / 512 threads, against the same phases as separate graph nodes.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_phase cluster_phase.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>

namespace cg = cooperative_groups;

// idx[i]: a permutation-like index array; x, y: vectors of n doubles
template <int MODE>
__global__ void k_phases(int phases, int n, const int* __restrict__ idx, double* x, double* y) {
  cg::cluster_group cluster = cg::this_cluster();
  const int T = (int)cluster.num_blocks() * blockDim.x;
  const int g = (int)cluster.block_rank() * blockDim.x + threadIdx.x;
  double* a = x;
  double* b = y;
  for (int p = 0; p < phases; ++p) {
    if (MODE >= 1) {
      for (int t = g; t < n; t += T) {
        double v = 1.0;
        if (MODE == 2) v = __ldcg(a + __ldg(idx + t));
        if (MODE == 3) { const int j = __ldcg(reinterpret_cast<const int*>(idx) + t); const int k = __ldcg(idx + j); v = __ldcg(a + k); }
        b[t] = v * 0.999 + 1e-3;
      }
    }
    cluster.sync();
    double* tmp = a; a = b; b = tmp;
  }
}

// the same phase as its own kernel (graph node)
template <int MODE>
__global__ void k_one(int n, const int* __restrict__ idx, const double* __restrict__ a, double* __restrict__ b) {
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    double v = 1.0;
    if (MODE == 2) v = a[idx[t]];
    if (MODE == 3) { const int j = idx[t]; const int k = idx[j]; v = a[k]; }
    b[t] = v * 0.999 + 1e-3;
  }
}

template <int MODE>
static float run_cluster(int csize, int threads, int phases, int n, const int* idx, double* x, double* y, cudaStream_t st) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(csize); cfg.blockDim = dim3(threads); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = csize; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  if (csize > 8) cudaFuncSetAttribute(k_phases<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaError_t e = cudaLaunchKernelEx(&cfg, k_phases<MODE>, phases, n, idx, x, y);
  if (e != cudaSuccess) { printf("launch failed (cluster %d): %s\n", csize, cudaGetErrorString(e)); cudaGetLastError(); return -1.f; }
  cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int r = 0; r < 10; ++r) cudaLaunchKernelEx(&cfg, k_phases<MODE>, phases, n, idx, x, y);
  cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return 1e3f * ms / (10.f * phases);
}

template <int MODE>
static float run_graph(int phases, int n, const int* idx, double* x, double* y, cudaStream_t st) {
  cudaGraph_t g; cudaGraphExec_t ge;
  cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
  const int grid = (n + 255) / 256 < 592 ? (n + 255) / 256 : 592;
  for (int p = 0; p < phases; ++p) k_one<MODE><<<grid, 256, 0, st>>>(n, idx, (p & 1) ? y : x, (p & 1) ? x : y);
  cudaStreamEndCapture(st, &g);
  cudaGraphInstantiate(&ge, g, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
  cudaEventRecord(e0, st);
  for (int r = 0; r < 10; ++r) cudaGraphLaunch(ge, st);
  cudaEventRecord(e1, st); cudaStreamSynchronize(st);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  cudaGraphExecDestroy(ge); cudaGraphDestroy(g);
  return 1e3f * ms / (10.f * phases);
}

int main() {
  const int phases = 200;
  cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
  for (int n : {4096, 16384}) {
    int* idx; double *x, *y;
    cudaMalloc(&idx, n * sizeof(int)); cudaMalloc(&x, n * sizeof(double)); cudaMalloc(&y, n * sizeof(double));
    int* h = new int[n];
    for (int i = 0; i < n; ++i) h[i] = (int)(((long long)i * 7919 + 13) % n);
    cudaMemcpy(idx, h, n * sizeof(int), cudaMemcpyHostToDevice);
    cudaMemset(x, 0, n * sizeof(double)); cudaMemset(y, 0, n * sizeof(double));
    printf("n = %d entries per phase\n", n);
    printf("  separate graph nodes:  store %.2f us   gather %.2f us   3 dependent loads %.2f us per phase\n",
           run_graph<1>(phases, n, idx, x, y, st), run_graph<2>(phases, n, idx, x, y, st), run_graph<3>(phases, n, idx, x, y, st));
    for (int cs : {8, 16}) {
      for (int th : {1024, 512}) {
        printf("  cluster %2d x %4d:  sync only %.2f us   store %.2f us   gather %.2f us   3 dependent loads %.2f us per phase\n", cs, th,
               run_cluster<0>(cs, th, phases, n, idx, x, y, st), run_cluster<1>(cs, th, phases, n, idx, x, y, st),
               run_cluster<2>(cs, th, phases, n, idx, x, y, st), run_cluster<3>(cs, th, phases, n, idx, x, y, st));
      }
    }
    cudaFree(idx); cudaFree(x); cudaFree(y); delete[] h;
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return e == cudaSuccess ? 0 : 1;
}
