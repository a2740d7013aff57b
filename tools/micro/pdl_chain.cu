// Micro-benchmark: cost per node of a CUDA-graph chain of tiny dependent kernels, with and without programmatic
// dependent launch (griddepcontrol).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pdl_chain pdl_chain.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>

template <int MODE>   // 0: plain, 1: wait only, 2: launch_dependents at start + wait
__global__ void k_axpy(int n, const double* __restrict__ x, double* __restrict__ y, double a) {
  if (MODE == 2) asm volatile("griddepcontrol.launch_dependents;");
  if (MODE >= 1) asm volatile("griddepcontrol.wait;" ::: "memory");
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) y[i] = fma(a, x[i], y[i]);
}

template <int MODE>
static void launch(int grid, int n, const double* x, double* y, cudaStream_t st, bool pdl) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(256); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = pdl ? 1 : 0;
  cudaLaunchKernelEx(&cfg, k_axpy<MODE>, n, x, y, 1e-9);
}

int main() {
  const int chain = 2000;
  for (int n : {2048, 32768, 262144, 2097152}) {
    double *x, *y;
    cudaMalloc(&x, n * sizeof(double)); cudaMalloc(&y, n * sizeof(double));
    cudaMemset(x, 0, n * sizeof(double)); cudaMemset(y, 0, n * sizeof(double));
    const int grid = (n + 1023) / 1024 < 592 ? (n + 1023) / 1024 : 592;
    for (int mode = 0; mode < 3; ++mode) {
      cudaStream_t st; cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking);
      cudaGraph_t g; cudaGraphExec_t ge;
      cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal);
      for (int i = 0; i < chain; ++i) {
        // ping-pong so that every kernel depends on its predecessor's output
        const double* a = (i & 1) ? y : x; double* b = (i & 1) ? x : y;
        if (mode == 0) launch<0>(grid, n, a, b, st, false);
        else if (mode == 1) launch<1>(grid, n, a, b, st, true);
        else launch<2>(grid, n, a, b, st, true);
      }
      cudaError_t e = cudaStreamEndCapture(st, &g);
      if (e != cudaSuccess) { printf("capture failed mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      e = cudaGraphInstantiate(&ge, g, 0);
      if (e != cudaSuccess) { printf("instantiate failed mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
      cudaGraphLaunch(ge, st); cudaStreamSynchronize(st);
      cudaEventRecord(e0, st);
      for (int r = 0; r < 5; ++r) cudaGraphLaunch(ge, st);
      cudaEventRecord(e1, st); cudaStreamSynchronize(st);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      printf("n=%8d grid=%4d mode=%d (%s): %.3f us per node\n", n, grid, mode,
             mode == 0 ? "plain" : mode == 1 ? "pdl wait" : "pdl trigger+wait", 1e3 * ms / (5.0 * chain));
      cudaGraphExecDestroy(ge); cudaGraphDestroy(g); cudaStreamDestroy(st);
    }
    cudaFree(x); cudaFree(y);
  }
  cudaError_t e = cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(e));
  return 0;
}
