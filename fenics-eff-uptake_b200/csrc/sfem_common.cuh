// Shared host/device helpers for libsulcusfem (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <string>
#include <atomic>

#include "../../include/sulcusfem.h"

namespace sfem {

void set_error(const std::string& msg);
int num_sms();
extern std::atomic<long long> g_launches;

// Optional per-launch timing (bench.py roofline): when enabled every instrumented launch is
// bracketed by a CUDA event pair on its stream; sfem_profile_stop() returns (category, algorithmic
// bytes, milliseconds) per launch.  Disabled -> a single relaxed load per launch.
enum ProfCat {
  PC_SPMV = 0, PC_SPMV_DOT = 1, PC_CHEB = 2, PC_RESID_D0 = 3, PC_ELEM = 4, PC_GATHER = 5, PC_VEC = 6,
  PC_OTHER = 7, PC_SPMV_STAGED = 8, PC_HALO = 9
};
struct Prof {
  Prof(int cat, double bytes, cudaStream_t st);
  ~Prof();
  int idx;
  cudaStream_t st;
};

#define SFEM_CUDA(call)                                                                   \
  do {                                                                                    \
    cudaError_t e_ = (call);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      sfem::set_error(std::string(#call) + ": " + cudaGetErrorString(e_));                \
      return SFEM_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

#define SFEM_LAUNCH_CHECK()                                                               \
  do {                                                                                    \
    sfem::g_launches.fetch_add(1, std::memory_order_relaxed);                             \
    cudaError_t e_ = cudaGetLastError();                                                  \
    if (e_ != cudaSuccess) {                                                              \
      sfem::set_error(std::string("kernel launch: ") + cudaGetErrorString(e_));           \
      return SFEM_ERR_CUDA;                                                               \
    }                                                                                     \
  } while (0)

#define SFEM_TRY(call)                                                                    \
  do {                                                                                    \
    int r_ = (call);                                                                      \
    if (r_ != SFEM_OK) return r_;                                                         \
  } while (0)

constexpr int kThreads = 256;
constexpr int kMaxPartials = 2048;   // upper bound on blocks of any reducing kernel

// Grid for a grid-stride kernel: enough blocks for the work, capped at a multiple of the SM count.
inline int grid_for(long long work_items, int items_per_block, int blocks_per_sm = 8) {
  long long need = (work_items + items_per_block - 1) / items_per_block;
  long long cap = (long long)num_sms() * blocks_per_sm;
  if (cap > kMaxPartials) cap = kMaxPartials;
  if (need < 1) need = 1;
  return (int)(need < cap ? need : cap);
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Deterministic block sum; result returned to every thread.  blockDim.x must be a multiple of 32
// and <= 1024.  `sh` needs 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* sh) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, nw = blockDim.x >> 5;
  v = warp_sum(v);
  __syncthreads();                       // protect sh from a previous use
  if (lane == 0) sh[wid] = v;
  __syncthreads();
  if (wid == 0) {
    double t = (lane < nw) ? sh[lane] : 0.0;
    t = warp_sum(t);
    if (lane == 0) sh[32] = t;
  }
  __syncthreads();
  return sh[32];
}

// Sum of a (small) partials array by one block, fixed order.
__device__ __forceinline__ double block_sum_array(const double* p, int n, double* sh) {
  double t = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) t += p[i];
  return block_sum(t, sh);
}

// Dot of CSR row `row` with x by LANES cooperating lanes; every lane of the group gets the sum.
// All 32 lanes of the warp must call this (invalid rows pass valid=false).
template <int LANES>
__device__ __forceinline__ double csr_row_dot(const int* __restrict__ rowptr, const int* __restrict__ cols,
                                              const double* __restrict__ vals, const double* __restrict__ x,
                                              int row, bool valid, int lane) {
  double acc = 0.0;
  if (valid) {
    const int s = rowptr[row], e = rowptr[row + 1];
    for (int k = s + lane; k < e; k += LANES) acc = fma(vals[k], __ldg(x + cols[k]), acc);
  }
#pragma unroll
  for (int o = LANES >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  return acc;
}

// 16-byte alignment of every pointer in the list (vectorised double2 streaming paths)
template <class... P>
inline bool aligned16(P... ptrs) {
  uintptr_t acc = 0;
  const uintptr_t v[] = {reinterpret_cast<uintptr_t>(ptrs)...};
  for (uintptr_t a : v) acc |= a;
  return (acc & 15u) == 0;
}

// lanes per row from the average row length
inline int pick_lanes(long long nnz, int nrows) {
  double avg = nrows > 0 ? (double)nnz / nrows : 1.0;
  if (avg <= 2.5) return 1;
  if (avg <= 5.0) return 2;
  if (avg <= 14.0) return 4;
  if (avg <= 28.0) return 8;
  if (avg <= 56.0) return 16;
  return 32;
}

}  // namespace sfem
