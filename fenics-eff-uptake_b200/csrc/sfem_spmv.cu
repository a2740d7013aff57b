// FP64 CSR sparse mat-vec family (plain, residual, accumulate, fused dot, fused Chebyshev step).
//
// Variant 1 ("vector"): LANES cooperating lanes per row, grid-stride over row groups, grid capped
// at a multiple of the SM count.  Consecutive rows of a warp are contiguous in CSR, so the value /
// index streams are read as contiguous spans; x is gathered through the read-only path.
//
// Variant 2 ("staged"): a CTA owns a block of consecutive rows; its contiguous span of vals/cols is
// streamed HBM -> shared memory with 1-D bulk async copies (TMA, cp.async.bulk) completing on an
// mbarrier, double buffered, then reduced from shared memory.  See sfem_spmv_staged.cu.
#include "sfem_common.cuh"
#include "sfem_internal.h"

namespace sfem {

// MODE 0: y = A x;  1: y = b - A x;  2: y += A x
template <int LANES, int MODE>
__global__ void __launch_bounds__(kThreads) k_spmv(int nrows, const int* __restrict__ rowptr,
                                                   const int* __restrict__ cols, const double* __restrict__ vals,
                                                   const double* __restrict__ x, const double* __restrict__ b,
                                                   double* __restrict__ y) {
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const double s = csr_row_dot<LANES>(rowptr, cols, vals, x, row, valid, lane);
    if (valid && lane == 0) {
      if (MODE == 0) y[row] = s;
      else if (MODE == 1) y[row] = b[row] - s;
      else y[row] += s;
    }
  }
}

// y = A x and partial sums of <x, y> (CG: p.Ap) or <y, y>; one partial per block.
template <int LANES>
__global__ void __launch_bounds__(kThreads) k_spmv_dot(int nrows, const int* __restrict__ rowptr,
                                                       const int* __restrict__ cols, const double* __restrict__ vals,
                                                       const double* __restrict__ x, double* __restrict__ y,
                                                       double* __restrict__ partial) {
  __shared__ double sh[33];
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const double s = csr_row_dot<LANES>(rowptr, cols, vals, x, row, valid, lane);
    if (valid && lane == 0) {
      y[row] = s;
      acc = fma(x[row], s, acc);
    }
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// One fused Chebyshev-Jacobi step (see sfem_mg.cu):
//   t = (A d_old)_i;  r_i -= t;  x_i += d_old_i (+ d_new_i when LAST);
//   d_new_i = c1 d_old_i + c2 dinv_i r_i
template <int LANES>
__global__ void __launch_bounds__(kThreads) k_cheb_step(int nrows, const int* __restrict__ rowptr,
                                                        const int* __restrict__ cols, const double* __restrict__ vals,
                                                        const double* __restrict__ dinv, const double* __restrict__ d_old,
                                                        double* __restrict__ d_new, double* __restrict__ r,
                                                        double* __restrict__ x, double c1, double c2, int last) {
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const double t = csr_row_dot<LANES>(rowptr, cols, vals, d_old, row, valid, lane);
    if (valid && lane == 0) {
      const double rn = r[row] - t;
      const double dold = d_old[row];
      const double dn = c1 * dold + c2 * dinv[row] * rn;
      r[row] = rn;
      d_new[row] = dn;
      x[row] += last ? (dold + dn) : dold;
    }
  }
}

// r = b - A x ; d = c0 * dinv * r      (start of a smoothing sweep with a non-zero iterate)
template <int LANES>
__global__ void __launch_bounds__(kThreads) k_resid_d0(int nrows, const int* __restrict__ rowptr,
                                                       const int* __restrict__ cols, const double* __restrict__ vals,
                                                       const double* __restrict__ dinv, const double* __restrict__ b,
                                                       const double* __restrict__ x, double* __restrict__ r,
                                                       double* __restrict__ d, double c0) {
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const double s = csr_row_dot<LANES>(rowptr, cols, vals, x, row, valid, lane);
    if (valid && lane == 0) {
      const double rr = b[row] - s;
      r[row] = rr;
      d[row] = c0 * dinv[row] * rr;
    }
  }
}

template <int LANES>
static int launch_spmv_lanes(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st) {
  const int grid = grid_for(A.nrows, kThreads / LANES);
  Prof prof(PC_SPMV, 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * A.ncols + 8.0 * A.nrows * (mode == 0 ? 1 : 2), st);
  if (mode == 0) k_spmv<LANES, 0><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  else if (mode == 1) k_spmv<LANES, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  else k_spmv<LANES, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

#define SFEM_DISPATCH_LANES(L, ...)       \
  switch (L) {                            \
    case 1: { constexpr int LN = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int LN = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int LN = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int LN = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int LN = 16; __VA_ARGS__; } break; \
    default: { constexpr int LN = 32; __VA_ARGS__; } break; \
  }

int spmv(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st) {
  if (A.nrows <= 0) return SFEM_OK;
  if (mode == 1 && b == nullptr) { set_error("spmv mode 1 needs b"); return SFEM_ERR_ARG; }
  if (A.tile_cap > 0 && mode != 2) return spmv_staged_plan(A, A.tile_rows, A.tile_cap, A.stages, x, b, y, mode, st);
  const int lanes = pick_lanes(A.nnz, A.nrows);
  SFEM_DISPATCH_LANES(lanes, return launch_spmv_lanes<LN>(A, x, b, y, mode, st));
  return SFEM_OK;
}

int spmv_dot(const Csr& A, const double* x, double* y, double* partial, int* nparts, cudaStream_t st) {
  const int lanes = pick_lanes(A.nnz, A.nrows);
  int grid = 1;
  Prof prof(PC_SPMV_DOT, 12.0 * A.nnz + 20.0 * A.nrows, st);
  SFEM_DISPATCH_LANES(lanes, {
    grid = grid_for(A.nrows, kThreads / LN);
    k_spmv_dot<LN><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, y, partial);
  });
  SFEM_LAUNCH_CHECK();
  *nparts = grid;
  return SFEM_OK;
}

int cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
              double c1, double c2, int last, cudaStream_t st) {
  const int lanes = pick_lanes(A.nnz, A.nrows);
  Prof prof(PC_CHEB, 12.0 * A.nnz + 60.0 * A.nrows, st);
  SFEM_DISPATCH_LANES(lanes, {
    const int grid = grid_for(A.nrows, kThreads / LN);
    k_cheb_step<LN><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, d_old, d_new, r, x, c1, c2, last);
  });
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
             double c0, cudaStream_t st) {
  const int lanes = pick_lanes(A.nnz, A.nrows);
  Prof prof(PC_RESID_D0, 12.0 * A.nnz + 44.0 * A.nrows, st);
  SFEM_DISPATCH_LANES(lanes, {
    const int grid = grid_for(A.nrows, kThreads / LN);
    k_resid_d0<LN><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, b, x, r, d, c0);
  });
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace sfem
