// FP64 CSR sparse mat-vec family (plain, residual, accumulate, fused dot, fused Chebyshev step).
//
// Variant 1 ("vector"): LANES cooperating lanes per row, grid-stride over row groups, grid capped
// at a multiple of the SM count.  Consecutive rows of a warp are contiguous in CSR, so the value /
// index streams are read as contiguous spans (streaming loads, evict-first); x is gathered through
// the read-only path.  The row loop is unrolled twice so that every lane has two independent
// (index -> gather) chains in flight.
//
// NB right-hand sides (1 or 2) are stored interleaved ([dof][NB]): with NB = 2 the matrix stream is
// read once for both vectors and every gather is one 16-byte load.  The Taylor-Hood velocity block
// is the same scalar stiffness matrix for u_x and u_y, so the Stokes solver runs everything that
// touches it (operator, smoother, transfers) with NB = 2.
//
// Variant 2 ("staged"): a CTA owns a block of consecutive rows; its contiguous span of vals/cols is
// streamed HBM -> shared memory with 1-D bulk async copies (TMA, cp.async.bulk) completing on an
// mbarrier, multi-buffered, then reduced from shared memory.  See sfem_spmv_staged.cu.
#include "sfem_common.cuh"
#include "sfem_internal.h"

namespace sfem {

template <int NB>
struct Acc {
  double v[NB];
};

template <int NB>
__device__ __forceinline__ void gather_fma(const double* __restrict__ x, int col, double a, Acc<NB>& acc);
template <>
__device__ __forceinline__ void gather_fma<1>(const double* __restrict__ x, int col, double a, Acc<1>& acc) {
  acc.v[0] = fma(a, __ldg(x + col), acc.v[0]);
}
template <>
__device__ __forceinline__ void gather_fma<2>(const double* __restrict__ x, int col, double a, Acc<2>& acc) {
  const double2 xv = __ldg(reinterpret_cast<const double2*>(x) + col);
  acc.v[0] = fma(a, xv.x, acc.v[0]);
  acc.v[1] = fma(a, xv.y, acc.v[1]);
}

// Dot of CSR row `row` with the NB interleaved vectors in x by LANES cooperating lanes; every lane
// of the group gets the sums.  All 32 lanes of the warp must call this (invalid rows: valid=false).
template <int LANES, int NB>
__device__ __forceinline__ Acc<NB> csr_row_dot_nb(const int* __restrict__ rowptr, const int* __restrict__ cols,
                                                  const double* __restrict__ vals, const double* __restrict__ x,
                                                  int row, bool valid, int lane) {
  Acc<NB> a0, a1;
#pragma unroll
  for (int c = 0; c < NB; ++c) { a0.v[c] = 0.0; a1.v[c] = 0.0; }
  if (valid) {
    const int s = rowptr[row], e = rowptr[row + 1];
    int k = s + lane;
    for (; k + LANES < e; k += 2 * LANES) {
      const int c0 = __ldcs(cols + k), c1 = __ldcs(cols + k + LANES);
      const double v0 = __ldcs(vals + k), v1 = __ldcs(vals + k + LANES);
      gather_fma<NB>(x, c0, v0, a0);
      gather_fma<NB>(x, c1, v1, a1);
    }
    if (k < e) gather_fma<NB>(x, __ldcs(cols + k), __ldcs(vals + k), a0);
  }
#pragma unroll
  for (int c = 0; c < NB; ++c) {
    double t = a0.v[c] + a1.v[c];
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
    a0.v[c] = t;
  }
  return a0;
}

// MODE 0: y = A x;  1: y = b - A x;  2: y += A x
template <int LANES, int NB, int MODE>
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
    const Acc<NB> s = csr_row_dot_nb<LANES, NB>(rowptr, cols, vals, x, row, valid, lane);
    if (valid && lane < NB) {
      const double sv = (NB == 2 && lane == 1) ? s.v[NB - 1] : s.v[0];
      const size_t i = (size_t)row * NB + lane;
      if (MODE == 0) y[i] = sv;
      else if (MODE == 1) y[i] = b[i] - sv;
      else y[i] += sv;
    }
  }
}

// y = A x and partial sums of <dx, y> (CG: dx = x, p.Ap); one partial per block.
template <int LANES, int NB>
__global__ void __launch_bounds__(kThreads) k_spmv_dot(int nrows, const int* __restrict__ rowptr,
                                                       const int* __restrict__ cols, const double* __restrict__ vals,
                                                       const double* __restrict__ x, const double* __restrict__ dx,
                                                       double* __restrict__ y, double* __restrict__ partial) {
  __shared__ double sh[33];
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const Acc<NB> s = csr_row_dot_nb<LANES, NB>(rowptr, cols, vals, x, row, valid, lane);
    if (valid && lane < NB) {
      const double sv = (NB == 2 && lane == 1) ? s.v[NB - 1] : s.v[0];
      const size_t i = (size_t)row * NB + lane;
      y[i] = sv;
      acc = fma(dx[i], sv, acc);
    }
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// One fused Chebyshev-Jacobi step (see sfem_mg.cu):
//   t = (A d_old)_i;  r_i -= t;  x_i += d_old_i (+ d_new_i when LAST);
//   d_new_i = c1 d_old_i + c2 dinv_i r_i
template <int LANES, int NB>
__global__ void __launch_bounds__(kThreads) k_cheb_step(int nrows, const int* __restrict__ rowptr,
                                                        const int* __restrict__ cols, const double* __restrict__ vals,
                                                        const double* __restrict__ dinv, const double* __restrict__ d_old,
                                                        double* __restrict__ d_new, double* __restrict__ r,
                                                        double* __restrict__ x, const double* __restrict__ c12, int last) {
  const double c1 = c12[0], c2 = c12[1];
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const Acc<NB> s = csr_row_dot_nb<LANES, NB>(rowptr, cols, vals, d_old, row, valid, lane);
    if (valid && lane < NB) {
      const double t = (NB == 2 && lane == 1) ? s.v[NB - 1] : s.v[0];
      const size_t i = (size_t)row * NB + lane;
      const double rn = r[i] - t;
      const double dold = d_old[i];
      const double dn = c1 * dold + c2 * dinv[row] * rn;
      r[i] = rn;
      d_new[i] = dn;
      x[i] += last ? (dold + dn) : dold;
    }
  }
}

// r = b - A x ; d = c0 * dinv * r      (start of a smoothing sweep with a non-zero iterate)
template <int LANES, int NB>
__global__ void __launch_bounds__(kThreads) k_resid_d0(int nrows, const int* __restrict__ rowptr,
                                                       const int* __restrict__ cols, const double* __restrict__ vals,
                                                       const double* __restrict__ dinv, const double* __restrict__ b,
                                                       const double* __restrict__ x, double* __restrict__ r,
                                                       double* __restrict__ d, const double* __restrict__ c0p) {
  const double c0 = c0p[0];
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * ROWS; base < nrows; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < nrows;
    const Acc<NB> s = csr_row_dot_nb<LANES, NB>(rowptr, cols, vals, x, row, valid, lane);
    if (valid && lane < NB) {
      const double sv = (NB == 2 && lane == 1) ? s.v[NB - 1] : s.v[0];
      const size_t i = (size_t)row * NB + lane;
      const double rr = b[i] - sv;
      r[i] = rr;
      d[i] = c0 * dinv[row] * rr;
    }
  }
}

// Taylor-Hood velocity rows of  y = A z  in block form (both components of dof `row` at once):
//   y_u[row] = K[row,:] z_u  +  BT[2 row + c, :] z_p ,  plus partial sums of <z_u, y_u>.
// K is the scalar P2 stiffness (Dirichlet rows = identity), BT the (interleaved-row) transpose of
// the divergence block.
template <int LANES>
__global__ void __launch_bounds__(kThreads) k_stokes_apply_u(int n2, const int* __restrict__ k_rowptr,
                                                             const int* __restrict__ k_cols, const double* __restrict__ k_vals,
                                                             const int* __restrict__ bt_rowptr, const int* __restrict__ bt_cols,
                                                             const double* __restrict__ bt_vals,
                                                             const double* __restrict__ zu, const double* __restrict__ zp,
                                                             double* __restrict__ yu, double* __restrict__ partial) {
  __shared__ double sh[33];
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  double acc = 0.0;
  for (long long base = (long long)blockIdx.x * ROWS; base < n2; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < n2;
    Acc<2> s = csr_row_dot_nb<LANES, 2>(k_rowptr, k_cols, k_vals, zu, row, valid, lane);
    double t0 = 0.0, t1 = 0.0;
    if (valid) {
      const int s0 = bt_rowptr[2 * row], s1 = bt_rowptr[2 * row + 1], s2 = bt_rowptr[2 * row + 2];
      for (int k = s0 + lane; k < s1; k += LANES) t0 = fma(__ldcs(bt_vals + k), __ldg(zp + __ldcs(bt_cols + k)), t0);
      for (int k = s1 + lane; k < s2; k += LANES) t1 = fma(__ldcs(bt_vals + k), __ldg(zp + __ldcs(bt_cols + k)), t1);
    }
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) {
      t0 += __shfl_xor_sync(0xffffffffu, t0, o);
      t1 += __shfl_xor_sync(0xffffffffu, t1, o);
    }
    if (valid && lane < 2) {
      const double sv = (lane == 1) ? (s.v[1] + t1) : (s.v[0] + t0);
      const size_t i = (size_t)row * 2 + lane;
      yu[i] = sv;
      acc = fma(zu[i], sv, acc);
    }
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
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

// with NB = 2 the two result lanes need LANES >= 2
static inline int lanes_for(const Csr& A, int nb) {
  int l = pick_lanes(A.nnz, A.nrows);
  if (l < nb) l = nb;
  return l;
}

static inline double spmv_bytes(const Csr& A, int nb, int vec_passes) {
  return 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * nb * ((double)A.ncols + (double)A.nrows * vec_passes);
}

template <int LANES, int NB>
static int launch_spmv(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st) {
  const int grid = grid_for(A.nrows, kThreads / LANES);
  Prof prof(PC_SPMV, spmv_bytes(A, NB, mode == 0 ? 1 : 2), st);
  if (mode == 0) k_spmv<LANES, NB, 0><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  else if (mode == 1) k_spmv<LANES, NB, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  else k_spmv<LANES, NB, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int spmv(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st, int nb) {
  if (A.nrows <= 0) return SFEM_OK;
  if (mode == 1 && b == nullptr) { set_error("spmv mode 1 needs b"); return SFEM_ERR_ARG; }
  if (nb != 1 && nb != 2) { set_error("spmv: nb must be 1 or 2"); return SFEM_ERR_ARG; }
  if (nb == 2 && (reinterpret_cast<uintptr_t>(x) & 15u)) { set_error("spmv: nb = 2 needs a 16-byte aligned x"); return SFEM_ERR_ARG; }
  if (nb == 1 && A.tile_cap > 0 && mode != 2)
    return spmv_staged_plan(A, A.tile_rows, A.tile_cap, A.stages, x, b, y, mode, st);
  const int lanes = lanes_for(A, nb);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, return (launch_spmv<LN, 1>(A, x, b, y, mode, st)));
  } else {
    SFEM_DISPATCH_LANES(lanes, return (launch_spmv<(LN < 2 ? 2 : LN), 2>(A, x, b, y, mode, st)));
  }
  return SFEM_OK;
}

int spmv_dot(const Csr& A, const double* x, double* y, double* partial, int* nparts, cudaStream_t st, int nb,
             const double* dotx) {
  if (dotx == nullptr) dotx = x;
  const int lanes = lanes_for(A, nb);
  int grid = 1;
  Prof prof(PC_SPMV_DOT, spmv_bytes(A, nb, 2), st);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, {
      grid = grid_for(A.nrows, kThreads / LN);
      k_spmv_dot<LN, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, dotx, y, partial);
    });
  } else {
    SFEM_DISPATCH_LANES(lanes, {
      constexpr int L2 = LN < 2 ? 2 : LN;
      grid = grid_for(A.nrows, kThreads / L2);
      k_spmv_dot<L2, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, dotx, y, partial);
    });
  }
  SFEM_LAUNCH_CHECK();
  *nparts = grid;
  return SFEM_OK;
}

int cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
              const double* c12, int last, cudaStream_t st, int nb) {
  const int lanes = lanes_for(A, nb);
  Prof prof(PC_CHEB, 12.0 * A.nnz + 12.0 * A.nrows + 48.0 * nb * A.nrows, st);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, {
      const int grid = grid_for(A.nrows, kThreads / LN);
      k_cheb_step<LN, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, d_old, d_new, r, x, c12, last);
    });
  } else {
    SFEM_DISPATCH_LANES(lanes, {
      constexpr int L2 = LN < 2 ? 2 : LN;
      const int grid = grid_for(A.nrows, kThreads / L2);
      k_cheb_step<L2, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, d_old, d_new, r, x, c12, last);
    });
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
             const double* c0, cudaStream_t st, int nb) {
  const int lanes = lanes_for(A, nb);
  Prof prof(PC_RESID_D0, 12.0 * A.nnz + 12.0 * A.nrows + 32.0 * nb * A.nrows, st);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, {
      const int grid = grid_for(A.nrows, kThreads / LN);
      k_resid_d0<LN, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, b, x, r, d, c0);
    });
  } else {
    SFEM_DISPATCH_LANES(lanes, {
      constexpr int L2 = LN < 2 ? 2 : LN;
      const int grid = grid_for(A.nrows, kThreads / L2);
      k_resid_d0<L2, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, b, x, r, d, c0);
    });
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int stokes_apply_u(const Csr& K, const Csr& BT, const double* zu, const double* zp, double* yu, double* partial,
                   int* nparts, cudaStream_t st) {
  int lanes = pick_lanes(K.nnz + BT.nnz / 2, K.nrows);
  if (lanes < 2) lanes = 2;
  int grid = 1;
  Prof prof(PC_SPMV_DOT, 12.0 * (K.nnz + BT.nnz) + 12.0 * K.nrows + 8.0 * (2.0 * K.nrows * 3 + BT.ncols), st);
  SFEM_DISPATCH_LANES(lanes, {
    constexpr int L2 = LN < 2 ? 2 : LN;
    grid = grid_for(K.nrows, kThreads / L2);
    k_stokes_apply_u<L2><<<grid, kThreads, 0, st>>>(K.nrows, K.rowptr, K.cols, K.vals, BT.rowptr, BT.cols, BT.vals, zu, zp,
                                                    yu, partial);
  });
  SFEM_LAUNCH_CHECK();
  *nparts = grid;
  return SFEM_OK;
}

}  // namespace sfem
