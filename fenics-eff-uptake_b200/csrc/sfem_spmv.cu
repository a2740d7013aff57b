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
// Variant 3 ("sliced ELL", sfem_spmv_sell.cu): matrices with a registered SELL-32 mirror are served by a
// one-lane-per-row kernel with perfectly coalesced matrix loads; it is tried first.
//
// Variant 2 ("staged"): a CTA owns a block of consecutive rows; its contiguous span of vals/cols is
// streamed HBM -> shared memory with 1-D bulk async copies (TMA, cp.async.bulk) completing on an
// mbarrier, multi-buffered, then reduced from shared memory.  See sfem_spmv_staged.cu.
#include "sfem_common.cuh"
#include "sfem_internal.h"
#include "sfem_row_engine.cuh"
#include "sfem_spmv_epi.cuh"
#include "sfem_dist.h"

#include <cstdlib>

namespace sfem {

// ------------------------------------------------------------------ kernels
template <int LANES, int NB, int MODE>
__global__ void __launch_bounds__(kThreads, 4) k_spmv(int nrows, const int* __restrict__ rowptr,
                                                   const int* __restrict__ cols, const double* __restrict__ vals,
                                                   const double* __restrict__ x, const double* __restrict__ b,
                                                   double* __restrict__ y) {
  EpiStore<NB, MODE> epi{b, y};
  row_engine<LANES, NB>(nrows, rowptr, cols, vals, x, epi);
}

// y = A x (MODE 0) or y += A x (MODE 2) and partial sums of <dx, y>; one partial per block.
template <int LANES, int NB, int MODE>
__global__ void __launch_bounds__(kThreads, 4) k_spmv_dot(int nrows, const int* __restrict__ rowptr,
                                                       const int* __restrict__ cols, const double* __restrict__ vals,
                                                       const double* __restrict__ x, const double* __restrict__ dx,
                                                       double* __restrict__ y, double* __restrict__ partial) {
  __shared__ double sh[33];
  EpiDot<NB, MODE> epi{dx, y, 0.0};
  row_engine<LANES, NB>(nrows, rowptr, cols, vals, x, epi);
  const double t = block_sum(epi.acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

template <int LANES, int NB>
__global__ void __launch_bounds__(kThreads, 4) k_cheb_step(int nrows, const int* __restrict__ rowptr,
                                                        const int* __restrict__ cols, const double* __restrict__ vals,
                                                        const double* __restrict__ dinv, const double* __restrict__ d_old,
                                                        double* __restrict__ d_new, double* __restrict__ r,
                                                        double* __restrict__ x, const double* __restrict__ c12, int last,
                                                        const double* __restrict__ b0) {
  EpiCheb<NB> epi{dinv, d_old, d_new, r, x, c12[0], c12[1], last, b0};
  row_engine<LANES, NB>(nrows, rowptr, cols, vals, d_old, epi);
}

template <int LANES, int NB>
__global__ void __launch_bounds__(kThreads, 4) k_resid_d0(int nrows, const int* __restrict__ rowptr,
                                                       const int* __restrict__ cols, const double* __restrict__ vals,
                                                       const double* __restrict__ dinv, const double* __restrict__ b,
                                                       const double* __restrict__ x, double* __restrict__ r,
                                                       double* __restrict__ d, const double* __restrict__ c0p) {
  EpiResidD0<NB> epi{dinv, b, r, d, c0p[0]};
  row_engine<LANES, NB>(nrows, rowptr, cols, vals, x, epi);
}

int engine_lanes(long long nnz, int nrows, int nb) {
  static int forced = -1;
  if (forced < 0) { const char* e = std::getenv("SFEM_LANES"); forced = e ? std::atoi(e) : 0; }
  const double avg = nrows > 0 ? (double)nnz / nrows : 1.0;
  int l = 1;
  while (l < 32 && 4.0 * l < avg) l <<= 1;
  if (forced > 0) l = forced;
  if (l < nb) l = nb;
  return l;
}
static inline int lanes_for(const Csr& A, int nb) { return engine_lanes(A.nnz, A.nrows, nb); }

static inline double spmv_bytes(const Csr& A, int nb, int vec_passes) {
  return 12.0 * A.nnz + 4.0 * A.nrows + 8.0 * nb * ((double)A.ncols + (double)A.nrows * vec_passes);
}

template <int LANES, int NB>
static int launch_spmv(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st) {
  const int grid = grid_for(A.nrows, kThreads / LANES, kSpmvBlocksPerSm);
  Prof prof(PC_SPMV, spmv_bytes(A, NB, mode == 0 ? 1 : 2), st);
  if (mode == 0) k_spmv<LANES, NB, 0><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  else if (mode == 1) k_spmv<LANES, NB, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  else k_spmv<LANES, NB, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, b, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int spmv(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st, int nb, bool exchange) {
  if (A.nrows <= 0) return SFEM_OK;
  if (mode == 1 && b == nullptr) { set_error("spmv mode 1 needs b"); return SFEM_ERR_ARG; }
  if (nb != 1 && nb != 2) { set_error("spmv: nb must be 1 or 2"); return SFEM_ERR_ARG; }
  if (nb == 2 && (reinterpret_cast<uintptr_t>(x) & 15u)) { set_error("spmv: nb = 2 needs a 16-byte aligned x"); return SFEM_ERR_ARG; }
  if (exchange) SFEM_TRY(halo_exchange(find_halo(A.rowptr), const_cast<double*>(x), nb, st));   // distributed matrix: fill the ghosts of x
  { const int took = sell_spmv(A, x, b, y, mode, nb, st); if (took != 0) return took < 0 ? took : SFEM_OK; }
  { const int took = staged_spmv(A, x, b, y, mode, nb, st); if (took != 0) return took < 0 ? took : SFEM_OK; }
  const int lanes = lanes_for(A, nb);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, return (launch_spmv<LN, 1>(A, x, b, y, mode, st)));
  } else {
    SFEM_DISPATCH_LANES(lanes, return (launch_spmv<(LN < 2 ? 2 : LN), 2>(A, x, b, y, mode, st)));
  }
  return SFEM_OK;
}

int spmv_dot(const Csr& A, const double* x, double* y, double* partial, int* nparts, cudaStream_t st, int nb,
             const double* dotx, int mode, bool exchange) {
  if (dotx == nullptr) dotx = x;
  if (mode != 0 && mode != 2) { set_error("spmv_dot: mode must be 0 or 2"); return SFEM_ERR_ARG; }
  if (exchange) SFEM_TRY(halo_exchange(find_halo(A.rowptr), const_cast<double*>(x), nb, st));
  { const int took = sell_spmv_dot(A, x, dotx, y, partial, nparts, mode, nb, st); if (took != 0) return took < 0 ? took : SFEM_OK; }
  { const int took = staged_spmv_dot(A, x, dotx, y, partial, nparts, mode, nb, st); if (took != 0) return took < 0 ? took : SFEM_OK; }
  const int lanes = lanes_for(A, nb);
  int grid = 1;
  Prof prof(PC_SPMV_DOT, spmv_bytes(A, nb, 2), st);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, {
      grid = grid_for(A.nrows, kThreads / LN, kSpmvBlocksPerSm);
      if (mode == 0) k_spmv_dot<LN, 1, 0><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, dotx, y, partial);
      else k_spmv_dot<LN, 1, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, dotx, y, partial);
    });
  } else {
    SFEM_DISPATCH_LANES(lanes, {
      constexpr int L2 = LN < 2 ? 2 : LN;
      grid = grid_for(A.nrows, kThreads / L2, kSpmvBlocksPerSm);
      if (mode == 0) k_spmv_dot<L2, 2, 0><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, dotx, y, partial);
      else k_spmv_dot<L2, 2, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, x, dotx, y, partial);
    });
  }
  SFEM_LAUNCH_CHECK();
  *nparts = grid;
  return SFEM_OK;
}

int cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
              const double* c12, int last, cudaStream_t st, int nb, const double* b0) {
  SFEM_TRY(halo_exchange(find_halo(A.rowptr), const_cast<double*>(d_old), nb, st));
  { const int took = sell_cheb_step(A, dinv, d_old, d_new, r, x, c12, last, nb, st, b0); if (took != 0) return took < 0 ? took : SFEM_OK; }
  { const int took = staged_cheb_step(A, dinv, d_old, d_new, r, x, c12, last, nb, st, b0); if (took != 0) return took < 0 ? took : SFEM_OK; }
  const int lanes = lanes_for(A, nb);
  Prof prof(PC_CHEB, 12.0 * A.nnz + 12.0 * A.nrows + (b0 ? 40.0 : 48.0) * nb * A.nrows, st);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, {
      const int grid = grid_for(A.nrows, kThreads / LN, kSpmvBlocksPerSm);
      k_cheb_step<LN, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, d_old, d_new, r, x, c12, last, b0);
    });
  } else {
    SFEM_DISPATCH_LANES(lanes, {
      constexpr int L2 = LN < 2 ? 2 : LN;
      const int grid = grid_for(A.nrows, kThreads / L2, kSpmvBlocksPerSm);
      k_cheb_step<L2, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, d_old, d_new, r, x, c12, last, b0);
    });
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
             const double* c0, cudaStream_t st, int nb) {
  SFEM_TRY(halo_exchange(find_halo(A.rowptr), const_cast<double*>(x), nb, st));
  { const int took = sell_resid_d0(A, dinv, b, x, r, d, c0, nb, st); if (took != 0) return took < 0 ? took : SFEM_OK; }
  { const int took = staged_resid_d0(A, dinv, b, x, r, d, c0, nb, st); if (took != 0) return took < 0 ? took : SFEM_OK; }
  const int lanes = lanes_for(A, nb);
  Prof prof(PC_RESID_D0, 12.0 * A.nnz + 12.0 * A.nrows + 32.0 * nb * A.nrows, st);
  if (nb == 1) {
    SFEM_DISPATCH_LANES(lanes, {
      const int grid = grid_for(A.nrows, kThreads / LN, kSpmvBlocksPerSm);
      k_resid_d0<LN, 1><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, b, x, r, d, c0);
    });
  } else {
    SFEM_DISPATCH_LANES(lanes, {
      constexpr int L2 = LN < 2 ? 2 : LN;
      const int grid = grid_for(A.nrows, kThreads / L2, kSpmvBlocksPerSm);
      k_resid_d0<L2, 2><<<grid, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, dinv, b, x, r, d, c0);
    });
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

// y_u = K z_u (both components, matrix read once) ; y_u += BT z_p with the partial sums of <z_u, y_u>
int stokes_apply_u(const Csr& K, const Csr& BT, const double* zu, const double* zp, double* yu, double* partial,
                   int* nparts, cudaStream_t st) {
  // row-partitioned: the velocity ghosts (K) and the pressure ghosts (B^T) of the vector travel in one exchange launch
  SFEM_TRY(halo_exchange_pair(find_halo(K.rowptr), const_cast<double*>(zu), 2, find_halo(BT.rowptr), const_cast<double*>(zp), 1, st));
  SFEM_TRY(spmv(K, zu, nullptr, yu, 0, st, 2, false));
  return spmv_dot(BT, zp, yu, partial, nparts, st, 1, zu, 2, false);
}

}  // namespace sfem
