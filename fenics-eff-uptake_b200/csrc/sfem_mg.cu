// Geometric multigrid V-cycle with fused Chebyshev-Jacobi smoothing (the preconditioner that lets
// Krylov iterations reach the residual level of the reference's sparse LU).
//
// Per level:   pre-smooth (zero initial guess)  ->  r = b - A x  ->  b_c = R r  ->  recurse
//              ->  x += P x_c  ->  post-smooth.          Coarsest level: dense inverse (gemv).
// Smoother, degree k on D^-1 A with eigenvalue window [lmax/ratio, lmax]:
//     r0 = b - A x0,  d0 = D^-1 r0 / theta
//     x_{i+1} = x_i + d_i;  r_{i+1} = r_i - A d_i;  d_{i+1} = rho_{i+1} rho_i d_i + (2 rho_{i+1}/delta) D^-1 r_{i+1}
// Each step after the first is ONE kernel (SpMV fused with all vector updates, ping-pong d).
// lmax is the Gershgorin bound of D^-1 A -- a guaranteed upper bound keeps the smoother (and with
// it the V-cycle) positive definite, which MINRES / CG rely on -- computed on the device; the
// coefficients stay in device memory, so set-up never synchronises with the host and captured
// CUDA graphs of a cycle stay valid when the operator is re-assembled.
// With nb = 2 a cycle carries two interleaved right-hand sides (see sfem_spmv.cu).
#include "sfem_mg.h"
#include "sfem_dist.h"

#include <cmath>
#include <cstdlib>
#include <string>
#include <vector>

namespace sfem {

namespace {

// n = rows * nb entries; dinv is per row
// NB == 2: one thread per dof streams the interleaved pair as double2 (b, r, d, x are 16-byte aligned)
// start of a sweep from x = 0:  d_0 = D^-1 b / theta.  FULL: also r = b, x = d_0 (a one-step sweep); otherwise only d_0 is
// written -- the first fused step reads its residual from b and takes x as zero (EpiCheb::b0).
template <int NB, bool FULL>
__global__ void __launch_bounds__(kThreads) k_cheb_init0(int nrows, const double* __restrict__ dinv,
                                                         const double* __restrict__ b, double* __restrict__ r,
                                                         double* __restrict__ d, double* __restrict__ x,
                                                         const double* __restrict__ coef) {
  const double c0 = coef[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nrows; i += gridDim.x * blockDim.x) {
    const double s = c0 * dinv[i];
    if (NB == 2) {
      const double2 bi = reinterpret_cast<const double2*>(b)[i];
      const double2 di = make_double2(s * bi.x, s * bi.y);
      reinterpret_cast<double2*>(d)[i] = di;
      if (FULL) {
        reinterpret_cast<double2*>(r)[i] = bi;
        reinterpret_cast<double2*>(x)[i] = di;
      }
    } else {
      const double bi = b[i];
      d[i] = s * bi;
      if (FULL) {
        r[i] = bi;
        x[i] = s * bi;
      }
    }
  }
}

// per-block max of sum_j |a_ij| * |dinv_i|  (Gershgorin bound on the spectrum of D^-1 A)
__global__ void __launch_bounds__(kThreads) k_gershgorin(int n, const int* __restrict__ rowptr,
                                                         const double* __restrict__ vals, const double* __restrict__ dinv,
                                                         double* __restrict__ partial) {
  __shared__ double sh[32];
  double mx = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) s += fabs(vals[k]);
    mx = fmax(mx, s * fabs(dinv[i]));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (blockDim.x >> 5); ++w) mx = fmax(mx, sh[w]);
    partial[blockIdx.x] = mx;
  }
}

__global__ void k_cheb_coef(const double* __restrict__ partial, int np, double fixed_lmax, double ratio, int degree,
                            double* __restrict__ coef, DistDev D) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double lmax = fixed_lmax;
  if (np > 0) {
    lmax = 0.0;
    for (int i = 0; i < np; ++i) lmax = fmax(lmax, partial[i]);
    if (D.nranks > 1) {
      // row-partitioned operator: the bound must be the same on every rank (a rank-dependent polynomial
      // would not be a symmetric preconditioner).  All-gather through the scalar all-reduce, then max.
      static_assert(kAllreduceMaxK >= kMaxRanks, "one all-reduce slot per rank");
      double v[kAllreduceMaxK];
      for (int q = 0; q < kAllreduceMaxK; ++q) v[q] = (q == D.rank) ? lmax : 0.0;
      dist_allreduce_scalars(D, v, D.nranks);
      for (int q = 0; q < D.nranks; ++q) lmax = fmax(lmax, v[q]);
    }
    if (!(lmax > 0.0)) lmax = 2.0;
  }
  const double lmin = lmax / ratio;
  const double theta = 0.5 * (lmax + lmin), delta = 0.5 * (lmax - lmin), sigma = theta / delta;
  double rho = 1.0 / sigma;
  coef[0] = lmax;
  coef[1] = 1.0 / theta;
  for (int i = 0; i + 1 < degree; ++i) {
    const double rho_new = 1.0 / (2.0 * sigma - rho);
    coef[2 + 2 * i] = rho_new * rho;
    coef[3 + 2 * i] = 2.0 * rho_new / delta;
    rho = rho_new;
  }
}

constexpr int kCoarseFallbackDegree = 12;

}  // namespace

int cheb_setup(const Csr* A, const double* dinv, double fixed_lmax, double ratio, int degree, double* scratch,
               double* coef, cudaStream_t st, bool distributed) {
  if (degree < 1 || degree > kMaxChebDegree) { set_error("chebyshev degree out of range"); return SFEM_ERR_ARG; }
  int np = 0;
  if (A != nullptr) {
    np = grid_for(A->nrows, kThreads, 4);
    k_gershgorin<<<np, kThreads, 0, st>>>(A->nrows, A->rowptr, A->vals, dinv, scratch);
    SFEM_LAUNCH_CHECK();
  }
  k_cheb_coef<<<1, 32, 0, st>>>(scratch, np, fixed_lmax, ratio, degree, coef, distributed ? dist_dev() : DistDev());
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int smooth(const Csr& A, const double* dinv, const double* coef, int degree, const double* b, double* x,
           double* r, double* d0, double* d1, bool zero_init, cudaStream_t st, int nb) {
  const int n = A.nrows * nb;
  const int gi = grid_for(A.nrows, kThreads * 2);
  if (zero_init) {
    if (degree <= 1) {                                        // one-step sweep: x = d_0, r = b
      Prof prof(PC_VEC, 32.0 * n + 8.0 * A.nrows, st);
      if (nb == 2) k_cheb_init0<2, true><<<gi, kThreads, 0, st>>>(A.nrows, dinv, b, r, d0, x, coef);
      else k_cheb_init0<1, true><<<gi, kThreads, 0, st>>>(A.nrows, dinv, b, r, d0, x, coef);
    } else {                                                  // only d_0: the first fused step reads b and takes x = 0
      Prof prof(PC_VEC, 16.0 * n + 8.0 * A.nrows, st);
      if (nb == 2) k_cheb_init0<2, false><<<gi, kThreads, 0, st>>>(A.nrows, dinv, b, r, d0, x, coef);
      else k_cheb_init0<1, false><<<gi, kThreads, 0, st>>>(A.nrows, dinv, b, r, d0, x, coef);
    }
    SFEM_LAUNCH_CHECK();
    if (degree <= 1) return SFEM_OK;
  } else {
    SFEM_TRY(resid_d0(A, dinv, b, x, r, d0, coef + 1, st, nb));
    if (degree <= 1) return vec_axpby(n, 1.0, d0, 1.0, x, st);
  }
  double* dold = d0;
  double* dnew = d1;
  for (int i = 0; i < degree - 1; ++i) {
    SFEM_TRY(cheb_step(A, dinv, dold, dnew, r, x, coef + 2 + 2 * i, i == degree - 2, st, nb,
                       (zero_init && i == 0) ? b : nullptr));
    double* t = dold; dold = dnew; dnew = t;
  }
  return SFEM_OK;
}

int mg_vcycle_level(sfem_mg* mg, int l, const double* b, double* x, cudaStream_t st) {
  if (l == mg_tail_start(mg)) return mg_tail_vcycle(mg, l, b, x, st);      // small levels: one fused launch (sfem_mg_tail.cu)
  MgLevel& L = mg->levels[l];
  const int last = (int)mg->levels.size() - 1;
  const int nb = mg->nb;
  if (l == last && mg->tail == nullptr) {
    if (mg->coarse_inv != nullptr) return dense_gemv(L.A.nrows, mg->coarse_inv, b, x, st, nb);
    // no dense inverse given: a long smoothing sweep stands in for the coarse solve
    return smooth(L.A, L.dinv, L.coef, kCoarseFallbackDegree, b, x, L.r, L.d0, L.d1, true, st, nb);
  }
  SFEM_TRY(smooth(L.A, L.dinv, L.coef, mg->degree, b, x, L.r, L.d0, L.d1, true, st, nb));
  SFEM_TRY(spmv(L.A, x, b, L.r, 1, st, nb));
  if (l == last) {
    // last row-partitioned level: restrict the owned part, sum over ranks, solve the replicated
    // hierarchy redundantly (identical on every rank), prolong to the owned rows
    SFEM_TRY(spmv(L.R, L.r, nullptr, mg->tail_b, 0, st, nb));
    SFEM_TRY(dist_allreduce_vec(active_dist(), mg->tail_b, mg->n_tail * nb, st));
    SFEM_TRY(mg_vcycle_level(mg->tail, 0, mg->tail_b, mg->tail_x, st));
    SFEM_TRY(spmv(L.P, mg->tail_x, nullptr, x, 2, st, nb));
  } else {
    MgLevel& C = mg->levels[l + 1];
    SFEM_TRY(spmv(L.R, L.r, nullptr, C.b, 0, st, nb));
    SFEM_TRY(mg_vcycle_level(mg, l + 1, C.b, C.x, st));
    SFEM_TRY(spmv(L.P, C.x, nullptr, x, 2, st, nb));
  }
  SFEM_TRY(smooth(L.A, L.dinv, L.coef, mg->degree, b, x, L.r, L.d0, L.d1, false, st, nb));
  return SFEM_OK;
}

}  // namespace sfem

using namespace sfem;

extern "C" {

sfem_mg_t sfem_mg_create(int nlevels, const int* h_n, const int* h_A_nnz,
                         const int* const* h_A_rowptr, const int* const* h_A_cols, const double* const* h_A_vals,
                         const int* h_P_nnz,
                         const int* const* h_P_rowptr, const int* const* h_P_cols, const double* const* h_P_vals,
                         const int* const* h_R_rowptr, const int* const* h_R_cols, const double* const* h_R_vals,
                         const double* coarse_inv, int cheb_degree, double eig_ratio, int nb) {
  if (nlevels < 1 || cheb_degree < 1 || cheb_degree > kMaxChebDegree || !(eig_ratio > 1.0) || (nb != 1 && nb != 2)) {
    set_error("sfem_mg_create: bad arguments");
    return nullptr;
  }
  graph_epoch_bump();
  sfem_mg* mg = new sfem_mg();
  mg->degree = cheb_degree;
  mg->ratio = eig_ratio;
  mg->nb = nb;
  mg->coarse_inv = coarse_inv;
  mg->levels.resize(nlevels);
  for (int l = 0; l < nlevels; ++l) {
    MgLevel& L = mg->levels[l];
    const int n = h_n[l];
    L.A.nrows = L.A.ncols = n;
    L.A.nnz = h_A_nnz[l];
    L.A.rowptr = h_A_rowptr[l]; L.A.cols = h_A_cols[l]; L.A.vals = h_A_vals[l];
    if (l < nlevels - 1) {
      const int nc = h_n[l + 1];
      L.P.nrows = n; L.P.ncols = nc; L.P.nnz = h_P_nnz[l];
      L.P.rowptr = h_P_rowptr[l]; L.P.cols = h_P_cols[l]; L.P.vals = h_P_vals[l];
      L.R.nrows = nc; L.R.ncols = n; L.R.nnz = h_P_nnz[l];
      L.R.rowptr = h_R_rowptr[l]; L.R.cols = h_R_cols[l]; L.R.vals = h_R_vals[l];
    }
    // a row-partitioned level keeps ghost entries behind the owned ones in every vector an SpMV reads
    const Halo* halo = find_halo(L.A.rowptr);
    const int n_alloc = (halo && halo->dev.n_loc > n) ? halo->dev.n_loc : n;
    const size_t bytes = (size_t)n_alloc * nb * sizeof(double);
    bool ok = cudaMalloc(&L.dinv, (size_t)n * sizeof(double)) == cudaSuccess && cudaMalloc(&L.r, bytes) == cudaSuccess &&
              cudaMalloc(&L.d0, bytes) == cudaSuccess && cudaMalloc(&L.d1, bytes) == cudaSuccess &&
              cudaMalloc(&L.coef, kChebCoefLen * sizeof(double)) == cudaSuccess;
    if (ok && l > 0) ok = cudaMalloc(&L.x, bytes) == cudaSuccess && cudaMalloc(&L.b, bytes) == cudaSuccess;
    if (!ok) {
      set_error("sfem_mg_create: device allocation failed");
      sfem_mg_destroy(mg);
      return nullptr;
    }
  }
  if (cudaMalloc(&mg->scratch, (kMaxPartials + 8) * sizeof(double)) != cudaSuccess) {
    set_error("sfem_mg_create: device allocation failed");
    sfem_mg_destroy(mg);
    return nullptr;
  }
  return mg;
}

int sfem_mg_setup(sfem_mg_t mg, void* stream) {
  if (!mg) { set_error("null mg handle"); return SFEM_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  SFEM_TRY(sell_ensure_all(st));            // operator values were just re-assembled: refresh the sliced-ELL mirrors
  const size_t nl = mg->levels.size();
  for (size_t l = 0; l < nl; ++l) {
    MgLevel& L = mg->levels[l];
    SFEM_TRY(extract_diag_inv(L.A, L.dinv, st));
    const bool dist = find_halo(L.A.rowptr) != nullptr || mg->tail != nullptr;
    if (l + 1 == nl && mg->tail == nullptr) {
      if (mg->coarse_inv == nullptr)
        SFEM_TRY(cheb_setup(&L.A, L.dinv, 0.0, 30.0, kCoarseFallbackDegree, mg->scratch, L.coef, st, dist));
      continue;
    }
    SFEM_TRY(cheb_setup(&L.A, L.dinv, 0.0, mg->ratio, mg->degree, mg->scratch, L.coef, st, dist));
  }
  mg->ready = true;
  return SFEM_OK;
}

/* Smoother set-up of the system level only (level 0: diagonal, Gershgorin bound, Chebyshev coefficients): for a
 * parameter sweep in which only boundary rows of the system matrix change (Robin coefficient mu) and the coarse levels
 * are kept from a nearby parameter value -- they are preconditioner data, the Krylov solve still converges to the
 * residual of the exact system.  The handle must have been set up in full once. */
int sfem_mg_setup_fine(sfem_mg_t mg, void* stream) {
  if (!mg || !mg->ready) { set_error("sfem_mg_setup_fine: the handle needs one full sfem_mg_setup first"); return SFEM_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  SFEM_TRY(sell_ensure_all(st));
  MgLevel& L = mg->levels[0];
  SFEM_TRY(extract_diag_inv(L.A, L.dinv, st));
  if (mg->levels.size() == 1 && mg->tail == nullptr) {
    if (mg->coarse_inv == nullptr)
      SFEM_TRY(cheb_setup(&L.A, L.dinv, 0.0, 30.0, kCoarseFallbackDegree, mg->scratch, L.coef, st, false));
    return SFEM_OK;
  }
  const bool dist = find_halo(L.A.rowptr) != nullptr || mg->tail != nullptr;
  return cheb_setup(&L.A, L.dinv, 0.0, mg->ratio, mg->degree, mg->scratch, L.coef, st, dist);
}

/* Multi-GPU: `mg` holds the row-partitioned levels; below its last level the replicated hierarchy `tail`
 * (n_tail dofs on its finest level) is solved redundantly.  P_last: n_own(last) x n_tail, R_last: n_tail x
 * n_own(last) restricted to the owned columns (the partial results are summed over the ranks). */
int sfem_mg_set_tail(sfem_mg_t mg, sfem_mg_t tail, int n_tail, int P_nnz, const int* P_rowptr, const int* P_cols,
                     const double* P_vals, int R_nnz, const int* R_rowptr, const int* R_cols, const double* R_vals) {
  if (!mg || !tail || n_tail <= 0 || tail->nb != mg->nb) { set_error("sfem_mg_set_tail: bad arguments"); return SFEM_ERR_ARG; }
  MgLevel& L = mg->levels.back();
  L.P.nrows = L.A.nrows; L.P.ncols = n_tail; L.P.nnz = P_nnz; L.P.rowptr = P_rowptr; L.P.cols = P_cols; L.P.vals = P_vals;
  L.R.nrows = n_tail; L.R.ncols = L.A.nrows; L.R.nnz = R_nnz; L.R.rowptr = R_rowptr; L.R.cols = R_cols; L.R.vals = R_vals;
  graph_epoch_bump();
  mg->tail = tail;
  mg->n_tail = n_tail;
  cudaFree(mg->tail_b); cudaFree(mg->tail_x);
  const size_t bytes = (size_t)n_tail * mg->nb * sizeof(double);
  SFEM_CUDA(cudaMalloc(&mg->tail_b, bytes));
  SFEM_CUDA(cudaMalloc(&mg->tail_x, bytes));
  return SFEM_OK;
}

int sfem_mg_vcycle(sfem_mg_t mg, const double* b, double* x, void* stream) {
  if (!mg || !mg->ready) { set_error("mg handle not set up"); return SFEM_ERR_ARG; }
  SFEM_TRY(sell_ensure_all((cudaStream_t)stream));
  return mg_vcycle_level(mg, 0, b, x, (cudaStream_t)stream);
}

int sfem_mg_lambda_max(sfem_mg_t mg, double* h_out) {
  if (!mg) { set_error("null mg handle"); return SFEM_ERR_ARG; }
  SFEM_CUDA(cudaDeviceSynchronize());
  for (size_t l = 0; l < mg->levels.size(); ++l) {
    h_out[l] = 2.0;
    const bool dense_coarse = (l + 1 == mg->levels.size()) && mg->coarse_inv != nullptr;
    if (!dense_coarse && mg->ready)
      SFEM_CUDA(cudaMemcpy(&h_out[l], mg->levels[l].coef, sizeof(double), cudaMemcpyDeviceToHost));
  }
  return SFEM_OK;
}

void sfem_mg_destroy(sfem_mg_t mg) {
  if (!mg) return;
  graph_epoch_bump();                       // graphs of OTHER handles may replay this one (Stokes velocity block, dist tail)
  for (MgLevel& L : mg->levels) {
    cudaFree(L.dinv); cudaFree(L.r); cudaFree(L.d0); cudaFree(L.d1); cudaFree(L.x); cudaFree(L.b); cudaFree(L.coef);
  }
  cudaFree(mg->scratch);
  cudaFree(mg->tail_b);
  cudaFree(mg->tail_x);
  delete mg;
}

}  // extern "C"
