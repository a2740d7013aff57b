// Krylov drivers: PCG, flexible GMRES and preconditioned MINRES (Taylor-Hood saddle point).
//
// All recurrence scalars live in device memory; dot products are two-stage (per-block partials +
// a one-block scalar kernel that also advances the recurrence), so an iteration is a fixed sequence
// of launches with no host arithmetic.  The host only polls one double (the residual estimate) to
// decide when to stop.  Summation order is fixed -> results are bit-reproducible run to run.
#include "sfem_mg.h"

#include <cmath>
#include <cstring>
#include <vector>

namespace sfem {

namespace {

// ------------------------------------------------------------------ grow-only device workspace
struct Workspace {
  double* ptr = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return SFEM_OK;
    if (ptr) cudaFree(ptr);
    ptr = nullptr; cap = 0;
    SFEM_CUDA(cudaMalloc(&ptr, n * sizeof(double)));
    cap = n;
    return SFEM_OK;
  }
};
thread_local Workspace t_ws;

int read_double(const double* dptr, double* h, cudaStream_t st) {
  SFEM_CUDA(cudaMemcpyAsync(h, dptr, sizeof(double), cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  return SFEM_OK;
}

int precond_apply(sfem_mg* mg, const double* dinv, int n, const double* r, double* z, cudaStream_t st) {
  if (mg) return mg_vcycle_level(mg, 0, r, z, st);
  return vec_mul_scale(n, 1.0, dinv, r, z, st);
}

// ================================================================== CG
// S: [0]=rz [1]=pq [2]=alpha [3]=beta [4]=rr [5]=bb
__global__ void k_cg_alpha(const double* __restrict__ partial, int np, double* __restrict__ S) {
  __shared__ double sh[33];
  const double pq = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) { S[1] = pq; S[2] = S[0] / pq; }
}

__global__ void __launch_bounds__(kThreads) k_cg_update(int n, const double* __restrict__ S, const double* __restrict__ p,
                                                        const double* __restrict__ q, double* __restrict__ x,
                                                        double* __restrict__ r, double* __restrict__ partial) {
  __shared__ double sh[33];
  const double alpha = S[2];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    x[i] = fma(alpha, p[i], x[i]);
    const double ri = fma(-alpha, q[i], r[i]);
    r[i] = ri;
    acc = fma(ri, ri, acc);
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

__global__ void k_cg_beta(const double* __restrict__ p_rz, int n_rz, const double* __restrict__ p_rr, int n_rr,
                          double* __restrict__ S, int first) {
  __shared__ double sh[33];
  const double rz = block_sum_array(p_rz, n_rz, sh);
  const double rr = (n_rr > 0) ? block_sum_array(p_rr, n_rr, sh) : 0.0;
  if (threadIdx.x == 0) {
    S[3] = first ? 0.0 : rz / S[0];
    S[0] = rz;
    if (n_rr > 0) S[4] = rr;
  }
}

__global__ void k_cg_p(int n, const double* __restrict__ S, const double* __restrict__ z, double* __restrict__ p) {
  const double beta = S[3];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) p[i] = fma(beta, p[i], z[i]);
}

__global__ void k_store_sum(const double* __restrict__ partial, int np, double* __restrict__ out) {
  __shared__ double sh[33];
  const double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) out[0] = t;
}

// ================================================================== FGMRES
// Hess layout (device): H[(m+1) x m] column-major (col j at H + j*(m+1)), cs[m], sn[m], g[m+1], misc[4]
__global__ void __launch_bounds__(kThreads) k_multidot(int n, const double* __restrict__ V, int nvec,
                                                       const double* __restrict__ w, double* __restrict__ partial, int gx) {
  __shared__ double sh[33];
  const double* v = V + (size_t)blockIdx.y * n;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc = fma(v[i], w[i], acc);
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gx + blockIdx.x] = t;
}

// one block per vector: hcol[i] (+)= sum partial[i][:]; hstep[i] = this pass's coefficient
__global__ void k_gs_coeff(const double* __restrict__ partial, int gx, double* __restrict__ hcol,
                           double* __restrict__ hstep, int accumulate) {
  __shared__ double sh[33];
  const int i = blockIdx.x;
  const double t = block_sum_array(partial + (size_t)i * gx, gx, sh);
  if (threadIdx.x == 0) {
    hstep[i] = t;
    hcol[i] = accumulate ? hcol[i] + t : t;
  }
}

__global__ void __launch_bounds__(kThreads) k_gs_update(int n, const double* __restrict__ V, int nvec,
                                                        const double* __restrict__ hstep, double* __restrict__ w,
                                                        double* __restrict__ partial) {
  __shared__ double sh[33];
  extern __shared__ double hs[];
  for (int k = threadIdx.x; k < nvec; k += blockDim.x) hs[k] = hstep[k];
  __syncthreads();
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double wi = w[i];
    for (int k = 0; k < nvec; ++k) wi = fma(-hs[k], V[(size_t)k * n + i], wi);
    w[i] = wi;
    acc = fma(wi, wi, acc);
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// single block: h_{j+1,j} = ||w||, apply old rotations to column j, new rotation, update g.
// misc[0] = 1/h_{j+1,j}, misc[1] = |g_{j+1}| (residual estimate)
__global__ void k_gmres_givens(const double* __restrict__ p_ww, int np, int j, int m, double* __restrict__ H,
                               double* __restrict__ cs, double* __restrict__ sn, double* __restrict__ g,
                               double* __restrict__ misc) {
  __shared__ double sh[33];
  const double ww = block_sum_array(p_ww, np, sh);
  if (threadIdx.x == 0) {
    double* h = H + (size_t)j * (m + 1);
    const double hn = sqrt(ww);
    h[j + 1] = hn;
    misc[0] = (hn > 0.0) ? 1.0 / hn : 0.0;
    for (int i = 0; i < j; ++i) {
      const double t = cs[i] * h[i] + sn[i] * h[i + 1];
      h[i + 1] = -sn[i] * h[i] + cs[i] * h[i + 1];
      h[i] = t;
    }
    const double a = h[j], b = h[j + 1];
    const double rr = sqrt(a * a + b * b);
    const double c = (rr > 0.0) ? a / rr : 1.0, s = (rr > 0.0) ? b / rr : 0.0;
    cs[j] = c; sn[j] = s;
    h[j] = rr; h[j + 1] = 0.0;
    const double gj = g[j];
    g[j] = c * gj;
    g[j + 1] = -s * gj;
    misc[1] = fabs(g[j + 1]);
  }
}

__global__ void k_scale_dev(int n, const double* __restrict__ scal, const double* __restrict__ w, double* __restrict__ v) {
  const double a = scal[0];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v[i] = a * w[i];
}

// beta = sqrt(sum); g[0] = beta; misc[0] = 1/beta; misc[1] = beta
__global__ void k_gmres_start(const double* __restrict__ partial, int np, double* __restrict__ g, double* __restrict__ misc) {
  __shared__ double sh[33];
  const double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    const double beta = sqrt(t);
    g[0] = beta;
    misc[0] = (beta > 0.0) ? 1.0 / beta : 0.0;
    misc[1] = beta;
  }
}

__global__ void k_lincomb_add(int n, const double* __restrict__ Z, int nvec, const double* __restrict__ y,
                              double* __restrict__ x) {
  extern __shared__ double ys[];
  for (int k = threadIdx.x; k < nvec; k += blockDim.x) ys[k] = y[k];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double xi = x[i];
    for (int k = 0; k < nvec; ++k) xi = fma(ys[k], Z[(size_t)k * n + i], xi);
    x[i] = xi;
  }
}

// ================================================================== MINRES
// S: 0 gamma_prev, 1 gamma, 2 gamma_next, 3 delta, 4 eta, 5 c_prev, 6 c, 7 s_prev, 8 s,
//    9 a1, 10 a2, 11 a3, 12 xcoef (= c_next * eta_old), 13 gamma1, 14 coefA (delta/gamma), 15 coefB (gamma/gamma_prev)
__global__ void k_minres_init(const double* __restrict__ partial, int np, double* __restrict__ S) {
  __shared__ double sh[33];
  const double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    const double gamma = sqrt(fabs(t));
    S[0] = 1.0; S[1] = gamma; S[2] = 0.0; S[3] = 0.0; S[4] = gamma;
    S[5] = 1.0; S[6] = 1.0; S[7] = 0.0; S[8] = 0.0; S[13] = gamma;
  }
}

// z /= gamma
__global__ void k_minres_scale(int n, const double* __restrict__ S, double* __restrict__ z) {
  const double inv = (S[1] > 0.0) ? 1.0 / S[1] : 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) z[i] *= inv;
}

__global__ void k_minres_delta(const double* __restrict__ partial, int np, double* __restrict__ S) {
  __shared__ double sh[33];
  const double delta = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    S[3] = delta;
    S[14] = delta / S[1];
    S[15] = S[1] / S[0];
  }
}

// v_next = Az - coefA v - coefB v_prev   (written over v_prev, which is then the new v)
__global__ void k_minres_vnext(int n, const double* __restrict__ S, const double* __restrict__ Az,
                               const double* __restrict__ v, double* __restrict__ v_prev) {
  const double a = S[14], b = S[15];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    v_prev[i] = Az[i] - a * v[i] - b * v_prev[i];
}

__global__ void k_minres_rot(const double* __restrict__ partial, int np, double* __restrict__ S) {
  __shared__ double sh[33];
  const double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    const double gamma = S[1], delta = S[3], eta = S[4];
    const double c_prev = S[5], c = S[6], s_prev = S[7], s = S[8];
    const double gamma_next = sqrt(fabs(t));
    const double a0 = c * delta - c_prev * s * gamma;
    const double a1 = sqrt(a0 * a0 + gamma_next * gamma_next);
    const double a2 = s * delta + c_prev * c * gamma;
    const double a3 = s_prev * gamma;
    const double c_next = a0 / a1, s_next = gamma_next / a1;
    S[9] = a1; S[10] = a2; S[11] = a3;
    S[12] = c_next * eta;
    S[4] = -s_next * eta;
    S[0] = gamma; S[1] = gamma_next;
    S[5] = c; S[6] = c_next; S[7] = s; S[8] = s_next;
  }
}

// w_next = (z - a3 w_prev - a2 w)/a1 (over w_prev); x += xcoef w_next
__global__ void k_minres_wx(int n, const double* __restrict__ S, const double* __restrict__ z,
                            const double* __restrict__ w, double* __restrict__ w_prev, double* __restrict__ x) {
  const double inv = 1.0 / S[9], a2 = S[10], a3 = S[11], xc = S[12];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double wn = (z[i] - a3 * w_prev[i] - a2 * w[i]) * inv;
    w_prev[i] = wn;
    x[i] = fma(xc, wn, x[i]);
  }
}

int true_relres(const Csr& A, const double* b, const double* x, double* r, double* scratch, double bnorm,
                double* out, cudaStream_t st) {
  SFEM_TRY(spmv(A, x, b, r, 1, st));
  double rr = 0.0;
  SFEM_TRY(vec_dot_host(A.nrows, r, r, scratch, &rr, st));
  *out = (bnorm > 0.0) ? std::sqrt(rr) / bnorm : std::sqrt(rr);
  return SFEM_OK;
}

}  // namespace

}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_krylov_cg(int n, int nnz, const int* rowptr, const int* cols, const double* vals, sfem_mg_t mg,
                   const double* b, double* x, double rtol, int maxit, double* h_info, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0) { set_error("cg: empty system"); return SFEM_ERR_ARG; }
  if (mg && !mg->ready) { set_error("cg: multigrid not set up"); return SFEM_ERR_ARG; }
  Csr A; A.nrows = A.ncols = n; A.nnz = nnz; A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  const size_t nn = (size_t)n;
  SFEM_TRY(t_ws.ensure(5 * nn + 3 * kMaxPartials + 64));
  double* r = t_ws.ptr; double* z = r + nn; double* p = z + nn; double* q = p + nn; double* dinv = q + nn;
  double* part0 = dinv + nn; double* part1 = part0 + kMaxPartials; double* scratch = part1 + kMaxPartials;
  double* S = scratch + kMaxPartials + 8;
  if (!mg) SFEM_TRY(extract_diag_inv(A, dinv, st));
  double bb = 0.0;
  SFEM_TRY(vec_dot_host(n, b, b, scratch, &bb, st));
  const double bnorm = std::sqrt(bb);
  SFEM_TRY(spmv(A, x, b, r, 1, st));
  double rr0 = 0.0;
  SFEM_TRY(vec_dot_host(n, r, r, scratch, &rr0, st));
  int it = 0;
  double rr = rr0;
  const double target = rtol * (bnorm > 0.0 ? bnorm : 1.0);
  if (std::sqrt(rr0) > target) {
    int np = 0;
    SFEM_TRY(precond_apply(mg, dinv, n, r, z, st));
    SFEM_TRY(vec_dot_partial(n, r, z, part0, &np, st));
    k_cg_beta<<<1, kThreads, 0, st>>>(part0, np, nullptr, 0, S, 1);
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(vec_copy(n, z, p, st));
    for (it = 1; it <= maxit; ++it) {
      int npq = 0, nrr = 0, nrz = 0;
      SFEM_TRY(spmv_dot(A, p, q, part0, &npq, st));
      k_cg_alpha<<<1, kThreads, 0, st>>>(part0, npq, S);
      SFEM_LAUNCH_CHECK();
      nrr = grid_for(n, kThreads * 4, 4);
      { Prof prof(PC_VEC, 48.0 * n, st);
      k_cg_update<<<nrr, kThreads, 0, st>>>(n, S, p, q, x, r, part1); }
      SFEM_LAUNCH_CHECK();
      SFEM_TRY(precond_apply(mg, dinv, n, r, z, st));
      SFEM_TRY(vec_dot_partial(n, r, z, part0, &nrz, st));
      k_cg_beta<<<1, kThreads, 0, st>>>(part0, nrz, part1, nrr, S, 0);
      SFEM_LAUNCH_CHECK();
      SFEM_TRY(read_double(S + 4, &rr, st));
      if (!(rr == rr)) { set_error("cg: NaN residual"); return SFEM_ERR_NOCONV; }
      if (std::sqrt(rr) <= target) break;
      k_cg_p<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, S, z, p);
      SFEM_LAUNCH_CHECK();
    }
    if (it > maxit) it = maxit;
  }
  double rel = 0.0;
  SFEM_TRY(true_relres(A, b, x, r, scratch, bnorm, &rel, st));
  h_info[0] = it; h_info[1] = rel; h_info[2] = (rel <= 10.0 * rtol) ? 1.0 : 0.0;
  h_info[3] = (bnorm > 0.0) ? std::sqrt(rr) / bnorm : std::sqrt(rr);
  return SFEM_OK;
}

int sfem_krylov_fgmres(int n, int nnz, const int* rowptr, const int* cols, const double* vals, sfem_mg_t mg,
                       const double* b, double* x, double rtol, int restart, int maxit, double* h_info,
                       void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 0 || restart < 1) { set_error("fgmres: bad arguments"); return SFEM_ERR_ARG; }
  if (mg && !mg->ready) { set_error("fgmres: multigrid not set up"); return SFEM_ERR_ARG; }
  Csr A; A.nrows = A.ncols = n; A.nnz = nnz; A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  const int m = restart;
  const size_t nn = (size_t)n;
  const int gx = grid_for(n, kThreads * 4, 2);
  const size_t small = (size_t)(m + 1) * m + 4 * (size_t)(m + 2) + 16;
  SFEM_TRY(t_ws.ensure((2 * (size_t)m + 3) * nn + (size_t)(m + 1) * gx + 2 * kMaxPartials + small + 64));
  double* V = t_ws.ptr;                       // (m+1) vectors
  double* Z = V + (size_t)(m + 1) * nn;       // m vectors
  double* w = Z + (size_t)m * nn;
  double* dinv = w + nn;
  double* partial = dinv + nn;                // (m+1)*gx
  double* part1 = partial + (size_t)(m + 1) * gx;
  double* scratch = part1 + kMaxPartials;
  double* H = scratch + kMaxPartials + 8;     // (m+1)*m
  double* cs = H + (size_t)(m + 1) * m;
  double* sn = cs + m + 1;
  double* g = sn + m + 1;
  double* hstep = g + m + 2;
  double* misc = hstep + m + 2;
  if (!mg) SFEM_TRY(extract_diag_inv(A, dinv, st));
  double bb = 0.0;
  SFEM_TRY(vec_dot_host(n, b, b, scratch, &bb, st));
  const double bnorm = std::sqrt(bb);
  const double target = rtol * (bnorm > 0.0 ? bnorm : 1.0);
  std::vector<double> hH((size_t)(m + 1) * m), hg(m + 2), hy(m);
  int total = 0;
  double est = 0.0;
  bool done = false;
  while (!done) {
    // r0 -> V0
    SFEM_TRY(spmv(A, x, b, w, 1, st));
    int np = 0;
    SFEM_TRY(vec_dot_partial(n, w, w, part1, &np, st));
    k_gmres_start<<<1, kThreads, 0, st>>>(part1, np, g, misc);
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(read_double(misc + 1, &est, st));
    if (est <= target || total >= maxit) break;
    k_scale_dev<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, misc, w, V);
    SFEM_LAUNCH_CHECK();
    int j = 0;
    for (; j < m && total < maxit; ++j) {
      double* vj = V + (size_t)j * nn;
      double* zj = Z + (size_t)j * nn;
      SFEM_TRY(precond_apply(mg, dinv, n, vj, zj, st));
      SFEM_TRY(spmv(A, zj, nullptr, w, 0, st));
      const int nvec = j + 1;
      double* hcol = H + (size_t)j * (m + 1);
      int nww = 0;
      for (int pass = 0; pass < 2; ++pass) {             // classical Gram-Schmidt, twice
        dim3 grid(gx, nvec);
        { Prof prof(PC_VEC, 8.0 * n * (nvec + 1), st);
        k_multidot<<<grid, kThreads, 0, st>>>(n, V, nvec, w, partial, gx); }
        SFEM_LAUNCH_CHECK();
        k_gs_coeff<<<nvec, kThreads, 0, st>>>(partial, gx, hcol, hstep, pass);
        SFEM_LAUNCH_CHECK();
        nww = grid_for(n, kThreads * 4, 4);
        { Prof prof(PC_VEC, 8.0 * n * (nvec + 2), st);
        k_gs_update<<<nww, kThreads, nvec * sizeof(double), st>>>(n, V, nvec, hstep, w, part1); }
        SFEM_LAUNCH_CHECK();
      }
      k_gmres_givens<<<1, kThreads, 0, st>>>(part1, nww, j, m, H, cs, sn, g, misc);
      SFEM_LAUNCH_CHECK();
      k_scale_dev<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, misc, w, V + (size_t)(j + 1) * nn);
      SFEM_LAUNCH_CHECK();
      ++total;
      SFEM_TRY(read_double(misc + 1, &est, st));
      if (!(est == est)) { set_error("fgmres: NaN residual"); return SFEM_ERR_NOCONV; }
      if (est <= target) { ++j; done = true; break; }
    }
    if (total >= maxit) done = true;
    const int k = j;                                      // columns built in this cycle
    if (k > 0) {
      SFEM_CUDA(cudaMemcpyAsync(hH.data(), H, hH.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
      SFEM_CUDA(cudaMemcpyAsync(hg.data(), g, (m + 1) * sizeof(double), cudaMemcpyDeviceToHost, st));
      SFEM_CUDA(cudaStreamSynchronize(st));
      for (int i = k - 1; i >= 0; --i) {                  // back substitution on the rotated Hessenberg
        double s = hg[i];
        for (int c = i + 1; c < k; ++c) s -= hH[(size_t)c * (m + 1) + i] * hy[c];
        hy[i] = s / hH[(size_t)i * (m + 1) + i];
      }
      SFEM_CUDA(cudaMemcpyAsync(hstep, hy.data(), k * sizeof(double), cudaMemcpyHostToDevice, st));
      k_lincomb_add<<<grid_for(n, kThreads * 4), kThreads, k * sizeof(double), st>>>(n, Z, k, hstep, x);
      SFEM_LAUNCH_CHECK();
      SFEM_CUDA(cudaStreamSynchronize(st));              // hy is reused by the next cycle
    }
  }
  double rel = 0.0;
  SFEM_TRY(true_relres(A, b, x, w, scratch, bnorm, &rel, st));
  h_info[0] = total; h_info[1] = rel; h_info[2] = (rel <= 10.0 * rtol) ? 1.0 : 0.0;
  h_info[3] = (bnorm > 0.0) ? est / bnorm : est;
  return SFEM_OK;
}

int sfem_krylov_minres_stokes(int n2, int nv, int nnz, const int* rowptr, const int* cols, const double* vals,
                              sfem_mg_t mg, int Mp_nnz, const int* Mp_rowptr, const int* Mp_cols,
                              const double* Mp_vals, const double* b, double* x, double rtol, int maxit,
                              double* h_info, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  if (n2 <= 0 || nv <= 0 || !mg || !mg->ready) { set_error("minres: bad arguments / multigrid not set up"); return SFEM_ERR_ARG; }
  if (mg->levels[0].A.nrows != n2) { set_error("minres: multigrid size != velocity component size"); return SFEM_ERR_ARG; }
  const int n = 2 * n2 + nv;
  Csr A; A.nrows = A.ncols = n; A.nnz = nnz; A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  Csr Mp; Mp.nrows = Mp.ncols = nv; Mp.nnz = Mp_nnz; Mp.rowptr = Mp_rowptr; Mp.cols = Mp_cols; Mp.vals = Mp_vals;
  const size_t nn = (size_t)n;
  SFEM_TRY(t_ws.ensure(6 * nn + 4 * (size_t)nv + 2 * kMaxPartials + 96));
  double* va = t_ws.ptr; double* vb = va + nn; double* z = vb + nn; double* Az = z + nn;
  double* wa = Az + nn; double* wb = wa + nn;
  double* mp_dinv = wb + nn; double* mp_r = mp_dinv + nv; double* mp_d0 = mp_r + nv; double* mp_d1 = mp_d0 + nv;
  double* part = mp_d1 + nv; double* scratch = part + kMaxPartials; double* S = scratch + kMaxPartials + 8;
  SFEM_TRY(extract_diag_inv(Mp, mp_dinv, st));
  auto precond = [&](const double* r, double* out) -> int {
    SFEM_TRY(mg_vcycle_level(mg, 0, r, out, st));
    SFEM_TRY(mg_vcycle_level(mg, 0, r + n2, out + n2, st));
    // P1 mass matrix with Jacobi scaling has spectrum in [1/2, 2]: 4 Chebyshev steps ~ exact solve
    return smooth(Mp, mp_dinv, 2.0, 4.0, 4, r + 2 * (size_t)n2, out + 2 * (size_t)n2, mp_r, mp_d0, mp_d1, true, st);
  };
  double bb = 0.0;
  SFEM_TRY(vec_dot_host(n, b, b, scratch, &bb, st));
  const double bnorm = std::sqrt(bb);
  // v = b - A x ; v_prev = 0 ; w = w_prev = 0
  double* v = va; double* v_prev = vb; double* w = wa; double* w_prev = wb;
  SFEM_TRY(spmv(A, x, b, v, 1, st));
  SFEM_CUDA(cudaMemsetAsync(v_prev, 0, nn * sizeof(double), st));
  SFEM_CUDA(cudaMemsetAsync(w, 0, nn * sizeof(double), st));
  SFEM_CUDA(cudaMemsetAsync(w_prev, 0, nn * sizeof(double), st));
  SFEM_TRY(precond(v, z));
  int np = 0;
  SFEM_TRY(vec_dot_partial(n, z, v, part, &np, st));
  k_minres_init<<<1, kThreads, 0, st>>>(part, np, S);
  SFEM_LAUNCH_CHECK();
  double gamma1 = 0.0, eta = 0.0;
  SFEM_TRY(read_double(S + 13, &gamma1, st));
  int it = 0;
  eta = gamma1;
  if (gamma1 > 0.0) {
    const int gv = grid_for(n, kThreads * 4);
    for (it = 1; it <= maxit; ++it) {
      { Prof prof(PC_VEC, 16.0 * n, st);
      k_minres_scale<<<gv, kThreads, 0, st>>>(n, S, z); }
      SFEM_LAUNCH_CHECK();
      SFEM_TRY(spmv_dot(A, z, Az, part, &np, st));
      k_minres_delta<<<1, kThreads, 0, st>>>(part, np, S);
      SFEM_LAUNCH_CHECK();
      { Prof prof(PC_VEC, 32.0 * n, st);
      k_minres_vnext<<<gv, kThreads, 0, st>>>(n, S, Az, v, v_prev); }
      SFEM_LAUNCH_CHECK();
      { double* t = v; v = v_prev; v_prev = t; }          // v now holds v_{j+1}, v_prev holds v_j
      // z_j is still needed for w_{j+1}: keep it in Az's place after the precond writes z_next
      SFEM_TRY(vec_copy(n, z, Az, st));
      SFEM_TRY(precond(v, z));
      SFEM_TRY(vec_dot_partial(n, z, v, part, &np, st));
      k_minres_rot<<<1, kThreads, 0, st>>>(part, np, S);
      SFEM_LAUNCH_CHECK();
      { Prof prof(PC_VEC, 48.0 * n, st);
      k_minres_wx<<<gv, kThreads, 0, st>>>(n, S, Az, w, w_prev, x); }
      SFEM_LAUNCH_CHECK();
      { double* t = w; w = w_prev; w_prev = t; }
      SFEM_TRY(read_double(S + 4, &eta, st));
      if (!(eta == eta)) { set_error("minres: NaN residual"); return SFEM_ERR_NOCONV; }
      if (std::fabs(eta) <= rtol * gamma1) break;
    }
    if (it > maxit) it = maxit;
  }
  double rel = 0.0;
  SFEM_TRY(true_relres(A, b, x, Az, scratch, bnorm, &rel, st));
  h_info[0] = it; h_info[1] = rel; h_info[2] = (std::fabs(eta) <= rtol * gamma1) ? 1.0 : 0.0;
  h_info[3] = (gamma1 > 0.0) ? std::fabs(eta) / gamma1 : 0.0;
  return SFEM_OK;
}

}  // extern "C"
