// Krylov drivers: PCG and flexible GMRES (the Taylor-Hood MINRES solver lives in sfem_stokes.cu).
//
// All recurrence scalars live in device memory; dot products are two-stage (per-block partials +
// a one-block scalar kernel that also advances the recurrence), so an iteration is a fixed sequence
// of launches with no host arithmetic.  The host only polls one double (the residual estimate) to
// decide when to stop.  Summation order is fixed -> results are bit-reproducible run to run.
#include "sfem_mg.h"
#include "sfem_graph.h"
#include "sfem_dist.h"

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
// (with a communicator the partial sums are all-reduced over the ranks inside these one-block kernels)
__global__ void k_cg_alpha(const double* __restrict__ partial, int np, double* __restrict__ S, DistDev D) {
  __shared__ double sh[33];
  double pq = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    dist_allreduce_scalars(D, &pq, 1);
    S[1] = pq; S[2] = S[0] / pq;
  }
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
                          double* __restrict__ S, int first, DistDev D) {
  __shared__ double sh[33];
  double rz = block_sum_array(p_rz, n_rz, sh);
  double rr = (n_rr > 0) ? block_sum_array(p_rr, n_rr, sh) : 0.0;
  if (threadIdx.x == 0) {
    double v[2] = {rz, rr};
    dist_allreduce_scalars(D, v, 2);
    rz = v[0]; rr = v[1];
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
// (ld = stride between the basis vectors: n on one GPU, n_loc = owned + ghosts when row-partitioned)
__global__ void __launch_bounds__(kThreads) k_multidot(int n, size_t ld, const double* __restrict__ V, int nvec,
                                                       const double* __restrict__ w, double* __restrict__ partial, int gx) {
  __shared__ double sh[33];
  const double* v = V + (size_t)blockIdx.y * ld;
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc = fma(v[i], w[i], acc);
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[(size_t)blockIdx.y * gx + blockIdx.x] = t;
}

// one block per vector: hcol[i] (+)= sum partial[i][:]; hstep[i] = this pass's coefficient
// (accumulate < 0: only hstep is written -- the multi-GPU path folds the GLOBAL sum into hcol in k_gs_allreduce)
__global__ void k_gs_coeff(const double* __restrict__ partial, int gx, double* __restrict__ hcol,
                           double* __restrict__ hstep, int accumulate) {
  __shared__ double sh[33];
  const int i = blockIdx.x;
  const double t = block_sum_array(partial + (size_t)i * gx, gx, sh);
  if (threadIdx.x == 0) {
    hstep[i] = t;
    if (accumulate >= 0) hcol[i] = accumulate ? hcol[i] + t : t;
  }
}

// multi-GPU: the nvec coefficients of one Gram-Schmidt pass are sums over the ranks; all-reduced in chunks of
// kAllreduceMaxK by one thread (hstep = this pass's local sums -> global), then folded into the Hessenberg column
__global__ void k_gs_allreduce(int nvec, double* __restrict__ hstep, double* __restrict__ hcol, int accumulate, DistDev D) {
  __shared__ double old[kAllreduceMaxK];
  for (int k0 = 0; k0 < nvec; k0 += kAllreduceMaxK) {
    const int K = min(kAllreduceMaxK, nvec - k0);
    if (threadIdx.x < K) old[threadIdx.x] = hstep[k0 + threadIdx.x];
    __syncthreads();
    if (threadIdx.x == 0) {
      double v[kAllreduceMaxK];
      for (int k = 0; k < K; ++k) v[k] = old[k];
      dist_allreduce_scalars(D, v, K);            // summed in rank order: bit-identical on every rank
      for (int k = 0; k < K; ++k) {
        hstep[k0 + k] = v[k];
        hcol[k0 + k] = accumulate ? hcol[k0 + k] + v[k] : v[k];
      }
    }
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kThreads) k_gs_update(int n, size_t ld, const double* __restrict__ V, int nvec,
                                                        const double* __restrict__ hstep, double* __restrict__ w,
                                                        double* __restrict__ partial) {
  __shared__ double sh[33];
  extern __shared__ double hs[];
  for (int k = threadIdx.x; k < nvec; k += blockDim.x) hs[k] = hstep[k];
  __syncthreads();
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double wi = w[i];
    for (int k = 0; k < nvec; ++k) wi = fma(-hs[k], V[(size_t)k * ld + i], wi);
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
                               double* __restrict__ misc, DistDev D) {
  __shared__ double sh[33];
  double ww = block_sum_array(p_ww, np, sh);
  if (threadIdx.x == 0) {
    dist_allreduce_scalars(D, &ww, 1);
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
__global__ void k_gmres_start(const double* __restrict__ partial, int np, double* __restrict__ g, double* __restrict__ misc,
                              DistDev D) {
  __shared__ double sh[33];
  double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    dist_allreduce_scalars(D, &t, 1);
    const double beta = sqrt(t);
    g[0] = beta;
    misc[0] = (beta > 0.0) ? 1.0 / beta : 0.0;
    misc[1] = beta;
  }
}

__global__ void k_lincomb_add(int n, size_t ld, const double* __restrict__ Z, int nvec, const double* __restrict__ y,
                              double* __restrict__ x) {
  extern __shared__ double ys[];
  for (int k = threadIdx.x; k < nvec; k += blockDim.x) ys[k] = y[k];
  __syncthreads();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double xi = x[i];
    for (int k = 0; k < nvec; ++k) xi = fma(ys[k], Z[(size_t)k * ld + i], xi);
    x[i] = xi;
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


// ------------------------------------------------------------------ lagged residual polling
// The residual estimate of iteration j is copied to pinned memory behind the iteration and read by the
// host while iteration j+1 already runs, so the device never idles on the host round trip.
struct Poller {
  double* pin = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init() {
    if (pin) return SFEM_OK;
    SFEM_CUDA(cudaMallocHost(&pin, 4 * sizeof(double)));
    SFEM_CUDA(cudaEventCreateWithFlags(&ev[0], cudaEventDisableTiming));
    SFEM_CUDA(cudaEventCreateWithFlags(&ev[1], cudaEventDisableTiming));
    return SFEM_OK;
  }
  int post(int slot, const double* dptr, cudaStream_t st) {
    SFEM_CUDA(cudaMemcpyAsync(pin + slot, dptr, sizeof(double), cudaMemcpyDeviceToHost, st));
    SFEM_CUDA(cudaEventRecord(ev[slot], st));
    return SFEM_OK;
  }
  int wait(int slot, double* out) {
    SFEM_CUDA(cudaEventSynchronize(ev[slot]));
    *out = pin[slot];
    return SFEM_OK;
  }
};
thread_local Poller t_poll;
thread_local WorkStream t_work;

// graph caches of the CG / FGMRES iteration bodies live in the multigrid handle that preconditions them (destroyed
// with it); Jacobi-preconditioned solves (no handle) use these thread-local ones.  Keys: sfem_graph.h.
thread_local GraphCache t_cg_graph, t_gmres_graph;

cudaStream_t pick_stream(cudaStream_t user, bool* forked) {
  *forked = false;
  if (user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread) {
    if (t_work.fork(user) == SFEM_OK) { *forked = true; return t_work.s; }
  }
  return user;
}

}  // namespace

}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_krylov_cg(int n, int nnz, const int* rowptr, const int* cols, const double* vals, sfem_mg_t mg,
                   const double* b, double* x, double rtol, int maxit, double* h_info, void* stream) {
  cudaStream_t user = (cudaStream_t)stream;
  SFEM_TRY(sell_ensure_all(user));          // dirty sliced-ELL mirrors are re-packed before any graph replay
  if (n <= 0) { set_error("cg: empty system"); return SFEM_ERR_ARG; }
  if (mg && (!mg->ready || mg->nb != 1)) { set_error("cg: multigrid not set up (or nb != 1)"); return SFEM_ERR_ARG; }
  Csr A; A.nrows = A.ncols = n; A.nnz = nnz; A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  // row-partitioned operator (multi-GPU): n = owned rows; vectors read by an SpMV carry the ghosts behind them
  const Halo* halo = find_halo(rowptr);
  if (halo) A.ncols = halo->dev.n_loc;
  const DistDev D = dist_dev();
  const size_t nn = (size_t)(halo ? halo->dev.n_loc : n);
  SFEM_TRY(t_ws.ensure(5 * nn + 3 * kMaxPartials + 64));
  SFEM_TRY(t_poll.init());
  bool forked = false;
  cudaStream_t st = pick_stream(user, &forked);
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
  // one iteration: q = A p, alpha, x/r update (+ r.r), z = M^-1 r, beta (stores r.r in S[4]), p update
  auto iteration = [&]() -> int {
    int npq = 0, nrz = 0;
    SFEM_TRY(spmv_dot(A, p, q, part0, &npq, st));
    k_cg_alpha<<<1, kThreads, 0, st>>>(part0, npq, S, D);
    SFEM_LAUNCH_CHECK();
    const int nrr = grid_for(n, kThreads * 4, 4);
    { Prof prof(PC_VEC, 48.0 * n, st);
    k_cg_update<<<nrr, kThreads, 0, st>>>(n, S, p, q, x, r, part1); }
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(precond_apply(mg, dinv, n, r, z, st));
    SFEM_TRY(vec_dot_partial(n, r, z, part0, &nrz, st));
    k_cg_beta<<<1, kThreads, 0, st>>>(part0, nrz, part1, nrr, S, 0, D);
    SFEM_LAUNCH_CHECK();
    { Prof prof(PC_VEC, 24.0 * n, st);
    k_cg_p<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, S, z, p); }
    SFEM_LAUNCH_CHECK();
    return SFEM_OK;
  };
  if (std::sqrt(rr0) > target && maxit > 0) {
    int np = 0;
    SFEM_TRY(precond_apply(mg, dinv, n, r, z, st));
    SFEM_TRY(vec_dot_partial(n, r, z, part0, &np, st));
    k_cg_beta<<<1, kThreads, 0, st>>>(part0, np, nullptr, 0, S, 1, D);
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(vec_copy(n, z, p, st));
    const bool use_graph = graphs_enabled();
    GraphCache& gc = mg ? mg->cg_graph : t_cg_graph;
    if (use_graph) {
      GraphKey key;
      key.a[0] = rowptr; key.a[1] = cols; key.a[2] = vals; key.a[3] = mg; key.a[4] = x; key.a[5] = t_ws.ptr;
      key.a[6] = b; key.a[7] = st;
      key.n = n; key.nnz = nnz; key.degree = mg ? mg->degree : 0; key.nranks = D.nranks; key.epoch = graph_epoch();
      if (!(gc.key == key) || gc.g.empty() || gc.g[0].exec == nullptr) {
        gc.invalidate(1);
        SFEM_TRY(graph_capture(st, gc.g[0], iteration));
        gc.key = key;
      }
    }
    bool done = false;
    for (it = 1; it <= maxit && !done; ++it) {
      if (use_graph) SFEM_TRY(graph_launch(gc.g[0], st));
      else SFEM_TRY(iteration());
      SFEM_TRY(t_poll.post(it & 1, S + 4, st));
      if (it > 1) {
        SFEM_TRY(t_poll.wait((it - 1) & 1, &rr));
        if (!(rr == rr)) { if (forked) t_work.join(user); set_error("cg: NaN residual"); return SFEM_ERR_NOCONV; }
        if (std::sqrt(rr) <= target) done = true;       // iteration `it` is already queued: keep its update
      }
    }
    --it;
    SFEM_TRY(t_poll.wait(it & 1, &rr));
  }
  double rel = 0.0;
  SFEM_TRY(true_relres(A, b, x, r, scratch, bnorm, &rel, st));
  if (forked) SFEM_TRY(t_work.join(user));
  h_info[0] = it; h_info[1] = rel; h_info[2] = (rel <= 10.0 * rtol) ? 1.0 : 0.0;
  h_info[3] = (bnorm > 0.0) ? std::sqrt(rr) / bnorm : std::sqrt(rr);
  return SFEM_OK;
}

int sfem_krylov_fgmres(int n, int nnz, const int* rowptr, const int* cols, const double* vals, sfem_mg_t mg,
                       const double* b, double* x, double rtol, int restart, int maxit, double* h_info,
                       void* stream) {
  cudaStream_t user = (cudaStream_t)stream;
  SFEM_TRY(sell_ensure_all(user));          // dirty sliced-ELL mirrors are re-packed before any graph replay
  if (n <= 0 || restart < 1) { set_error("fgmres: bad arguments"); return SFEM_ERR_ARG; }
  if (mg && (!mg->ready || mg->nb != 1)) { set_error("fgmres: multigrid not set up (or nb != 1)"); return SFEM_ERR_ARG; }
  Csr A; A.nrows = A.ncols = n; A.nnz = nnz; A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  // row-partitioned operator (multi-GPU): n = owned rows; every vector an SpMV reads (x, the z_j) carries the ghosts
  // behind the owned entries, so all basis vectors use the stride n_loc and the reductions run over the owned part
  const Halo* halo = find_halo(rowptr);
  if (halo) A.ncols = halo->dev.n_loc;
  const DistDev D = dist_dev();
  if (D.nranks > 1 && !halo) { set_error("fgmres: a communicator is active but the matrix is not row-partitioned"); return SFEM_ERR_ARG; }
  const int m = restart;
  const size_t nn = (size_t)(halo ? halo->dev.n_loc : n);
  const int gx = grid_for(n, kThreads * 4, 2);
  const size_t small = (size_t)(m + 1) * m + 4 * (size_t)(m + 2) + 16;
  SFEM_TRY(t_ws.ensure((2 * (size_t)m + 3) * nn + (size_t)(m + 1) * gx + 2 * kMaxPartials + small + 64));
  SFEM_TRY(t_poll.init());
  bool forked = false;
  cudaStream_t st = pick_stream(user, &forked);
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
  // Arnoldi step j: z_j = M^-1 v_j, w = A z_j, two classical Gram-Schmidt passes, Givens, v_{j+1}
  auto arnoldi = [&](int j) -> int {
    double* vj = V + (size_t)j * nn;
    double* zj = Z + (size_t)j * nn;
    SFEM_TRY(precond_apply(mg, dinv, n, vj, zj, st));
    SFEM_TRY(spmv(A, zj, nullptr, w, 0, st));
    const int nvec = j + 1;
    double* hcol = H + (size_t)j * (m + 1);
    int nww = 0;
    for (int pass = 0; pass < 2; ++pass) {
      dim3 grid(gx, nvec);
      { Prof prof(PC_VEC, 8.0 * n * (nvec + 1), st);
      k_multidot<<<grid, kThreads, 0, st>>>(n, nn, V, nvec, w, partial, gx); }
      SFEM_LAUNCH_CHECK();
      k_gs_coeff<<<nvec, kThreads, 0, st>>>(partial, gx, hcol, hstep, D.nranks > 1 ? -1 : pass);
      SFEM_LAUNCH_CHECK();
      if (D.nranks > 1) {
        k_gs_allreduce<<<1, 32, 0, st>>>(nvec, hstep, hcol, pass, D);
        SFEM_LAUNCH_CHECK();
      }
      nww = grid_for(n, kThreads * 4, 4);
      { Prof prof(PC_VEC, 8.0 * n * (nvec + 2), st);
      k_gs_update<<<nww, kThreads, nvec * sizeof(double), st>>>(n, nn, V, nvec, hstep, w, part1); }
      SFEM_LAUNCH_CHECK();
    }
    k_gmres_givens<<<1, kThreads, 0, st>>>(part1, nww, j, m, H, cs, sn, g, misc, D);
    SFEM_LAUNCH_CHECK();
    { Prof prof(PC_VEC, 16.0 * n, st);
    k_scale_dev<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, misc, w, V + (size_t)(j + 1) * nn); }
    SFEM_LAUNCH_CHECK();
    return SFEM_OK;
  };
  const bool use_graph = graphs_enabled();
  GraphCache& gc = mg ? mg->gmres_graph : t_gmres_graph;
  if (use_graph) {
    GraphKey key;
    key.a[0] = rowptr; key.a[1] = cols; key.a[2] = vals; key.a[3] = mg; key.a[4] = t_ws.ptr; key.a[7] = st;
    key.n = n; key.m = m; key.nnz = nnz; key.degree = mg ? mg->degree : 0; key.nranks = dist_dev().nranks;
    key.epoch = graph_epoch();
    if (!(gc.key == key) || (int)gc.g.size() != m) {
      gc.invalidate(m);
      gc.key = key;
    }
  }
  int total = 0;
  double est = 0.0;
  bool done = false;
  while (!done) {
    // r0 -> V0
    SFEM_TRY(spmv(A, x, b, w, 1, st));
    int np = 0;
    SFEM_TRY(vec_dot_partial(n, w, w, part1, &np, st));
    k_gmres_start<<<1, kThreads, 0, st>>>(part1, np, g, misc, D);
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(read_double(misc + 1, &est, st));
    if (est <= target || total >= maxit) break;
    k_scale_dev<<<grid_for(n, kThreads * 4), kThreads, 0, st>>>(n, misc, w, V);
    SFEM_LAUNCH_CHECK();
    int queued = 0, conv_at = -1;
    for (int j = 0; j < m && total + j < maxit; ++j) {
      if (use_graph) {
        GraphExec& ge = gc.g[j];
        if (ge.exec == nullptr) SFEM_TRY(graph_capture(st, ge, [&]() { return arnoldi(j); }));
        SFEM_TRY(graph_launch(ge, st));
      } else {
        SFEM_TRY(arnoldi(j));
      }
      SFEM_TRY(t_poll.post(j & 1, misc + 1, st));
      queued = j + 1;
      if (j > 0) {
        SFEM_TRY(t_poll.wait((j - 1) & 1, &est));
        if (!(est == est)) { if (forked) t_work.join(user); set_error("fgmres: NaN residual"); return SFEM_ERR_NOCONV; }
        if (est <= target) { conv_at = j - 1; break; }
      }
    }
    if (conv_at < 0 && queued > 0) {
      SFEM_TRY(t_poll.wait((queued - 1) & 1, &est));
      if (!(est == est)) { if (forked) t_work.join(user); set_error("fgmres: NaN residual"); return SFEM_ERR_NOCONV; }
      if (est <= target) conv_at = queued - 1;
    }
    const int k = (conv_at >= 0) ? conv_at + 1 : queued;  // columns used (a speculative extra one is ignored)
    total += k;
    if (conv_at >= 0 || total >= maxit) done = true;
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
      k_lincomb_add<<<grid_for(n, kThreads * 4), kThreads, k * sizeof(double), st>>>(n, nn, Z, k, hstep, x);
      SFEM_LAUNCH_CHECK();
      SFEM_CUDA(cudaStreamSynchronize(st));              // hy is reused by the next cycle
    }
  }
  double rel = 0.0;
  SFEM_TRY(true_relres(A, b, x, w, scratch, bnorm, &rel, st));
  if (forked) SFEM_TRY(t_work.join(user));
  h_info[0] = total; h_info[1] = rel; h_info[2] = (rel <= 10.0 * rtol) ? 1.0 : 0.0;
  h_info[3] = (bnorm > 0.0) ? est / bnorm : est;
  return SFEM_OK;
}

}  // extern "C"
