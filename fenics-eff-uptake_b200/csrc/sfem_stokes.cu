// Taylor-Hood Stokes solver (reference solvers.py:237-306: what dolfin's solve() hands to sparse LU).
//
// Unknowns: [ velocity, interleaved (ux_0, uy_0, ux_1, uy_1, ...) | pressure ].  The operator is kept
// in block form
//        [ K   0   Bx^T ]
//    A = [ 0   K   By^T ]          K : scalar P2 stiffness (same for both components, read ONCE per
//        [ Bx  By  0    ]              apply with the two components as interleaved right-hand sides)
// whose values are copied out of the assembled Taylor-Hood CSR (dolfin clique pattern, Dirichlet rows
// and columns eliminated), i.e. it is the assembled matrix minus its structural zeros.
//
// Preconditioned MINRES with the block-diagonal SPD preconditioner
//    M^-1 = diag( V-cycle(K) [both components in one 2-RHS cycle],  S^-1 ),
//    S^-1 = Cheb_4(Mp) + Z C Z^T,
// Mp = P1 pressure mass matrix (4 Chebyshev-Jacobi steps ~ exact), Z = 1-D hat functions in x over the
// channel and C >= 0 a small dense matrix built on the host from the lubrication (Reynolds) operator:
// it lifts the O((H/L)^2) eigenvalues of the long-channel Schur complement that would otherwise
// cost MINRES ~60% more iterations.  C = 0 (nz = 0) gives the classical mass-matrix preconditioner.
//
// The iteration body is a fixed launch sequence with all recurrence scalars in device memory; it is
// captured into two CUDA graphs (even / odd buffer parity) and replayed.  The host polls the
// residual estimate one iteration late, so the device never waits for the host.
#include "sfem_mg.h"
#include "sfem_graph.h"
#include "sfem_dist.h"

#include <cmath>
#include <cstdlib>
#include <vector>

namespace sfem {

namespace {

// S: 0 gamma_prev, 1 gamma, 3 delta, 4 eta, 5 c_prev, 6 c, 7 s_prev, 8 s, 9 a1, 10 a2, 11 a3,
//    12 xcoef, 13 gamma1, 14 delta/gamma, 15 gamma/gamma_prev, 16 1/gamma (of the current z)
__global__ void k_sm_init(const double* __restrict__ partial, int np, double* __restrict__ S, DistDev D) {
  __shared__ double sh[33];
  double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    dist_allreduce_scalars(D, &t, 1);
    const double gamma = sqrt(fabs(t));
    S[0] = 1.0; S[1] = gamma; S[2] = 0.0; S[3] = 0.0; S[4] = gamma;
    S[5] = 1.0; S[6] = 1.0; S[7] = 0.0; S[8] = 0.0; S[13] = gamma;
    S[16] = (gamma > 0.0) ? 1.0 / gamma : 0.0;
  }
}

__global__ void k_sm_delta(const double* __restrict__ pu, int npu, const double* __restrict__ pp, int npp,
                           double* __restrict__ S, DistDev D) {
  __shared__ double sh[33];
  const double du = block_sum_array(pu, npu, sh);
  const double dp = block_sum_array(pp, npp, sh);
  if (threadIdx.x == 0) {
    double dsum = du + dp;
    dist_allreduce_scalars(D, &dsum, 1);
    const double gamma = S[1];
    const double invg = (gamma > 0.0) ? 1.0 / gamma : 0.0;
    const double delta = dsum * invg * invg;
    S[3] = delta;
    S[14] = delta * invg;
    S[15] = gamma / S[0];
    S[16] = invg;
  }
}

// v_prev <- Az/gamma - (delta/gamma) v - (gamma/gamma_prev) v_prev     (then v_prev is the new v)
// (VEC: 16-byte aligned arrays are streamed as double2 pairs, two pairs in flight per thread)
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_sm_vnext(int n, const double* __restrict__ S, const double* __restrict__ Az,
                                                       const double* __restrict__ v, double* __restrict__ v_prev) {
  const double invg = S[16], a = S[14], b = S[15];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (VEC) {
    const int n2 = n >> 1;
    const double2* A2 = reinterpret_cast<const double2*>(Az);
    const double2* V2 = reinterpret_cast<const double2*>(v);
    double2* P2 = reinterpret_cast<double2*>(v_prev);
    for (int i = tid; i < n2; i += 2 * nt) {
      const int j = i + nt;
      const bool two = j < n2;
      const double2 a0 = A2[i], v0 = V2[i], p0 = P2[i];
      double2 a1 = a0, v1 = v0, p1 = p0;
      if (two) { a1 = A2[j]; v1 = V2[j]; p1 = P2[j]; }
      P2[i] = make_double2(a0.x * invg - a * v0.x - b * p0.x, a0.y * invg - a * v0.y - b * p0.y);
      if (two) P2[j] = make_double2(a1.x * invg - a * v1.x - b * p1.x, a1.y * invg - a * v1.y - b * p1.y);
    }
    if ((n & 1) && tid == 0) v_prev[n - 1] = Az[n - 1] * invg - a * v[n - 1] - b * v_prev[n - 1];
  } else {
    for (int i = tid; i < n; i += nt) v_prev[i] = Az[i] * invg - a * v[i] - b * v_prev[i];
  }
}

__global__ void k_sm_rot(const double* __restrict__ partial, int np, double* __restrict__ S, DistDev D) {
  __shared__ double sh[33];
  double t = block_sum_array(partial, np, sh);
  if (threadIdx.x == 0) {
    dist_allreduce_scalars(D, &t, 1);
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

// w_prev <- (z/gamma - a3 w_prev - a2 w)/a1 ; x += xcoef w_prev        (then w_prev is the new w)
template <bool VEC>
__global__ void __launch_bounds__(kThreads) k_sm_wx(int n, const double* __restrict__ S, const double* __restrict__ z,
                                                    const double* __restrict__ w, double* __restrict__ w_prev,
                                                    double* __restrict__ x) {
  const double invg = S[16], inv = 1.0 / S[9], a2 = S[10], a3 = S[11], xc = S[12];
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nt = gridDim.x * blockDim.x;
  if (VEC) {
    const int n2 = n >> 1;
    const double2* Z2 = reinterpret_cast<const double2*>(z);
    const double2* W2 = reinterpret_cast<const double2*>(w);
    double2* P2 = reinterpret_cast<double2*>(w_prev);
    double2* X2 = reinterpret_cast<double2*>(x);
    for (int i = tid; i < n2; i += 2 * nt) {
      const int j = i + nt;
      const bool two = j < n2;
      const double2 z0 = Z2[i], w0 = W2[i], p0 = P2[i], x0 = X2[i];
      double2 z1 = z0, w1 = w0, p1 = p0, x1 = x0;
      if (two) { z1 = Z2[j]; w1 = W2[j]; p1 = P2[j]; x1 = X2[j]; }
      const double2 n0 = make_double2((z0.x * invg - a3 * p0.x - a2 * w0.x) * inv, (z0.y * invg - a3 * p0.y - a2 * w0.y) * inv);
      P2[i] = n0;
      X2[i] = make_double2(fma(xc, n0.x, x0.x), fma(xc, n0.y, x0.y));
      if (two) {
        const double2 n1 = make_double2((z1.x * invg - a3 * p1.x - a2 * w1.x) * inv, (z1.y * invg - a3 * p1.y - a2 * w1.y) * inv);
        P2[j] = n1;
        X2[j] = make_double2(fma(xc, n1.x, x1.x), fma(xc, n1.y, x1.y));
      }
    }
    if ((n & 1) && tid == 0) {
      const int i = n - 1;
      const double wn = (z[i] * invg - a3 * w_prev[i] - a2 * w[i]) * inv;
      w_prev[i] = wn;
      x[i] = fma(xc, wn, x[i]);
    }
  } else {
    for (int i = tid; i < n; i += nt) {
      const double wn = (z[i] * invg - a3 * w_prev[i] - a2 * w[i]) * inv;
      w_prev[i] = wn;
      x[i] = fma(xc, wn, x[i]);
    }
  }
}

// out = a - b, partial sums of out.out
__global__ void __launch_bounds__(kThreads) k_sub_norm(int n, const double* __restrict__ a, const double* __restrict__ b,
                                                       double* __restrict__ out, double* __restrict__ partial) {
  __shared__ double sh[33];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double r = a[i] - b[i];
    out[i] = r;
    acc = fma(r, r, acc);
  }
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}

// ---- coarse pressure correction  z_p += Z C Z^T r_p
// stage 1: partial[k][chunk] = sum over a chunk of row k of Z^T (CSR) of val * r[col]
__global__ void __launch_bounds__(kThreads) k_zt_partial(const int* __restrict__ zt_rowptr, const int* __restrict__ zt_cols,
                                                         const double* __restrict__ zt_vals, const double* __restrict__ r,
                                                         int nchunks, double* __restrict__ partial) {
  __shared__ double sh[33];
  const int k = blockIdx.y;
  const int s = zt_rowptr[k], e = zt_rowptr[k + 1];
  const int len = e - s;
  const int per = (len + nchunks - 1) / nchunks;
  const int c0 = s + blockIdx.x * per;
  const int c1 = min(e, c0 + per);
  double acc = 0.0;
  for (int i = c0 + threadIdx.x; i < c1; i += blockDim.x) acc = fma(zt_vals[i], r[zt_cols[i]], acc);
  const double t = block_sum(acc, sh);
  if (threadIdx.x == 0) partial[(size_t)k * nchunks + blockIdx.x] = t;
}

// stage 2 (one block): t = sum of partials per row; coef = C t
__global__ void __launch_bounds__(kThreads) k_zt_coef(int nz, int nchunks, const double* __restrict__ partial,
                                                      const double* __restrict__ Cc, double* __restrict__ coef, DistDev D) {
  extern __shared__ double t[];
  for (int k = threadIdx.x; k < nz; k += blockDim.x) {
    double s = 0.0;
    for (int c = 0; c < nchunks; ++c) s += partial[(size_t)k * nchunks + c];
    t[k] = s;
  }
  __syncthreads();
  if (D.nranks > 1) {                       // row-partitioned pressure: Z^T r is a sum over the ranks' owned dofs
    if (threadIdx.x == 0)
      for (int k0 = 0; k0 < nz; k0 += kAllreduceMaxK) dist_allreduce_scalars(D, t + k0, min(kAllreduceMaxK, nz - k0));
    __syncthreads();
  }
  for (int k = threadIdx.x; k < nz; k += blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < nz; ++j) s = fma(Cc[(size_t)k * nz + j], t[j], s);
    coef[k] = s;
  }
}

// stage 3: z[i] += w_i coef[idx_i] + (1 - w_i) coef[idx_i + 1]
__global__ void k_z_apply(int nv, const int* __restrict__ zidx, const double* __restrict__ zw,
                          const double* __restrict__ coef, double* __restrict__ z) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += gridDim.x * blockDim.x) {
    const int k = zidx[i];
    const double w = zw[i];
    z[i] += w * coef[k] + (1.0 - w) * coef[k + 1];
  }
}

constexpr int kZtChunks = 64;     // 64 x nz blocks: the 16-chunk version ran 24 us at r = 2 (11 hats x 16 = 176 CTAs for 0.94 M entries)

}  // namespace

}  // namespace sfem

using namespace sfem;

struct sfem_stokes {
  int n2 = 0, nv = 0, n = 0;           // owned velocity dofs (per component), owned pressure dofs, n = 2 n2 + nv
  // Row-partitioned (multi-GPU) layout of every Stokes vector: [u owned, interleaved | p owned | pad | u ghosts | p ghosts].
  // The owned part is contiguous (all vector kernels and dot products run over it and never see a ghost); K's local
  // column numbering has a hole of ceil(nv / 2) pairs between owned and ghost velocity dofs, B^T's ghost pressure
  // columns sit behind the velocity ghosts (the partition planner, sulcusfem/dist.py, numbers them so).  Single GPU:
  // n_alloc = n, nothing changes.
  size_t n_alloc = 0;
  size_t nv_alloc = 0;                 // length of the pressure work vectors (owned + hole + ghosts)
  Csr K, B, BT, Mp;
  sfem_mg* mg = nullptr;
  int nz = 0;
  const int* zt_rowptr = nullptr; const int* zt_cols = nullptr; const double* zt_vals = nullptr;
  const int* zidx = nullptr; const double* zw = nullptr; const double* Cc = nullptr;
  // owned work space
  double* buf = nullptr;
  double *v[2] = {nullptr, nullptr}, *z[2] = {nullptr, nullptr}, *w[2] = {nullptr, nullptr};
  double *Az = nullptr, *mp_dinv = nullptr, *mp_r = nullptr, *mp_d0 = nullptr, *mp_d1 = nullptr;
  double *part_u = nullptr, *part_p = nullptr, *part = nullptr, *zt_part = nullptr, *zcoef = nullptr, *S = nullptr;
  double* mp_coef = nullptr;
  double* h_pin = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  WorkStream ws;
  cudaStream_t side = nullptr;         // second stream: the pressure block of the preconditioner (st_precond)
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  GraphExec iter[2];
  const double* graph_x = nullptr;     // the x pointer baked into the graphs
  unsigned long long graph_epoch = 0;  // registry epoch of their capture (sfem_graph.h)
  int graph_nranks = 0;
  int npu = 0, npp = 0;
};

namespace {

int st_apply(sfem_stokes* h, const double* zin, double* y, cudaStream_t st) {
  const size_t nu = 2 * (size_t)h->n2;
  SFEM_TRY(stokes_apply_u(h->K, h->BT, zin, zin + nu, y, h->part_u, &h->npu, st));
  SFEM_TRY(spmv_dot(h->B, zin, y + nu, h->part_p, &h->npp, st, 1, zin + nu));
  return SFEM_OK;
}

// S^-1 r_p: pressure-mass Chebyshev sweep + lubrication coarse correction (7 short launches on nv unknowns)
int st_precond_pressure(sfem_stokes* h, const double* r, double* out, cudaStream_t st) {
  const size_t nu = 2 * (size_t)h->n2;
  // P1 mass matrix with Jacobi scaling has its spectrum in [1/2, 2]: 4 Chebyshev steps ~ exact solve
  SFEM_TRY(smooth(h->Mp, h->mp_dinv, h->mp_coef, 4, r + nu, out + nu, h->mp_r, h->mp_d0, h->mp_d1, true, st));
  if (h->nz > 0) {
    dim3 grid(kZtChunks, h->nz);
    k_zt_partial<<<grid, kThreads, 0, st>>>(h->zt_rowptr, h->zt_cols, h->zt_vals, r + nu, kZtChunks, h->zt_part);
    SFEM_LAUNCH_CHECK();
    k_zt_coef<<<1, kThreads, h->nz * sizeof(double), st>>>(h->nz, kZtChunks, h->zt_part, h->Cc, h->zcoef, dist_dev());
    SFEM_LAUNCH_CHECK();
    { Prof prof(PC_VEC, 28.0 * h->nv, st);
    k_z_apply<<<grid_for(h->nv, kThreads * 2), kThreads, 0, st>>>(h->nv, h->zidx, h->zw, h->zcoef, out + nu); }
    SFEM_LAUNCH_CHECK();
  }
  // experiment knob: relative scaling of the Schur block of the block-diagonal preconditioner (1 = none; measured flat,
  // profiles/r02_ncu_hot_kernels.md)
  static const double schur_scale = [] { const char* e = std::getenv("SFEM_SCHUR_SCALE"); return e ? std::atof(e) : 1.0; }();
  if (schur_scale != 1.0 && schur_scale > 0.0) SFEM_TRY(vec_axpby(h->nv, schur_scale, out + nu, 0.0, out + nu, st));
  return SFEM_OK;
}

// M^-1 r = ( V-cycle(K) r_u , S^-1 r_p ).  The two blocks touch disjoint data, so the pressure block (7 latency-bound
// launches, ~70 us at r = 2) runs on a second stream next to the velocity V-cycle; inside a capture the fork / join
// events become graph edges, i.e. the replayed iteration has two parallel branches.  Same kernels, same order inside
// each branch: results are bit-identical to the serial sequence.  (Row-partitioned: the pressure branch owns the scalar
// all-reduce and the pressure halo, the velocity branch the velocity halos and the vector all-reduce -- separate mailbox
// channels and sequence counters.)
int st_precond(sfem_stokes* h, const double* r, double* out, cudaStream_t st) {
  static const bool fork = [] { const char* e = std::getenv("SFEM_STOKES_FORK"); return !(e && e[0] == '0'); }();
  if (!fork || profiling_active() || h->side == nullptr) {
    SFEM_TRY(mg_vcycle_level(h->mg, 0, r, out, st));
    return st_precond_pressure(h, r, out, st);
  }
  SFEM_CUDA(cudaEventRecord(h->ev_fork, st));
  SFEM_CUDA(cudaStreamWaitEvent(h->side, h->ev_fork, 0));
  SFEM_TRY(st_precond_pressure(h, r, out, h->side));
  SFEM_CUDA(cudaEventRecord(h->ev_join, h->side));
  SFEM_TRY(mg_vcycle_level(h->mg, 0, r, out, st));
  SFEM_CUDA(cudaStreamWaitEvent(st, h->ev_join, 0));
  return SFEM_OK;
}

// one MINRES iteration with buffer parity q (z[q] holds the current unscaled z, v[q] the current v)
int st_iteration(sfem_stokes* h, int q, double* x, cudaStream_t st) {
  const int n = h->n;
  const int gv = grid_for(n, kThreads * 4, 4);
  double* zc = h->z[q];     double* zn = h->z[q ^ 1];
  double* vc = h->v[q];     double* vp = h->v[q ^ 1];
  double* wc = h->w[q];     double* wp = h->w[q ^ 1];
  SFEM_TRY(st_apply(h, zc, h->Az, st));
  k_sm_delta<<<1, kThreads, 0, st>>>(h->part_u, h->npu, h->part_p, h->npp, h->S, dist_dev());
  SFEM_LAUNCH_CHECK();
  { Prof prof(PC_VEC, 32.0 * n, st);
  if (aligned16(h->Az, vc, vp)) k_sm_vnext<true><<<gv, kThreads, 0, st>>>(n, h->S, h->Az, vc, vp);
  else k_sm_vnext<false><<<gv, kThreads, 0, st>>>(n, h->S, h->Az, vc, vp); }
  SFEM_LAUNCH_CHECK();
  SFEM_TRY(st_precond(h, vp, zn, st));                       // vp now holds v_{j+1}
  int np = 0;
  SFEM_TRY(vec_dot_partial(n, zn, vp, h->part, &np, st));
  k_sm_rot<<<1, kThreads, 0, st>>>(h->part, np, h->S, dist_dev());
  SFEM_LAUNCH_CHECK();
  { Prof prof(PC_VEC, 48.0 * n, st);
  if (aligned16(zc, wc, wp, x)) k_sm_wx<true><<<gv, kThreads, 0, st>>>(n, h->S, zc, wc, wp, x);
  else k_sm_wx<false><<<gv, kThreads, 0, st>>>(n, h->S, zc, wc, wp, x); }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace

extern "C" {

sfem_stokes_t sfem_stokes_create(int n2, int nv,
                                 int K_nnz, const int* K_rowptr, const int* K_cols, const double* K_vals,
                                 int B_nnz, const int* B_rowptr, const int* B_cols, const double* B_vals,
                                 const int* BT_rowptr, const int* BT_cols, const double* BT_vals,
                                 int Mp_nnz, const int* Mp_rowptr, const int* Mp_cols, const double* Mp_vals,
                                 sfem_mg_t mg,
                                 int nz, const int* zt_rowptr, const int* zt_cols, const double* zt_vals,
                                 const int* zidx, const double* zw, const double* Cc) {
  return sfem_stokes_create_part(n2, nv, K_nnz, K_rowptr, K_cols, K_vals, B_nnz, B_rowptr, B_cols, B_vals, BT_rowptr, BT_cols,
                                 BT_vals, Mp_nnz, Mp_rowptr, Mp_cols, Mp_vals, mg, nz, zt_rowptr, zt_cols, zt_vals, zidx, zw, Cc,
                                 B_nnz, 2LL * n2 + nv, nv);
}

/* Row-partitioned variant (one rank of a multi-GPU solve): n2 / nv = OWNED velocity / pressure dofs, all matrices hold
 * the owned rows with columns in the local numbering described at `struct sfem_stokes`; n_alloc = length of every
 * Stokes vector (b, x and the work vectors: owned part + hole + ghosts), nv_alloc = length of a pressure work vector. */
sfem_stokes_t sfem_stokes_create_part(int n2, int nv,
                                      int K_nnz, const int* K_rowptr, const int* K_cols, const double* K_vals,
                                      int B_nnz, const int* B_rowptr, const int* B_cols, const double* B_vals,
                                      const int* BT_rowptr, const int* BT_cols, const double* BT_vals,
                                      int Mp_nnz, const int* Mp_rowptr, const int* Mp_cols, const double* Mp_vals,
                                      sfem_mg_t mg,
                                      int nz, const int* zt_rowptr, const int* zt_cols, const double* zt_vals,
                                      const int* zidx, const double* zw, const double* Cc,
                                      int BT_nnz, long long n_alloc, long long nv_alloc) {
  if (n2 <= 0 || nv <= 0 || !mg || mg->nb != 2 || mg->levels[0].A.nrows != n2 || nz < 0 || nz > 1024 ||
      n_alloc < 2LL * n2 + nv || nv_alloc < nv) {
    set_error("sfem_stokes_create: bad arguments (the multigrid handle must be built with nb = 2 on the velocity block)");
    return nullptr;
  }
  sfem_stokes* h = new sfem_stokes();
  h->n2 = n2; h->nv = nv; h->n = 2 * n2 + nv;
  h->K.nrows = h->K.ncols = n2; h->K.nnz = K_nnz; h->K.rowptr = K_rowptr; h->K.cols = K_cols; h->K.vals = K_vals;
  h->B.nrows = nv; h->B.ncols = 2 * n2; h->B.nnz = B_nnz; h->B.rowptr = B_rowptr; h->B.cols = B_cols; h->B.vals = B_vals;
  h->BT.nrows = 2 * n2; h->BT.ncols = nv; h->BT.nnz = BT_nnz; h->BT.rowptr = BT_rowptr; h->BT.cols = BT_cols; h->BT.vals = BT_vals;
  h->Mp.nrows = h->Mp.ncols = nv; h->Mp.nnz = Mp_nnz; h->Mp.rowptr = Mp_rowptr; h->Mp.cols = Mp_cols; h->Mp.vals = Mp_vals;
  h->mg = mg;
  h->nz = nz; h->zt_rowptr = zt_rowptr; h->zt_cols = zt_cols; h->zt_vals = zt_vals; h->zidx = zidx; h->zw = zw; h->Cc = Cc;
  const size_t nn = ((size_t)n_alloc + 1) & ~(size_t)1;  // even stride: velocity parts are read as double2
  const size_t nvp = (size_t)nv_alloc;
  h->n_alloc = nn; h->nv_alloc = nvp;
  const size_t total = 7 * nn + 4 * nvp + 3 * (size_t)kMaxPartials + (size_t)(nz + 1) * (kZtChunks + 1) + kChebCoefLen + 64;
  if (cudaMalloc(&h->buf, total * sizeof(double)) != cudaSuccess || cudaMemset(h->buf, 0, total * sizeof(double)) != cudaSuccess ||
      cudaMallocHost(&h->h_pin, 8 * sizeof(double)) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev[0], cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev[1], cudaEventDisableTiming) != cudaSuccess || h->ws.init() != SFEM_OK ||
      cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
      cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming) != cudaSuccess) {
    set_error("sfem_stokes_create: allocation failed");
    sfem_stokes_destroy(h);
    return nullptr;
  }
  double* p = h->buf;
  h->v[0] = p; p += nn; h->v[1] = p; p += nn; h->z[0] = p; p += nn; h->z[1] = p; p += nn;
  h->w[0] = p; p += nn; h->w[1] = p; p += nn; h->Az = p; p += nn;
  h->mp_dinv = p; p += nvp; h->mp_r = p; p += nvp; h->mp_d0 = p; p += nvp; h->mp_d1 = p; p += nvp;
  h->part_u = p; p += kMaxPartials; h->part_p = p; p += kMaxPartials; h->part = p; p += kMaxPartials;
  h->zt_part = p; p += (size_t)(nz + 1) * kZtChunks; h->zcoef = p; p += nz + 1; h->mp_coef = p; p += kChebCoefLen; h->S = p;
  return h;
}

void sfem_stokes_destroy(sfem_stokes_t h) {
  if (!h) return;
  h->iter[0].reset(); h->iter[1].reset();
  if (h->ev[0]) cudaEventDestroy(h->ev[0]);
  if (h->ev[1]) cudaEventDestroy(h->ev[1]);
  if (h->h_pin) cudaFreeHost(h->h_pin);
  cudaFree(h->buf);
  h->ws.destroy();
  if (h->ev_fork) cudaEventDestroy(h->ev_fork);
  if (h->ev_join) cudaEventDestroy(h->ev_join);
  if (h->side) cudaStreamDestroy(h->side);
  delete h;
}

/* xref != NULL: the stopping level is anchored to the preconditioned residual norm of `xref` (the plain initial guess:
 * Dirichlet values, zero elsewhere) instead of the one of the actual starting vector x, so a better starting vector
 * saves iterations without changing the absolute residual level the solve stops at (and with it the accuracy the
 * tolerance was chosen for, profiles/r01_tolerance_study.md).  Costs one extra operator + preconditioner application. */
static int stokes_solve_impl(sfem_stokes_t h, const double* b, double* x, const double* xref, double rtol, int maxit,
                             double* h_info, void* stream) {
  if (!h || !h->mg->ready) { set_error("stokes solve: handle / multigrid not set up"); return SFEM_ERR_ARG; }
  if (dist_dev().nranks > 1 && (find_halo(h->K.rowptr) == nullptr || h->n_alloc == (((size_t)h->n + 1) & ~(size_t)1))) {
    set_error("stokes solve: a communicator is active but this handle is not row-partitioned (sfem_stokes_create_part + halos)");
    return SFEM_ERR_ARG;
  }
  if ((reinterpret_cast<uintptr_t>(b) | reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(xref)) & 15u) {
    set_error("stokes solve: b, x and xref must be 16-byte aligned (interleaved velocity is read as double2)");
    return SFEM_ERR_ARG;
  }
  cudaStream_t user = (cudaStream_t)stream;
  SFEM_TRY(sell_ensure_all(user));          // dirty sliced-ELL mirrors are re-packed before any graph replay
  SFEM_TRY(h->ws.fork(user));
  cudaStream_t st = h->ws.s;
  const int n = h->n;
  const size_t nn = h->n_alloc;              // whole vectors (ghost areas included) are cleared
  const int gv = grid_for(n, kThreads * 4, 4);
  SFEM_TRY(extract_diag_inv(h->Mp, h->mp_dinv, st));
  SFEM_TRY(cheb_setup(nullptr, nullptr, 2.0, 4.0, 4, h->part, h->mp_coef, st));
  double bb = 0.0;
  SFEM_TRY(vec_dot_host(n, b, b, h->part, &bb, st));
  const double bnorm = std::sqrt(bb);
  // reference level of the stopping test: gamma(xref) = sqrt(v^T M^-1 v), v = b - A xref
  double gamma_ref = 0.0;
  if (xref != nullptr) {
    int npr = 0;
    SFEM_TRY(st_apply(h, const_cast<double*>(xref), h->Az, st));
    k_sub_norm<<<gv, kThreads, 0, st>>>(n, b, h->Az, h->v[0], h->part);
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(st_precond(h, h->v[0], h->z[0], st));
    SFEM_TRY(vec_dot_partial(n, h->z[0], h->v[0], h->part, &npr, st));
    k_sm_init<<<1, kThreads, 0, st>>>(h->part, npr, h->S, dist_dev());
    SFEM_LAUNCH_CHECK();
    SFEM_CUDA(cudaMemcpyAsync(&gamma_ref, h->S + 13, sizeof(double), cudaMemcpyDeviceToHost, st));
    SFEM_CUDA(cudaStreamSynchronize(st));
  }
  // v = b - A x ; z = M^-1 v ; gamma_1
  SFEM_TRY(st_apply(h, x, h->Az, st));
  k_sub_norm<<<gv, kThreads, 0, st>>>(n, b, h->Az, h->v[0], h->part);
  SFEM_LAUNCH_CHECK();
  SFEM_CUDA(cudaMemsetAsync(h->v[1], 0, nn * sizeof(double), st));
  SFEM_CUDA(cudaMemsetAsync(h->w[0], 0, nn * sizeof(double), st));
  SFEM_CUDA(cudaMemsetAsync(h->w[1], 0, nn * sizeof(double), st));
  SFEM_TRY(st_precond(h, h->v[0], h->z[0], st));
  int np = 0;
  SFEM_TRY(vec_dot_partial(n, h->z[0], h->v[0], h->part, &np, st));
  k_sm_init<<<1, kThreads, 0, st>>>(h->part, np, h->S, dist_dev());
  SFEM_LAUNCH_CHECK();
  double gamma1 = 0.0;
  SFEM_CUDA(cudaMemcpyAsync(&gamma1, h->S + 13, sizeof(double), cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  int it = 0;
  double eta = gamma1;
  const bool use_graph = graphs_enabled();
  if (gamma1 > 0.0 && maxit > 0) {
    if (use_graph && (h->iter[0].exec == nullptr || h->graph_x != x || h->graph_epoch != graph_epoch() ||
                      h->graph_nranks != dist_dev().nranks)) {
      for (int q = 0; q < 2; ++q) SFEM_TRY(graph_capture(st, h->iter[q], [&]() { return st_iteration(h, q, x, st); }));
      h->graph_x = x;
      h->graph_epoch = graph_epoch();
      h->graph_nranks = dist_dev().nranks;
    }
    const double target = rtol * ((xref != nullptr && gamma_ref > 0.0) ? gamma_ref : gamma1);   // (a zero reference level: plain test)
    bool done = false;
    for (it = 1; it <= maxit && !done; ++it) {
      const int q = (it - 1) & 1;
      if (use_graph) SFEM_TRY(graph_launch(h->iter[q], st));
      else SFEM_TRY(st_iteration(h, q, x, st));
      SFEM_CUDA(cudaMemcpyAsync(h->h_pin + (it & 1), h->S + 4, sizeof(double), cudaMemcpyDeviceToHost, st));
      SFEM_CUDA(cudaEventRecord(h->ev[it & 1], st));
      if (it > 1) {                                      // poll the previous iteration (device keeps running)
        SFEM_CUDA(cudaEventSynchronize(h->ev[(it - 1) & 1]));
        eta = h->h_pin[(it - 1) & 1];
        if (!(eta == eta)) { h->ws.join(user); set_error("minres: NaN residual"); return SFEM_ERR_NOCONV; }
        if (std::fabs(eta) <= target) done = true;       // iteration `it` is already queued: keep its update
      }
    }
    --it;
    SFEM_CUDA(cudaEventSynchronize(h->ev[it & 1]));
    eta = h->h_pin[it & 1];
  }
  // true residual
  SFEM_TRY(st_apply(h, x, h->Az, st));
  k_sub_norm<<<gv, kThreads, 0, st>>>(n, b, h->Az, h->Az, h->part);
  SFEM_LAUNCH_CHECK();
  double rr = 0.0;
  {
    double* out = h->part + kMaxPartials - 1;            // last slot of the partial array is free (gv < kMaxPartials)
    // reuse the generic one-block sum
    SFEM_TRY(vec_sum_partials(h->part, gv, out, st));
    SFEM_CUDA(cudaMemcpyAsync(&rr, out, sizeof(double), cudaMemcpyDeviceToHost, st));
    SFEM_CUDA(cudaStreamSynchronize(st));
  }
  SFEM_TRY(h->ws.join(user));
  const double rel = (bnorm > 0.0) ? std::sqrt(rr) / bnorm : std::sqrt(rr);
  h_info[0] = it; h_info[1] = rel;
  const double glevel = (xref != nullptr && gamma_ref > 0.0) ? gamma_ref : gamma1;
  h_info[2] = (gamma1 == 0.0 || std::fabs(eta) <= rtol * glevel) ? 1.0 : 0.0;
  h_info[3] = (glevel > 0.0) ? std::fabs(eta) / glevel : 0.0;
  return SFEM_OK;
}

int sfem_stokes_solve(sfem_stokes_t h, const double* b, double* x, double rtol, int maxit, double* h_info, void* stream) {
  return stokes_solve_impl(h, b, x, nullptr, rtol, maxit, h_info, stream);
}

int sfem_stokes_solve_from(sfem_stokes_t h, const double* b, double* x, const double* xref, double rtol, int maxit,
                           double* h_info, void* stream) {
  if (xref == nullptr) { set_error("stokes solve_from: xref required"); return SFEM_ERR_ARG; }
  return stokes_solve_impl(h, b, x, xref, rtol, maxit, h_info, stream);
}

}  // extern "C"
