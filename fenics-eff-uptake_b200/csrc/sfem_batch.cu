// Batched multi-parameter PCG: the Robin-coefficient sweep of ONE geometry in ONE Krylov loop.
//
// The reference's mu sweep (no_advection_analysis_A.py:1306-1347; Phase B: no_advection_analysis_B.py:86-200) solves
//     A(mu_c) x_c = b(mu_c),     A(mu) = D K + mu M_Gamma,   b(mu) = b_0 + mu b_M     (Dirichlet data lifted)
// serially, one sparse LU per mu.  Only the boundary rows of A differ between the cases, and at the reference's mesh
// size (120 k P2 dofs) a single solve is launch-latency bound on a B200 (2.1 ms = 13 CG iterations x ~60 graph nodes of
// 2-3 us each).  Here nb <= 8 cases share every launch:
//   * vectors are interleaved [dof][nb]; thread t of a row kernel owns entry (row, c) = (t / nb, t % nb), so the nb
//     lanes of a row read the same matrix entry (one broadcast transaction) and gather / store nb consecutive doubles;
//   * the operator of column c is applied as (vals0[k] + mu_c valsM[k]) on the shared pattern -- two value arrays, no
//     per-case assembly;
//   * ONE multigrid hierarchy (an nb = 1 sfem_mg handle set up for a reference mu inside the batch's range) preconditions
//     all columns -- preconditioner data only, every column iterates on its exact operator to the same true residual.
//     On the SYSTEM level the smoother of column c works on that column's own operator A0 + mu_c M (its own D^-1, one
//     Gershgorin bound that holds for every mu of the batch: the row bound is convex-over-linear in mu, hence
//     quasi-convex, so its maximum over [mu_min, mu_max] sits at an end point); only the coarse levels are shared.  Smoothing with A(mu_ref) on the
//     system level as well cost 22 instead of 13 iterations for a batch spanning mu_max / mu_min = 8 (measured);
//   * every column carries its own CG scalars (alpha_c, beta_c in device memory); the loop runs until the slowest column
//     has converged (extra iterations only lower the residual of the others; exact-zero residuals are guarded).
// Single GPU only (sweeps shard by case, BASELINE config 4); row-partitioned matrices are rejected.
// Summation orders are fixed (per-block partials in block order, rows of a block in order), so results are
// bit-reproducible run to run.
#include "sfem_mg.h"
#include "sfem_graph.h"
#include "sfem_dist.h"

#include <cmath>
#include <cstdlib>
#include <vector>

namespace sfem {

namespace {

constexpr int kBatchMax = 16;           // right-hand sides per batch
constexpr int kSStride = 8;             // doubles of CG state per column: [0]=rz [1]=pq [2]=alpha [3]=beta [4]=rr
constexpr int kCoarseFallbackDegreeB = 12;
constexpr int kCoarseChebMax = 8;       // steps of the Chebyshev solve on the coarsest level

// ------------------------------------------------------------------ row kernels (thread = one (row, column) entry)
// sum_k (v0[k] + m vM[k]) x[cols[k]][c]   -- four independent gathers in flight, two accumulators
template <bool PARAM>
__device__ __forceinline__ double row_dot_b(const int* __restrict__ rowptr, const int* __restrict__ cols,
                                            const double* __restrict__ v0, const double* __restrict__ vM, double m,
                                            const double* __restrict__ x, int row, int c, int nb) {
  const int s = rowptr[row], e = rowptr[row + 1];
  double a0 = 0.0, a1 = 0.0;
  int k = s;
  for (; k + 3 < e; k += 4) {
    const int j0 = cols[k], j1 = cols[k + 1], j2 = cols[k + 2], j3 = cols[k + 3];
    double w0 = v0[k], w1 = v0[k + 1], w2 = v0[k + 2], w3 = v0[k + 3];
    if (PARAM) {
      w0 = fma(m, vM[k], w0); w1 = fma(m, vM[k + 1], w1); w2 = fma(m, vM[k + 2], w2); w3 = fma(m, vM[k + 3], w3);
    }
    const double x0 = x[(size_t)j0 * nb + c], x1 = x[(size_t)j1 * nb + c];
    const double x2 = x[(size_t)j2 * nb + c], x3 = x[(size_t)j3 * nb + c];
    a0 = fma(w0, x0, a0); a1 = fma(w1, x1, a1); a0 = fma(w2, x2, a0); a1 = fma(w3, x3, a1);
  }
  for (; k < e; ++k) {
    double w = v0[k];
    if (PARAM) w = fma(m, vM[k], w);
    a0 = fma(w, x[(size_t)cols[k] * nb + c], a0);
  }
  return a0 + a1;
}

#define SFEM_BATCH_ENTRY_LOOP(total)                                                                           \
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < (total);                            \
       t += (long long)gridDim.x * blockDim.x)

// MODE 0: y = A x;  1: y = b - A x;  2: y += A x      (PARAM: A = A0 + mu_c M, column-dependent)
template <int MODE, bool PARAM>
__global__ void __launch_bounds__(kThreads) kb_spmv(int nrows, int nb, const int* __restrict__ rowptr,
                                                    const int* __restrict__ cols, const double* __restrict__ v0,
                                                    const double* __restrict__ vM, const double* __restrict__ mu,
                                                    const double* __restrict__ x, const double* __restrict__ b,
                                                    double* __restrict__ y) {
  const long long total = (long long)nrows * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int row = (int)(t / nb), c = (int)(t - (long long)row * nb);
    const double m = PARAM ? mu[c] : 0.0;
    const double s = row_dot_b<PARAM>(rowptr, cols, v0, vM, m, x, row, c, nb);
    y[t] = (MODE == 0) ? s : (MODE == 1 ? b[t] - s : y[t] + s);
  }
}

// start of a smoothing sweep from x = 0:  d_0 = D^-1 b / theta  (FULL: also r = b, x = d_0 -- a one-step sweep)
// (PC: dinv is per (row, column) -- the system level of a batch; otherwise per row)
template <bool FULL, bool PC>
__global__ void __launch_bounds__(kThreads) kb_cheb_init0(int nrows, int nb, const double* __restrict__ dinv,
                                                          const double* __restrict__ b, double* __restrict__ r,
                                                          double* __restrict__ d, double* __restrict__ x,
                                                          const double* __restrict__ coef) {
  const double c0 = coef[1];
  const long long total = (long long)nrows * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int row = (int)(t / nb);
    const double bi = b[t];
    const double di = c0 * dinv[PC ? t : (long long)row] * bi;
    d[t] = di;
    if (FULL) { r[t] = bi; x[t] = di; }
  }
}

// one fused Chebyshev-Jacobi step (same update as EpiCheb, sfem_spmv_epi.cuh)
template <bool PC>
__global__ void __launch_bounds__(kThreads) kb_cheb_step(int nrows, int nb, const int* __restrict__ rowptr,
                                                         const int* __restrict__ cols, const double* __restrict__ vals,
                                                         const double* __restrict__ valsM, const double* __restrict__ mu,
                                                         const double* __restrict__ dinv,
                                                         const double* __restrict__ d_old, double* __restrict__ d_new,
                                                         double* __restrict__ r, double* __restrict__ xx,
                                                         const double* __restrict__ c12, int last,
                                                         const double* __restrict__ b0) {
  const double c1 = c12[0], c2 = c12[1];
  const long long total = (long long)nrows * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int row = (int)(t / nb), c = (int)(t - (long long)row * nb);
    const double s = row_dot_b<PC>(rowptr, cols, vals, valsM, PC ? mu[c] : 0.0, d_old, row, c, nb);
    const double dd = d_old[t];
    double rin, xin = 0.0;
    if (b0 != nullptr) rin = b0[t];
    else { rin = r[t]; xin = xx[t]; }
    const double rn = rin - s;
    const double dn = c1 * dd + c2 * dinv[PC ? t : (long long)row] * rn;
    r[t] = rn;
    d_new[t] = dn;
    xx[t] = xin + (last ? (dd + dn) : dd);
  }
}

// r = b - A x ; d = c0 D^-1 r
template <bool PC>
__global__ void __launch_bounds__(kThreads) kb_resid_d0(int nrows, int nb, const int* __restrict__ rowptr,
                                                        const int* __restrict__ cols, const double* __restrict__ vals,
                                                        const double* __restrict__ valsM, const double* __restrict__ mu,
                                                        const double* __restrict__ dinv, const double* __restrict__ b,
                                                        const double* __restrict__ x, double* __restrict__ r,
                                                        double* __restrict__ d, const double* __restrict__ c0p) {
  const double c0 = c0p[0];
  const long long total = (long long)nrows * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int row = (int)(t / nb), c = (int)(t - (long long)row * nb);
    const double s = row_dot_b<PC>(rowptr, cols, vals, valsM, PC ? mu[c] : 0.0, x, row, c, nb);
    const double rr = b[t] - s;
    r[t] = rr;
    d[t] = c0 * dinv[PC ? t : (long long)row] * rr;
  }
}

// x = M b for the dense inverse of the coarsest operator: one warp per row, all columns at once; four matrix entries
// (and their 4 x nb right-hand-side values) are requested per round so the short row is not a chain of L2 round trips
__global__ void __launch_bounds__(kThreads) kb_dense_gemv(int n, int nb, const double* __restrict__ M,
                                                          const double* __restrict__ b, double* __restrict__ x) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n; row += gridDim.x * warps_per_block) {
    const double* m = M + (size_t)row * n;
    double acc[kBatchMax];
#pragma unroll
    for (int c = 0; c < kBatchMax; ++c) acc[c] = 0.0;
    for (int j0 = lane; j0 < n; j0 += 128) {
      double mj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) mj[u] = (j0 + 32 * u < n) ? m[j0 + 32 * u] : 0.0;
#pragma unroll
      for (int c = 0; c < kBatchMax; ++c) {
        if (c < nb) {
          double bv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) bv[u] = (j0 + 32 * u < n) ? b[(size_t)(j0 + 32 * u) * nb + c] : 0.0;
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[c] = fma(mj[u], bv[u], acc[c]);
        }
      }
    }
    double mine = 0.0;
#pragma unroll
    for (int c = 0; c < kBatchMax; ++c) {
      const double sum = warp_sum(acc[c]);
      if (c == lane) mine = sum;
    }
    if (lane < nb) x[(size_t)row * nb + lane] = mine;
  }
}

// Coarsest level with per-column operators: the dense inverse belongs to the REFERENCE operator A(mu_ref); column c
// solves A(mu_c) x = b with a Chebyshev iteration preconditioned by it.  The eigenvalues of A(mu_ref)^-1 A(mu_c) lie
// in [min(1, mu_c / mu_ref), max(1, mu_c / mu_ref)] (A(mu) = A0 + mu M is monotone in mu in the Loewner order), so the
// window is known exactly and a few steps reduce the error by 20x -- a fixed SPD polynomial per column.
// One step = this kernel: z = M rin, then  first: d = k0_c z, x = d;   else: d = a_c d + b_c z, x += d
// (coefficients k: [k0] or [a | b], nb each).
__global__ void __launch_bounds__(kThreads) kb_dense_gemv_cheb(int n, int nb, const double* __restrict__ M,
                                                               const double* __restrict__ rin, double* __restrict__ d,
                                                               double* __restrict__ x, const double* __restrict__ k,
                                                               int first) {
  const int warps_per_block = blockDim.x >> 5;
  const int lane = threadIdx.x & 31;
  for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < n; row += gridDim.x * warps_per_block) {
    const double* m = M + (size_t)row * n;
    double acc[kBatchMax];
#pragma unroll
    for (int c = 0; c < kBatchMax; ++c) acc[c] = 0.0;
    for (int j0 = lane; j0 < n; j0 += 128) {
      double mj[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) mj[u] = (j0 + 32 * u < n) ? m[j0 + 32 * u] : 0.0;
#pragma unroll
      for (int c = 0; c < kBatchMax; ++c) {
        if (c < nb) {
          double bv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) bv[u] = (j0 + 32 * u < n) ? rin[(size_t)(j0 + 32 * u) * nb + c] : 0.0;
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[c] = fma(mj[u], bv[u], acc[c]);
        }
      }
    }
    double z = 0.0;
#pragma unroll
    for (int c = 0; c < kBatchMax; ++c) {
      const double sum = warp_sum(acc[c]);
      if (c == lane) z = sum;
    }
    if (lane < nb) {
      const size_t t = (size_t)row * nb + lane;
      if (first) {
        const double dn = k[lane] * z;
        d[t] = dn;
        x[t] = dn;
      } else {
        const double dn = fma(k[lane], d[t], k[nb + lane] * z);
        d[t] = dn;
        x[t] += dn;
      }
    }
  }
}

// system level of a batch: D^-1 of every column's own operator, dinv[row][c] = 1 / (A0 + mu_c M)_row,row
__global__ void __launch_bounds__(kThreads) kb_diag_inv(int nrows, int nb, const int* __restrict__ rowptr,
                                                        const int* __restrict__ cols, const double* __restrict__ v0,
                                                        const double* __restrict__ vM, const double* __restrict__ mu,
                                                        double* __restrict__ dinv) {
  const long long total = (long long)nrows * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int row = (int)(t / nb), c = (int)(t - (long long)row * nb);
    double d = 0.0;
    for (int k = rowptr[row]; k < rowptr[row + 1]; ++k)
      if (cols[k] == row) d = fma(mu[c], vM[k], v0[k]);
    dinv[t] = (d != 0.0) ? 1.0 / d : 1.0;
  }
}

// per-block max over the rows of the Gershgorin bound of D^-1 (A0 + mu M) at mu = mu_lo and mu = mu_hi: the row bound
// (sum_j |a_ij + mu m_ij|) / (a_ii + mu m_ii) has a convex numerator and a positive linear denominator, i.e. it is
// quasi-convex in mu, so on [mu_lo, mu_hi] it is largest at an end point
__global__ void __launch_bounds__(kThreads) kb_gershgorin(int n, const int* __restrict__ rowptr, const int* __restrict__ cols,
                                                          const double* __restrict__ v0, const double* __restrict__ vM,
                                                          double mu_lo, double mu_hi, double* __restrict__ partial) {
  __shared__ double sh[32];
  double mx = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s_lo = 0.0, s_hi = 0.0, d_lo = 0.0, d_hi = 0.0;
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) {
      const double a = v0[k], m = vM[k];
      const double lo = fma(mu_lo, m, a), hi = fma(mu_hi, m, a);
      s_lo += fabs(lo); s_hi += fabs(hi);
      if (cols[k] == i) { d_lo = lo; d_hi = hi; }
    }
    if (d_lo != 0.0) mx = fmax(mx, s_lo / fabs(d_lo));
    if (d_hi != 0.0) mx = fmax(mx, s_hi / fabs(d_hi));
  }
  for (int o = 16; o > 0; o >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) mx = fmax(mx, sh[w]);
    partial[blockIdx.x] = mx;
  }
}

// Chebyshev coefficients for the window [lmax / ratio, lmax] (layout and recurrences of k_cheb_coef, sfem_mg.cu)
__global__ void kb_cheb_coef(const double* __restrict__ partial, int np, double ratio, int degree, double* __restrict__ coef) {
  if (blockIdx.x != 0 || threadIdx.x != 0) return;
  double lmax = 0.0;
  for (int i = 0; i < np; ++i) lmax = fmax(lmax, partial[i]);
  if (!(lmax > 0.0)) lmax = 2.0;
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

// out[row][c] = a[row] + mu_c m[row]   (m may be NULL): broadcast of the shared initial guess / right-hand side
__global__ void __launch_bounds__(kThreads) kb_expand(int nrows, int nb, const double* __restrict__ a,
                                                      const double* __restrict__ m, const double* __restrict__ mu,
                                                      double* __restrict__ out) {
  const long long total = (long long)nrows * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int row = (int)(t / nb), c = (int)(t - (long long)row * nb);
    out[t] = (m != nullptr) ? fma(mu[c], m[row], a[row]) : a[row];
  }
}

// out[row] = X[row][c]
__global__ void __launch_bounds__(kThreads) kb_column(int nrows, int nb, int c, const double* __restrict__ X,
                                                      double* __restrict__ out) {
  for (int row = blockIdx.x * blockDim.x + threadIdx.x; row < nrows; row += gridDim.x * blockDim.x)
    out[row] = X[(size_t)row * nb + c];
}

// ------------------------------------------------------------------ per-column reductions
// A block holds rpb = blockDim / nb consecutive rows (blockDim is a multiple of nb): thread -> (row lr, column c),
// consecutive threads touch consecutive memory.  The rpb partial sums of a column are added in row order by one thread.
__device__ __forceinline__ void block_column_sums(double acc, int nb, double* sh, double* __restrict__ partial) {
  sh[threadIdx.x] = acc;
  __syncthreads();
  if ((int)threadIdx.x < nb) {
    const int rpb = blockDim.x / nb;
    double s = 0.0;
    for (int j = 0; j < rpb; ++j) s += sh[j * nb + threadIdx.x];
    partial[(size_t)blockIdx.x * nb + threadIdx.x] = s;
  }
}

// (four rows per thread and round: 8 / 16 independent loads in flight instead of a chain of L2 round trips)
__global__ void __launch_bounds__(kThreads) kb_dot(int n, int nb, const double* __restrict__ x,
                                                   const double* __restrict__ y, double* __restrict__ partial) {
  __shared__ double sh[kThreads];
  const int rpb = blockDim.x / nb;
  const int c = threadIdx.x % nb, lr = threadIdx.x / nb;
  const long long stride = (long long)gridDim.x * rpb;
  double acc = 0.0;
  for (long long row = (long long)blockIdx.x * rpb + lr; row < n; row += 4 * stride) {
    double xv[4], yv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ru = row + u * stride;
      const bool ok = ru < n;
      const size_t t = (size_t)(ok ? ru : row) * nb + c;
      xv[u] = ok ? x[t] : 0.0;
      yv[u] = ok ? y[t] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) acc = fma(xv[u], yv[u], acc);
  }
  block_column_sums(acc, nb, sh, partial);
}

// x_c += alpha_c p_c ; r_c -= alpha_c q_c ; partial sums of r_c . r_c
__global__ void __launch_bounds__(kThreads) kb_cg_update(int n, int nb, const double* __restrict__ S,
                                                         const double* __restrict__ p, const double* __restrict__ q,
                                                         double* __restrict__ x, double* __restrict__ r,
                                                         double* __restrict__ partial) {
  __shared__ double sh[kThreads];
  const int rpb = blockDim.x / nb;
  const int c = threadIdx.x % nb, lr = threadIdx.x / nb;
  const long long stride = (long long)gridDim.x * rpb;
  const double alpha = S[c * kSStride + 2];
  double acc = 0.0;
  for (long long row = (long long)blockIdx.x * rpb + lr; row < n; row += 4 * stride) {
    double pv[4], qv[4], xv[4], rv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ru = row + u * stride;
      if (ru < n) {
        const size_t t = (size_t)ru * nb + c;
        pv[u] = p[t]; qv[u] = q[t]; xv[u] = x[t]; rv[u] = r[t];
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long ru = row + u * stride;
      if (ru < n) {
        const size_t t = (size_t)ru * nb + c;
        x[t] = fma(alpha, pv[u], xv[u]);
        const double ri = fma(-alpha, qv[u], rv[u]);
        r[t] = ri;
        acc = fma(ri, ri, acc);
      }
    }
  }
  block_column_sums(acc, nb, sh, partial);
}

__global__ void __launch_bounds__(kThreads) kb_cg_p(int n, int nb, const double* __restrict__ S,
                                                    const double* __restrict__ z, double* __restrict__ p) {
  const long long total = (long long)n * nb;
  SFEM_BATCH_ENTRY_LOOP(total) {
    const int c = (int)(t % nb);
    p[t] = fma(S[c * kSStride + 3], p[t], z[t]);
  }
}

// ------------------------------------------------------------------ one-block scalar kernels: warp c owns column c
__device__ __forceinline__ double column_total(const double* __restrict__ partial, int np, int nb, int c, int lane) {
  double s0 = 0.0, s1 = 0.0;
  int j = lane;
  for (; j + 96 < np; j += 128) {                        // four independent loads per round
    const double a0 = partial[(size_t)j * nb + c], a1 = partial[(size_t)(j + 32) * nb + c];
    const double a2 = partial[(size_t)(j + 64) * nb + c], a3 = partial[(size_t)(j + 96) * nb + c];
    s0 += a0; s1 += a1; s0 += a2; s1 += a3;
  }
  for (; j < np; j += 32) s0 += partial[(size_t)j * nb + c];
  return warp_sum(s0 + s1);
}

__global__ void kb_store_sums(const double* __restrict__ partial, int np, int nb, double* __restrict__ out) {
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (c >= nb) return;
  const double s = column_total(partial, np, nb, c, lane);
  if (lane == 0) out[c] = s;
}

__global__ void kb_cg_alpha(const double* __restrict__ partial, int np, int nb, double* __restrict__ S) {
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (c >= nb) return;
  const double pq = column_total(partial, np, nb, c, lane);
  if (lane == 0) {
    S[c * kSStride + 1] = pq;
    S[c * kSStride + 2] = (pq > 0.0) ? S[c * kSStride + 0] / pq : 0.0;      // a column that reached r = 0 stays put
  }
}

// beta_c, rz_c, rr_c; S[nb * kSStride] = max_c rr_c / bb_c  (NaN if any column is NaN): the one double the host polls
__global__ void kb_cg_beta(const double* __restrict__ p_rz, int n_rz, const double* __restrict__ p_rr, int n_rr, int nb,
                           double* __restrict__ S, int first, const double* __restrict__ BB) {
  const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (c < nb) {
    const double rz = column_total(p_rz, n_rz, nb, c, lane);
    const double rr = column_total(p_rr, n_rr, nb, c, lane);
    if (lane == 0) {
      const double old = S[c * kSStride + 0];
      S[c * kSStride + 3] = (first || !(old > 0.0)) ? 0.0 : rz / old;
      S[c * kSStride + 0] = rz;
      S[c * kSStride + 4] = rr;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    bool bad = false;
    for (int k = 0; k < nb; ++k) {
      const double bb = BB[k];
      const double v = S[k * kSStride + 4] / (bb > 0.0 ? bb : 1.0);
      if (!(v == v)) bad = true;
      else if (v > m) m = v;
    }
    S[nb * kSStride] = bad ? nan("") : m;
  }
}

// ------------------------------------------------------------------ host side
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

// lagged polling of the residual estimate (same scheme as sfem_krylov.cu)
struct Poller {
  double* pin = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};
  int init() {
    if (pin) return SFEM_OK;
    SFEM_CUDA(cudaMallocHost(&pin, (4 + 3 * kBatchMax) * sizeof(double)));
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

thread_local Workspace t_bws;
thread_local Poller t_bpoll;
thread_local WorkStream t_bwork;
thread_local GraphCache t_batch_graph;

struct LevelVecs { double* x = nullptr; double* b = nullptr; double* r = nullptr; double* d0 = nullptr; double* d1 = nullptr; };

inline int entry_grid(long long total) { return grid_for(total, kThreads, 8); }

// the system level of a batch: column c smooths with its own operator A0 + mu_c M
struct FineOp {
  const double* valsM = nullptr;     // second value array on the pattern of the level's matrix (NULL: plain level)
  const double* mu = nullptr;        // device [nb]
  const double* dinv = nullptr;      // device [rows][nb]
  const double* coef = nullptr;      // Chebyshev coefficients valid for every mu of the batch
};

template <int MODE>
int spmv_b(const Csr& A, const double* x, const double* b, double* y, int nb, cudaStream_t st, const FineOp* F = nullptr) {
  if (A.nrows <= 0) return SFEM_OK;
  const long long total = (long long)A.nrows * nb;
  Prof prof(PC_SPMV, (F ? 20.0 : 12.0) * A.nnz + 4.0 * A.nrows + 8.0 * nb * ((double)A.ncols + (double)A.nrows * (MODE == 0 ? 1 : 2)), st);
  if (F) kb_spmv<MODE, true><<<entry_grid(total), kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, F->valsM, F->mu, x, b, y);
  else kb_spmv<MODE, false><<<entry_grid(total), kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, nullptr, nullptr, x, b, y);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

// Chebyshev-Jacobi sweep (smooth() of sfem_mg.cu) on nb columns; F != NULL: per-column operator, D^-1 and coefficients
int smooth_b(const Csr& A, const double* dinv, const double* coef, int degree, const double* b, double* x, double* r,
             double* d0, double* d1, bool zero_init, int nb, cudaStream_t st, const FineOp* F = nullptr) {
  const long long total = (long long)A.nrows * nb;
  const int g = entry_grid(total);
  if (F) { dinv = F->dinv; coef = F->coef; }
  if (zero_init) {
    if (degree <= 1) {
      if (F) kb_cheb_init0<true, true><<<g, kThreads, 0, st>>>(A.nrows, nb, dinv, b, r, d0, x, coef);
      else kb_cheb_init0<true, false><<<g, kThreads, 0, st>>>(A.nrows, nb, dinv, b, r, d0, x, coef);
    } else {
      if (F) kb_cheb_init0<false, true><<<g, kThreads, 0, st>>>(A.nrows, nb, dinv, b, r, d0, x, coef);
      else kb_cheb_init0<false, false><<<g, kThreads, 0, st>>>(A.nrows, nb, dinv, b, r, d0, x, coef);
    }
    SFEM_LAUNCH_CHECK();
    if (degree <= 1) return SFEM_OK;
  } else {
    if (F) kb_resid_d0<true><<<g, kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, F->valsM, F->mu, dinv, b, x, r, d0, coef + 1);
    else kb_resid_d0<false><<<g, kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, nullptr, nullptr, dinv, b, x, r, d0, coef + 1);
    SFEM_LAUNCH_CHECK();
    if (degree <= 1) return vec_axpby((int)total, 1.0, d0, 1.0, x, st);
  }
  double* dold = d0;
  double* dnew = d1;
  for (int i = 0; i < degree - 1; ++i) {
    Prof prof(PC_CHEB, (F ? 20.0 : 12.0) * A.nnz + 12.0 * A.nrows + 48.0 * nb * A.nrows, st);
    const int last = i == degree - 2 ? 1 : 0;
    const double* b0 = (zero_init && i == 0) ? b : nullptr;
    if (F) kb_cheb_step<true><<<g, kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, F->valsM, F->mu, dinv, dold, dnew, r, x,
                                                      coef + 2 + 2 * i, last, b0);
    else kb_cheb_step<false><<<g, kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, nullptr, nullptr, dinv, dold, dnew, r, x,
                                                     coef + 2 + 2 * i, last, b0);
    SFEM_LAUNCH_CHECK();
    double* t = dold; dold = dnew; dnew = t;
  }
  return SFEM_OK;
}

// Chebyshev solve of the coarsest level (kb_dense_gemv_cheb): m steps, coefficients on the device
struct CoarseCheb {
  int m = 0;                         // 0: plain product with the dense inverse (shared operator)
  const double* coef = nullptr;      // device [(2 m - 1) * nb]: k0 | (a_i | b_i), i = 2..m
};

// the V-cycle of sfem_mg.cu (mg_vcycle_level) on nb interleaved columns; level vectors come from the batch work space.
// AL[l] / FL[l]: matrix of level l with the value array vals0_l and the per-column data (valsM_l, mu, D^-1,
// coefficients); FL[l].valsM == NULL: the level uses the handle's operator (assembled for the reference mu) for
// every column.  Structure, transfers and the dense coarsest inverse are the handle's.
int vcycle_b(sfem_mg* mg, const std::vector<LevelVecs>& V, int l, const double* b, double* x, int nb, cudaStream_t st,
             const std::vector<Csr>& AL, const std::vector<FineOp>& FL, const CoarseCheb& CC) {
  MgLevel& L = mg->levels[l];
  const int last = (int)mg->levels.size() - 1;
  const LevelVecs& W = V[l];
  const bool own = FL[l].valsM != nullptr;
  const Csr& A = own ? AL[l] : L.A;
  const FineOp* F = own ? &FL[l] : nullptr;
  if (l == last) {
    if (mg->coarse_inv != nullptr) {
      const int g = grid_for(L.A.nrows, kThreads / 32);
      if (!own || CC.m <= 0) {
        kb_dense_gemv<<<g, kThreads, 0, st>>>(L.A.nrows, nb, mg->coarse_inv, b, x);
        SFEM_LAUNCH_CHECK();
        return SFEM_OK;
      }
      kb_dense_gemv_cheb<<<g, kThreads, 0, st>>>(L.A.nrows, nb, mg->coarse_inv, b, W.d0, x, CC.coef, 1);
      SFEM_LAUNCH_CHECK();
      for (int i = 1; i < CC.m; ++i) {
        SFEM_TRY(spmv_b<1>(A, x, b, W.r, nb, st, F));
        kb_dense_gemv_cheb<<<g, kThreads, 0, st>>>(L.A.nrows, nb, mg->coarse_inv, W.r, W.d0, x,
                                                   CC.coef + (size_t)nb * (1 + 2 * (i - 1)), 0);
        SFEM_LAUNCH_CHECK();
      }
      return SFEM_OK;
    }
    if (own) {
      set_error("cg_batch: per-column coarse operators need the dense coarsest inverse");
      return SFEM_ERR_ARG;
    }
    return smooth_b(L.A, L.dinv, L.coef, kCoarseFallbackDegreeB, b, x, W.r, W.d0, W.d1, true, nb, st);
  }
  SFEM_TRY(smooth_b(A, L.dinv, L.coef, mg->degree, b, x, W.r, W.d0, W.d1, true, nb, st, F));
  SFEM_TRY(spmv_b<1>(A, x, b, W.r, nb, st, F));
  const LevelVecs& C = V[l + 1];
  SFEM_TRY(spmv_b<0>(L.R, W.r, nullptr, C.b, nb, st));
  SFEM_TRY(vcycle_b(mg, V, l + 1, C.b, C.x, nb, st, AL, FL, CC));
  SFEM_TRY(spmv_b<2>(L.P, C.x, nullptr, x, nb, st));
  return smooth_b(A, L.dinv, L.coef, mg->degree, b, x, W.r, W.d0, W.d1, false, nb, st, F);
}

}  // namespace

}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_krylov_cg_batch(int n, int nnz, const int* rowptr, const int* cols, int nlevels, const double* const* lvl_vals0,
                         const double* const* lvl_valsM, int nb, const double* h_mu, double mu_ref, sfem_mg_t mg,
                         const double* b0, const double* bM, const double* x0, double* X, double rtol, int maxit,
                         double* h_info, void* stream) {
  cudaStream_t user = (cudaStream_t)stream;
  if (n <= 0 || nb < 1 || nb > kBatchMax || !rowptr || !cols || !lvl_vals0 || !lvl_valsM || nlevels < 1 || !lvl_vals0[0] ||
      !lvl_valsM[0] || !h_mu || !b0 || !bM || !x0 || !X || !h_info) {
    set_error("cg_batch: bad arguments (1 <= nb <= 16, all pointers required)");
    return SFEM_ERR_ARG;
  }
  if (!mg || !mg->ready || mg->nb != 1 || mg->tail != nullptr || mg->levels.empty() || mg->levels[0].A.nrows != n ||
      (int)mg->levels.size() != nlevels) {
    set_error("cg_batch: needs a set-up single-GPU multigrid handle (nb = 1) of the same system size and depth");
    return SFEM_ERR_ARG;
  }
  if (find_halo(rowptr) != nullptr || dist_dev().nranks > 1) {
    set_error("cg_batch: row-partitioned operators are not supported (sweeps shard by case)");
    return SFEM_ERR_ARG;
  }
  const double* vals0 = lvl_vals0[0];
  const double* valsM = lvl_valsM[0];
  double mu_lo = h_mu[0], mu_hi = h_mu[0];
  for (int c = 1; c < nb; ++c) { mu_lo = std::fmin(mu_lo, h_mu[c]); mu_hi = std::fmax(mu_hi, h_mu[c]); }
  if (!(mu_lo >= 0.0) || !(mu_ref >= 0.0) || (mu_hi > 0.0 && !(mu_ref > 0.0))) {
    set_error("cg_batch: coefficients must be >= 0 and the reference coefficient > 0 unless all are 0");
    return SFEM_ERR_ARG;
  }
  const size_t nn = (size_t)n * nb;
  const int nl = (int)mg->levels.size();
  size_t lv = 0;                                         // level vectors + per-column D^-1 + coefficients of every level
  for (int l = 0; l < nl; ++l) lv += (size_t)mg->levels[l].A.nrows * nb * (l == 0 ? 4 : 6) + kChebCoefLen;
  const int rpb = kThreads / nb;                         // rows per block of the reducing kernels
  const int bd = rpb * nb;
  int gd = (int)(((long long)n + rpb - 1) / rpb);
  if (gd > 2 * num_sms()) gd = 2 * num_sms();
  if (gd < 1) gd = 1;
  const size_t small = 2 * (size_t)gd * nb + (size_t)nb * kSStride + 8 + 3 * (size_t)kBatchMax + kMaxPartials +
                       2 * (size_t)kCoarseChebMax * kBatchMax + 64;
  SFEM_TRY(t_bws.ensure(5 * nn + lv + small));
  SFEM_TRY(t_bpoll.init());
  bool forked = false;
  cudaStream_t st = user;
  if (user == nullptr || user == cudaStreamLegacy || user == cudaStreamPerThread) {
    if (t_bwork.fork(user) == SFEM_OK) { forked = true; st = t_bwork.s; }
  }
  double* w = t_bws.ptr;
  double* r = w; w += nn;
  double* z = w; w += nn;
  double* p = w; w += nn;
  double* q = w; w += nn;
  double* B = w; w += nn;
  std::vector<LevelVecs> V(nl);
  std::vector<double*> dinvL(nl), coefL(nl);
  for (int l = 0; l < nl; ++l) {
    const size_t m = (size_t)mg->levels[l].A.nrows * nb;
    V[l].r = w; w += m; V[l].d0 = w; w += m; V[l].d1 = w; w += m;
    if (l > 0) { V[l].x = w; w += m; V[l].b = w; w += m; }
    dinvL[l] = w; w += m;
    coefL[l] = w; w += kChebCoefLen;
  }
  double* part0 = w; w += (size_t)gd * nb;
  double* part1 = w; w += (size_t)gd * nb;
  double* S = w; w += (size_t)nb * kSStride + 8;
  double* mu = w; w += kBatchMax;
  double* BB = w; w += kBatchMax;
  double* RR = w; w += kBatchMax;
  double* gpart = w; w += kMaxPartials;
  double* ccoef = w; w += 2 * (size_t)kCoarseChebMax * kBatchMax;
  const int ge = entry_grid((long long)nn);
  const int sb = 32 * nb;                                // one warp per column in the scalar kernels
  auto apply = [&](const double* xin, const double* bin, double* yout, int mode) -> int {
    Prof prof(PC_SPMV, 20.0 * nnz + 4.0 * n + 8.0 * nb * (double)n * (mode == 0 ? 2 : 3), st);
    if (mode == 0) kb_spmv<0, true><<<ge, kThreads, 0, st>>>(n, nb, rowptr, cols, vals0, valsM, mu, xin, bin, yout);
    else kb_spmv<1, true><<<ge, kThreads, 0, st>>>(n, nb, rowptr, cols, vals0, valsM, mu, xin, bin, yout);
    SFEM_LAUNCH_CHECK();
    return SFEM_OK;
  };
  auto dot = [&](const double* a, const double* b, double* part) -> int {
    Prof prof(PC_VEC, 16.0 * nn, st);
    kb_dot<<<gd, bd, 0, st>>>(n, nb, a, b, part);
    SFEM_LAUNCH_CHECK();
    return SFEM_OK;
  };
  SFEM_CUDA(cudaMemcpyAsync(mu, h_mu, nb * sizeof(double), cudaMemcpyHostToDevice, st));
  // every level given with two value arrays: column c smooths with its own operator A0_l + mu_c M_l (own D^-1; one
  // eigenvalue bound per level that holds for the whole batch)
  std::vector<Csr> AL(nl);
  std::vector<FineOp> FL(nl);
  for (int l = 0; l < nl; ++l) {
    AL[l] = mg->levels[l].A;
    if (lvl_vals0[l] == nullptr || lvl_valsM[l] == nullptr) continue;
    AL[l].vals = lvl_vals0[l];
    FL[l].valsM = lvl_valsM[l]; FL[l].mu = mu; FL[l].dinv = dinvL[l]; FL[l].coef = coefL[l];
    if (l == nl - 1 && mg->coarse_inv != nullptr) continue;            // the coarsest level is not smoothed
    const Csr& A = AL[l];
    kb_diag_inv<<<entry_grid((long long)A.nrows * nb), kThreads, 0, st>>>(A.nrows, nb, A.rowptr, A.cols, A.vals, FL[l].valsM, mu, dinvL[l]);
    SFEM_LAUNCH_CHECK();
    const int gg = grid_for(A.nrows, kThreads, 4);
    kb_gershgorin<<<gg, kThreads, 0, st>>>(A.nrows, A.rowptr, A.cols, A.vals, FL[l].valsM, mu_lo, mu_hi, gpart);
    SFEM_LAUNCH_CHECK();
    kb_cheb_coef<<<1, 32, 0, st>>>(gpart, gg, mg->ratio, mg->degree, coefL[l]);
    SFEM_LAUNCH_CHECK();
  }
  // coarsest level: Chebyshev iteration preconditioned by the dense inverse of A(mu_ref); window of column c =
  // [min(1, mu_c / mu_ref), max(1, mu_c / mu_ref)] with 1 % slack; as many steps as the widest window needs for a
  // 20-fold error reduction (0.02 / 0.05 / 0.1 give the same 13 iterations; 0.05 saves one product per cycle, measured)
  CoarseCheb CC;
  if (FL[nl - 1].valsM != nullptr && mg->coarse_inv != nullptr) {
    constexpr double eps = 0.01;
    double lo[kBatchMax], hi[kBatchMax], kmax = 1.0;
    for (int c = 0; c < nb; ++c) {
      double ratio = 1.0;
      if (mu_ref > 0.0) ratio = h_mu[c] > 0.0 ? h_mu[c] / mu_ref : 0.1;       // mu_c = 0: spectrum in (0, 1], see below
      lo[c] = std::fmin(1.0, ratio) * (1.0 - eps);
      hi[c] = std::fmax(1.0, ratio) * (1.0 + eps);
      kmax = std::fmax(kmax, hi[c] / lo[c]);
    }
    // (an eigenvalue BELOW the window keeps the polynomial preconditioner positive -- the residual polynomial decreases
    //  monotonically from 1 at 0 to the window -- so the guessed lower end for mu_c = 0 is safe; the upper end is exact)
    const double sq = std::sqrt(kmax), rr = (sq - 1.0) / (sq + 1.0);
    static const double tol = []() {                       // SFEM_BATCH_COARSE_TOL: experiments with the coarse accuracy
      const char* e = std::getenv("SFEM_BATCH_COARSE_TOL");
      const double v = e ? std::atof(e) : 0.05;
      return (v > 0.0 && v < 1.0) ? v : 0.05;
    }();
    int m = 1;
    while (m < kCoarseChebMax && 2.0 * std::pow(rr, m) / (1.0 + std::pow(rr, 2 * m)) > tol) ++m;
    std::vector<double> hc((size_t)(2 * m - 1) * nb);
    for (int c = 0; c < nb; ++c) {
      const double theta = 0.5 * (hi[c] + lo[c]), delta = 0.5 * (hi[c] - lo[c]), sigma = theta / delta;
      double rho = 1.0 / sigma;
      hc[c] = 1.0 / theta;
      for (int i = 1; i < m; ++i) {
        const double rho_new = 1.0 / (2.0 * sigma - rho);
        hc[(size_t)nb * (1 + 2 * (i - 1)) + c] = rho_new * rho;
        hc[(size_t)nb * (2 + 2 * (i - 1)) + c] = 2.0 * rho_new / delta;
        rho = rho_new;
      }
    }
    SFEM_CUDA(cudaMemcpyAsync(ccoef, hc.data(), hc.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    SFEM_CUDA(cudaStreamSynchronize(st));                  // hc leaves scope
    CC.m = m;
    CC.coef = ccoef;
  }
  kb_expand<<<ge, kThreads, 0, st>>>(n, nb, x0, nullptr, mu, X);
  SFEM_LAUNCH_CHECK();
  kb_expand<<<ge, kThreads, 0, st>>>(n, nb, b0, bM, mu, B);
  SFEM_LAUNCH_CHECK();
  SFEM_TRY(dot(B, B, part0));
  kb_store_sums<<<1, sb, 0, st>>>(part0, gd, nb, BB);
  SFEM_LAUNCH_CHECK();
  SFEM_TRY(apply(X, B, r, 1));
  // first search direction: z = M^-1 r, rz, rr (beta = 0), p = z
  SFEM_TRY(vcycle_b(mg, V, 0, r, z, nb, st, AL, FL, CC));
  SFEM_TRY(dot(r, z, part0));
  SFEM_TRY(dot(r, r, part1));
  kb_cg_beta<<<1, sb, 0, st>>>(part0, gd, part1, gd, nb, S, 1, BB);
  SFEM_LAUNCH_CHECK();
  SFEM_TRY(vec_copy((int)nn, z, p, st));
  double est = 0.0;                                      // max_c rr_c / bb_c
  SFEM_CUDA(cudaMemcpyAsync(t_bpoll.pin + 2, S + (size_t)nb * kSStride, sizeof(double), cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  est = t_bpoll.pin[2];
  if (!(est == est)) { if (forked) t_bwork.join(user); set_error("cg_batch: NaN in the initial residual"); return SFEM_ERR_NOCONV; }
  auto iteration = [&]() -> int {
    SFEM_TRY(apply(p, nullptr, q, 0));
    SFEM_TRY(dot(p, q, part0));
    kb_cg_alpha<<<1, sb, 0, st>>>(part0, gd, nb, S);
    SFEM_LAUNCH_CHECK();
    { Prof prof(PC_VEC, 48.0 * nn, st);
    kb_cg_update<<<gd, bd, 0, st>>>(n, nb, S, p, q, X, r, part1); }
    SFEM_LAUNCH_CHECK();
    SFEM_TRY(vcycle_b(mg, V, 0, r, z, nb, st, AL, FL, CC));
    SFEM_TRY(dot(r, z, part0));
    kb_cg_beta<<<1, sb, 0, st>>>(part0, gd, part1, gd, nb, S, 0, BB);
    SFEM_LAUNCH_CHECK();
    { Prof prof(PC_VEC, 24.0 * nn, st);
    kb_cg_p<<<ge, kThreads, 0, st>>>(n, nb, S, z, p); }
    SFEM_LAUNCH_CHECK();
    return SFEM_OK;
  };
  const double target2 = rtol * rtol;
  int it = 0;
  if (est > target2 && maxit > 0) {
    const bool use_graph = graphs_enabled();
    GraphCache& gc = t_batch_graph;
    if (use_graph) {
      GraphKey key;
      key.a[0] = rowptr; key.a[1] = vals0; key.a[2] = valsM; key.a[3] = mg; key.a[4] = X; key.a[5] = t_bws.ptr;
      key.a[6] = cols; key.a[7] = st;
      key.n = n; key.m = nb; key.nnz = nnz; key.degree = mg->degree * 100 + CC.m; key.nranks = 1; key.epoch = graph_epoch();
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
      SFEM_TRY(t_bpoll.post(it & 1, S + (size_t)nb * kSStride, st));
      if (it > 1) {
        SFEM_TRY(t_bpoll.wait((it - 1) & 1, &est));
        if (!(est == est)) { if (forked) t_bwork.join(user); set_error("cg_batch: NaN residual"); return SFEM_ERR_NOCONV; }
        if (est <= target2) done = true;                 // iteration `it` is already queued: keep its update
      }
    }
    --it;
    SFEM_TRY(t_bpoll.wait(it & 1, &est));
  }
  // true residuals per column
  SFEM_TRY(apply(X, B, r, 1));
  SFEM_TRY(dot(r, r, part0));
  kb_store_sums<<<1, sb, 0, st>>>(part0, gd, nb, RR);
  SFEM_LAUNCH_CHECK();
  double* pin = t_bpoll.pin + 4;                         // [RR | BB | recurrence rr] x kBatchMax
  SFEM_CUDA(cudaMemcpyAsync(pin, RR, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpyAsync(pin + kBatchMax, BB, nb * sizeof(double), cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaMemcpy2DAsync(pin + 2 * kBatchMax, sizeof(double), S + 4, kSStride * sizeof(double), sizeof(double), nb,
                              cudaMemcpyDeviceToHost, st));
  SFEM_CUDA(cudaStreamSynchronize(st));
  if (forked) SFEM_TRY(t_bwork.join(user));
  for (int c = 0; c < nb; ++c) {
    const double bn = pin[kBatchMax + c] > 0.0 ? std::sqrt(pin[kBatchMax + c]) : 1.0;
    const double rel = std::sqrt(pin[c]) / bn;
    h_info[4 * c + 0] = it;
    h_info[4 * c + 1] = rel;
    h_info[4 * c + 2] = (rel <= 10.0 * rtol) ? 1.0 : 0.0;
    h_info[4 * c + 3] = std::sqrt(pin[2 * kBatchMax + c]) / bn;
  }
  return SFEM_OK;
}

int sfem_batch_column(int n, int nb, int c, const double* X, double* out, void* stream) {
  if (n <= 0) return SFEM_OK;
  if (nb < 1 || c < 0 || c >= nb || !X || !out) { set_error("batch_column: bad arguments"); return SFEM_ERR_ARG; }
  kb_column<<<grid_for(n, kThreads * 2), kThreads, 0, (cudaStream_t)stream>>>(n, nb, c, X, out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
