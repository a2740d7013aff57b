// Fused tail of the multigrid V-cycle: all levels with at most SFEM_TAIL_ROWS unknowns in ONE launch.
// EVALUATED AND OFF BY DEFAULT (round 2): correct to rounding, but slower than the separate launches it replaces.
//
// The coarse levels of the hierarchy (7.5 k, 2.2 k rows ... and the dense 556-row coarsest solve at the default geometry)
// are L2-resident and launch-latency bound: per level the V-cycle issues 7 launches of ~4-5 us each whose work is a few
// hundred nanoseconds (profiles/r02_launch_shares_bench.md: 15 such launches = ~70 us of every cycle, the same on 1 and
// on 8 GPUs, the same for a 120 k-dof sweep case and for the 6 M-dof bench case).  Programmatic dependent launch does not
// help (profiles/r02_pdl_microbench.md: a graph edge already costs only 1.6 us).  This kernel runs the whole sub-cycle
//     smooth -> residual -> restrict -> ... -> dense coarsest solve -> ... -> prolong -> smooth
// as a sequence of PHASES inside one thread-block cluster (8 CTAs x 1024 threads on 8 SMs of one GPC): a phase is a
// row-parallel operation over the (row, right-hand side) entries of one level, phases are separated by the hardware
// cluster barrier (cooperative_groups cluster.sync(): MEMBAR.ALL.GPU + UCGABAR arrive / wait + CCTL.IVALL in SASS)
// instead of by kernel boundaries.  Vectors written in one phase are read in the next with ld.global.cg (L2),
// matrices through the read-only path.
// Arithmetic is that of the separate kernels (same Chebyshev-Jacobi recurrences and coefficients, sfem_mg.cu); only the
// summation order inside a row differs (one thread per entry here, LANES lanes per row there): measured agreement
// 2e-16 relative on the cycle's output, identical Krylov iteration counts, bit-reproducible run to run
// (tests/test_gpu_core.py::test_fused_tail_vcycle_matches_separate_launches).
// Measured on B200 (profiles/r02_fused_tail.md): launches per bench step 14 928 -> 10 266, but the step gets SLOWER --
// 18.25 -> 20.97 ms at 385 k dofs, 92.4 -> 95.4 ms at 6.1 M dofs, i.e. the fused sub-cycle takes ~95 us where the 15
// launches took ~70 us.  Every one of its ~20 phases is a chain of four dependent L2 round trips (rowptr -> cols / vals
// -> gather -> store acknowledged by the fence) on 8 SMs with 8 192 threads: two entries per thread on the 7.5 k level,
// three rows per warp in the dense solve, ~4 us per phase -- no better than a launch that spreads the same work over
// 148 SMs.  The micro-benchmark tools/micro/cluster_phase.cu settles it: a dependent CUDA-graph node with a store / a
// gather / three dependent loads costs 0.85 / 1.45 / 1.70 us, the same phase between two cluster barriers 0.75 / 1.33 /
// 1.90 us (cluster.sync() alone 0.3 us) -- there is nothing to gain from fusing launches on this machine, with or
// without the matrices in shared memory.  The knob stays at 0; sfem_mg_set_tail_rows / SFEM_TAIL_ROWS switch the kernel
// on for experiments.
#include "sfem_mg.h"

#include <cooperative_groups.h>
#include <cstdlib>

namespace cg = cooperative_groups;

namespace sfem {

namespace {

constexpr int kTailMaxLevels = 6;
constexpr int kTailThreads = 1024;
constexpr int kTailCluster = 8;

struct TailLevel {
  int n;
  const int* a_rp; const int* a_ci; const double* a_v;
  const int* p_rp; const int* p_ci; const double* p_v;      // prolongation from the next (coarser) level: n rows
  const int* r_rp; const int* r_ci; const double* r_v;      // restriction to the next level: n_next rows
  const double* dinv; const double* coef;
  double* x; const double* b; double* r; double* d0; double* d1;
};
struct TailArgs {
  int nlev;                    // sparse levels L[0 .. nlev-2] + the dense coarsest level L[nlev-1]
  int degree;
  int nb;
  const double* inv;           // dense inverse of the coarsest operator, row-major
  TailLevel L[kTailMaxLevels];
};

__device__ __forceinline__ double ldv(const double* p) { return __ldcg(p); }      // vectors: L2 (written by other SMs)

__device__ __forceinline__ double row_dot_t(const int* __restrict__ rp, const int* __restrict__ ci,
                                            const double* __restrict__ v, const double* x, int row, int c, int nb) {
  const int s = __ldg(rp + row), e = __ldg(rp + row + 1);
  double a0 = 0.0, a1 = 0.0;
  int k = s;
  for (; k + 3 < e; k += 4) {
    const int j0 = __ldg(ci + k), j1 = __ldg(ci + k + 1), j2 = __ldg(ci + k + 2), j3 = __ldg(ci + k + 3);
    const double w0 = __ldg(v + k), w1 = __ldg(v + k + 1), w2 = __ldg(v + k + 2), w3 = __ldg(v + k + 3);
    const double x0 = ldv(x + (size_t)j0 * nb + c), x1 = ldv(x + (size_t)j1 * nb + c);
    const double x2 = ldv(x + (size_t)j2 * nb + c), x3 = ldv(x + (size_t)j3 * nb + c);
    a0 = fma(w0, x0, a0); a1 = fma(w1, x1, a1); a0 = fma(w2, x2, a0); a1 = fma(w3, x3, a1);
  }
  for (; k < e; ++k) a0 = fma(__ldg(v + k), ldv(x + (size_t)__ldg(ci + k) * nb + c), a0);
  return a0 + a1;
}

#define TAIL_ENTRIES(L)                                                                            \
  for (int t = gtid, total_ = (L).n * nb; t < total_; t += nthreads)

// d0 = c0 D^-1 b  (FULL: also r = b, x = d0)
__device__ __forceinline__ void ph_init0(const TailLevel& L, int nb, int gtid, int nthreads, bool full) {
  const double c0 = __ldg(L.coef + 1);
  TAIL_ENTRIES(L) {
    const int row = t / nb;
    const double bi = ldv(L.b + t);
    const double di = c0 * __ldg(L.dinv + row) * bi;
    L.d0[t] = di;
    if (full) { L.r[t] = bi; L.x[t] = di; }
  }
}

__device__ __forceinline__ void ph_cheb(const TailLevel& L, int nb, int gtid, int nthreads, const double* dold, double* dnew,
                                        int step, bool last, bool from_b) {
  const double c1 = __ldg(L.coef + 2 + 2 * step), c2 = __ldg(L.coef + 3 + 2 * step);
  TAIL_ENTRIES(L) {
    const int row = t / nb, c = t - row * nb;
    const double s = row_dot_t(L.a_rp, L.a_ci, L.a_v, dold, row, c, nb);
    const double dd = ldv(dold + t);
    double rin, xin = 0.0;
    if (from_b) rin = ldv(L.b + t);
    else { rin = ldv(L.r + t); xin = ldv(L.x + t); }
    const double rn = rin - s;
    const double dn = c1 * dd + c2 * __ldg(L.dinv + row) * rn;
    L.r[t] = rn;
    dnew[t] = dn;
    L.x[t] = xin + (last ? (dd + dn) : dd);
  }
}

// r = b - A x  (d0 = c0 D^-1 r as well when with_d0)
__device__ __forceinline__ void ph_resid(const TailLevel& L, int nb, int gtid, int nthreads, bool with_d0) {
  const double c0 = __ldg(L.coef + 1);
  TAIL_ENTRIES(L) {
    const int row = t / nb, c = t - row * nb;
    const double s = row_dot_t(L.a_rp, L.a_ci, L.a_v, L.x, row, c, nb);
    const double rr = ldv(L.b + t) - s;
    L.r[t] = rr;
    if (with_d0) L.d0[t] = c0 * __ldg(L.dinv + row) * rr;
  }
}

// b_next = R r
__device__ __forceinline__ void ph_restrict(const TailLevel& L, const TailLevel& N, int nb, int gtid, int nthreads) {
  for (int t = gtid, total_ = N.n * nb; t < total_; t += nthreads) {
    const int row = t / nb, c = t - row * nb;
    const_cast<double*>(N.b)[t] = row_dot_t(L.r_rp, L.r_ci, L.r_v, L.r, row, c, nb);
  }
}

// x += P x_next
__device__ __forceinline__ void ph_prolong(const TailLevel& L, const TailLevel& N, int nb, int gtid, int nthreads) {
  TAIL_ENTRIES(L) {
    const int row = t / nb, c = t - row * nb;
    L.x[t] = ldv(L.x + t) + row_dot_t(L.p_rp, L.p_ci, L.p_v, N.x, row, c, nb);
  }
}

// x += d0
__device__ __forceinline__ void ph_add_d0(const TailLevel& L, int nb, int gtid, int nthreads) {
  TAIL_ENTRIES(L) L.x[t] = ldv(L.x + t) + ldv(L.d0 + t);
}

// x = inv b on the coarsest level: one warp per row, both right-hand sides at once, four entries per round
__device__ __forceinline__ void ph_dense(const TailLevel& L, const double* __restrict__ inv, int nb, int gtid, int nthreads) {
  const int lane = gtid & 31, n = L.n;
  for (int row = gtid >> 5; row < n; row += nthreads >> 5) {
    const double* m = inv + (size_t)row * n;
    double a0 = 0.0, a1 = 0.0;
    for (int j0 = lane; j0 < n; j0 += 128) {
      double mj[4], b0[4], b1[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + 32 * u;
        const bool ok = j < n;
        mj[u] = ok ? __ldg(m + j) : 0.0;
        b0[u] = ok ? ldv(L.b + (size_t)j * nb) : 0.0;
        b1[u] = (ok && nb == 2) ? ldv(L.b + (size_t)j * nb + 1) : 0.0;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) { a0 = fma(mj[u], b0[u], a0); a1 = fma(mj[u], b1[u], a1); }
    }
    a0 = warp_sum(a0);
    a1 = warp_sum(a1);
    if (lane == 0) {
      L.x[(size_t)row * nb] = a0;
      if (nb == 2) L.x[(size_t)row * nb + 1] = a1;
    }
  }
}

__global__ void __launch_bounds__(kTailThreads, 1) k_mg_tail(const TailArgs T) {
  cg::cluster_group cluster = cg::this_cluster();
  const int nthreads = (int)cluster.num_blocks() * blockDim.x;
  const int gtid = (int)cluster.block_rank() * blockDim.x + threadIdx.x;
  const int nb = T.nb, deg = T.degree, last = T.nlev - 1;
  // every thread of every CTA executes every phase loop and every barrier (no early exit anywhere)
  for (int k = 0; k < last; ++k) {
    const TailLevel& L = T.L[k];
    // pre-smoothing from x = 0
    ph_init0(L, nb, gtid, nthreads, deg <= 1);
    cluster.sync();
    const double* dold = L.d0;
    double* dnew = L.d1;
    for (int i = 0; i < deg - 1; ++i) {
      ph_cheb(L, nb, gtid, nthreads, dold, dnew, i, i == deg - 2, i == 0);
      cluster.sync();
      const double* tmp = dold; dold = dnew; dnew = const_cast<double*>(tmp);
    }
    ph_resid(L, nb, gtid, nthreads, false);
    cluster.sync();
    ph_restrict(L, T.L[k + 1], nb, gtid, nthreads);
    cluster.sync();
  }
  ph_dense(T.L[last], T.inv, nb, gtid, nthreads);
  cluster.sync();
  for (int k = last - 1; k >= 0; --k) {
    const TailLevel& L = T.L[k];
    ph_prolong(L, T.L[k + 1], nb, gtid, nthreads);
    cluster.sync();
    ph_resid(L, nb, gtid, nthreads, true);
    cluster.sync();
    if (deg <= 1) {
      ph_add_d0(L, nb, gtid, nthreads);
    } else {
      const double* dold = L.d0;
      double* dnew = L.d1;
      for (int i = 0; i < deg - 1; ++i) {
        if (i > 0) cluster.sync();
        ph_cheb(L, nb, gtid, nthreads, dold, dnew, i, i == deg - 2, false);
        const double* tmp = dold; dold = dnew; dnew = const_cast<double*>(tmp);
      }
    }
    if (k > 0) cluster.sync();
  }
}

std::atomic<int> g_tail_rows{-2};          // -2: read SFEM_TAIL_ROWS on first use; 0: fusion off
std::atomic<int> g_tail_ok{-1};            // can a cluster of this shape be scheduled on this device?

int tail_rows() {
  int r = g_tail_rows.load(std::memory_order_relaxed);
  if (r == -2) {
    const char* e = std::getenv("SFEM_TAIL_ROWS");
    r = e ? std::atoi(e) : 0;             // off by default: slower than the separate launches (see the header)
    if (r < 0) r = 0;
    g_tail_rows.store(r);
  }
  return r;
}

bool tail_supported() {
  int ok = g_tail_ok.load(std::memory_order_relaxed);
  if (ok < 0) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(kTailCluster); cfg.blockDim = dim3(kTailThreads);
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = kTailCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int nclusters = 0;
    const cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_mg_tail, &cfg);
    if (e != cudaSuccess) cudaGetLastError();
    ok = (e == cudaSuccess && nclusters > 0) ? 1 : 0;
    g_tail_ok.store(ok);
  }
  return ok == 1;
}

}  // namespace

// First level of `mg` from which the rest of the hierarchy can run in the fused kernel (-1: none): every level from
// there on has at most tail_rows() unknowns, the coarsest solve is the dense inverse, at least one sparse level.
int mg_tail_start(const sfem_mg* mg) {
  const int rows = tail_rows();
  if (rows <= 0 || mg->tail != nullptr || mg->coarse_inv == nullptr) return -1;
  const int nl = (int)mg->levels.size();
  int first = nl - 1;
  while (first > 0 && mg->levels[first - 1].A.nrows <= rows && nl - (first - 1) <= kTailMaxLevels) --first;
  if (nl - first < 2) return -1;
  return tail_supported() ? first : -1;
}

// x = V-cycle of the levels l .. last applied to b, one launch (b, x: the vectors mg_vcycle_level would get for level l)
int mg_tail_vcycle(sfem_mg* mg, int l, const double* b, double* x, cudaStream_t st) {
  const int nl = (int)mg->levels.size();
  TailArgs T;
  T.nlev = nl - l;
  T.degree = mg->degree;
  T.nb = mg->nb;
  T.inv = mg->coarse_inv;
  for (int k = 0; k < T.nlev; ++k) {
    const MgLevel& M = mg->levels[l + k];
    TailLevel& L = T.L[k];
    L.n = M.A.nrows;
    L.a_rp = M.A.rowptr; L.a_ci = M.A.cols; L.a_v = M.A.vals;
    L.p_rp = M.P.rowptr; L.p_ci = M.P.cols; L.p_v = M.P.vals;
    L.r_rp = M.R.rowptr; L.r_ci = M.R.cols; L.r_v = M.R.vals;
    L.dinv = M.dinv; L.coef = M.coef;
    L.x = (k == 0) ? x : M.x;
    L.b = (k == 0) ? b : M.b;
    L.r = M.r; L.d0 = M.d0; L.d1 = M.d1;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kTailCluster); cfg.blockDim = dim3(kTailThreads); cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = kTailCluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  Prof prof(PC_OTHER, 0.0, st);
  SFEM_CUDA(cudaLaunchKernelEx(&cfg, k_mg_tail, T));
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // namespace sfem

using namespace sfem;

extern "C" {

/* levels with at most `rows` unknowns run in the fused tail kernel (0: off); returns the previous setting */
int sfem_mg_set_tail_rows(int rows) {
  const int old = tail_rows();
  g_tail_rows.store(rows < 0 ? 0 : rows);
  graph_epoch_bump();                      // captured cycles baked in the launch sequence
  return old;
}

}  // extern "C"
