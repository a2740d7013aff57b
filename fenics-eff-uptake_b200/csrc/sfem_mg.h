// Multigrid handle layout (internal).
#pragma once
#include "sfem_common.cuh"
#include "sfem_internal.h"
#include "sfem_graph.h"

#include <vector>

namespace sfem {

constexpr int kMaxChebDegree = 16;
constexpr int kChebCoefLen = 2 + 2 * kMaxChebDegree;   // [lmax, 1/theta, (c1, c2) per step]

struct MgLevel {
  Csr A, P, R;
  double* dinv = nullptr;
  double* x = nullptr;     // level iterate (levels > 0)
  double* b = nullptr;     // level right-hand side (levels > 0)
  double* r = nullptr;
  double* d0 = nullptr;
  double* d1 = nullptr;
  double* coef = nullptr;  // device: Chebyshev coefficients of this level (written by sfem_mg_setup)
};

}  // namespace sfem

struct sfem_mg {
  std::vector<sfem::MgLevel> levels;
  const double* coarse_inv = nullptr;
  int degree = 2;
  int nb = 1;                         // interleaved right-hand sides per V-cycle (1 or 2)
  double ratio = 8.0;
  double* scratch = nullptr;
  bool ready = false;
  // multi-GPU: levels above are row-partitioned; below them a replicated hierarchy (`tail`) is solved
  // redundantly by every rank after an all-reduce of the restricted residual
  sfem_mg* tail = nullptr;
  int n_tail = 0;
  double* tail_b = nullptr;
  double* tail_x = nullptr;
  // captured Krylov iteration bodies preconditioned by this handle (owned here: destroyed with the handle)
  sfem::GraphCache cg_graph, gmres_graph;
};

namespace sfem {
// Chebyshev-Jacobi sweep of `degree` steps on D^-1 A; coefficients are read from device memory
// (coef, layout above) so that captured graphs stay valid when the operator values change.
int smooth(const Csr& A, const double* dinv, const double* coef, int degree, const double* b, double* x,
           double* r, double* d0, double* d1, bool zero_init, cudaStream_t st, int nb = 1);
// coef <- coefficients for the window [lmax/ratio, lmax]; lmax = Gershgorin bound of D^-1 A when A is
// given (scratch: kMaxPartials doubles), else fixed_lmax.  No host synchronisation.
int cheb_setup(const Csr* A, const double* dinv, double fixed_lmax, double ratio, int degree, double* scratch,
               double* coef, cudaStream_t st, bool distributed = false);
int mg_vcycle_level(sfem_mg* mg, int l, const double* b, double* x, cudaStream_t st);
// sfem_mg_tail.cu: the levels l >= mg_tail_start(mg) (-1: none) run as one fused cluster kernel
int mg_tail_start(const sfem_mg* mg);
int mg_tail_vcycle(sfem_mg* mg, int l, const double* b, double* x, cudaStream_t st);
}  // namespace sfem
