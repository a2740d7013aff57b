// Multigrid handle layout (internal).
#pragma once
#include "sfem_common.cuh"
#include "sfem_internal.h"

#include <vector>

namespace sfem {

struct MgLevel {
  Csr A, P, R;
  double* dinv = nullptr;
  double* x = nullptr;     // level iterate (levels > 0)
  double* b = nullptr;     // level right-hand side (levels > 0)
  double* r = nullptr;
  double* d0 = nullptr;
  double* d1 = nullptr;
  double lmax = 2.0;
};

}  // namespace sfem

struct sfem_mg {
  std::vector<sfem::MgLevel> levels;
  const double* coarse_inv = nullptr;
  int degree = 2;
  double ratio = 8.0;
  double* scratch = nullptr;
  bool ready = false;
  bool use_power_iteration = false;   // SFEM_LMAX=power: estimate instead of the Gershgorin bound
};

namespace sfem {
int smooth(const Csr& A, const double* dinv, double lmax, double ratio, int degree, const double* b, double* x,
           double* r, double* d0, double* d1, bool zero_init, cudaStream_t st);
int mg_vcycle_level(sfem_mg* mg, int l, const double* b, double* x, cudaStream_t st);
int estimate_lambda_max(const Csr& A, const double* dinv, double* v, double* w, double* scratch, double* out,
                        cudaStream_t st);
}  // namespace sfem
