// Row epilogues shared by the two SpMV engines (vector: sfem_spmv.cu, TMA-staged: sfem_spmv_staged.cu).
// An epilogue turns the row sums s = (A x)_row into the fused update of its kernel.
//   Pre pre(row, lane, active)           operand loads, issued early (before the row sum is known)
//   void fin(row, lane, value, pre)      called for lane < NB of valid rows; lane = right-hand side index
#pragma once
#include "sfem_common.cuh"

namespace sfem {

// ---- epilogues
// MODE 0: y = s;  1: y = b - s;  2: y += s
// (TX = storage type of the vectors: double, or float in the FP32-storage preconditioner; values are FP64)
template <int NB, int MODE, class TX = double>
struct EpiStore {
  const TX* __restrict__ b;
  TX* __restrict__ y;
  struct Pre { double a; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.a = 0.0;
    if (active) {
      const size_t i = (size_t)row * NB + lane;
      if (MODE == 1) p.a = b[i];
      if (MODE == 2) p.a = y[i];
    }
    return p;
  }
  __device__ __forceinline__ void fin(int row, int lane, double sv, const Pre& p) const {
    const size_t i = (size_t)row * NB + lane;
    y[i] = (TX)((MODE == 0) ? sv : (MODE == 1 ? p.a - sv : p.a + sv));
  }
};

// y = s (MODE 0) or y += s (MODE 2); acc += dx[i] * y_new
template <int NB, int MODE>
struct EpiDot {
  const double* __restrict__ dx;
  double* __restrict__ y;
  double acc;
  struct Pre { double d, yo; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.d = 0.0; p.yo = 0.0;
    if (active) {
      const size_t i = (size_t)row * NB + lane;
      p.d = dx[i];
      if (MODE == 2) p.yo = y[i];
    }
    return p;
  }
  __device__ __forceinline__ void fin(int row, int lane, double sv, const Pre& p) {
    const size_t i = (size_t)row * NB + lane;
    const double yn = (MODE == 2) ? p.yo + sv : sv;
    y[i] = yn;
    acc = fma(p.d, yn, acc);
  }
};

// One fused Chebyshev-Jacobi step (see sfem_mg.cu), x of the engine = d_old:
//   t = (A d_old)_i;  r_i -= t;  x_i += d_old_i (+ d_new_i when LAST);  d_new_i = c1 d_old_i + c2 dinv_i r_i
template <int NB, class TX = double>
struct EpiCheb {
  const TX* __restrict__ dinv;
  const TX* __restrict__ d_old;
  TX* __restrict__ d_new;
  TX* __restrict__ r;
  TX* __restrict__ xx;
  double c1, c2;
  int last;
  struct Pre { double r, d, di, x; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.r = p.d = p.di = p.x = 0.0;
    if (active) {
      const size_t i = (size_t)row * NB + lane;
      p.r = r[i]; p.d = d_old[i]; p.di = dinv[row]; p.x = xx[i];
    }
    return p;
  }
  __device__ __forceinline__ void fin(int row, int lane, double t, const Pre& p) const {
    const size_t i = (size_t)row * NB + lane;
    const double rn = p.r - t;
    const double dn = c1 * p.d + c2 * p.di * rn;
    r[i] = (TX)rn;
    d_new[i] = (TX)dn;
    xx[i] = (TX)(p.x + (last ? (p.d + dn) : p.d));
  }
};

// r = b - A x ; d = c0 * dinv * r      (start of a smoothing sweep with a non-zero iterate)
template <int NB, class TX = double>
struct EpiResidD0 {
  const TX* __restrict__ dinv;
  const TX* __restrict__ b;
  TX* __restrict__ r;
  TX* __restrict__ d;
  double c0;
  struct Pre { double b, di; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.b = p.di = 0.0;
    if (active) { p.b = b[(size_t)row * NB + lane]; p.di = dinv[row]; }
    return p;
  }
  __device__ __forceinline__ void fin(int row, int lane, double sv, const Pre& p) const {
    const size_t i = (size_t)row * NB + lane;
    const double rr = p.b - sv;
    r[i] = (TX)rr;
    d[i] = (TX)(c0 * p.di * rr);
  }
};


}  // namespace sfem
