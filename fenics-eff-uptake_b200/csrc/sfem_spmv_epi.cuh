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
  // NB = 2, one thread per row (sliced-ELL engine): both right-hand sides as one 16-byte access (double vectors,
  // 16-byte aligned -- checked by the launcher)
  struct Pre2 { double2 a; };
  __device__ __forceinline__ Pre2 pre2(int row, bool active) const {
    Pre2 p; p.a = make_double2(0.0, 0.0);
    if (active) {
      if (MODE == 1) p.a = reinterpret_cast<const double2*>(b)[row];
      if (MODE == 2) p.a = reinterpret_cast<const double2*>(y)[row];
    }
    return p;
  }
  __device__ __forceinline__ void fin2(int row, double s0, double s1, const Pre2& p) const {
    reinterpret_cast<double2*>(y)[row] = (MODE == 0) ? make_double2(s0, s1)
                                                     : (MODE == 1 ? make_double2(p.a.x - s0, p.a.y - s1)
                                                                  : make_double2(p.a.x + s0, p.a.y + s1));
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
  struct Pre2 { double2 d, yo; };
  __device__ __forceinline__ Pre2 pre2(int row, bool active) const {
    Pre2 p; p.d = make_double2(0.0, 0.0); p.yo = make_double2(0.0, 0.0);
    if (active) {
      p.d = reinterpret_cast<const double2*>(dx)[row];
      if (MODE == 2) p.yo = reinterpret_cast<const double2*>(y)[row];
    }
    return p;
  }
  __device__ __forceinline__ void fin2(int row, double s0, double s1, const Pre2& p) {
    const double y0 = (MODE == 2) ? p.yo.x + s0 : s0;
    const double y1 = (MODE == 2) ? p.yo.y + s1 : s1;
    reinterpret_cast<double2*>(y)[row] = make_double2(y0, y1);
    acc = fma(p.d.x, y0, acc);
    acc = fma(p.d.y, y1, acc);
  }
};

// One fused Chebyshev-Jacobi step (see sfem_mg.cu), x of the engine = d_old:
//   t = (A d_old)_i;  r_i -= t;  x_i += d_old_i (+ d_new_i when LAST);  d_new_i = c1 d_old_i + c2 dinv_i r_i
// FIRST step of a sweep that starts from x = 0 (b0 != nullptr): the residual is read from the right-hand side b0 and x is
// taken as zero, so the sweep's set-up kernel writes only d_0 (no copy r = b, no x = 0) -- 48 B per unknown less per sweep.
template <int NB, class TX = double>
struct EpiCheb {
  const TX* __restrict__ dinv;
  const TX* __restrict__ d_old;
  TX* __restrict__ d_new;
  TX* __restrict__ r;
  TX* __restrict__ xx;
  double c1, c2;
  int last;
  const TX* __restrict__ b0;
  struct Pre { double r, d, di, x; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.r = p.d = p.di = p.x = 0.0;
    if (active) {
      const size_t i = (size_t)row * NB + lane;
      p.d = d_old[i]; p.di = dinv[row];
      if (b0 != nullptr) p.r = b0[i];
      else { p.r = r[i]; p.x = xx[i]; }
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


// Variants for engines that take the epilogue by value as a kernel parameter (staged, sliced-ELL): the Chebyshev
// coefficients live in device memory (no host synchronisation in multigrid set-up), so they are fetched through
// these pointers in pre().
template <int NB>
struct EpiChebPtr {
  const double* __restrict__ dinv;
  const double* __restrict__ d_old;
  double* __restrict__ d_new;
  double* __restrict__ r;
  double* __restrict__ xx;
  const double* __restrict__ c12;
  int last;
  const double* __restrict__ b0;       // first step of a sweep from x = 0: residual read from b0, x taken as zero (see EpiCheb)
  struct Pre { double r, d, di, x; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.r = p.d = p.di = p.x = 0.0;
    if (active) {
      const size_t i = (size_t)row * NB + lane;
      p.d = d_old[i]; p.di = dinv[row];
      if (b0 != nullptr) p.r = b0[i];
      else { p.r = r[i]; p.x = xx[i]; }
    }
    return p;
  }
  __device__ __forceinline__ void fin(int row, int lane, double t, const Pre& p) const {
    const size_t i = (size_t)row * NB + lane;
    const double rn = p.r - t;
    const double dn = __ldg(c12) * p.d + __ldg(c12 + 1) * p.di * rn;    // coefficients: L1-resident broadcast loads
    r[i] = rn;
    d_new[i] = dn;
    xx[i] = p.x + (last ? (p.d + dn) : p.d);
  }
  struct Pre2 { double2 r, d, x; double di; };
  __device__ __forceinline__ Pre2 pre2(int row, bool active) const {
    Pre2 p; p.r = p.d = p.x = make_double2(0.0, 0.0); p.di = 0.0;
    if (active) {
      p.d = reinterpret_cast<const double2*>(d_old)[row];
      p.di = dinv[row];
      if (b0 != nullptr) p.r = reinterpret_cast<const double2*>(b0)[row];
      else {
        p.r = reinterpret_cast<const double2*>(r)[row];
        p.x = reinterpret_cast<const double2*>(xx)[row];
      }
    }
    return p;
  }
  __device__ __forceinline__ void fin2(int row, double t0, double t1, const Pre2& p) const {
    const double c1 = __ldg(c12), c2 = __ldg(c12 + 1) * p.di;
    const double r0 = p.r.x - t0, r1 = p.r.y - t1;
    const double d0 = c1 * p.d.x + c2 * r0, d1 = c1 * p.d.y + c2 * r1;
    reinterpret_cast<double2*>(r)[row] = make_double2(r0, r1);
    reinterpret_cast<double2*>(d_new)[row] = make_double2(d0, d1);
    reinterpret_cast<double2*>(xx)[row] = last ? make_double2(p.x.x + (p.d.x + d0), p.x.y + (p.d.y + d1))
                                               : make_double2(p.x.x + p.d.x, p.x.y + p.d.y);
  }
};

template <int NB>
struct EpiResidD0Ptr {
  const double* __restrict__ dinv;
  const double* __restrict__ b;
  double* __restrict__ r;
  double* __restrict__ d;
  const double* __restrict__ c0p;
  struct Pre { double b, di; };
  __device__ __forceinline__ Pre pre(int row, int lane, bool active) const {
    Pre p; p.b = p.di = 0.0;
    if (active) { p.b = b[(size_t)row * NB + lane]; p.di = dinv[row]; }
    return p;
  }
  __device__ __forceinline__ void fin(int row, int lane, double sv, const Pre& p) const {
    const size_t i = (size_t)row * NB + lane;
    const double rr = p.b - sv;
    r[i] = rr;
    d[i] = __ldg(c0p) * p.di * rr;
  }
  struct Pre2 { double2 b; double di; };
  __device__ __forceinline__ Pre2 pre2(int row, bool active) const {
    Pre2 p; p.b = make_double2(0.0, 0.0); p.di = 0.0;
    if (active) { p.b = reinterpret_cast<const double2*>(b)[row]; p.di = dinv[row]; }
    return p;
  }
  __device__ __forceinline__ void fin2(int row, double s0, double s1, const Pre2& p) const {
    const double r0 = p.b.x - s0, r1 = p.b.y - s1;
    const double c = __ldg(c0p) * p.di;
    reinterpret_cast<double2*>(r)[row] = make_double2(r0, r1);
    reinterpret_cast<double2*>(d)[row] = make_double2(c * r0, c * r1);
  }
};

// dot contribution accumulated by an epilogue (0 for the non-reducing ones)
template <int NB, int MODE>
__device__ __forceinline__ double epi_acc_of(const EpiDot<NB, MODE>& e) { return e.acc; }
template <class E>
__device__ __forceinline__ double epi_acc_of(const E&) { return 0.0; }

}  // namespace sfem
