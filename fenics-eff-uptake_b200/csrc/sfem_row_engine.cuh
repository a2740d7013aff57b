// Software-pipelined CSR row engine of the FP64 SpMV family (sfem_spmv.cu); its accumulator / gather helpers are
// also used by the sliced-ELL engine (sfem_spmv_sell.cu).  TV = storage type of the matrix values, TX = storage type
// of the vectors (both double in the shipped kernels; an FP32-storage preconditioner was measured and dropped, see
// profiles/r01_spmv_microbench.md); all arithmetic is FP64.
#pragma once
#include "sfem_common.cuh"

namespace sfem {

// ------------------------------------------------------------------ pipelined row engine
// A warp owns kThreads/LANES... rows per pass: LANES lanes per row, UNROLL entries per lane and pass.
// Per row the dependent chain is  rowptr -> (cols, vals) -> gather x -> reduce -> epilogue loads;
// executed naively that is three exposed memory latencies per row group and the kernel is
// latency-bound (measured 3.6 TB/s on B200).  The engine software-pipelines it: while the gathers
// of row group i are in flight it already holds the (cols, vals) of group i+1 in registers and has
// the rowptr entries of group i+2 and the epilogue operands of group i+1 requested.
template <int NB>
struct Acc {
  double v[NB];
};

constexpr int kUnroll = 4;
constexpr int kSpmvBlocksPerSm = 4;   // matches __launch_bounds__(kThreads, 4): one resident wave, persistent grid-stride

// gathered x entry (NB interleaved values) of storage type TX; arithmetic is always FP64
template <int NB, class TX>
struct XVal;
template <>
struct XVal<1, double> {
  double a;
  __device__ __forceinline__ void load(const double* __restrict__ x, int col, bool ok) { a = ok ? __ldg(x + col) : 0.0; }
  __device__ __forceinline__ void fma_into(double v, Acc<1>& acc) const { acc.v[0] = fma(v, a, acc.v[0]); }
};
template <>
struct XVal<2, double> {
  double2 a;
  __device__ __forceinline__ void load(const double* __restrict__ x, int col, bool ok) {
    a = ok ? __ldg(reinterpret_cast<const double2*>(x) + col) : make_double2(0.0, 0.0);
  }
  __device__ __forceinline__ void fma_into(double v, Acc<2>& acc) const {
    acc.v[0] = fma(v, a.x, acc.v[0]);
    acc.v[1] = fma(v, a.y, acc.v[1]);
  }
};
template <>
struct XVal<1, float> {
  float a;
  __device__ __forceinline__ void load(const float* __restrict__ x, int col, bool ok) { a = ok ? __ldg(x + col) : 0.f; }
  __device__ __forceinline__ void fma_into(double v, Acc<1>& acc) const { acc.v[0] = fma(v, (double)a, acc.v[0]); }
};
template <>
struct XVal<2, float> {
  float2 a;
  __device__ __forceinline__ void load(const float* __restrict__ x, int col, bool ok) {
    a = ok ? __ldg(reinterpret_cast<const float2*>(x) + col) : make_float2(0.f, 0.f);
  }
  __device__ __forceinline__ void fma_into(double v, Acc<2>& acc) const {
    acc.v[0] = fma(v, (double)a.x, acc.v[0]);
    acc.v[1] = fma(v, (double)a.y, acc.v[1]);
  }
};

template <int LANES, class TV>
__device__ __forceinline__ void load_pass(const int* __restrict__ cols, const TV* __restrict__ vals, int k, int e,
                                          int (&cc)[kUnroll], double (&vv)[kUnroll]) {
#pragma unroll
  for (int j = 0; j < kUnroll; ++j) {
    const int kk = k + j * LANES;
    const bool ok = kk < e;
    cc[j] = ok ? __ldcs(cols + kk) : -1;
    vv[j] = ok ? (double)__ldcs(vals + kk) : 0.0;
  }
}

// Epi: struct with   Pre pre(int row, int lane, bool active)   (loads issued early)
//                    void fin(int row, int lane, double value, const Pre&)   (called for lane < NB of valid rows)
template <int LANES, int NB, class Epi, class TV = double, class TX = double>
__device__ __forceinline__ void row_engine(int nrows, const int* __restrict__ rowptr, const int* __restrict__ cols,
                                           const TV* __restrict__ vals, const TX* __restrict__ x, Epi& epi) {
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  const long long stride = (long long)gridDim.x * ROWS;
  long long base = (long long)blockIdx.x * ROWS;
  if (base >= nrows) return;
  // prologue: row group 0 fully fetched, rowptr of group 1 requested
  long long row = base + sub;
  bool valid = row < nrows;
  int s = 0, e = 0;
  if (valid) { s = rowptr[row]; e = rowptr[row + 1]; }
  int cc[kUnroll];
  double vv[kUnroll];
  load_pass<LANES, TV>(cols, vals, s + lane, e, cc, vv);
  typename Epi::Pre pre = epi.pre((int)row, lane, valid && lane < NB);
  long long nrow = row + stride;
  bool nvalid = nrow < nrows;
  int ns = 0, ne = 0;
  if (nvalid) { ns = rowptr[nrow]; ne = rowptr[nrow + 1]; }
  for (; base < nrows; base += stride) {
    // rowptr of the group after next
    const long long nnrow = nrow + stride;
    const bool nnvalid = nnrow < nrows;
    int nns = 0, nne = 0;
    if (nnvalid) { nns = rowptr[nnrow]; nne = rowptr[nnrow + 1]; }
    // gathers of the current group (first pass)
    XVal<NB, TX> xv[kUnroll];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) xv[j].load(x, cc[j], cc[j] >= 0);
    // (cols, vals) and epilogue operands of the next group
    int ncc[kUnroll];
    double nvv[kUnroll];
    load_pass<LANES, TV>(cols, vals, ns + lane, ne, ncc, nvv);
    typename Epi::Pre npre = epi.pre((int)nrow, lane, nvalid && lane < NB);
    Acc<NB> a0, a1;
#pragma unroll
    for (int c = 0; c < NB; ++c) { a0.v[c] = 0.0; a1.v[c] = 0.0; }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) xv[j].fma_into(vv[j], (j & 1) ? a1 : a0);
    // rows longer than one pass (rare for FEM patterns)
    for (int k = s + lane + kUnroll * LANES; k < e; k += kUnroll * LANES) {
      load_pass<LANES, TV>(cols, vals, k, e, cc, vv);
#pragma unroll
      for (int j = 0; j < kUnroll; ++j) xv[j].load(x, cc[j], cc[j] >= 0);
#pragma unroll
      for (int j = 0; j < kUnroll; ++j) xv[j].fma_into(vv[j], (j & 1) ? a1 : a0);
    }
#pragma unroll
    for (int c = 0; c < NB; ++c) {
      double t = a0.v[c] + a1.v[c];
#pragma unroll
      for (int o = LANES >> 1; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      a0.v[c] = t;
    }
    if (valid && lane < NB) epi.fin((int)row, lane, (NB == 2 && lane == 1) ? a0.v[NB - 1] : a0.v[0], pre);
    // rotate the pipeline
    row = nrow; valid = nvalid; s = ns; e = ne;
    nrow = nnrow; nvalid = nnvalid; ns = nns; ne = nne;
    pre = npre;
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) { cc[j] = ncc[j]; vv[j] = nvv[j]; }
  }
}


#define SFEM_DISPATCH_LANES(L, ...)       \
  switch (L) {                            \
    case 1: { constexpr int LN = 1; __VA_ARGS__; } break;   \
    case 2: { constexpr int LN = 2; __VA_ARGS__; } break;   \
    case 4: { constexpr int LN = 4; __VA_ARGS__; } break;   \
    case 8: { constexpr int LN = 8; __VA_ARGS__; } break;   \
    case 16: { constexpr int LN = 16; __VA_ARGS__; } break; \
    default: { constexpr int LN = 32; __VA_ARGS__; } break; \
  }

// lanes per row: 4 entries per lane and pass -> the smallest lane group that covers an average row in one
// pass; with NB = 2 the two result lanes need LANES >= 2.  SFEM_LANES overrides (experiments).
int engine_lanes(long long nnz, int nrows, int nb);

}  // namespace sfem
