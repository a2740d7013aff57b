// C-ABI entry points for the sparse mat-vec kernels.
#include "sfem_common.cuh"
#include "sfem_internal.h"

using namespace sfem;

extern "C" {

int sfem_spmv_csr_f64(int nrows, int nnz, const int* rowptr, const int* cols, const double* vals, const double* x,
                      const double* b, double* y, int mode, void* stream) {
  if (nrows < 0 || nnz < 0 || mode < 0 || mode > 2) { set_error("spmv: bad arguments"); return SFEM_ERR_ARG; }
  Csr A;
  A.nrows = nrows; A.ncols = nrows; A.nnz = nnz;
  A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  return spmv(A, x, b, y, mode, (cudaStream_t)stream);
}

int sfem_spmv_csr_f64_nb(int nrows, int ncols, int nnz, const int* rowptr, const int* cols, const double* vals,
                         const double* x, const double* b, double* y, int mode, int nb, void* stream) {
  if (nrows < 0 || ncols < 0 || nnz < 0 || mode < 0 || mode > 2) { set_error("spmv: bad arguments"); return SFEM_ERR_ARG; }
  Csr A;
  A.nrows = nrows; A.ncols = ncols; A.nnz = nnz;
  A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  return spmv(A, x, b, y, mode, (cudaStream_t)stream, nb);
}

int sfem_spmv_csr_f64_staged(int nrows, int ncols, int nnz, const int* rowptr, const int* cols, const double* vals,
                             const double* x, const double* b, double* y, int mode, int nb, void* stream) {
  if (nrows < 0 || nnz < 0 || mode < 0 || mode > 2 || (nb != 1 && nb != 2)) { set_error("staged spmv: bad arguments"); return SFEM_ERR_ARG; }
  if (mode == 1 && b == nullptr) { set_error("staged spmv: mode 1 needs b"); return SFEM_ERR_ARG; }
  Csr A;
  A.nrows = nrows; A.ncols = ncols; A.nnz = nnz;
  A.rowptr = rowptr; A.cols = cols; A.vals = vals;
  const int r = staged_spmv(A, x, b, y, mode, nb, (cudaStream_t)stream);
  if (r == 0) { set_error("staged spmv: no tile plan registered for this matrix (or fewer tiles than SFEM_STAGED_MIN_TILES)"); return SFEM_ERR_ARG; }
  return r < 0 ? r : SFEM_OK;
}

}  // extern "C"

// ------------------------------------------------------------------ small dense inverse (coarsest level)
// Blocked in-place Gauss-Jordan inversion without pivoting (coarsest-level operators are small and diagonally dominant
// after Dirichlet rows are set to identity), 32 x 32 blocks, every phase spread over the whole GPU:
//   for each block column k:   P = inv(A_kk)                       (one CTA, shared memory)
//                              A_kj <- P A_kj (j != k), A_kk <- P  (row panel, one CTA per block column; the old block
//                                                                   column A_ik is saved to F on the way)
//                              A_ij <- A_ij - F_i A_kj (i, j != k), A_ik <- -F_i P     (one CTA per block)
// = 3 n / 32 short launches instead of one CTA sweeping the whole matrix n times (1.3 ms for n ~ 450 -> ~0.1 ms; the
// inverse is rebuilt at every re-assembly, i.e. once per case of a mu sweep).
namespace {

constexpr int kGjB = 32;

__global__ void __launch_bounds__(256) k_gj_load(int n, int np, const int* __restrict__ rowptr, const int* __restrict__ cols,
                                                 const double* __restrict__ vals, double* __restrict__ W) {
  for (int i = blockIdx.x; i < np; i += gridDim.x) {
    double* row = W + (size_t)i * np;
    for (int j = threadIdx.x; j < np; j += blockDim.x) row[j] = (i >= n && j == i) ? 1.0 : 0.0;
    __syncthreads();
    if (i < n)
      for (int k = rowptr[i] + threadIdx.x; k < rowptr[i + 1]; k += blockDim.x) row[cols[k]] = vals[k];
    __syncthreads();
  }
}

__global__ void __launch_bounds__(kGjB * kGjB) k_gj_pivot(int np, int k, const double* __restrict__ W, double* __restrict__ Pout) {
  __shared__ double a[kGjB][kGjB + 1];
  __shared__ double rowk[kGjB], colk[kGjB];
  const int i = threadIdx.y, j = threadIdx.x;
  a[i][j] = W[(size_t)(k * kGjB + i) * np + k * kGjB + j];
  __syncthreads();
  for (int s = 0; s < kGjB; ++s) {
    const double ipiv = 1.0 / a[s][s];
    __syncthreads();
    if (i == 0) rowk[j] = ((j == s) ? 1.0 : a[s][j]) * ipiv;
    if (j == 0) colk[i] = (i == s) ? 0.0 : a[i][s];
    __syncthreads();
    const double old = (j == s) ? 0.0 : a[i][j];
    a[i][j] = (i == s) ? rowk[j] : fma(-colk[i], rowk[j], old);
    __syncthreads();
  }
  Pout[i * kGjB + j] = a[i][j];
}

// block b: saves F_b = A_bk (old block column) and rewrites the row-panel block A_kb <- P A_kb (A_kk <- P)
__global__ void __launch_bounds__(kGjB * kGjB) k_gj_panel(int np, int k, double* __restrict__ W, const double* __restrict__ Pm,
                                                         double* __restrict__ F) {
  __shared__ double p[kGjB][kGjB + 1], t[kGjB][kGjB + 1];
  const int b = blockIdx.x, i = threadIdx.y, j = threadIdx.x;
  F[((size_t)b * kGjB + i) * kGjB + j] = W[(size_t)(b * kGjB + i) * np + k * kGjB + j];
  p[i][j] = Pm[i * kGjB + j];
  double* blk = W + (size_t)(k * kGjB) * np + b * kGjB;
  t[i][j] = blk[(size_t)i * np + j];
  __syncthreads();
  if (b == k) { blk[(size_t)i * np + j] = p[i][j]; return; }
  double acc = 0.0;
#pragma unroll
  for (int m = 0; m < kGjB; ++m) acc = fma(p[i][m], t[m][j], acc);
  blk[(size_t)i * np + j] = acc;
}

__global__ void __launch_bounds__(kGjB * kGjB) k_gj_update(int np, int k, double* __restrict__ W, const double* __restrict__ Pm,
                                                          const double* __restrict__ F) {
  const int bi = blockIdx.y, bj = blockIdx.x;
  if (bi == k) return;
  __shared__ double f[kGjB][kGjB + 1], t[kGjB][kGjB + 1];
  const int i = threadIdx.y, j = threadIdx.x;
  f[i][j] = F[((size_t)bi * kGjB + i) * kGjB + j];
  t[i][j] = (bj == k) ? Pm[i * kGjB + j] : W[(size_t)(k * kGjB + i) * np + bj * kGjB + j];
  __syncthreads();
  double acc = 0.0;
#pragma unroll
  for (int m = 0; m < kGjB; ++m) acc = fma(f[i][m], t[m][j], acc);
  double* dst = W + (size_t)(bi * kGjB + i) * np + bj * kGjB + j;
  *dst = (bj == k) ? -acc : *dst - acc;
}

__global__ void __launch_bounds__(256) k_gj_store(int n, int np, const double* __restrict__ W, double* __restrict__ M) {
  for (int i = blockIdx.x; i < n; i += gridDim.x)
    for (int j = threadIdx.x; j < n; j += blockDim.x) M[(size_t)i * n + j] = W[(size_t)i * np + j];
}

struct GjWorkspace {
  double* ptr = nullptr;
  size_t cap = 0;
};
thread_local GjWorkspace t_gj;

}  // namespace

extern "C" {

int sfem_dense_inverse_csr(int n, const int* rowptr, const int* cols, const double* vals, double* out, void* stream) {
  if (n <= 0 || n > 2048) { set_error("dense inverse: n must be in 1..2048"); return SFEM_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int nbk = (n + kGjB - 1) / kGjB, np = nbk * kGjB;
  const size_t need = (size_t)np * np + (size_t)kGjB * kGjB + (size_t)np * kGjB;
  if (need > t_gj.cap) {
    if (t_gj.ptr) { SFEM_CUDA(cudaStreamSynchronize(st)); cudaFree(t_gj.ptr); t_gj.ptr = nullptr; t_gj.cap = 0; }
    SFEM_CUDA(cudaMalloc(&t_gj.ptr, need * sizeof(double)));
    t_gj.cap = need;
  }
  double* W = t_gj.ptr;
  double* Pm = W + (size_t)np * np;
  double* F = Pm + kGjB * kGjB;
  const dim3 tb(kGjB, kGjB);
  k_gj_load<<<grid_for(np, 1, 8), 256, 0, st>>>(n, np, rowptr, cols, vals, W);
  SFEM_LAUNCH_CHECK();
  for (int k = 0; k < nbk; ++k) {
    k_gj_pivot<<<1, tb, 0, st>>>(np, k, W, Pm);
    SFEM_LAUNCH_CHECK();
    k_gj_panel<<<nbk, tb, 0, st>>>(np, k, W, Pm, F);
    SFEM_LAUNCH_CHECK();
    k_gj_update<<<dim3(nbk, nbk), tb, 0, st>>>(np, k, W, Pm, F);
    SFEM_LAUNCH_CHECK();
  }
  k_gj_store<<<grid_for(n, 1, 8), 256, 0, st>>>(n, np, W, out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_vec_pointwise_mul(int n, double a, const double* d, const double* x, double* y, void* stream) {
  return vec_mul_scale(n, a, d, x, y, (cudaStream_t)stream);
}

int sfem_vec_set(int n, double a, double* x, void* stream) { return vec_set(n, a, x, (cudaStream_t)stream); }

}  // extern "C"
