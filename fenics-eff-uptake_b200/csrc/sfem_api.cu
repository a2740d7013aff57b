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
namespace {

// In-place Gauss-Jordan inversion without pivoting (coarsest-level operators are small and
// diagonally dominant after Dirichlet rows are set to identity).  One CTA; M is n x n row-major.
__global__ void __launch_bounds__(1024) k_dense_inverse(int n, const int* __restrict__ rowptr, const int* __restrict__ cols,
                                                        const double* __restrict__ vals, double* __restrict__ M) {
  extern __shared__ double sm[];
  double* rowk = sm;
  double* fcol = sm + n;
  const size_t nn = (size_t)n * n;
  for (size_t t = threadIdx.x; t < nn; t += blockDim.x) M[t] = 0.0;
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x)
    for (int k = rowptr[i]; k < rowptr[i + 1]; ++k) M[(size_t)i * n + cols[k]] = vals[k];
  __syncthreads();
  for (int k = 0; k < n; ++k) {
    const double piv = M[(size_t)k * n + k];
    __syncthreads();
    const double ipiv = 1.0 / piv;
    for (int j = threadIdx.x; j < n; j += blockDim.x) {
      const double a = (j == k) ? 1.0 : M[(size_t)k * n + j];
      const double v = a * ipiv;
      rowk[j] = v;
      fcol[j] = (j == k) ? 0.0 : M[(size_t)j * n + k];
    }
    __syncthreads();
    for (size_t t = threadIdx.x; t < nn; t += blockDim.x) {
      const int i = (int)(t / n), j = (int)(t % n);
      if (i == k) { M[t] = rowk[j]; continue; }
      const double a = (j == k) ? 0.0 : M[t];
      M[t] = fma(-fcol[i], rowk[j], a);
    }
    __syncthreads();
  }
}

}  // namespace

extern "C" {

int sfem_dense_inverse_csr(int n, const int* rowptr, const int* cols, const double* vals, double* out, void* stream) {
  if (n <= 0 || n > 2048) { set_error("dense inverse: n must be in 1..2048"); return SFEM_ERR_ARG; }
  k_dense_inverse<<<1, 1024, 2 * (size_t)n * sizeof(double), (cudaStream_t)stream>>>(n, rowptr, cols, vals, out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_vec_pointwise_mul(int n, double a, const double* d, const double* x, double* y, void* stream) {
  return vec_mul_scale(n, a, d, x, y, (cudaStream_t)stream);
}

int sfem_vec_set(int n, double a, double* x, void* stream) { return vec_set(n, a, x, (cudaStream_t)stream); }

}  // extern "C"
