// Internal (non-ABI) declarations shared by the translation units of libsulcusfem.
#pragma once
#include <cuda_runtime.h>

namespace sfem {

struct Csr {
  int nrows = 0;
  int ncols = 0;
  long long nnz = 0;
  const int* rowptr = nullptr;
  const int* cols = nullptr;
  const double* vals = nullptr;
};

// sfem_spmv.cu  (nb = number of interleaved right-hand sides, 1 or 2)
// (exchange = false: the ghost entries of x are already current -- the caller ran the halo exchange itself)
int spmv(const Csr& A, const double* x, const double* b, double* y, int mode, cudaStream_t st, int nb = 1,
         bool exchange = true);
int spmv_dot(const Csr& A, const double* x, double* y, double* partial, int* nparts, cudaStream_t st, int nb = 1,
             const double* dotx = nullptr, int mode = 0, bool exchange = true);   // partial sums of <dotx, y>; mode 2: y += A x first
// Chebyshev coefficients are read from device memory: c12 -> {c1, c2}, c0 -> {c0}
// (b0 != nullptr: first step of a sweep that starts from x = 0 -- the residual is read from b0, x is taken as zero)
int cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
              const double* c12, int last, cudaStream_t st, int nb = 1, const double* b0 = nullptr);
int resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
             const double* c0, cudaStream_t st, int nb = 1);
// y_u = K z_u + BT z_p (interleaved velocity), partial sums of <z_u, y_u>
int stokes_apply_u(const Csr& K, const Csr& BT, const double* zu, const double* zp, double* yu, double* partial,
                   int* nparts, cudaStream_t st);
// sfem_spmv_staged.cu: TMA-staged engine; each returns 1 when it took the launch (a tile plan is registered for
// A.rowptr and the matrix is large enough), 0 when the caller should use the vector engine, < 0 on error
int staged_spmv(const Csr& A, const double* x, const double* b, double* y, int mode, int nb, cudaStream_t st);
int staged_spmv_dot(const Csr& A, const double* x, const double* dx, double* y, double* partial, int* nparts, int mode,
                    int nb, cudaStream_t st);
int staged_cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
                     const double* c12, int last, int nb, cudaStream_t st, const double* b0 = nullptr);
int staged_resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
                    const double* c0, int nb, cudaStream_t st);

// sfem_spmv_sell.cu: sliced-ELL mirror engine; same return convention as the staged engine
int sell_spmv(const Csr& A, const double* x, const double* b, double* y, int mode, int nb, cudaStream_t st);
int sell_spmv_dot(const Csr& A, const double* x, const double* dx, double* y, double* partial, int* nparts, int mode,
                  int nb, cudaStream_t st);
int sell_cheb_step(const Csr& A, const double* dinv, const double* d_old, double* d_new, double* r, double* x,
                   const double* c12, int last, int nb, cudaStream_t st, const double* b0 = nullptr);
int sell_resid_d0(const Csr& A, const double* dinv, const double* b, const double* x, double* r, double* d,
                  const double* c0, int nb, cudaStream_t st);
void sell_mark_dirty(const double* csr_vals);     // called by every entry that writes CSR values
int sell_ensure_all(cudaStream_t st);             // re-pack every dirty mirror (solver entries, before graph replays)

// sfem_vector.cu
unsigned long long graph_epoch();                 // see sfem_graph.h
void graph_epoch_bump();
int vec_set(int n, double a, double* x, cudaStream_t st);
int vec_copy(int n, const double* x, double* y, cudaStream_t st);
int vec_axpby(int n, double a, const double* x, double b, double* y, cudaStream_t st);
int vec_mul_scale(int n, double a, const double* d, const double* x, double* y, cudaStream_t st);   // y = a d.*x
int vec_dot_partial(int n, const double* x, const double* y, double* partial, int* nparts, cudaStream_t st);
int vec_sum_partials(const double* partial, int n, double* out, cudaStream_t st);   // out[0] = sum, one block
int vec_dot_host(int n, const double* x, const double* y, double* scratch, double* h_out, cudaStream_t st);
int extract_diag_inv(const Csr& A, double* dinv, cudaStream_t st);
int dense_gemv(int n, const double* M, const double* x, double* y, cudaStream_t st, int nb = 1);

}  // namespace sfem
