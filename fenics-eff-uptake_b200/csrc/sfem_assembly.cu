// Element / facet kernels, gather-style CSR assembly and Dirichlet application.
//
// Layout: per-cell geometry is SoA (geo[k][cell], k = x0,y0,x1,y1,x2,y2) so a warp reads six
// contiguous spans; element matrices are written AoS ([cell][i][j], row-major) so the gather
// kernel -- one thread per CSR slot, contributions in fixed (family, cell) order, no atomics --
// finds the 6 entries of a matrix row of one cell inside one 48-byte span.
//
// Element conventions: reference triangle (0,0),(1,0),(0,1), barycentrics
// l0=1-x-y, l1=x, l2=y; P2 dofs 0..2 at vertices (l_i(2l_i-1)), 3..5 on the edges opposite
// vertex 0,1,2 (4 l1 l2, 4 l0 l2, 4 l0 l1); affine cells with |det J| (cells are stored with
// ascending vertex ids so orientation may be negative).
#include "sfem_common.cuh"
#include "sfem_internal.h"

namespace sfem {

namespace {

struct TriGeom {
  double gx[3], gy[3];   // physical gradients of the barycentrics
  double adet;           // |det J| = 2 * area
};

__device__ __forceinline__ TriGeom load_geom(const double* __restrict__ geo, int nc, int c) {
  const double x0 = geo[0 * (size_t)nc + c], y0 = geo[1 * (size_t)nc + c];
  const double x1 = geo[2 * (size_t)nc + c], y1 = geo[3 * (size_t)nc + c];
  const double x2 = geo[4 * (size_t)nc + c], y2 = geo[5 * (size_t)nc + c];
  // explicit roundings: the compiler may not pick a different multiply-add contraction in different kernels, so every
  // kernel that inlines this helper sees bit-identical geometry
  const double det = fma(x1 - x0, y2 - y0, -__dmul_rn(x2 - x0, y1 - y0));
  const double inv = 1.0 / det;
  TriGeom g;
  g.gx[0] = (y1 - y2) * inv; g.gy[0] = (x2 - x1) * inv;
  g.gx[1] = (y2 - y0) * inv; g.gy[1] = (x0 - x2) * inv;
  g.gx[2] = (y0 - y1) * inv; g.gy[2] = (x1 - x0) * inv;
  g.adet = fabs(det);
  return g;
}

__device__ __forceinline__ void p2_basis(double l0, double l1, double l2, double* phi) {
  phi[0] = l0 * (2.0 * l0 - 1.0);
  phi[1] = l1 * (2.0 * l1 - 1.0);
  phi[2] = l2 * (2.0 * l2 - 1.0);
  phi[3] = 4.0 * l1 * l2;
  phi[4] = 4.0 * l0 * l2;
  phi[5] = 4.0 * l0 * l1;
}

// physical gradients of the six P2 basis functions at barycentric point (l0,l1,l2)
__device__ __forceinline__ void p2_grads(const TriGeom& g, double l0, double l1, double l2, double* Gx, double* Gy) {
  const double a0 = 4.0 * l0 - 1.0, a1 = 4.0 * l1 - 1.0, a2 = 4.0 * l2 - 1.0;
  Gx[0] = a0 * g.gx[0]; Gy[0] = a0 * g.gy[0];
  Gx[1] = a1 * g.gx[1]; Gy[1] = a1 * g.gy[1];
  Gx[2] = a2 * g.gx[2]; Gy[2] = a2 * g.gy[2];
  // (explicit fma / rounded products, see load_geom)
  Gx[3] = 4.0 * fma(l2, g.gx[1], __dmul_rn(l1, g.gx[2])); Gy[3] = 4.0 * fma(l2, g.gy[1], __dmul_rn(l1, g.gy[2]));
  Gx[4] = 4.0 * fma(l2, g.gx[0], __dmul_rn(l0, g.gx[2])); Gy[4] = 4.0 * fma(l2, g.gy[0], __dmul_rn(l0, g.gy[2]));
  Gx[5] = 4.0 * fma(l1, g.gx[0], __dmul_rn(l0, g.gx[1])); Gy[5] = 4.0 * fma(l1, g.gy[0], __dmul_rn(l0, g.gy[1]));
}

// degree-2 rule: 3 interior points, weights 1/6 (reference area 1/2)
__constant__ double kQ2[3][3] = {{2.0 / 3.0, 1.0 / 6.0, 1.0 / 6.0}, {1.0 / 6.0, 1.0 / 6.0, 2.0 / 3.0}, {1.0 / 6.0, 2.0 / 3.0, 1.0 / 6.0}};
// degree-5 rule: 7 points (Radon); barycentrics and weights (sum 1/2)
#define SFEM_S15 3.872983346207417
#define SFEM_A1 ((6.0 - SFEM_S15) / 21.0)
#define SFEM_A2 ((6.0 + SFEM_S15) / 21.0)
#define SFEM_W1 ((155.0 - SFEM_S15) / 2400.0)
#define SFEM_W2 ((155.0 + SFEM_S15) / 2400.0)
__constant__ double kQ5[7][4] = {
    {1.0 / 3.0, 1.0 / 3.0, 1.0 / 3.0, 9.0 / 80.0},
    {1.0 - 2.0 * SFEM_A1, SFEM_A1, SFEM_A1, SFEM_W1}, {SFEM_A1, SFEM_A1, 1.0 - 2.0 * SFEM_A1, SFEM_W1}, {SFEM_A1, 1.0 - 2.0 * SFEM_A1, SFEM_A1, SFEM_W1},
    {1.0 - 2.0 * SFEM_A2, SFEM_A2, SFEM_A2, SFEM_W2}, {SFEM_A2, SFEM_A2, 1.0 - 2.0 * SFEM_A2, SFEM_W2}, {SFEM_A2, 1.0 - 2.0 * SFEM_A2, SFEM_A2, SFEM_W2}};
// Gauss-Legendre on [0,1]: 4 points (exact to degree 7) and 3 points (exact to degree 5)
__constant__ double kGL4x[4] = {0.06943184420297371, 0.33000947820757187, 0.66999052179242813, 0.93056815579702629};
__constant__ double kGL4w[4] = {0.17392742256872692, 0.32607257743127308, 0.32607257743127308, 0.17392742256872692};

// ------------------------------------------------------------------ P2 advection-diffusion
// One thread per cell computes the 6 x 6 element matrix in registers; the 32 matrices of a warp are then staged through
// shared memory (row stride 37: conflict-free) so that every store instruction writes 32 CONSECUTIVE doubles of the AoS
// element buffer (thread-per-cell stores at a 288-byte stride ran at 1.5 TB/s, ncu r02_hot; k_elem_th was fixed the same way).
template <bool ADV>
__global__ void __launch_bounds__(128) k_elem_p2(int nc, const double* __restrict__ geo, const int* __restrict__ celldofs,
                                                 double D, const double* __restrict__ ux, const double* __restrict__ uy,
                                                 double* __restrict__ E) {
  __shared__ double stage[4][32 * 37];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int base = blockIdx.x * blockDim.x + warp * 32; base < nc; base += gridDim.x * blockDim.x) {
    const bool valid = base + lane < nc;
    const int c = valid ? base + lane : nc - 1;
    const TriGeom g = load_geom(geo, nc, c);
    double Ke[36];
#pragma unroll
    for (int k = 0; k < 36; ++k) Ke[k] = 0.0;
    const double wk = __ddiv_rn(__dmul_rn(D, g.adet), 6.0);
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      double Gx[6], Gy[6];
      p2_grads(g, kQ2[q][0], kQ2[q][1], kQ2[q][2], Gx, Gy);
#pragma unroll
      for (int i = 0; i < 6; ++i)
#pragma unroll
        for (int j = 0; j < 6; ++j) Ke[i * 6 + j] = fma(wk, fma(Gx[i], Gx[j], __dmul_rn(Gy[i], Gy[j])), Ke[i * 6 + j]);
    }
    if (ADV) {
      double vx[6], vy[6];
#pragma unroll
      for (int k = 0; k < 6; ++k) {
        const int d = celldofs[k * (size_t)nc + c];
        vx[k] = ux[d];
        vy[k] = uy[d];
      }
#pragma unroll 1
      for (int q = 0; q < 7; ++q) {
        const double l0 = kQ5[q][0], l1 = kQ5[q][1], l2 = kQ5[q][2], w = kQ5[q][3] * g.adet;
        double phi[6], Gx[6], Gy[6];
        p2_basis(l0, l1, l2, phi);
        p2_grads(g, l0, l1, l2, Gx, Gy);
        double uqx = 0.0, uqy = 0.0;
#pragma unroll
        for (int k = 0; k < 6; ++k) { uqx = fma(phi[k], vx[k], uqx); uqy = fma(phi[k], vy[k], uqy); }
#pragma unroll
        for (int j = 0; j < 6; ++j) {
          const double aj = w * fma(uqx, Gx[j], uqy * Gy[j]);
#pragma unroll
          for (int i = 0; i < 6; ++i) Ke[i * 6 + j] = fma(phi[i], aj, Ke[i * 6 + j]);
        }
      }
    }
    double* mine = stage[warp] + lane * 37;
#pragma unroll
    for (int k = 0; k < 36; ++k) mine[k] = Ke[k];
    __syncwarp();
    const int ncell = min(32, nc - base);
    double* out = E + (size_t)base * 36;
    for (int t = lane; t < ncell * 36; t += 32) out[t] = stage[warp][(t / 36) * 37 + (t % 36)];
    __syncwarp();
  }
}

// ------------------------------------------------------------------ P1 (multigrid coarse levels)
__global__ void __launch_bounds__(kThreads) k_elem_p1(int nc, const double* __restrict__ geo,
                                                      const int* __restrict__ cellverts, double D,
                                                      const double* __restrict__ ux, const double* __restrict__ uy,
                                                      int upwind, double* __restrict__ E) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const TriGeom g = load_geom(geo, nc, c);
    const double area = 0.5 * g.adet;
    double Dc = D, ucx = 0.0, ucy = 0.0;
    if (ux != nullptr) {
      for (int k = 0; k < 3; ++k) {
        const int v = cellverts[k * (size_t)nc + c];
        ucx += ux[v];
        ucy += uy[v];
      }
      ucx *= (1.0 / 3.0);
      ucy *= (1.0 / 3.0);
      if (upwind) {
        const double pe = sqrt(ucx * ucx + ucy * ucy) * sqrt(g.adet) / (2.0 * D);
        Dc = D * fmax(1.0, pe);
      }
    }
    double* out = E + (size_t)c * 9;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) {
        double v = Dc * area * (g.gx[i] * g.gx[j] + g.gy[i] * g.gy[j]);
        if (ux != nullptr) v += (area / 3.0) * (ucx * g.gx[j] + ucy * g.gy[j]);
        out[i * 3 + j] = v;
      }
  }
}

__global__ void __launch_bounds__(kThreads) k_elem_p1_mass(int nc, const double* __restrict__ geo, double* __restrict__ E) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
    const TriGeom g = load_geom(geo, nc, c);
    const double a = g.adet / 24.0;        // area/12
    double* out = E + (size_t)c * 9;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) out[i * 3 + j] = (i == j) ? 2.0 * a : a;
  }
}

// ------------------------------------------------------------------ Taylor-Hood Stokes
// a = grad u:grad v - div(v) p - q div(u); cell layout [ux x6, uy x6, p x3]; 15x15 row-major.
// One thread computes one cell, but a thread-per-cell store of 225 doubles at a stride of 1800 bytes makes every
// store instruction touch 32 different lines (measured: 0.62 TB/s).  The element matrices of a warp's 32 cells are
// therefore staged through shared memory in three row groups (rows 0-5, 6-11, 12-14) and written out cell by
// cell as contiguous 720 / 360-byte runs, 32 consecutive doubles per store instruction.
constexpr int kThWarps = 4;
constexpr int kThStride = 91;          // 90 staged doubles per cell + 1: odd stride, conflict-free column writes

__device__ __forceinline__ void th_copy_out(const double* __restrict__ ws, double* __restrict__ dst, int len, int ncell,
                                            int lane) {
  for (int cell = 0; cell < ncell; ++cell)
    for (int k = lane; k < len; k += 32) dst[(size_t)cell * 225 + k] = ws[cell * kThStride + k];
}

__global__ void __launch_bounds__(kThWarps * 32, 2) k_elem_th(int nc, const double* __restrict__ geo, double* __restrict__ E) {
  extern __shared__ double th_stage[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = th_stage + (size_t)warp * 32 * kThStride;
  double* mine = ws + lane * kThStride;
  for (long long base = ((long long)blockIdx.x * kThWarps + warp) * 32; base < nc; base += (long long)gridDim.x * kThWarps * 32) {
    const int c = (int)base + lane;
    const bool valid = c < nc;
    const int ncell = (nc - base) < 32 ? (int)(nc - base) : 32;
    double K[36], Bx[18], By[18];
#pragma unroll
    for (int k = 0; k < 36; ++k) K[k] = 0.0;
#pragma unroll
    for (int k = 0; k < 18; ++k) { Bx[k] = 0.0; By[k] = 0.0; }
    if (valid) {
      const TriGeom g = load_geom(geo, nc, c);
      const double wk = __ddiv_rn(g.adet, 6.0);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        double Gx[6], Gy[6];
        const double l[3] = {kQ2[q][0], kQ2[q][1], kQ2[q][2]};
        p2_grads(g, l[0], l[1], l[2], Gx, Gy);
#pragma unroll
        for (int i = 0; i < 6; ++i)
#pragma unroll
          for (int j = 0; j < 6; ++j) K[i * 6 + j] = fma(wk, fma(Gx[i], Gx[j], __dmul_rn(Gy[i], Gy[j])), K[i * 6 + j]);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            Bx[k * 6 + j] = fma(-__dmul_rn(wk, l[k]), Gx[j], Bx[k * 6 + j]);
            By[k * 6 + j] = fma(-__dmul_rn(wk, l[k]), Gy[j], By[k * 6 + j]);
          }
      }
    }
    double* out = E + (size_t)base * 225;
    // rows 0..5 (u_x test functions): [K row | 0 | Bx column]
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { mine[i * 15 + j] = K[i * 6 + j]; mine[i * 15 + 6 + j] = 0.0; }
#pragma unroll
      for (int k = 0; k < 3; ++k) mine[i * 15 + 12 + k] = Bx[k * 6 + i];
    }
    __syncwarp();
    th_copy_out(ws, out, 90, ncell, lane);
    __syncwarp();
    // rows 6..11 (u_y test functions): [0 | K row | By column]
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { mine[i * 15 + j] = 0.0; mine[i * 15 + 6 + j] = K[i * 6 + j]; }
#pragma unroll
      for (int k = 0; k < 3; ++k) mine[i * 15 + 12 + k] = By[k * 6 + i];
    }
    __syncwarp();
    th_copy_out(ws, out + 90, 90, ncell, lane);
    __syncwarp();
    // rows 12..14 (pressure test functions): [Bx row | By row | 0]
#pragma unroll
    for (int k = 0; k < 3; ++k) {
#pragma unroll
      for (int j = 0; j < 6; ++j) { mine[k * 15 + j] = Bx[k * 6 + j]; mine[k * 15 + 6 + j] = By[k * 6 + j]; }
#pragma unroll
      for (int m = 0; m < 3; ++m) mine[k * 15 + 12 + m] = 0.0;
    }
    __syncwarp();
    th_copy_out(ws, out + 180, 45, ncell, lane);
    __syncwarp();
  }
}

// Divergence blocks only: EB[c][k][comp * 6 + j] = -int l_k d(phi_j)/dx_comp  (3 x 12 per cell: the B rows of the
// cell's three pressure dofs against [u_x x6 | u_y x6]); the same numbers, in the same order of operations, as the
// Bx / By entries of k_elem_th.  With the P2 stiffness kernel (D = 1) for K this is everything the block-form Stokes
// solver needs -- 72 doubles per cell instead of 225.  Staged through shared memory like k_elem_th.
constexpr int kDivStride = 37;

__global__ void __launch_bounds__(kThWarps * 32) k_elem_th_div(int nc, const double* __restrict__ geo, double* __restrict__ EB) {
  __shared__ double stage[kThWarps * 32 * kDivStride];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double* ws = stage + warp * 32 * kDivStride;
  double* mine = ws + lane * kDivStride;
  for (long long base = ((long long)blockIdx.x * kThWarps + warp) * 32; base < nc; base += (long long)gridDim.x * kThWarps * 32) {
    const int c = (int)base + lane;
    const int ncell = (nc - base) < 32 ? (int)(nc - base) : 32;
    double Bx[18], By[18];
#pragma unroll
    for (int k = 0; k < 18; ++k) { Bx[k] = 0.0; By[k] = 0.0; }
    if (c < nc) {
      const TriGeom g = load_geom(geo, nc, c);
      const double wk = __ddiv_rn(g.adet, 6.0);
#pragma unroll
      for (int q = 0; q < 3; ++q) {
        double Gx[6], Gy[6];
        const double l[3] = {kQ2[q][0], kQ2[q][1], kQ2[q][2]};
        p2_grads(g, l[0], l[1], l[2], Gx, Gy);
#pragma unroll
        for (int k = 0; k < 3; ++k)
#pragma unroll
          for (int j = 0; j < 6; ++j) {
            Bx[k * 6 + j] = fma(-__dmul_rn(wk, l[k]), Gx[j], Bx[k * 6 + j]);
            By[k * 6 + j] = fma(-__dmul_rn(wk, l[k]), Gy[j], By[k * 6 + j]);
          }
      }
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int j = 0; j < 6; ++j) { mine[k * 12 + j] = Bx[k * 6 + j]; mine[k * 12 + 6 + j] = By[k * 6 + j]; }
    __syncwarp();
    double* out = EB + (size_t)base * 36;
    for (int cell = 0; cell < ncell; ++cell)
      for (int k = lane; k < 36; k += 32) out[(size_t)cell * 36 + k] = ws[cell * kDivStride + k];
    __syncwarp();
  }
}

// ------------------------------------------------------------------ Robin facets
// P2 trace basis on a facet parametrised by t in [0,1] from vertex a to vertex b: [a, b, mid]
template <int NDOF>
__global__ void __launch_bounds__(kThreads) k_facet_robin(int nf, const double* __restrict__ fgeo,
                                                          const int* __restrict__ fdofs, double mu_const,
                                                          const double* __restrict__ mu_nodal, int clamp,
                                                          double* __restrict__ F) {
  for (int f = blockIdx.x * blockDim.x + threadIdx.x; f < nf; f += gridDim.x * blockDim.x) {
    const double xa = fgeo[0 * (size_t)nf + f], ya = fgeo[1 * (size_t)nf + f];
    const double xb = fgeo[2 * (size_t)nf + f], yb = fgeo[3 * (size_t)nf + f];
    const double len = sqrt((xb - xa) * (xb - xa) + (yb - ya) * (yb - ya));
    double mun[NDOF];
    for (int k = 0; k < NDOF; ++k) mun[k] = mu_nodal ? mu_nodal[fdofs[k * (size_t)nf + f]] : mu_const;
    double M[NDOF * NDOF];
    for (int k = 0; k < NDOF * NDOF; ++k) M[k] = 0.0;
    for (int q = 0; q < 4; ++q) {
      const double t = kGL4x[q];
      double phi[NDOF];
      if (NDOF == 3) {
        phi[0] = (1.0 - t) * (1.0 - 2.0 * t);
        phi[1] = t * (2.0 * t - 1.0);
        phi[2] = 4.0 * t * (1.0 - t);
      } else {
        phi[0] = 1.0 - t;
        phi[1] = t;
      }
      double mu = 0.0;
      for (int k = 0; k < NDOF; ++k) mu = fma(phi[k], mun[k], mu);
      if (clamp && !(mu >= 0.0)) mu = 0.0;
      const double w = kGL4w[q] * len * mu;
      for (int i = 0; i < NDOF; ++i)
        for (int j = 0; j < NDOF; ++j) M[i * NDOF + j] = fma(w * phi[i], phi[j], M[i * NDOF + j]);
    }
    double* out = F + (size_t)f * (NDOF * NDOF);
    for (int k = 0; k < NDOF * NDOF; ++k) out[k] = M[k];
  }
}

// vals[k] = 0 for every entry whose row or column is flagged (Dirichlet elimination of the rectangular Stokes
// blocks: bc columns of B, bc rows of B^T)
__global__ void __launch_bounds__(kThreads) k_csr_zero_flagged(int nrows, const int* __restrict__ rowptr, const int* __restrict__ cols,
                                                               double* __restrict__ vals, const unsigned char* __restrict__ row_flag,
                                                               const unsigned char* __restrict__ col_flag) {
  const int lane = threadIdx.x & 3;
  for (long long row = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 2; row < nrows;
       row += ((long long)gridDim.x * blockDim.x) >> 2) {
    const bool rz = row_flag != nullptr && row_flag[row];
    for (int k = rowptr[row] + lane; k < rowptr[row + 1]; k += 4)
      if (rz || (col_flag != nullptr && col_flag[cols[k]])) vals[k] = 0.0;
  }
}

// ------------------------------------------------------------------ gather + Dirichlet
__global__ void __launch_bounds__(kThreads) k_gather(int nnz, const int* __restrict__ cptr, const int* __restrict__ code,
                                                     const double* __restrict__ E, double* __restrict__ vals) {
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nnz; s += gridDim.x * blockDim.x) {
    double acc = 0.0;
    const int e = cptr[s + 1];
    for (int k = cptr[s]; k < e; ++k) acc += E[code[k]];
    vals[s] = acc;
  }
}

// dst[i] = src[slot[i]]  (copy a sub-block of an assembled CSR into its own CSR, e.g. K / B / B^T of
// the Taylor-Hood matrix)
__global__ void __launch_bounds__(kThreads) k_extract(int n, const int* __restrict__ slot, const double* __restrict__ src,
                                                      double* __restrict__ dst) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) dst[i] = src[slot[i]];
}

// blocked [c][n] <-> interleaved [n][c] layouts of a two-component nodal field
__global__ void k_interleave2(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    reinterpret_cast<double2*>(out)[i] = make_double2(a[i], b[i]);
}
__global__ void k_deinterleave2(int n, const double* __restrict__ in, double* __restrict__ a, double* __restrict__ b) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const double2 v = reinterpret_cast<const double2*>(in)[i];
    a[i] = v.x;
    b[i] = v.y;
  }
}

template <int LANES>
__global__ void __launch_bounds__(kThreads) k_dirichlet(int n, const int* __restrict__ rowptr, const int* __restrict__ cols,
                                                        double* __restrict__ vals, double* __restrict__ rhs,
                                                        const unsigned char* __restrict__ flag,
                                                        const double* __restrict__ g, int mode) {
  constexpr int ROWS = kThreads / LANES;
  const int lane = threadIdx.x % LANES;
  const int sub = threadIdx.x / LANES;
  for (long long base = (long long)blockIdx.x * ROWS; base < n; base += (long long)gridDim.x * ROWS) {
    const int row = (int)base + sub;
    const bool valid = row < n;
    double lift = 0.0;
    bool isbc = false;
    if (valid) {
      isbc = flag[row] != 0;
      const int s = rowptr[row], e = rowptr[row + 1];
      for (int k = s + lane; k < e; k += LANES) {
        const int cj = cols[k];
        if (isbc) {
          vals[k] = (cj == row) ? 1.0 : 0.0;
        } else if (mode == 1 && flag[cj]) {
          lift = fma(vals[k], g[cj], lift);
          vals[k] = 0.0;
        }
      }
    }
#pragma unroll
    for (int o = LANES >> 1; o > 0; o >>= 1) lift += __shfl_xor_sync(0xffffffffu, lift, o);
    if (valid && lane == 0) {
      if (isbc) rhs[row] = g[row];
      else if (mode == 1) rhs[row] -= lift;
    }
  }
}

}  // namespace

}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_elem_p2_advdiff(int nc, const double* geo, const int* celldofs, double D, const double* ux,
                         const double* uy, double* E, void* stream) {
  if (nc <= 0) return SFEM_OK;
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(nc, 128, 16);
  Prof prof(PC_ELEM, (double)nc * (48.0 + 288.0 + ((ux && uy) ? 24.0 + 96.0 : 0.0)), st);
  if (ux != nullptr && uy != nullptr) {
    if (celldofs == nullptr) { set_error("celldofs required with a velocity field"); return SFEM_ERR_ARG; }
    k_elem_p2<true><<<grid, 128, 0, st>>>(nc, geo, celldofs, D, ux, uy, E);
  } else {
    k_elem_p2<false><<<grid, 128, 0, st>>>(nc, geo, celldofs, D, ux, uy, E);
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_elem_p1_advdiff(int nc, const double* geo, const int* cellverts, double D, const double* ux,
                         const double* uy, int upwind, double* E, void* stream) {
  if (nc <= 0) return SFEM_OK;
  if ((ux != nullptr) != (uy != nullptr)) { set_error("ux/uy must both be given"); return SFEM_ERR_ARG; }
  k_elem_p1<<<grid_for(nc, kThreads), kThreads, 0, (cudaStream_t)stream>>>(nc, geo, cellverts, D, ux, uy, upwind, E);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_elem_th_stokes(int nc, const double* geo, double* E, void* stream) {
  if (nc <= 0) return SFEM_OK;
  Prof prof(PC_ELEM, (double)nc * (48.0 + 1800.0), (cudaStream_t)stream);
  constexpr int smem = kThWarps * 32 * kThStride * (int)sizeof(double);
  static thread_local bool configured = false;
  if (!configured) {
    SFEM_CUDA(cudaFuncSetAttribute(k_elem_th, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  k_elem_th<<<grid_for(nc, kThWarps * 32, 2), kThWarps * 32, smem, (cudaStream_t)stream>>>(nc, geo, E);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_elem_th_div(int nc, const double* geo, double* EB, void* stream) {
  if (nc <= 0) return SFEM_OK;
  Prof prof(PC_ELEM, (double)nc * (48.0 + 288.0), (cudaStream_t)stream);
  k_elem_th_div<<<grid_for(nc, kThWarps * 32, 8), kThWarps * 32, 0, (cudaStream_t)stream>>>(nc, geo, EB);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_csr_zero_flagged(int nrows, const int* rowptr, const int* cols, double* vals, const unsigned char* row_flag,
                          const unsigned char* col_flag, void* stream) {
  if (nrows <= 0 || (row_flag == nullptr && col_flag == nullptr)) return SFEM_OK;
  sell_mark_dirty(vals);
  k_csr_zero_flagged<<<grid_for(nrows, kThreads / 4), kThreads, 0, (cudaStream_t)stream>>>(nrows, rowptr, cols, vals, row_flag, col_flag);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_elem_p1_mass(int nc, const double* geo, double* E, void* stream) {
  if (nc <= 0) return SFEM_OK;
  k_elem_p1_mass<<<grid_for(nc, kThreads), kThreads, 0, (cudaStream_t)stream>>>(nc, geo, E);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_facet_p2_robin(int nf, const double* fgeo, const int* fdofs, double mu_const, const double* mu_nodal,
                        int clamp, double* F, void* stream) {
  if (nf <= 0) return SFEM_OK;
  k_facet_robin<3><<<grid_for(nf, kThreads), kThreads, 0, (cudaStream_t)stream>>>(nf, fgeo, fdofs, mu_const, mu_nodal, clamp, F);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_facet_p1_robin(int nf, const double* fgeo, const int* fdofs, double mu_const, const double* mu_nodal,
                        int clamp, double* F, void* stream) {
  if (nf <= 0) return SFEM_OK;
  k_facet_robin<2><<<grid_for(nf, kThreads), kThreads, 0, (cudaStream_t)stream>>>(nf, fgeo, fdofs, mu_const, mu_nodal, clamp, F);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_gather_csr(int nnz, const int* contrib_ptr, const int* contrib_code, const double* E, double* vals,
                    void* stream) {
  if (nnz <= 0) return SFEM_OK;
  sell_mark_dirty(vals);
  Prof prof(PC_GATHER, (double)nnz * 12.0, (cudaStream_t)stream);   // + 12 B per contribution, unknown here
  k_gather<<<grid_for(nnz, kThreads, 16), kThreads, 0, (cudaStream_t)stream>>>(nnz, contrib_ptr, contrib_code, E, vals);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_csr_extract(int n, const int* slot, const double* src_vals, double* dst_vals, void* stream) {
  if (n <= 0) return SFEM_OK;
  sell_mark_dirty(dst_vals);
  Prof prof(PC_GATHER, (double)n * 20.0, (cudaStream_t)stream);
  k_extract<<<grid_for(n, kThreads * 2, 16), kThreads, 0, (cudaStream_t)stream>>>(n, slot, src_vals, dst_vals);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_vec_interleave2(int n, const double* a, const double* b, double* out, void* stream) {
  if (n <= 0) return SFEM_OK;
  k_interleave2<<<grid_for(n, kThreads * 2), kThreads, 0, (cudaStream_t)stream>>>(n, a, b, out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_vec_deinterleave2(int n, const double* in, double* a, double* b, void* stream) {
  if (n <= 0) return SFEM_OK;
  k_deinterleave2<<<grid_for(n, kThreads * 2), kThreads, 0, (cudaStream_t)stream>>>(n, in, a, b);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_apply_dirichlet(int n, int nnz, const int* rowptr, const int* cols, double* vals, double* rhs,
                         const unsigned char* bc_flag, const double* bc_val, int mode, void* stream) {
  if (n <= 0) return SFEM_OK;
  if (mode != 0 && mode != 1) { set_error("dirichlet mode must be 0 or 1"); return SFEM_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  sell_mark_dirty(vals);
  const int lanes = pick_lanes(nnz, n);
  switch (lanes) {
    case 1: case 2:
      k_dirichlet<2><<<grid_for(n, kThreads / 2), kThreads, 0, st>>>(n, rowptr, cols, vals, rhs, bc_flag, bc_val, mode); break;
    case 4:
      k_dirichlet<4><<<grid_for(n, kThreads / 4), kThreads, 0, st>>>(n, rowptr, cols, vals, rhs, bc_flag, bc_val, mode); break;
    case 8:
      k_dirichlet<8><<<grid_for(n, kThreads / 8), kThreads, 0, st>>>(n, rowptr, cols, vals, rhs, bc_flag, bc_val, mode); break;
    default:
      k_dirichlet<16><<<grid_for(n, kThreads / 16), kThreads, 0, st>>>(n, rowptr, cols, vals, rhs, bc_flag, bc_val, mode); break;
  }
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
