// Flux / concentration / mass functionals (reference analysis.py) as reductions.
//
// Facet functionals: the host groups facet entries (cell whose trace is used, local facet) by
// measure -- ds_bc(1..4), ds_bottom(5,6,7), ds_y0(10), dS_y0(10) with the channel-side cell -- and
// one launch evaluates every group: one CTA per group, threads stride over the group's facets, a
// fixed-order block reduction produces the eight integrals of the group.  All polynomial integrands
// use the 3-point Gauss-Legendre rule (exact to degree 5 >= the degree FFC would pick); the
// non-smooth ones (|q|, q+, q-) use the same 3 points FFC uses for their estimated degree 4.
#include "sfem_common.cuh"
#include "sfem_internal.h"

namespace sfem {

namespace {

__constant__ double kGL3x[3] = {0.11270166537925831, 0.5, 0.88729833462074169};
__constant__ double kGL3w[3] = {0.27777777777777779, 0.44444444444444442, 0.27777777777777779};

__device__ __forceinline__ void p2_eval(const double* l, double* phi) {
  phi[0] = l[0] * (2.0 * l[0] - 1.0);
  phi[1] = l[1] * (2.0 * l[1] - 1.0);
  phi[2] = l[2] * (2.0 * l[2] - 1.0);
  phi[3] = 4.0 * l[1] * l[2];
  phi[4] = 4.0 * l[0] * l[2];
  phi[5] = 4.0 * l[0] * l[1];
}

__global__ void __launch_bounds__(128) k_facet_functionals(const int* __restrict__ grp_ptr, const int* __restrict__ ent_cell,
                                                           const int* __restrict__ ent_local, const double* __restrict__ geo,
                                                           const int* __restrict__ celldofs, int nc,
                                                           const double* __restrict__ cvec, const double* __restrict__ ux,
                                                           const double* __restrict__ uy, double D, double mu_const,
                                                           const double* __restrict__ mu_nodal, double* __restrict__ out) {
  __shared__ double sh[33];
  const int g = blockIdx.x;
  const int e0 = grp_ptr[g], e1 = grp_ptr[g + 1];
  double acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.0;
  for (int e = e0 + threadIdx.x; e < e1; e += blockDim.x) {
    const int c = ent_cell[e];
    const int lf = ent_local[e];
    const double x0 = geo[0 * (size_t)nc + c], y0 = geo[1 * (size_t)nc + c];
    const double x1 = geo[2 * (size_t)nc + c], y1 = geo[3 * (size_t)nc + c];
    const double x2 = geo[4 * (size_t)nc + c], y2 = geo[5 * (size_t)nc + c];
    const double det = (x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0);
    const double inv = 1.0 / det;
    double gx[3], gy[3];
    gx[0] = (y1 - y2) * inv; gy[0] = (x2 - x1) * inv;
    gx[1] = (y2 - y0) * inv; gy[1] = (x0 - x2) * inv;
    gx[2] = (y0 - y1) * inv; gy[2] = (x1 - x0) * inv;
    // facet lf is opposite local vertex lf; its end points are the other two vertices (a < b)
    const int a = (lf == 0) ? 1 : 0;
    const int b = (lf == 2) ? 1 : 2;
    const double px[3] = {x0, x1, x2}, py[3] = {y0, y1, y2};
    const double len = sqrt((px[b] - px[a]) * (px[b] - px[a]) + (py[b] - py[a]) * (py[b] - py[a]));
    const double gn = sqrt(gx[lf] * gx[lf] + gy[lf] * gy[lf]);
    const double nx = -gx[lf] / gn, ny = -gy[lf] / gn;      // outward normal of this cell
    double cl[6], vx[6], vy[6], ml[6];
    for (int k = 0; k < 6; ++k) {
      const int d = celldofs[k * (size_t)nc + c];
      cl[k] = cvec[d];
      vx[k] = ux ? ux[d] : 0.0;
      vy[k] = uy ? uy[d] : 0.0;
      ml[k] = mu_nodal ? mu_nodal[d] : mu_const;
    }
    for (int q = 0; q < 3; ++q) {
      const double t = kGL3x[q], w = kGL3w[q] * len;
      double l[3] = {0.0, 0.0, 0.0};
      l[a] = 1.0 - t;
      l[b] = t;
      double phi[6];
      p2_eval(l, phi);
      double cq = 0.0, uq = 0.0, vq = 0.0, mq = 0.0;
      for (int k = 0; k < 6; ++k) {
        cq = fma(phi[k], cl[k], cq);
        uq = fma(phi[k], vx[k], uq);
        vq = fma(phi[k], vy[k], vq);
        mq = fma(phi[k], ml[k], mq);
      }
      // grad c = sum_k c_k grad phi_k
      const double a0 = 4.0 * l[0] - 1.0, a1 = 4.0 * l[1] - 1.0, a2 = 4.0 * l[2] - 1.0;
      double dcx = cl[0] * a0 * gx[0] + cl[1] * a1 * gx[1] + cl[2] * a2 * gx[2];
      double dcy = cl[0] * a0 * gy[0] + cl[1] * a1 * gy[1] + cl[2] * a2 * gy[2];
      dcx += 4.0 * (cl[3] * (l[2] * gx[1] + l[1] * gx[2]) + cl[4] * (l[2] * gx[0] + l[0] * gx[2]) + cl[5] * (l[1] * gx[0] + l[0] * gx[1]));
      dcy += 4.0 * (cl[3] * (l[2] * gy[1] + l[1] * gy[2]) + cl[4] * (l[2] * gy[0] + l[0] * gy[2]) + cl[5] * (l[1] * gy[0] + l[0] * gy[1]));
      const double qd = -D * (dcx * nx + dcy * ny);
      const double qa = (uq * nx + vq * ny) * cq;
      const double qq = qd + qa;
      acc[0] = fma(w, qd, acc[0]);
      acc[1] = fma(w, qa, acc[1]);
      acc[2] = fma(w, mq * cq, acc[2]);
      acc[3] = fma(w, cq, acc[3]);
      acc[4] += w;
      acc[5] = fma(w, fabs(qq), acc[5]);
      acc[6] = fma(w, (qq >= 0.0) ? qq : 0.0, acc[6]);
      acc[7] = fma(w, (qq <= 0.0) ? -qq : 0.0, acc[7]);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const double t = block_sum(acc[k], sh);
    if (threadIdx.x == 0) out[(size_t)g * 8 + k] = t;
  }
}

// per block partials: partial[block][nmarkers][2]
__global__ void __launch_bounds__(kThreads) k_cell_functionals(int nc, const double* __restrict__ geo,
                                                               const int* __restrict__ celldofs,
                                                               const int* __restrict__ marker, int nmarkers,
                                                               const double* __restrict__ cvec, double* __restrict__ partial) {
  __shared__ double sh[33];
  for (int m = 0; m < nmarkers; ++m) {
    double mass = 0.0, area = 0.0;
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < nc; c += gridDim.x * blockDim.x) {
      if (marker != nullptr && marker[c] != m) continue;
      const double x0 = geo[0 * (size_t)nc + c], y0 = geo[1 * (size_t)nc + c];
      const double x1 = geo[2 * (size_t)nc + c], y1 = geo[3 * (size_t)nc + c];
      const double x2 = geo[4 * (size_t)nc + c], y2 = geo[5 * (size_t)nc + c];
      const double adet = fabs((x1 - x0) * (y2 - y0) - (x2 - x0) * (y1 - y0));
      // int phi_vertex = 0, int phi_edge = area/3  (P2, exact)
      const double ce = cvec[celldofs[3 * (size_t)nc + c]] + cvec[celldofs[4 * (size_t)nc + c]] + cvec[celldofs[5 * (size_t)nc + c]];
      mass = fma(adet * (1.0 / 6.0), ce, mass);
      area = fma(0.5, adet, area);
    }
    const double tm = block_sum(mass, sh);
    const double ta = block_sum(area, sh);
    if (threadIdx.x == 0) {
      partial[((size_t)blockIdx.x * nmarkers + m) * 2 + 0] = tm;
      partial[((size_t)blockIdx.x * nmarkers + m) * 2 + 1] = ta;
    }
  }
}

__global__ void k_cell_functionals_final(const double* __restrict__ partial, int nblocks, int nvals, double* __restrict__ out) {
  __shared__ double sh[33];
  for (int v = 0; v < nvals; ++v) {
    double t = 0.0;
    for (int b = threadIdx.x; b < nblocks; b += blockDim.x) t += partial[(size_t)b * nvals + v];
    t = block_sum(t, sh);
    if (threadIdx.x == 0) out[v] = t;
  }
}

struct Scratch {
  double* ptr = nullptr;
  size_t cap = 0;
};
thread_local Scratch t_scratch;

}  // namespace

}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_facet_functionals(int ngroups, const int* grp_ptr, const int* ent_cell, const int* ent_local,
                           const double* geo, const int* celldofs, int nc, const double* c, const double* ux,
                           const double* uy, double D, double mu_const, const double* mu_nodal, double* out,
                           void* stream) {
  if (ngroups <= 0) return SFEM_OK;
  if ((ux != nullptr) != (uy != nullptr)) { set_error("ux/uy must both be given"); return SFEM_ERR_ARG; }
  k_facet_functionals<<<ngroups, 128, 0, (cudaStream_t)stream>>>(grp_ptr, ent_cell, ent_local, geo, celldofs, nc, c, ux,
                                                                uy, D, mu_const, mu_nodal, out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

int sfem_cell_functionals(int nc, const double* geo, const int* celldofs, const int* cell_marker, int nmarkers,
                          const double* c, double* out, void* stream) {
  if (nc <= 0 || nmarkers <= 0 || nmarkers > 16) { set_error("cell functionals: bad arguments"); return SFEM_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(nc, kThreads * 2, 4);
  const size_t need = (size_t)grid * nmarkers * 2;
  if (need > t_scratch.cap) {
    if (t_scratch.ptr) cudaFree(t_scratch.ptr);
    t_scratch.ptr = nullptr; t_scratch.cap = 0;
    SFEM_CUDA(cudaMalloc(&t_scratch.ptr, need * sizeof(double)));
    t_scratch.cap = need;
  }
  k_cell_functionals<<<grid, kThreads, 0, st>>>(nc, geo, celldofs, cell_marker, nmarkers, c, t_scratch.ptr);
  SFEM_LAUNCH_CHECK();
  k_cell_functionals_final<<<1, kThreads, 0, st>>>(t_scratch.ptr, grid, nmarkers * 2, out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
