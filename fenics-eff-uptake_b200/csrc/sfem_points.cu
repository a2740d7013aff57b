// Point location + Lagrange evaluation at arbitrary points (reference analysis.py:341-632, 721-830: the
// `mesh.bounding_box_tree().compute_first_entity_collision(Point)` / `c(Point)` / `u(Point)` loops behind the line
// profiles and the velocity metrics).
//
// The reference walks dolfin's bounding-box tree once per point from Python.  Here all sample points of a call go
// through one launch: a uniform bin grid over the mesh bounding box (host-built once per mesh: for every bin the
// ascending list of cells whose bounding box overlaps it) turns location into a short candidate scan.  One thread
// per point: bin -> candidates -> barycentric coordinates; the containing cell is the LOWEST-numbered candidate
// with min(lambda) >= -tol (deterministic where dolfin returns "the first collision" of its tree order; on shared
// edges and vertices every incident cell gives the same value of a continuous function up to rounding).  Up to
// four nodal fields are evaluated with the same basis values (c; u_x, u_y; ...).  Points in no cell get cell = -1
// and value 0 -- the reference skips them (`valid_points`).
#include "sfem_common.cuh"
#include "sfem_internal.h"

namespace sfem {

namespace {

constexpr int kMaxPointFields = 4;

struct PointFields {
  const double* f[kMaxPointFields];
};

template <int DEGREE>
__global__ void __launch_bounds__(kThreads)
    k_eval_points(int npts, const double* __restrict__ pts, int nbx, int nby, double x0, double y0, double inv_hx,
                  double inv_hy, const int* __restrict__ bin_ptr, const int* __restrict__ bin_cells,
                  const double* __restrict__ geo, int nc, const int* __restrict__ celldofs, int nfields, PointFields F,
                  double tol, double* __restrict__ out, int* __restrict__ cell_out) {
  const int stride = gridDim.x * blockDim.x;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < npts; i += stride) {
    const double px = pts[2 * (size_t)i], py = pts[2 * (size_t)i + 1];
    int found = -1;
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
    // bin of the point; points a hair outside the grid are clamped into the border bins (the inclusion test decides)
    const double fx = (px - x0) * inv_hx, fy = (py - y0) * inv_hy;
    if (fx >= -1.0 && fy >= -1.0 && fx <= (double)nbx + 1.0 && fy <= (double)nby + 1.0) {
      int bx = (int)floor(fx), by = (int)floor(fy);
      bx = bx < 0 ? 0 : (bx >= nbx ? nbx - 1 : bx);
      by = by < 0 ? 0 : (by >= nby ? nby - 1 : by);
      const int b = by * nbx + bx;
      for (int k = bin_ptr[b]; k < bin_ptr[b + 1]; ++k) {
        const int c = bin_cells[k];
        const double ax = geo[0 * (size_t)nc + c], ay = geo[1 * (size_t)nc + c];
        const double bxv = geo[2 * (size_t)nc + c], byv = geo[3 * (size_t)nc + c];
        const double cx = geo[4 * (size_t)nc + c], cy = geo[5 * (size_t)nc + c];
        const double det = (bxv - ax) * (cy - ay) - (cx - ax) * (byv - ay);
        const double m1 = ((px - ax) * (cy - ay) - (cx - ax) * (py - ay)) / det;
        const double m2 = ((bxv - ax) * (py - ay) - (px - ax) * (byv - ay)) / det;
        const double m0 = 1.0 - m1 - m2;
        if (fmin(m0, fmin(m1, m2)) >= -tol) {
          found = c; l0 = m0; l1 = m1; l2 = m2;
          break;                                    // candidates are ascending: lowest-numbered containing cell
        }
      }
    }
    cell_out[i] = found;
    double phi[DEGREE == 2 ? 6 : 3];
    if (DEGREE == 2) {
      phi[0] = l0 * (2.0 * l0 - 1.0); phi[1] = l1 * (2.0 * l1 - 1.0); phi[2] = l2 * (2.0 * l2 - 1.0);
      phi[3] = 4.0 * l1 * l2; phi[4] = 4.0 * l0 * l2; phi[5] = 4.0 * l0 * l1;
    } else {
      phi[0] = l0; phi[1] = l1; phi[2] = l2;
    }
    constexpr int ND = DEGREE == 2 ? 6 : 3;
    int dof[ND];
#pragma unroll
    for (int k = 0; k < ND; ++k) dof[k] = found >= 0 ? celldofs[k * (size_t)nc + found] : 0;
    for (int f = 0; f < nfields; ++f) {
      double v = 0.0;
      if (found >= 0) {
#pragma unroll
        for (int k = 0; k < ND; ++k) v = fma(phi[k], F.f[f][dof[k]], v);     // fixed order: bit-reproducible
      }
      out[(size_t)f * npts + i] = v;
    }
  }
}

}  // namespace

}  // namespace sfem

using namespace sfem;

extern "C" {

int sfem_eval_points(int degree, int npts, const double* pts, int nbx, int nby, double x0, double y0, double hx, double hy,
                     const int* bin_ptr, const int* bin_cells, const double* geo, int nc, const int* celldofs,
                     int nfields, const double* const* h_fields, double tol, double* out, int* cell_out, void* stream) {
  if ((degree != 1 && degree != 2) || npts < 0 || nbx < 1 || nby < 1 || !(hx > 0.0) || !(hy > 0.0) || nc < 1 ||
      nfields < 0 || nfields > kMaxPointFields || !bin_ptr || !bin_cells || !geo || !celldofs || !cell_out ||
      (nfields > 0 && (!h_fields || !out)) || !(tol >= 0.0)) {
    set_error("eval_points: bad arguments");
    return SFEM_ERR_ARG;
  }
  if (npts == 0) return SFEM_OK;
  PointFields F;
  for (int f = 0; f < kMaxPointFields; ++f) F.f[f] = f < nfields ? h_fields[f] : nullptr;
  for (int f = 0; f < nfields; ++f)
    if (!F.f[f]) { set_error("eval_points: null field"); return SFEM_ERR_ARG; }
  cudaStream_t st = (cudaStream_t)stream;
  const int grid = grid_for(npts, kThreads);
  Prof prof(PC_OTHER, (double)npts * (16.0 + 8.0 * nfields + 4.0), st);
  if (degree == 2)
    k_eval_points<2><<<grid, kThreads, 0, st>>>(npts, pts, nbx, nby, x0, y0, 1.0 / hx, 1.0 / hy, bin_ptr, bin_cells, geo, nc,
                                                celldofs, nfields, F, tol, out, cell_out);
  else
    k_eval_points<1><<<grid, kThreads, 0, st>>>(npts, pts, nbx, nby, x0, y0, 1.0 / hx, 1.0 / hy, bin_ptr, bin_cells, geo, nc,
                                                celldofs, nfields, F, tol, out, cell_out);
  SFEM_LAUNCH_CHECK();
  return SFEM_OK;
}

}  // extern "C"
