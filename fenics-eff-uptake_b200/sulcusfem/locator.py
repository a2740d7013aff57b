"""Uniform-grid cell locator (host plan) + device point evaluation of nodal fields.

Replaces dolfin's ``BoundingBoxTree`` walk behind ``mesh.bounding_box_tree().compute_first_entity_collision(p)``
and ``c(Point)`` / ``u(Point)`` in the reference's profile extractors (``analysis.py:341-419, 544-632``) and
velocity metrics (``analysis.py:721-830``): all sample points of a call are located and evaluated in ONE launch of
``sfem_eval_points`` (csrc/sfem_points.cu).  The plan -- for every bin of a uniform grid over the mesh bounding
box the ascending list of cells whose (slightly inflated) bounding box overlaps it -- is built once per mesh.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import capi
from . import dofmap as dm
from .hostmesh import HostMesh

TOL = 1e-12          # inclusion tolerance on the barycentric coordinates (closed cells, like dolfin's collision test)


@dataclass
class BinGrid:
    nbx: int
    nby: int
    x0: float
    y0: float
    hx: float
    hy: float
    bin_ptr: np.ndarray      # int32 [nbx * nby + 1]
    bin_cells: np.ndarray    # int32, ascending inside every bin


def build_bins(mesh: HostMesh, cells_per_bin: float = 2.0) -> BinGrid:
    """Bins of edge ~ sqrt(cells_per_bin * mean cell area * 2): a handful of candidate cells per bin."""
    X = mesh.coords
    lo, hi = X.min(axis=0), X.max(axis=0)
    span = np.maximum(hi - lo, 1e-300)
    area = float(np.abs(mesh.signed_areas()).mean())
    edge = float(np.sqrt(2.0 * area * cells_per_bin))
    nbx = int(max(1, min(4096, np.ceil(span[0] / edge))))
    nby = int(max(1, min(4096, np.ceil(span[1] / edge))))
    hx, hy = float(span[0] / nbx), float(span[1] / nby)
    P = X[mesh.cells.astype(np.int64)]                       # [nc,3,2]
    pad = 1e-9 * float(span.max())
    cmin, cmax = P.min(axis=1) - pad, P.max(axis=1) + pad
    ix0 = np.clip(np.floor((cmin[:, 0] - lo[0]) / hx).astype(np.int64), 0, nbx - 1)
    ix1 = np.clip(np.floor((cmax[:, 0] - lo[0]) / hx).astype(np.int64), 0, nbx - 1)
    iy0 = np.clip(np.floor((cmin[:, 1] - lo[1]) / hy).astype(np.int64), 0, nby - 1)
    iy1 = np.clip(np.floor((cmax[:, 1] - lo[1]) / hy).astype(np.int64), 0, nby - 1)
    wx, wy = ix1 - ix0 + 1, iy1 - iy0 + 1
    cnt = wx * wy
    nc = mesh.num_cells
    cell = np.repeat(np.arange(nc, dtype=np.int64), cnt)
    off = np.arange(int(cnt.sum()), dtype=np.int64) - np.repeat(np.cumsum(cnt) - cnt, cnt)
    w = np.repeat(wx, cnt)
    bx = np.repeat(ix0, cnt) + off % w
    by = np.repeat(iy0, cnt) + off // w
    b = by * nbx + bx
    order = np.lexsort((cell, b))                            # by bin, cells ascending inside a bin
    b, cell = b[order], cell[order]
    bin_ptr = np.zeros(nbx * nby + 1, dtype=np.int64)
    np.add.at(bin_ptr, b + 1, 1)
    np.cumsum(bin_ptr, out=bin_ptr)
    return BinGrid(nbx, nby, float(lo[0]), float(lo[1]), hx, hy, bin_ptr.astype(np.int32), cell.astype(np.int32))


class DeviceLocator:
    """Bin grid + geometry + dof maps of one mesh in HBM; ``eval`` runs ``sfem_eval_points``."""

    def __init__(self, mesh: HostMesh, ctx=None):
        from .device import Context, cell_geometry
        self.ctx = ctx or Context.get()
        self.mesh = mesh
        g = build_bins(mesh)
        self.grid = g
        up = self.ctx.up
        self.bin_ptr, self.bin_cells = up(g.bin_ptr, np.int32), up(g.bin_cells, np.int32)
        self.geo = up(cell_geometry(mesh), np.float64)
        self.nc = mesh.num_cells
        self._dofs = {}

    def _celldofs(self, degree):
        if degree not in self._dofs:
            cd = dm.p2_cell_dofs(self.mesh) if degree == 2 else dm.p1_cell_dofs(self.mesh)
            self._dofs[degree] = self.ctx.up(np.ascontiguousarray(cd.T), np.int32)
        return self._dofs[degree]

    def eval(self, pts: np.ndarray, fields: Sequence, degree: int = 2, tol: float = TOL):
        """fields: device tensors of nodal values (<= 4).  Returns (values [nfields, npts] numpy, cell [npts] numpy,
        -1 = outside the mesh)."""
        import torch
        ctx, g = self.ctx, self.grid
        pts = np.ascontiguousarray(np.asarray(pts, dtype=np.float64).reshape(-1, 2))
        n, nf = len(pts), len(fields)
        dp = ctx.up(pts.ravel(), np.float64)
        out = ctx.empty(max(nf * n, 1))
        cell = torch.empty(max(n, 1), dtype=torch.int32, device=ctx.device)
        arr = (C.c_void_p * max(nf, 1))(*[f.data_ptr() for f in fields])
        capi.check(ctx.lib.sfem_eval_points(int(degree), n, capi.ptr(dp), g.nbx, g.nby, g.x0, g.y0, g.hx, g.hy,
                                            capi.ptr(self.bin_ptr), capi.ptr(self.bin_cells), capi.ptr(self.geo), self.nc,
                                            capi.ptr(self._celldofs(degree)), nf, arr, float(tol), capi.ptr(out),
                                            capi.ptr(cell), ctx.stream), 'sfem_eval_points')
        return out[:nf * n].cpu().numpy().reshape(nf, n), cell[:n].cpu().numpy()


def locator_for(mesh: HostMesh) -> DeviceLocator:
    loc = getattr(mesh, '_device_locator', None)
    if loc is None:
        loc = DeviceLocator(mesh)
        mesh._device_locator = loc
    return loc
