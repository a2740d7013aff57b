"""Coarse pressure correction of the Stokes Schur-complement preconditioner (host set-up, once per mesh).

The reference solves the Taylor-Hood system with a sparse LU (``solvers.py:298``).  The device path
runs MINRES preconditioned by ``diag(MG(K), MG(K), S^-1)``; with the classical choice ``S = Mp``
(pressure mass matrix) the iteration count on the 10:1 channel is governed by the O((H/L)^2) small
eigenvalues of ``Mp^-1 B K^-1 B^T`` -- the long-wave pressure modes p(x), for which the Schur
complement acts like the 1-D lubrication (Reynolds) operator  ``-d/dx( H(x)^3/12 dp/dx )``.

This module builds, from the mesh alone,

    S^-1 = Mp^-1 + Z C Z^T,        C = psd( E^-1 - G^-1 )

* ``Z``  1-D hat functions in x (spacing ~ channel height) evaluated at the pressure nodes,
* ``E``  the lubrication stiffness on those hats with the local gap H(x) measured from the boundary
         facets (ids 3 = top, 4 = bottom), pinned at the outflow end (natural outflow fixes p there),
* ``G``  = Z^T Mp Z, so that on span(Z) the pair (Mp^-1 + Z C Z^T) acts like E^-1,
* ``psd``  clips negative eigenvalues, which keeps the preconditioner symmetric positive definite
           whatever the quality of the lubrication estimate (MINRES needs SPD).

It only changes the iteration count (about -40% on the default sulcus channel), never the solution.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

from .hostmesh import HostMesh
from . import dofmap as dm


@dataclass
class SchurCorrection:
    nz: int
    zidx: np.ndarray        # [nv] left hat index of every pressure node
    zw: np.ndarray          # [nv] weight of the left hat (right hat: 1 - zw)
    zt_rowptr: np.ndarray   # CSR of Z^T (nz rows)
    zt_cols: np.ndarray
    zt_vals: np.ndarray
    C: np.ndarray           # [nz, nz] symmetric positive semi-definite


def _boundary_profile(mesh: HostMesh, markers: np.ndarray, marker_id: int, xs: np.ndarray, reduce) -> Optional[np.ndarray]:
    f = np.flatnonzero((markers == marker_id) & mesh.edge_on_boundary)
    if len(f) == 0:
        return None
    v = np.unique(mesh.edges[f].ravel())
    p = mesh.coords[v]
    order = np.argsort(p[:, 0], kind='stable')
    x, y = p[order, 0], p[order, 1]
    # several boundary vertices may share an x (vertical walls): keep the extreme one
    ux, inv = np.unique(x, return_inverse=True)
    uy = np.full(len(ux), -np.inf if reduce is np.maximum else np.inf)
    reduce.at(uy, inv, y)
    return np.interp(xs, ux, uy)


def lubrication_correction(mesh: HostMesh, bc_markers: np.ndarray, spacing: Optional[float] = None) -> Optional[SchurCorrection]:
    X = mesh.coords
    x0, x1 = float(X[:, 0].min()), float(X[:, 0].max())
    L = x1 - x0
    xs = np.linspace(x0, x1, 4001)
    xm = 0.5 * (xs[1:] + xs[:-1])
    top = _boundary_profile(mesh, bc_markers, 3, xm, np.maximum)
    bot = _boundary_profile(mesh, bc_markers, 4, xm, np.minimum)
    if top is None or bot is None:
        return None
    gap = np.maximum(top - bot, 1e-12)
    Href = float(np.median(gap))
    hz = spacing if spacing is not None else Href
    nz = int(round(L / hz)) + 1
    if nz < 3:
        return None                                   # short domain: the mass matrix alone is fine
    nz = min(nz, 256)
    zn = np.linspace(x0, x1, nz)
    hz = zn[1] - zn[0]
    # hats at the pressure nodes
    t = (X[:, 0] - x0) / hz
    zidx = np.clip(np.floor(t).astype(np.int64), 0, nz - 2)
    zw = 1.0 - (t - zidx)
    nv = mesh.num_vertices
    # G = Z^T Mp Z with the P1 mass matrix (area/12 * (1 + delta_ij)) accumulated cell by cell
    import scipy.sparse as sp
    area = np.abs(mesh.signed_areas())
    c = mesh.cells.astype(np.int64)
    Me = (area / 12.0)[:, None, None] * (np.ones((3, 3)) + np.eye(3))[None]
    Mp = sp.coo_matrix((Me.ravel(), (np.repeat(c[:, :, None], 3, 2).ravel(), np.repeat(c[:, None, :], 3, 1).ravel())),
                       shape=(nv, nv)).tocsr()
    ar = np.arange(nv)
    Z = sp.coo_matrix((np.concatenate([zw, 1.0 - zw]), (np.concatenate([ar, ar]), np.concatenate([zidx, zidx + 1]))),
                      shape=(nv, nz)).tocsr()
    G = (Z.T @ (Mp @ Z)).toarray()
    # lubrication stiffness  E_ij = int gap^3/12 z_i' z_j' dx  (piecewise-constant hat slopes)
    dx = xs[1] - xs[0]
    seg = np.clip(np.floor((xm - x0) / hz).astype(np.int64), 0, nz - 2)
    kseg = np.bincount(seg, weights=gap ** 3 / 12.0 * dx, minlength=nz - 1) / hz ** 2
    E = np.zeros((nz, nz))
    i = np.arange(nz - 1)
    E[i, i] += kseg
    E[i + 1, i + 1] += kseg
    E[i, i + 1] -= kseg
    E[i + 1, i] -= kseg
    E[-1, -1] += 1e3 * np.abs(E).max()                # outflow end: pressure level fixed by the natural condition
    Cm = np.linalg.inv(E) - np.linalg.inv(G)
    Cm = 0.5 * (Cm + Cm.T)
    ev, V = np.linalg.eigh(Cm)
    Cm = (V * np.maximum(ev, 0.0)) @ V.T
    Cm = 0.5 * (Cm + Cm.T)
    Zt = Z.T.tocsr()
    Zt.sort_indices()
    return SchurCorrection(nz, zidx.astype(np.int32), zw.astype(np.float64), Zt.indptr.astype(np.int32),
                           Zt.indices.astype(np.int32), Zt.data.astype(np.float64), Cm)
