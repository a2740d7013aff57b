"""Mesh hierarchy + inter-level transfer operators for the multigrid preconditioner (host, once).

The reference solves every system with a sparse direct LU (``solve(a == L, ...)`` with dolfin
defaults, reference ``solvers.py:55,84,151,213,298``).  A direct factorisation has no B200-native
analogue worth building; the device path instead runs Krylov iterations preconditioned by a
geometric multigrid V-cycle to LU-level residuals.  This module prepares the *integer / geometric*
side of that preconditioner on the host:

level 0            the P2 space of the fine mesh (the system being solved)
level 1            the P1 space of the same mesh (p-coarsening; P2 nodes of mesh l = P1 nodes of
                   its uniform refinement, so the transfer is the edge-midpoint average)
level 2, 3, ...    P1 spaces of coarser meshes: the refinement parents when the mesh was produced
                   by ``refine`` (nested, topological transfer), then independently generated
                   coarser synthetic meshes of the same geometry (non-nested, transfer = P1
                   interpolation at the finer vertices through point location)

Transfers are plain CSR matrices (prolongation ``P`` and its explicit transpose ``R``) so the
device applies them with the same CSR SpMV kernel as the operators.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from .hostmesh import HostMesh, rectangle_mesh, sulcus_mesh, build_markers
from . import dofmap as dm


@dataclass
class Transfer:
    n_fine: int
    n_coarse: int
    rowptr: np.ndarray      # P: [n_fine+1]
    cols: np.ndarray
    vals: np.ndarray
    t_rowptr: np.ndarray    # R = P^T: [n_coarse+1]
    t_cols: np.ndarray
    t_vals: np.ndarray
    t_perm: np.ndarray = None   # t_vals = vals[t_perm]
    nested: bool = False        # coarse dofs are a prefix of the fine dofs (midpoint transfer)


def _finish(n_fine, n_coarse, rows, cols, vals, nested=False, presorted=False) -> Transfer:
    if not presorted:                                      # (rows ascending, columns ascending inside a row)
        order = np.lexsort((cols, rows))
        rows, cols, vals = rows[order], cols[order], vals[order]
    rowptr = np.concatenate([[0], np.cumsum(np.bincount(rows, minlength=n_fine))]).astype(np.int32)
    cols = cols.astype(np.int32)
    t_rowptr, t_cols, perm = dm.transpose_csr(n_fine, n_coarse, rowptr, cols)
    return Transfer(n_fine, n_coarse, rowptr, cols, vals.astype(np.float64),
                    t_rowptr, t_cols, vals[perm].astype(np.float64), perm, nested)


def midpoint_transfer(mesh: HostMesh) -> Transfer:
    """P1(mesh) -> {P2(mesh) or P1(refine(mesh))}: identity on vertices, average on edge nodes."""
    nv, ne = mesh.num_vertices, mesh.num_edges
    rows = np.concatenate([np.arange(nv), nv + np.repeat(np.arange(ne), 2)]).astype(np.int64)
    cols = np.concatenate([np.arange(nv), mesh.edges.astype(np.int64).ravel()])
    vals = np.concatenate([np.ones(nv), np.full(2 * ne, 0.5)])
    # rows come out in order and an edge lists its vertices as (low, high): already sorted, no lexsort of 3 nv entries
    presorted = bool(ne == 0 or np.all(mesh.edges[:, 0] < mesh.edges[:, 1]))
    return _finish(nv + ne, nv, rows, cols, vals, nested=True, presorted=presorted)


def locate_points(mesh: HostMesh, pts: np.ndarray, k: int = 16):
    """For each point the best containing (or nearest) cell and its barycentric coordinates."""
    from scipy.spatial import cKDTree
    tree = cKDTree(mesh.cell_midpoints())
    k = min(k, mesh.num_cells)
    _, cand = tree.query(pts, k=k)
    cand = cand.reshape(len(pts), k)
    p = mesh.coords[mesh.cells[cand]]                              # [n,k,3,2]
    x0, x1, x2 = p[:, :, 0], p[:, :, 1], p[:, :, 2]
    det = (x1[..., 0] - x0[..., 0]) * (x2[..., 1] - x0[..., 1]) - (x2[..., 0] - x0[..., 0]) * (x1[..., 1] - x0[..., 1])
    q = pts[:, None, :]
    l1 = ((q[..., 0] - x0[..., 0]) * (x2[..., 1] - x0[..., 1]) - (x2[..., 0] - x0[..., 0]) * (q[..., 1] - x0[..., 1])) / det
    l2 = ((x1[..., 0] - x0[..., 0]) * (q[..., 1] - x0[..., 1]) - (q[..., 0] - x0[..., 0]) * (x1[..., 1] - x0[..., 1])) / det
    l0 = 1.0 - l1 - l2
    lam = np.stack([l0, l1, l2], axis=-1)                          # [n,k,3]
    score = lam.min(axis=-1)
    best = np.argmax(score, axis=1)
    ar = np.arange(len(pts))
    return cand[ar, best], lam[ar, best]


def interpolation_transfer(fine: HostMesh, coarse: HostMesh) -> Transfer:
    """P1(coarse) -> P1(fine) by evaluating the coarse hat functions at the fine vertices."""
    cell, lam = locate_points(coarse, fine.coords)
    lam = np.clip(lam, 0.0, None)                                  # points a hair outside: project
    lam /= lam.sum(axis=1, keepdims=True)
    rows = np.repeat(np.arange(fine.num_vertices, dtype=np.int64), 3)
    cols = coarse.cells[cell].astype(np.int64).ravel()
    vals = lam.ravel()
    keep = vals > 1e-14
    # merge duplicates is unnecessary (3 distinct vertices per cell)
    return _finish(fine.num_vertices, coarse.num_vertices, rows[keep], cols[keep], vals[keep])


def _mesh_h(mesh: HostMesh) -> float:
    return float(np.sqrt(2.0 * np.abs(mesh.signed_areas()).mean()))


def coarser_synthetic(mesh: HostMesh, h: float) -> Optional[HostMesh]:
    g = mesh.geometry
    if str(g.get('mesher', '')).startswith('delaunay'):
        from .unstructured import mesh_domain
        try:
            # a graded fine mesh (reference refinement_factor > 1) is coarsened with the same grading: every level
            # doubles the spacing everywhere, near the sulcus too
            return mesh_domain(g['L'], g['H'], g.get('w', 0.5), g.get('d', 1.0), h, g['domain_type'],
                               refinement_factor=g.get('refinement_factor', 1))
        except RuntimeError:
            pass
    if g.get('domain_type') == 'sulcus':
        return sulcus_mesh(g['L'], g['H'], g['w'], g['d'], h)
    if g.get('domain_type') == 'rectangular':
        nx = max(2, int(round(g['L'] / h)))
        ny = max(2, int(round(g['H'] / h)))
        return rectangle_mesh(g['L'], g['H'], nx, ny)
    return None


@dataclass
class Hierarchy:
    meshes: List[HostMesh]          # meshes[0] = fine mesh, then coarser
    transfers: List[Transfer]       # transfers[0]: P1(mesh0)->P2(mesh0); transfers[l]: P1(mesh_l)->P1(mesh_{l-1})


def build_hierarchy(mesh: HostMesh, coarsest_vertices: int = 700, max_levels: int = 12) -> Hierarchy:
    meshes = [mesh]
    transfers = [midpoint_transfer(mesh)]
    m = mesh
    while m.parent is not None and len(meshes) < max_levels:
        transfers.append(midpoint_transfer(m.parent))
        m = m.parent
        meshes.append(m)
    h = float(m.geometry['h']) if m.geometry.get('mesher') == 'delaunay-graded' else _mesh_h(m)
    H = m.geometry.get('H', None)
    while (m.num_vertices > coarsest_vertices and len(meshes) < max_levels
           and m.geometry.get('domain_type') in ('sulcus', 'rectangular')):
        h = 2.0 * h
        if H is not None and h > 0.51 * H:
            break
        c = coarser_synthetic(m, h)
        if c is None or c.num_vertices >= 0.6 * m.num_vertices:
            break
        transfers.append(interpolation_transfer(m, c))
        meshes.append(c)
        m = c
    return Hierarchy(meshes, transfers)


def level_markers(mesh: HostMesh) -> dict:
    """bc markers of a hierarchy mesh, from its recorded geometry."""
    g = mesh.geometry
    if g.get('domain_type') == 'sulcus':
        return build_markers(mesh, g['L'], g['H'], g['xL'], g['xR'], 'sulcus')
    return build_markers(mesh, g['L'], g['H'], g['L'] / 2, g['L'] / 2, 'rectangular')
