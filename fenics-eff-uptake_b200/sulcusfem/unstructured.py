"""Isotropic unstructured mesher for the sulcus / rectangular domains (host, numpy + Qhull).

Stands in for the reference's ``gmsh -2 -algo del2d -smooth 1`` call (reference ``mesh.py:353``) on
the geometry its ``.geo`` file describes (``mesh.py:263-348``): channel ``[0,L]x[0,H]``, cavity floor
through ``y=-d sin(pi x_rel)`` (``mesh.py:147-154``), mouth line embedded in the surface
(``mesh.py:310-311``) so that mesh edges lie on y=0 across the mouth.

Method: boundary and mouth line sampled at spacing ~h; interior points on a hexagonal lattice kept
away from the sampled lines (so every sampled segment satisfies the Gabriel condition and is
therefore a Delaunay edge); Delaunay triangulation (scipy / Qhull); triangles outside the domain
dropped; a few Laplacian smoothing + re-triangulation passes (the analogue of ``-smooth 1``).
The result is validated: every prescribed segment must be an edge and the boundary of the
triangulation must be exactly the prescribed boundary.  Mesh *generation* is outside the hot path
(north_star: "Host code stays Python: it reads the Gmsh/XML mesh").
"""
from __future__ import annotations

import math

import numpy as np
from scipy.spatial import Delaunay, cKDTree

from .hostmesh import HostMesh, sulcus_floor


def _sample_segment(p, q, h, include_end=False):
    n = max(1, int(round(np.linalg.norm(np.subtract(q, p)) / h)))
    t = np.linspace(0.0, 1.0, n + 1)
    if not include_end:
        t = t[:-1]
    out = np.outer(1 - t, p) + np.outer(t, q)
    for k in (0, 1):                                   # axis-aligned walls: exact coordinate (marker predicates use ~1e-16 tolerances)
        if p[k] == q[k]:
            out[:, k] = p[k]
    return out


def _sample_floor(xL, w, d, h):
    """Points on the cavity curve from (xL,0) to (xR,0), ~equal arc length, ends included."""
    s = np.linspace(0.0, 1.0, 4001)
    x = xL + s * w
    y = sulcus_floor(x, xL, w, d)
    y[0] = y[-1] = 0.0
    seg = np.hypot(np.diff(x), np.diff(y))
    arc = np.concatenate([[0.0], np.cumsum(seg)])
    n = max(4, int(round(arc[-1] / h)))
    if n % 2:
        n += 1                                     # keep the tip (x_rel=0.5) a sample
    target = np.linspace(0.0, arc[-1], n + 1)
    sx = np.interp(target, arc, s)
    px = xL + sx * w
    py = sulcus_floor(px, xL, w, d)
    px[0], px[-1] = xL, xL + w
    py[0] = py[-1] = 0.0
    return np.stack([px, py], axis=1)


def _inside_polygon(poly, pts):
    x, y = pts[:, 0], pts[:, 1]
    inside = np.zeros(len(pts), dtype=bool)
    n = len(poly)
    for i in range(n):
        x0, y0 = poly[i]
        x1, y1 = poly[(i + 1) % n]
        if y0 == y1:
            continue
        cond = ((y0 > y) != (y1 > y))
        xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
        inside ^= cond & (x < xi)
    return inside


def _triangulate(points, poly, h_min):
    tri = Delaunay(points)
    t = tri.simplices
    p = points[t]
    area = 0.5 * np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                        - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
    cen = p.mean(axis=1)
    keep = (area > 1e-10 * h_min * h_min) & _inside_polygon(poly, cen)
    return t[keep]


def _edge_set(cells):
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [0, 2]]], axis=0)
    e = np.sort(e, axis=1)
    key = e[:, 0].astype(np.int64) * (cells.max() + 1) + e[:, 1]
    uk, cnt = np.unique(key, return_counts=True)
    return uk, cnt, cells.max() + 1


def _validate(cells, n_points, boundary_segs, interior_segs):
    if len(np.unique(cells)) != n_points:
        return False
    uk, cnt, base = _edge_set(cells)
    bkey = np.sort(boundary_segs, axis=1)
    bkey = bkey[:, 0].astype(np.int64) * base + bkey[:, 1]
    if not np.array_equal(np.sort(uk[cnt == 1]), np.sort(bkey)):
        return False
    if len(interior_segs):
        ikey = np.sort(interior_segs, axis=1)
        ikey = ikey[:, 0].astype(np.int64) * base + ikey[:, 1]
        pos = np.searchsorted(uk, ikey)
        if np.any(pos >= len(uk)) or not np.array_equal(uk[np.minimum(pos, len(uk) - 1)], ikey):
            return False
    return True


def _smooth(points, cells, fixed_mask):
    n = len(points)
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [0, 2]]], axis=0)
    e = np.unique(np.sort(e, axis=1), axis=0)
    acc = np.zeros((n, 2))
    deg = np.zeros(n)
    np.add.at(acc, e[:, 0], points[e[:, 1]])
    np.add.at(acc, e[:, 1], points[e[:, 0]])
    np.add.at(deg, e[:, 0], 1)
    np.add.at(deg, e[:, 1], 1)
    new = points.copy()
    free = ~fixed_mask & (deg > 0)
    new[free] = acc[free] / deg[free, None]
    return new


def mesh_domain(L=10.0, H=1.0, w=0.5, d=1.0, h=0.02, domain_type='sulcus', smooth_passes=2) -> HostMesh:
    """Unstructured isotropic mesh of the sulcus or rectangular domain."""
    xL, xR = L / 2 - w / 2, L / 2 + w / 2
    if domain_type == 'sulcus':
        h_cav = min(h, w / 3.0)
        floor = _sample_floor(xL, w, d, h_cav)
        bpts = np.concatenate([
            _sample_segment((0.0, 0.0), (xL, 0.0), h),
            floor[:-1],
            _sample_segment((xR, 0.0), (L, 0.0), h),
            _sample_segment((L, 0.0), (L, H), h),
            _sample_segment((L, H), (0.0, H), h),
            _sample_segment((0.0, H), (0.0, 0.0), h)], axis=0)
        mouth_inner = _sample_segment((xL, 0.0), (xR, 0.0), h_cav)[1:]
    else:
        bpts = np.concatenate([
            _sample_segment((0.0, 0.0), (L, 0.0), h),
            _sample_segment((L, 0.0), (L, H), h),
            _sample_segment((L, H), (0.0, H), h),
            _sample_segment((0.0, H), (0.0, 0.0), h)], axis=0)
        mouth_inner = np.zeros((0, 2))
        h_cav = h
    nb = len(bpts)
    bsegs = np.stack([np.arange(nb), (np.arange(nb) + 1) % nb], axis=1)
    isegs = np.zeros((0, 2), dtype=np.int64)
    if len(mouth_inner):
        # mouth chain: corner(xL,0) -> inner points -> corner(xR,0); corners are boundary points
        iL = int(np.argmin(np.hypot(bpts[:, 0] - xL, bpts[:, 1])))
        iR = int(np.argmin(np.hypot(bpts[:, 0] - xR, bpts[:, 1])))
        ids = np.concatenate([[iL], nb + np.arange(len(mouth_inner)), [iR]])
        isegs = np.stack([ids[:-1], ids[1:]], axis=1)
    fixed = np.concatenate([bpts, mouth_inner], axis=0)
    # hexagonal lattice over the bounding box
    y0 = -d if domain_type == 'sulcus' else 0.0
    dy = h * math.sqrt(3.0) / 2.0
    ny = int(math.ceil((H - y0) / dy)) + 1
    nx = int(math.ceil(L / h)) + 2
    jj, ii = np.meshgrid(np.arange(ny), np.arange(nx), indexing='ij')
    lat = np.stack([(ii + 0.5 * (jj % 2)) * h - 0.25 * h, y0 + jj * dy + 0.37 * dy], axis=-1).reshape(-1, 2)
    if domain_type == 'sulcus' and h_cav < h:
        # extra finer lattice inside / around a sub-h cavity
        dyc = h_cav * math.sqrt(3.0) / 2.0
        nyc = int(math.ceil((d + 2 * h) / dyc)) + 1
        nxc = int(math.ceil((w + 4 * h) / h_cav)) + 2
        jj, ii = np.meshgrid(np.arange(nyc), np.arange(nxc), indexing='ij')
        fine = np.stack([xL - 2 * h + (ii + 0.5 * (jj % 2)) * h_cav, -d + jj * dyc + 0.37 * dyc], axis=-1).reshape(-1, 2)
        near = (fine[:, 1] < 1.5 * h)
        lat = np.concatenate([lat[~((lat[:, 1] < 1.5 * h) & (lat[:, 0] > xL - 2 * h) & (lat[:, 0] < xR + 2 * h))], fine[near]])
    poly = bpts
    lat = lat[_inside_polygon(poly, lat)]
    tree = cKDTree(fixed)
    dist, idx = tree.query(lat)
    # local spacing of the nearest fixed sample: cavity samples use h_cav
    loc_h = np.where((fixed[idx, 1] <= 0.0) & (fixed[idx, 0] >= xL) & (fixed[idx, 0] <= xR), h_cav, h) \
        if domain_type == 'sulcus' else np.full(len(lat), h)
    lat = lat[dist >= 0.8 * loc_h]
    points = np.concatenate([fixed, lat], axis=0)
    fixed_mask = np.zeros(len(points), dtype=bool)
    fixed_mask[:len(fixed)] = True
    cells = _triangulate(points, poly, h_cav)
    if not _validate(cells, len(points), bsegs, isegs):
        raise RuntimeError("unstructured mesher: boundary/mouth recovery failed")
    for _ in range(smooth_passes):
        trial = _smooth(points, cells, fixed_mask)
        tcells = _triangulate(trial, poly, h_cav)
        if _validate(tcells, len(trial), bsegs, isegs):
            points, cells = trial, tcells
        else:
            break
    geo = {'domain_type': domain_type, 'L': float(L), 'H': float(H), 'h': float(h), 'mesher': 'delaunay'}
    if domain_type == 'sulcus':
        geo.update({'w': float(w), 'd': float(d), 'xL': float(xL), 'xR': float(xR)})
    return HostMesh(points, cells, geo).check()
