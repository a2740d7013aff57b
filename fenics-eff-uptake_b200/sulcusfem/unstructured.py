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


def _inside_channel_domain(poly, pts, box):
    """``_inside_polygon`` for the channel polygons of this module, same answers, ~20x cheaper: above the channel floor
    (y > 0) the polygon is the rectangle (0, L) x (0, H) -- the ray cast only ever meets the two side walls there -- and
    below it (y < 0) only the span of the sulcus can be inside; the edge loop runs on those few points (and on y == 0)."""
    L, H, xL, xR = box
    x, y = pts[:, 0], pts[:, 1]
    inside = (y > 0.0) & (y < H) & ((x < L) != (x < 0.0))
    rest = np.flatnonzero((y == 0.0) | ((y < 0.0) & (x >= xL) & (x <= xR)))
    if len(rest):
        inside[rest] = _inside_polygon(poly, pts[rest])
    return inside


def _triangulate(points, poly, h_min, box=None):
    tri = Delaunay(points)
    t = tri.simplices
    p = points[t]
    area = 0.5 * np.abs((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                        - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
    cen = p.mean(axis=1)
    keep = (area > 1e-10 * h_min * h_min) & (_inside_polygon(poly, cen) if box is None else _inside_channel_domain(poly, cen, box))
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
    e = np.concatenate([cells[:, [0, 1]], cells[:, [1, 2]], cells[:, [0, 2]]], axis=0).astype(np.int64)
    e.sort(axis=1)
    key = np.unique(e[:, 0] * n + e[:, 1])                     # unique edges in lexicographic (v0, v1) order
    e0, e1 = key // n, key % n
    # neighbour sums: first-endpoint contributions, then second-endpoint contributions, each in edge order -- one
    # sequential bincount over the concatenated lists adds in exactly the order the np.add.at version did (same bits)
    idx, nbr = np.concatenate([e0, e1]), np.concatenate([e1, e0])
    acc = np.stack([np.bincount(idx, weights=points[nbr, c], minlength=n) for c in range(2)], axis=1)
    deg = np.bincount(idx, minlength=n).astype(np.float64)
    new = points.copy()
    free = ~fixed_mask & (deg > 0)
    new[free] = acc[free] / deg[free, None]
    return new


# ------------------------------------------------------------------------------------ graded meshes
def threshold_size_field(pts, nodes, lc, lc_fine, dist_min, dist_max):
    """Gmsh ``Field[1] = Distance`` (to the node list) + ``Field[2] = Threshold`` exactly as the reference's ``.geo``
    file sets them (``mesh.py:330-337``): ``LcMin`` within ``DistMin`` of the nearest sulcus node, ``LcMax`` beyond
    ``DistMax``, linear in between."""
    dist, _ = cKDTree(np.asarray(nodes)).query(np.asarray(pts))
    t = np.clip((dist - dist_min) / max(dist_max - dist_min, 1e-300), 0.0, 1.0)
    return lc_fine + t * (lc - lc_fine)


def reference_sulcus_nodes(L, w, d, n_segments=20):
    """The 21 floor samples the reference lists in ``Field[1].NodesList`` (``mesh.py:139-155``), for BOTH domain types
    (on the rectangle they lie below the floor: the "imaginary sulcus")."""
    x_rel = np.arange(n_segments + 1) / n_segments
    x = (L / 2 - w / 2) + x_rel * w
    y = -d * np.sin(np.pi * x_rel)
    y[0] = y[-1] = 0.0
    return np.stack([x, y], axis=1)


def _dyadic_level(size, h_f, lmax):
    """Largest lattice level whose spacing h_f 2^l does not exceed the requested size."""
    return np.clip(np.floor(np.log2(np.maximum(size, h_f) / h_f + 1e-9)).astype(np.int64), 0, lmax)


def _sample_polyline_graded(dense, size_fn):
    """Points along the dense polyline ``dense`` ([M, 2], end points included) spaced by ``size_fn``: the metric length
    int ds / size is split into an integer number of equal parts.  Returns the samples WITHOUT the end point."""
    seg = np.hypot(*np.diff(dense, axis=0).T)
    mid = 0.5 * (dense[1:] + dense[:-1])
    m = np.concatenate([[0.0], np.cumsum(seg / size_fn(mid))])
    n = max(1, int(round(m[-1])))
    tgt = np.linspace(0.0, m[-1], n + 1)[:-1]
    arc = np.concatenate([[0.0], np.cumsum(seg)])
    s = np.interp(tgt, m, arc)
    out = np.stack([np.interp(s, arc, dense[:, 0]), np.interp(s, arc, dense[:, 1])], axis=1)
    out[0] = dense[0]
    return out


def _mesh_domain_graded(L, H, w, d, h, domain_type, refinement_factor, smooth_passes):
    """Locally graded mesh: the reference's ``refinement_factor`` (``lc_fine = lc / refinement_factor`` within w/10 of
    the sulcus nodes, ``lc`` beyond w/2, ``mesh.py:266,330-337``).  Interior points come from ONE hexagonal lattice of
    spacing lc_fine and its nested dyadic sub-lattices (spacing lc_fine 2^l): a lattice point of level l is kept where
    the size field asks for a spacing >= lc_fine 2^l' with l' <= l, so the spacing follows the Threshold field in
    octaves and is never coarser than it."""
    xL, xR = L / 2 - w / 2, L / 2 + w / 2
    rf = float(refinement_factor)
    h_f = h / rf
    lmax = int(math.floor(math.log2(rf) + 1e-9))
    nodes = reference_sulcus_nodes(L, w, d)
    h_cav = min(h, w / 3.0)

    def size(p):
        s = threshold_size_field(p, nodes, h, h_f, w / 10.0, w / 2.0)
        if domain_type == 'sulcus' and h_cav < h:          # narrow cavities: at least three cells across (as ungraded)
            near = (p[:, 1] < 1.5 * h) & (p[:, 0] > xL - 2 * h) & (p[:, 0] < xR + 2 * h)
            s = np.where(near, np.minimum(s, h_cav), s)
        return s

    def qsize(p):
        return h_f * 2.0 ** _dyadic_level(size(p), h_f, lmax)

    def line(p, q):
        n = max(2, int(math.ceil(np.linalg.norm(np.subtract(q, p)) / (0.05 * h_f))))
        t = np.linspace(0.0, 1.0, n + 1)
        out = np.outer(1 - t, p) + np.outer(t, q)
        for k in (0, 1):
            if p[k] == q[k]:
                out[:, k] = p[k]
        return out

    def sample(p, q):
        out = _sample_polyline_graded(line(p, q), qsize)
        for k in (0, 1):                                   # axis-aligned walls: exact coordinate
            if p[k] == q[k]:
                out[:, k] = p[k]
        return out
    if domain_type == 'sulcus':
        sx = np.linspace(0.0, 1.0, 8001)
        fx = xL + sx * w
        fy = sulcus_floor(fx, xL, w, d)
        fy[0] = fy[-1] = 0.0
        floor = _sample_polyline_graded(np.stack([fx, fy], axis=1), qsize)
        floor[:, 1] = sulcus_floor(floor[:, 0], xL, w, d)
        floor[0] = (xL, 0.0)
        bpts = np.concatenate([sample((0.0, 0.0), (xL, 0.0)), floor, sample((xR, 0.0), (L, 0.0)),
                               sample((L, 0.0), (L, H)), sample((L, H), (0.0, H)), sample((0.0, H), (0.0, 0.0))], axis=0)
        mouth_inner = sample((xL, 0.0), (xR, 0.0))[1:]
    else:
        bpts = np.concatenate([sample((0.0, 0.0), (L, 0.0)), sample((L, 0.0), (L, H)), sample((L, H), (0.0, H)),
                               sample((0.0, H), (0.0, 0.0))], axis=0)
        mouth_inner = np.zeros((0, 2))
    nb = len(bpts)
    bsegs = np.stack([np.arange(nb), (np.arange(nb) + 1) % nb], axis=1)
    isegs = np.zeros((0, 2), dtype=np.int64)
    if len(mouth_inner):
        iL = int(np.argmin(np.hypot(bpts[:, 0] - xL, bpts[:, 1])))
        iR = int(np.argmin(np.hypot(bpts[:, 0] - xR, bpts[:, 1])))
        ids = np.concatenate([[iL], nb + np.arange(len(mouth_inner)), [iR]])
        isegs = np.stack([ids[:-1], ids[1:]], axis=1)
    fixed = np.concatenate([bpts, mouth_inner], axis=0)
    # nested hexagonal lattices: fine integer coordinates (X in units of h_f / 2, Y in units of dy)
    y0 = -d if domain_type == 'sulcus' else 0.0
    dy = h_f * math.sqrt(3.0) / 2.0
    ny = int(math.ceil((H - y0) / dy)) + 2
    nx = int(math.ceil(L / h_f)) + 3
    # keep only the rows / columns a level can use before materialising the finest lattice everywhere: build per level
    pts, lvl = [], []
    for l in range(lmax, -1, -1):
        step = 1 << l
        jj = np.arange(0, ny, step)
        ii = np.arange(0, 2 * nx, step)
        Y, X = np.meshgrid(jj, ii, indexing='ij')
        on = (((X >> l) - (Y >> l)) % 2 == 0)
        if l < lmax:                                       # points of the coarser lattices were emitted already
            up = step << 1
            on &= ~((X % up == 0) & (Y % up == 0) & ((((X >> (l + 1)) - (Y >> (l + 1))) % 2) == 0))
        P = np.stack([X[on] * (0.5 * h_f) - 0.25 * h_f, y0 + Y[on] * dy + 0.37 * dy], axis=1)
        P = P[(P[:, 0] > 0) & (P[:, 0] < L) & (P[:, 1] < H)]
        if l < lmax:                                       # a finer point is only needed where the field asks for it
            P = P[_dyadic_level(size(P), h_f, lmax) <= l]
        pts.append(P)
        lvl.append(np.full(len(P), l))
    lat = np.concatenate(pts, axis=0)
    poly = bpts
    box = (float(L), float(H), float(xL), float(xR)) if domain_type == 'sulcus' else (float(L), float(H), 1.0, 0.0)
    lat = lat[_inside_channel_domain(poly, lat, box)]
    dist, _ = cKDTree(fixed).query(lat)
    lat = lat[dist >= 0.8 * qsize(lat)]
    points = np.concatenate([fixed, lat], axis=0)
    fixed_mask = np.zeros(len(points), dtype=bool)
    fixed_mask[:len(fixed)] = True
    h_min = min(h_f, h_cav)
    cells = _triangulate(points, poly, h_min, box)
    if not _validate(cells, len(points), bsegs, isegs):
        raise RuntimeError("graded mesher: boundary/mouth recovery failed")
    for _ in range(smooth_passes):
        trial = _smooth(points, cells, fixed_mask)
        tcells = _triangulate(trial, poly, h_min, box)
        if _validate(tcells, len(trial), bsegs, isegs):
            points, cells = trial, tcells
        else:
            break
    geo = {'domain_type': domain_type, 'L': float(L), 'H': float(H), 'h': float(h), 'mesher': 'delaunay-graded',
           'refinement_factor': rf, 'lc_fine': h_f}
    if domain_type == 'sulcus':
        geo.update({'w': float(w), 'd': float(d), 'xL': float(xL), 'xR': float(xR)})
    return HostMesh(points, cells, geo).check()


def mesh_domain(L=10.0, H=1.0, w=0.5, d=1.0, h=0.02, domain_type='sulcus', smooth_passes=2,
                refinement_factor=1) -> HostMesh:
    """Unstructured isotropic mesh of the sulcus or rectangular domain; ``refinement_factor`` > 1 grades it towards
    the sulcus like the reference's Distance / Threshold size field."""
    if float(refinement_factor) > 1.0:
        return _mesh_domain_graded(L, H, w, d, h, domain_type, refinement_factor, smooth_passes)
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
    box = (float(L), float(H), float(xL), float(xR)) if domain_type == 'sulcus' else (float(L), float(H), 1.0, 0.0)
    lat = lat[_inside_channel_domain(poly, lat, box)]
    tree = cKDTree(fixed)
    dist, idx = tree.query(lat)
    # local spacing of the nearest fixed sample: cavity samples use h_cav
    loc_h = np.where((fixed[idx, 1] <= 0.0) & (fixed[idx, 0] >= xL) & (fixed[idx, 0] <= xR), h_cav, h) \
        if domain_type == 'sulcus' else np.full(len(lat), h)
    lat = lat[dist >= 0.8 * loc_h]
    points = np.concatenate([fixed, lat], axis=0)
    fixed_mask = np.zeros(len(points), dtype=bool)
    fixed_mask[:len(fixed)] = True
    cells = _triangulate(points, poly, h_cav, box)
    if not _validate(cells, len(points), bsegs, isegs):
        raise RuntimeError("unstructured mesher: boundary/mouth recovery failed")
    for _ in range(smooth_passes):
        trial = _smooth(points, cells, fixed_mask)
        tcells = _triangulate(trial, poly, h_cav, box)
        if _validate(tcells, len(trial), bsegs, isegs):
            points, cells = trial, tcells
        else:
            break
    geo = {'domain_type': domain_type, 'L': float(L), 'H': float(H), 'h': float(h), 'mesher': 'delaunay'}
    if domain_type == 'sulcus':
        geo.update({'w': float(w), 'd': float(d), 'xL': float(xL), 'xR': float(xR)})
    return HostMesh(points, cells, geo).check()
