"""Host-side triangle meshes for the sulcus transport model (numpy only).

The reference obtains its meshes from Gmsh -> meshio -> ``dolfin.Mesh`` (reference
``mesh.py:350-391,421,487``) and marks facets with ``SubDomain.mark`` lambdas
(``mesh.py:196-256``).  Neither Gmsh nor dolfin travels with this repo, so the host layer
provides

* :class:`HostMesh` -- vertices, cells (vertices sorted ascending per cell, the convention dolfin
  applies when it orders a mesh on XML read), edges numbered lexicographically by their sorted
  vertex pair, cell->edge and edge->cell connectivity;
* deterministic synthetic meshers for the two reference domains (``rectangle_mesh``,
  ``sulcus_mesh``): channel ``[0,L]x[0,H]`` with mesh columns aligned to the sulcus mouth
  ``[xL,xR]`` and a cavity floor ``y=-d sin(pi x_rel)`` (``mesh.py:139-155``) whose mouth line is
  made of mesh edges like the reference's embedded ``Line{7} In Surface{1}`` (``mesh.py:310-311``);
* uniform red refinement (mesh-convergence study, BASELINE config 5) with nested vertex numbering
  (level l+1 vertices = level l vertices followed by level l edge midpoints, i.e. the P2 node set
  of level l *is* the P1 node set of level l+1);
* marker construction with the exact ``SubDomain.mark`` semantics (a facet gets the id iff the
  predicate holds at both vertices and at the midpoint; later marks overwrite earlier ones).

Everything here is integer/index bookkeeping done once per mesh; no solve-path arithmetic.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, Dict, Optional

import numpy as np

DOLFIN_EPS = 3.0e-16          # dolfin/common/constants.h (used by mesh.py:50,201-213)
TOLERANCE = 2.0 * DOLFIN_EPS  # MeshGenerator.TOLERANCE, mesh.py:50

MARKERS = {                    # mesh.py:43-47 / analysis.py:17-21
    'left': 1, 'right': 2, 'top': 3, 'bottom': 4,
    'bottom_left': 5, 'sulcus': 6, 'bottom_right': 7, 'sulcus_opening': 8,
    'y0_line': 10,
}


class _CallableInt(int):
    def __call__(self):
        return int(self)


class HostMesh:
    """2-D triangle mesh with dolfin-style ordered cells and derived connectivity."""

    def __init__(self, coords: np.ndarray, cells: np.ndarray, geometry: Optional[dict] = None):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int64)
        if coords.ndim != 2 or coords.shape[1] != 2:
            raise ValueError("coords must be [nv,2]")
        if cells.ndim != 2 or cells.shape[1] != 3:
            raise ValueError("cells must be [nc,3]")
        if cells.size and (cells.min() < 0 or cells.max() >= len(coords)):
            raise ValueError("cell vertex index out of range")
        self.coords = coords
        self.cells = np.sort(cells, axis=1).astype(np.int32)   # dolfin mesh.order()
        self.geometry = dict(geometry or {})
        self.parent: Optional["HostMesh"] = None                # set by refine()
        self._build_connectivity()

    # ------------------------------------------------------------------ connectivity
    def _build_connectivity(self):
        c = self.cells.astype(np.int64)
        nv = self.num_vertices
        # local edge i is opposite local vertex i (UFC): e0=(v1,v2), e1=(v0,v2), e2=(v0,v1)
        pairs = np.stack([c[:, [1, 2]], c[:, [0, 2]], c[:, [0, 1]]], axis=1)   # [nc,3,2], sorted
        keys = (pairs[..., 0] * nv + pairs[..., 1]).ravel()
        # ONE stable sort of the 3 nc keys gives everything: the unique edges in lexicographic order (= np.unique), the
        # edge id of every (cell, local edge) slot and the slots of an edge grouped in cell order (what a stable argsort
        # of those ids followed by two searchsorted passes produced before -- same arrays, half the time)
        order = np.argsort(keys, kind='stable')
        sk = keys[order]
        new_edge = np.ones(len(sk), dtype=bool)
        new_edge[1:] = sk[1:] != sk[:-1]
        first = np.flatnonzero(new_edge)
        uniq = sk[first]
        ne = len(uniq)
        inv = np.empty(len(keys), dtype=np.int64)
        inv[order] = np.cumsum(new_edge) - 1
        self.edges = np.stack([uniq // nv, uniq % nv], axis=1).astype(np.int32)
        self.cell_edges = inv.reshape(-1, 3).astype(np.int32)
        # edge -> (first cell, second cell or -1), first = lowest cell index ('+' side in dolfin)
        cell_of = (order // 3).astype(np.int32)
        loc_of = (order % 3).astype(np.int8)
        count = np.diff(np.append(first, len(sk)))
        if count.max(initial=0) > 2:
            raise ValueError("non-manifold mesh: an edge has more than two cells")
        self.edge_cells = np.full((ne, 2), -1, dtype=np.int32)
        self.edge_local = np.full((ne, 2), -1, dtype=np.int8)
        self.edge_cells[:, 0] = cell_of[first]
        self.edge_local[:, 0] = loc_of[first]
        two = count == 2
        self.edge_cells[two, 1] = cell_of[first[two] + 1]
        self.edge_local[two, 1] = loc_of[first[two] + 1]
        self.edge_on_boundary = ~two

    # ------------------------------------------------------------------ sizes
    # plain ints that can also be *called* like the dolfin methods (simulation.py:245-246 uses
    # mesh.num_vertices() / mesh.num_cells())
    @property
    def num_vertices(self) -> int:
        return _CallableInt(self.coords.shape[0])

    @property
    def num_cells(self) -> int:
        return _CallableInt(self.cells.shape[0])

    @property
    def num_edges(self) -> int:
        return _CallableInt(self.edges.shape[0])

    num_facets = num_edges

    # dolfin-like accessors used by the reference callers (simulation.py:245-248)
    def coordinates(self) -> np.ndarray:
        return self.coords

    def hmin(self) -> float:
        return float(self._cell_diameters().min())

    def hmax(self) -> float:
        return float(self._cell_diameters().max())

    def _cell_diameters(self) -> np.ndarray:
        # dolfin 2019.1 Cell::h() for simplices = largest vertex-to-vertex distance
        p = self.coords[self.cells]
        d01 = np.linalg.norm(p[:, 0] - p[:, 1], axis=1)
        d02 = np.linalg.norm(p[:, 0] - p[:, 2], axis=1)
        d12 = np.linalg.norm(p[:, 1] - p[:, 2], axis=1)
        return np.maximum(np.maximum(d01, d02), d12)

    def cell_midpoints(self) -> np.ndarray:
        return self.coords[self.cells].mean(axis=1)

    def edge_midpoints(self) -> np.ndarray:
        return 0.5 * (self.coords[self.edges[:, 0]] + self.coords[self.edges[:, 1]])

    def signed_areas(self) -> np.ndarray:
        """Signed cell areas (memoised: check(), the hierarchy, the locator and the Schur set-up all ask; the arrays of
        a mesh are never modified in place).  The returned array is read-only."""
        a = getattr(self, '_signed_areas', None)
        if a is None:
            p = self.coords[self.cells]
            a = 0.5 * ((p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1])
                       - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1]))
            a.setflags(write=False)
            self._signed_areas = a
        return a

    def check(self):
        a = np.abs(self.signed_areas())
        if not np.all(a > 0):
            raise ValueError("degenerate (zero-area) cell in mesh")
        return self


# ====================================================================== facet / cell markers
@dataclass
class MeshMarkers:
    """Integer labels per facet (edge) or per cell, like ``MeshFunction('size_t')``."""
    values: np.ndarray
    dim: int

    def array(self) -> np.ndarray:        # solvers.py:242 calls bc_markers.array()
        return self.values

    def __getitem__(self, i):
        return int(self.values[int(i)])


def _near(a, b, eps):
    return (b - eps <= a) & (a <= b + eps)    # dolfin::near = between(x, (x0-eps, x0+eps))


def boundary_predicates(width, height, xL, xR) -> Dict[str, Callable]:
    """Vectorised restatement of the lambdas in reference ``mesh.py:200-214``.

    Each predicate maps (x[n], y[n], on_boundary[n]) -> bool[n].
    """
    return {
        'left':   lambda x, y, ob: ob & _near(x, 0.0, DOLFIN_EPS),
        'right':  lambda x, y, ob: ob & _near(x, width, DOLFIN_EPS),
        'top':    lambda x, y, ob: ob & _near(y, height, DOLFIN_EPS),
        'bottom': lambda x, y, ob: ob & (y <= 0.0),
        'bottom_left':  lambda x, y, ob: ob & _near(y, 0.0, TOLERANCE) & (x <= xL - DOLFIN_EPS),
        'bottom_right': lambda x, y, ob: ob & _near(y, 0.0, TOLERANCE) & (x >= xR + DOLFIN_EPS),
        'sulcus': lambda x, y, ob: ob & (xL <= x) & (x <= xR) & (y < -DOLFIN_EPS),
        'sulcus_opening': lambda x, y, ob: (_near(y, 0.0, TOLERANCE)
                                            & (xL + DOLFIN_EPS < x) & (x < xR - DOLFIN_EPS)),
        'y0_line': lambda x, y, ob: _near(y, 0.0, TOLERANCE),
    }


def mark_facets(mesh: HostMesh, names, predicates) -> MeshMarkers:
    """``SubDomain.mark`` on facets (reference ``mesh.py:217-256``).

    A facet receives the id iff ``inside`` is true for both vertices *and* the midpoint, with
    ``on_boundary`` = "the facet has exactly one cell"; ids are applied in list order so later
    marks overwrite earlier ones.
    """
    vals = np.zeros(mesh.num_edges, dtype=np.int32)
    ob = mesh.edge_on_boundary
    pts = getattr(mesh, '_facet_points', None)          # end points + midpoints of all facets, shared by the marker sets
    if pts is None:
        p0 = mesh.coords[mesh.edges[:, 0]]
        p1 = mesh.coords[mesh.edges[:, 1]]
        pts = mesh._facet_points = (p0, p1, 0.5 * (p0 + p1))
    p0, p1, pm = pts
    for name in names:
        f = predicates[name]
        # the three evaluations are AND-ed: the second end point and the midpoint are only looked at where the first end
        # point is inside (a few thousand of the millions of facets) -- same marks
        idx = np.flatnonzero(f(p0[:, 0], p0[:, 1], ob))
        if len(idx):
            obi = ob[idx]
            keep = f(p1[idx, 0], p1[idx, 1], obi) & f(pm[idx, 0], pm[idx, 1], obi)
            vals[idx[keep]] = MARKERS[name]
    return MeshMarkers(vals, 1)


def build_markers(mesh: HostMesh, width: float, height: float, xL: float, xR: float,
                  domain_type: str) -> dict:
    """The three facet-marker sets + cell markers of reference ``mesh.py:425-453,494-496``."""
    pred = boundary_predicates(width, height, xL, xR)
    out = {'bc_markers': mark_facets(mesh, ['left', 'right', 'top', 'bottom'], pred)}
    if domain_type == 'sulcus':
        out['bottom_segment_markers'] = mark_facets(
            mesh, ['bottom_left', 'bottom_right', 'sulcus', 'sulcus_opening'], pred)
        out['y0_markers'] = mark_facets(mesh, ['y0_line'], pred)
        cy = mesh.cell_midpoints()[:, 1]
        out['domain_markers'] = MeshMarkers(np.where(cy <= 0.0, 1, 2).astype(np.int32), 2)
    mesh._facet_points = None                    # 100 MB at 1.4 M facets: only needed while the sets are being marked
    mesh._sfem_markers = out                 # analysis functions whose reference signature carries only `measures`
    return out


# ====================================================================== synthetic meshers
def _grid_nodes(n_total: int, length: float) -> np.ndarray:
    return np.linspace(0.0, length, n_total + 1)


def rectangle_mesh(L: float = 10.0, H: float = 1.0, nx: int = 500, ny: int = 50,
                   x_nodes: Optional[np.ndarray] = None) -> HostMesh:
    """Structured right-diagonal triangulation of ``[0,L]x[0,H]`` (SURVEY 8(d) ``rect(r)``)."""
    xs = _grid_nodes(nx, L) if x_nodes is None else np.asarray(x_nodes, dtype=np.float64)
    nx = len(xs) - 1
    ys = _grid_nodes(ny, H)
    X, Y = np.meshgrid(xs, ys, indexing='ij')              # vertex id = i*(ny+1)+j
    coords = np.stack([X.ravel(), Y.ravel()], axis=1)
    i, j = np.meshgrid(np.arange(nx), np.arange(ny), indexing='ij')
    v00 = (i * (ny + 1) + j).ravel()
    v10 = v00 + (ny + 1)
    v01 = v00 + 1
    v11 = v10 + 1
    cells = np.concatenate([np.stack([v00, v10, v11], axis=1),
                            np.stack([v00, v11, v01], axis=1)], axis=0)
    geo = {'domain_type': 'rectangular', 'L': float(L), 'H': float(H)}
    return HostMesh(coords, cells, geo).check()


def _mouth_aligned_x_nodes(L, xL, xR, h, min_mouth_cols=2):
    n1 = max(1, int(round(xL / h)))
    n2 = max(min_mouth_cols, int(round((xR - xL) / h)))
    n3 = max(1, int(round((L - xR) / h)))
    xs = np.concatenate([np.linspace(0.0, xL, n1 + 1)[:-1],
                         np.linspace(xL, xR, n2 + 1)[:-1],
                         np.linspace(xR, L, n3 + 1)])
    xs[n1] = xL
    xs[n1 + n2] = xR
    return xs, n1, n2, n3


def sulcus_floor(x, xL, w, d):
    """Cavity floor ``y=-d sin(pi (x-xL)/w)`` sampled by the reference at 21 points (mesh.py:147-154)."""
    return -d * np.sin(np.pi * (x - xL) / w)


def sulcus_mesh(L: float = 10.0, H: float = 1.0, w: float = 0.5, d: float = 1.0,
                h: float = 0.02) -> HostMesh:
    """Channel + one sinusoidal cavity, columns aligned with the mouth, mouth edges on y=0.

    Channel: structured right-diagonal grid.  Cavity: vertical node columns under each mouth grid
    line with ``max(1, round(depth/h))`` layers, neighbouring columns zipped into triangles.
    """
    if not (0 < w < L) or d <= 0 or h <= 0:
        raise ValueError("invalid sulcus geometry")
    xL, xR = L / 2 - w / 2, L / 2 + w / 2           # mesh.py:100-101
    xs, n1, n2, n3 = _mouth_aligned_x_nodes(L, xL, xR, h)
    ny = max(1, int(round(H / h)))
    chan = rectangle_mesh(L, H, len(xs) - 1, ny, x_nodes=xs)
    coords = [chan.coords]
    cells = [chan.cells.astype(np.int64)]
    next_id = chan.num_vertices
    col_ids = []
    for i in range(n2 + 1):
        top = (n1 + i) * (ny + 1)                    # channel vertex on y=0 of this column
        if i == 0 or i == n2:
            col_ids.append(np.array([top], dtype=np.int64))
            continue
        x = xs[n1 + i]
        depth = d * math.sin(math.pi * i / n2)
        m = max(1, int(round(depth / h)))
        yy = -depth * np.arange(1, m + 1) / m
        ids = np.arange(next_id, next_id + m, dtype=np.int64)
        next_id += m
        coords.append(np.stack([np.full(m, x), yy], axis=1))
        col_ids.append(np.concatenate([[top], ids]))
    tri = []
    for i in range(n2):
        a, b = col_ids[i], col_ids[i + 1]
        m, n = len(a) - 1, len(b) - 1
        p = q = 0
        while p < m or q < n:
            adv_a = (q == n) or (p < m and (p + 1) * max(n, 1) <= (q + 1) * max(m, 1))
            if adv_a:
                tri.append((a[p], a[p + 1], b[q]))
                p += 1
            else:
                tri.append((a[p], b[q], b[q + 1]))
                q += 1
    cells.append(np.asarray(tri, dtype=np.int64).reshape(-1, 3))
    geo = {'domain_type': 'sulcus', 'L': float(L), 'H': float(H), 'w': float(w), 'd': float(d),
           'xL': float(xL), 'xR': float(xR), 'h': float(h)}
    return HostMesh(np.concatenate(coords, axis=0), np.concatenate(cells, axis=0), geo).check()


# ====================================================================== uniform refinement
def refine(mesh: HostMesh, project_curved_boundary: bool = True) -> HostMesh:
    """Uniform red refinement; new vertex ``nv + e`` sits on the midpoint of edge ``e``.

    With ``project_curved_boundary`` the new vertices of boundary edges on the cavity floor are
    moved onto ``y=-d sin(pi x_rel)`` (the channel walls are straight, so nothing else moves).
    """
    nv = mesh.num_vertices
    mid = mesh.edge_midpoints()
    geo = mesh.geometry
    if project_curved_boundary and geo.get('domain_type') == 'sulcus':
        e = mesh.edges
        below = (mesh.coords[e[:, 0], 1] <= 0.0) & (mesh.coords[e[:, 1], 1] <= 0.0)
        strictly = (mesh.coords[e[:, 0], 1] < 0.0) | (mesh.coords[e[:, 1], 1] < 0.0)
        on_curve = mesh.edge_on_boundary & below & strictly
        mid[on_curve, 1] = sulcus_floor(mid[on_curve, 0], geo['xL'], geo['w'], geo['d'])
    coords = np.concatenate([mesh.coords, mid], axis=0)
    c = mesh.cells.astype(np.int64)
    m = mesh.cell_edges.astype(np.int64) + nv                 # m[:,i] = midpoint opposite vertex i
    cells = np.concatenate([
        np.stack([c[:, 0], m[:, 2], m[:, 1]], axis=1),
        np.stack([c[:, 1], m[:, 2], m[:, 0]], axis=1),
        np.stack([c[:, 2], m[:, 1], m[:, 0]], axis=1),
        np.stack([m[:, 0], m[:, 1], m[:, 2]], axis=1)], axis=0)
    fine = HostMesh(coords, cells, geo).check()
    fine.parent = mesh
    return fine


def refine_n(mesh: HostMesh, n: int, project_curved_boundary: bool = True) -> HostMesh:
    for _ in range(int(n)):
        mesh = refine(mesh, project_curved_boundary)
    return mesh


# ====================================================================== mesh file readers
def read_dolfin_xml(path: str) -> HostMesh:
    """Reader for the legacy dolfin-XML triangle meshes the reference writes (mesh.py:384)."""
    import xml.etree.ElementTree as ET
    root = ET.parse(path).getroot()
    m = root.find('mesh')
    if m is None or m.get('celltype') != 'triangle':
        raise ValueError("expected a dolfin-XML triangle mesh")
    verts = m.find('vertices')
    cs = m.find('cells')
    coords = np.zeros((int(verts.get('size')), 2))
    for v in verts:
        coords[int(v.get('index'))] = (float(v.get('x')), float(v.get('y')))
    cells = np.zeros((int(cs.get('size')), 3), dtype=np.int64)
    for t in cs:
        cells[int(t.get('index'))] = (int(t.get('v0')), int(t.get('v1')), int(t.get('v2')))
    return HostMesh(coords, cells).check()


def write_dolfin_xml(mesh: HostMesh, path: str):
    with open(path, 'w') as f:
        f.write('<?xml version="1.0"?>\n<dolfin xmlns:dolfin="http://fenicsproject.org">\n')
        f.write('  <mesh celltype="triangle" dim="2">\n')
        f.write(f'    <vertices size="{mesh.num_vertices}">\n')
        for i, (x, y) in enumerate(mesh.coords):
            f.write(f'      <vertex index="{i}" x="{float(x)!r}" y="{float(y)!r}" />\n')
        f.write(f'    </vertices>\n    <cells size="{mesh.num_cells}">\n')
        for i, (a, b, c) in enumerate(mesh.cells):
            f.write(f'      <triangle index="{i}" v0="{a}" v1="{b}" v2="{c}" />\n')
        f.write('    </cells>\n  </mesh>\n</dolfin>\n')


def read_gmsh_msh2(path: str) -> HostMesh:
    """Reader for Gmsh ``-format msh2`` ASCII files (reference mesh.py:353); z is dropped (mesh.py:381-382)."""
    with open(path) as f:
        lines = [ln.strip() for ln in f]
    i = lines.index('$Nodes')
    n = int(lines[i + 1])
    ids = np.zeros(n, dtype=np.int64)
    coords = np.zeros((n, 2))
    for k in range(n):
        t = lines[i + 2 + k].split()
        ids[k] = int(t[0])
        coords[k] = (float(t[1]), float(t[2]))
    remap = {int(g): k for k, g in enumerate(ids)}
    i = lines.index('$Elements')
    ne = int(lines[i + 1])
    tris = []
    for k in range(ne):
        t = lines[i + 2 + k].split()
        if int(t[1]) == 2:                                  # 3-node triangle
            ntags = int(t[2])
            tris.append([remap[int(v)] for v in t[3 + ntags:3 + ntags + 3]])
    cells = np.asarray(tris, dtype=np.int64)
    used = np.unique(cells)                                 # meshio keeps all points; prune orphans
    if len(used) != n:
        lut = -np.ones(n, dtype=np.int64)
        lut[used] = np.arange(len(used))
        coords, cells = coords[used], lut[cells]
    return HostMesh(coords, cells).check()
