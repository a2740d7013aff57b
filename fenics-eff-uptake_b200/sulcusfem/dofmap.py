"""DOF maps, CSR sparsity patterns and gather-style scatter maps (host side, built once per mesh).

Replaces what dolfin's ``DofMapBuilder`` / ``SparsityPatternBuilder`` do when the reference calls
``FunctionSpace(mesh,"CG",2)`` (reference ``simulation.py:146``) and
``VectorFunctionSpace``/``MixedElement`` (``simulation.py:128-130``).

Numbering contract: the UFC numbering dolfin starts from *before* its
build-dependent graph reordering -- P1 dof = vertex index; P2 dof = vertex index for the three
vertex dofs and ``n_v + edge index`` for the three edge dofs (edges numbered lexicographically by
sorted vertex pair); the Taylor-Hood space concatenates ``[u_x (P2), u_y (P2), p (P1)]`` with
offsets and uses the cell layout ``[u_x x6, u_y x6, p x3]``.

Sparsity (App. B.2): union over cells of the full clique of the element dofs, *including* the
structural zeros (u_x-u_y and p-p blocks of the Stokes matrix); exterior-facet integrals add
nothing new.  Column indices are sorted inside each row.

The scatter map is stored gather-style: for every CSR slot the list of element-buffer positions
that contribute to it, sorted by (family, cell) so device assembly is a fixed-order sum with no
atomics and is bit-reproducible.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .hostmesh import HostMesh

# local vertices of local facet l (the facet opposite local vertex l), UFC ordering
FACET_LOCAL_VERTS = np.array([[1, 2], [0, 2], [0, 1]], dtype=np.int64)


def p1_cell_dofs(mesh: HostMesh) -> np.ndarray:
    return mesh.cells.astype(np.int32)


def p2_cell_dofs(mesh: HostMesh) -> np.ndarray:
    nv = mesh.num_vertices
    return np.concatenate([mesh.cells, mesh.cell_edges + nv], axis=1).astype(np.int32)


def p2_num_dofs(mesh: HostMesh) -> int:
    return mesh.num_vertices + mesh.num_edges


def th_num_dofs(mesh: HostMesh) -> int:
    return 2 * p2_num_dofs(mesh) + mesh.num_vertices


def th_cell_dofs(mesh: HostMesh) -> np.ndarray:
    n2 = p2_num_dofs(mesh)
    c2 = p2_cell_dofs(mesh).astype(np.int64)
    return np.concatenate([c2, c2 + n2, mesh.cells.astype(np.int64) + 2 * n2], axis=1).astype(np.int32)


def p2_dof_coordinates(mesh: HostMesh) -> np.ndarray:
    """Coordinates of the P2 dofs (vertices, then edge midpoints); cached on the mesh (read-only)."""
    X = getattr(mesh, '_p2_dof_coords', None)
    if X is None:
        X = np.concatenate([mesh.coords, mesh.edge_midpoints()], axis=0)
        X.setflags(write=False)
        mesh._p2_dof_coords = X
    return X


def p2_facet_dofs(mesh: HostMesh, facets: np.ndarray) -> np.ndarray:
    """Global P2 dofs [va, vb, edge] of each facet (va < vb)."""
    facets = np.asarray(facets, dtype=np.int64)
    nv = mesh.num_vertices
    return np.concatenate([mesh.edges[facets].astype(np.int64), (facets + nv)[:, None]], axis=1).astype(np.int32)


def p2_facet_local_dofs(local_facet: np.ndarray) -> np.ndarray:
    """Cell-local P2 dof indices [a, b, 3+l] of local facet l."""
    l = np.asarray(local_facet, dtype=np.int64)
    return np.concatenate([FACET_LOCAL_VERTS[l], (3 + l)[:, None]], axis=1).astype(np.int32)


@dataclass
class CsrPattern:
    nrows: int
    ncols: int
    rowptr: np.ndarray          # int32 [nrows+1]
    cols: np.ndarray            # int32 [nnz]
    contrib_ptr: np.ndarray     # int32 [nnz+1]   gather map: slot -> range in contrib_code
    contrib_code: np.ndarray    # int32 [total]   positions in the element buffer
    family_base: List[int]      # element-buffer offset of each family
    buffer_len: int             # total doubles in the element buffer

    @property
    def nnz(self) -> int:
        return int(self.cols.shape[0])


# Set by ``device.Context`` (never in the host-only worker processes of ``simulation.prefetch_meshes``): the CUDA device
# on which large patterns are sorted.  The key sort is >= 85 % of the host set-up time of a refined mesh (numpy's stable
# int64 sort: ~10 s for the 34 M element entries of the bench mesh); on the device it takes milliseconds.  Both paths are
# the same stable sort of the same keys, hence bit-identical outputs (tests/test_gpu_core.py).
DEVICE_SORT = None
DEVICE_SORT_MIN = 1 << 20


def _build_pattern_device(nrows, ncols, families, device) -> CsrPattern:
    import torch
    key_parts, bases = [], []
    base = 0
    for rows, cols in families:
        n, a = np.asarray(rows).shape
        b = np.asarray(cols).shape[1]
        bases.append(base)
        if n:
            r = torch.from_numpy(np.ascontiguousarray(rows, dtype=np.int64)).to(device)
            c = torch.from_numpy(np.ascontiguousarray(cols, dtype=np.int64)).to(device)
            key_parts.append((r[:, :, None] * int(ncols) + c[:, None, :]).reshape(-1))
        base += n * a * b
    if base >= 2 ** 31:
        raise ValueError("element buffer exceeds int32 addressing")
    keys = torch.cat(key_parts) if len(key_parts) > 1 else key_parts[0]
    del key_parts
    # element-buffer codes are base_f + u a_f b_f + i b_f + j = the running index of the concatenated key list, so the
    # sorted codes are the sort permutation itself
    skeys, order = torch.sort(keys, stable=True)
    del keys
    new = torch.ones(skeys.numel(), dtype=torch.bool, device=device)
    new[1:] = skeys[1:] != skeys[:-1]
    starts = torch.nonzero(new).reshape(-1)
    ukeys = skeys[starts]
    del skeys, new
    contrib_ptr = torch.cat([starts, torch.tensor([order.numel()], device=device, dtype=starts.dtype)]).to(torch.int32)
    contrib_code = order.to(torch.int32)
    r = torch.div(ukeys, int(ncols), rounding_mode='floor')
    c = (ukeys - r * int(ncols)).to(torch.int32)
    rowptr = torch.zeros(nrows + 1, dtype=torch.int64, device=device)
    rowptr[1:] = torch.cumsum(torch.bincount(r, minlength=nrows), 0)
    return CsrPattern(nrows, ncols, rowptr.to(torch.int32).cpu().numpy(), c.cpu().numpy(), contrib_ptr.cpu().numpy(),
                      contrib_code.cpu().numpy(), bases, base)


def build_pattern(nrows: int, ncols: int,
                  families: Sequence[Tuple[np.ndarray, np.ndarray]]) -> CsrPattern:
    """CSR pattern + gather map for a list of element families.

    ``families[f] = (row_dofs [n_f, a_f], col_dofs [n_f, b_f])``.  The element buffer is AoS per
    family: entry (unit u, local i, local j) of family f lives at
    ``family_base[f] + u*a_f*b_f + i*b_f + j`` (one row-major element matrix per unit), so the
    gather kernel finds the entries of one matrix row of one cell inside one short span.
    """
    if DEVICE_SORT is not None:
        total = sum(int(np.asarray(r).shape[0]) * int(np.asarray(r).shape[1]) * int(np.asarray(c).shape[1]) for r, c in families)
        if total >= DEVICE_SORT_MIN:
            return _build_pattern_device(nrows, ncols, families, DEVICE_SORT)
    key_parts, code_parts, bases = [], [], []
    base = 0
    for rows, cols in families:
        rows = np.asarray(rows, dtype=np.int64)
        cols = np.asarray(cols, dtype=np.int64)
        n, a = rows.shape
        b = cols.shape[1]
        bases.append(base)
        if n:
            keys = (rows[:, :, None] * ncols + cols[:, None, :]).reshape(n, a * b)
            codes = base + np.arange(n, dtype=np.int64)[:, None] * (a * b) + np.arange(a * b, dtype=np.int64)[None, :]
            key_parts.append(keys.ravel())
            code_parts.append(codes.ravel())
        base += n * a * b
    if base >= 2 ** 31:
        raise ValueError("element buffer exceeds int32 addressing")
    keys = np.concatenate(key_parts) if key_parts else np.zeros(0, dtype=np.int64)
    codes = np.concatenate(code_parts) if code_parts else np.zeros(0, dtype=np.int64)
    order = np.argsort(keys, kind='stable')       # stable: contributions stay in (family, unit) order
    skeys = keys[order]
    new = np.ones(len(skeys), dtype=bool)
    new[1:] = skeys[1:] != skeys[:-1]
    starts = np.flatnonzero(new)
    ukeys = skeys[starts]
    contrib_ptr = np.concatenate([starts, [len(skeys)]]).astype(np.int32)
    contrib_code = codes[order].astype(np.int32)
    r = ukeys // ncols
    c = (ukeys % ncols).astype(np.int32)
    rowptr = np.zeros(nrows + 1, dtype=np.int64)
    rowptr[1:] = np.cumsum(np.bincount(r, minlength=nrows))
    return CsrPattern(nrows, ncols, rowptr.astype(np.int32), c, contrib_ptr, contrib_code, bases, base)


def memo_pattern(mesh: HostMesh, key, nrows: int, ncols: int, families) -> CsrPattern:
    """``build_pattern`` memoised on the mesh object (``mesh._pattern_cache``).  Patterns are pure functions of the mesh
    and the family definition named by ``key``; the cache travels with the mesh when it is pickled, which is how the
    worker processes of ``simulation.prefetch_meshes`` hand finished patterns to the solving process."""
    cache = mesh.__dict__.setdefault('_pattern_cache', {})
    pat = cache.get(key)
    if pat is None or pat.nrows != nrows or pat.ncols != ncols:
        pat = build_pattern(nrows, ncols, families)
        cache[key] = pat
    return pat


def scalar_level_plan(mesh: HostMesh, bc_markers: np.ndarray, degree: int, robin_id: Optional[int]):
    """Host plan of one scalar Lagrange level (device.ScalarLevel): cell dofs, Robin facets and their dofs, and the
    CSR pattern + gather map (memoised).  Returns ``(n, cd, facets, facet_dofs, pattern)``; facets / facet_dofs are
    None without a Robin boundary."""
    if degree == 2:
        cd, n = p2_cell_dofs(mesh), p2_num_dofs(mesh)
    else:
        cd, n = p1_cell_dofs(mesh), int(mesh.num_vertices)
    fam = [(cd, cd)]
    f = fd = None
    key = ('scalar', int(degree), None)
    if robin_id is not None:
        f, _, _ = boundary_facets(mesh, bc_markers, robin_id)
        fd = p2_facet_dofs(mesh, f) if degree == 2 else mesh.edges[f].astype(np.int32)
        fam.append((fd, fd))
        import hashlib                       # (not hash(): that is salted per process, and the cache crosses processes)
        digest = hashlib.blake2b(np.ascontiguousarray(f).tobytes(), digest_size=8).hexdigest()
        key = ('scalar', int(degree), (int(robin_id), len(f), digest))
    return n, cd, f, fd, memo_pattern(mesh, key, n, n, fam)


def stokes_block_plans(mesh: HostMesh):
    """Host plans of the directly assembled Stokes blocks (device.StokesProblem): patterns + gather maps of
    B (nv x 2 n2) and B^T (2 n2 x nv), velocity in interleaved numbering, both reading ONE element buffer
    EB [nc][3][12] (rows = the cell's pressure dofs, columns = [u_x x6 | u_y x6]), and of the P1 pressure mass matrix.
    Returns ``(pb, pbt, bt_code, pmass)``; ``bt_code`` = the gather codes of B^T re-addressed to B's buffer layout."""
    n2, nv = p2_num_dofs(mesh), int(mesh.num_vertices)
    c2 = p2_cell_dofs(mesh).astype(np.int64)
    il = np.concatenate([2 * c2, 2 * c2 + 1], axis=1)                       # [nc, 12] interleaved velocity dofs
    c1v = p1_cell_dofs(mesh).astype(np.int64)                               # [nc, 3]
    pb = memo_pattern(mesh, ('stokes_B',), nv, 2 * n2, [(c1v, il)])         # codes: cell*36 + k*12 + m
    pbt = memo_pattern(mesh, ('stokes_BT',), 2 * n2, nv, [(il, c1v)])       # codes: cell*36 + m*3 + k  -> remap
    code = np.asarray(pbt.contrib_code, dtype=np.int32)                     # < 36 nc: int32 arithmetic throughout
    loc = code % np.int32(36)
    bt_code = (code - loc) + (loc % np.int32(3)) * np.int32(12) + loc // np.int32(3)
    c1 = p1_cell_dofs(mesh)
    pmass = memo_pattern(mesh, ('p1_mass',), nv, nv, [(c1, c1)])
    return pb, pbt, bt_code, pmass


def transpose_csr(nrows, ncols, rowptr, cols, vals=None):
    """CSR transpose; returns (rowptr_t, cols_t, perm) with ``vals_t = vals[perm]``."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    rows = np.repeat(np.arange(nrows, dtype=np.int64), np.diff(rowptr))
    perm = np.argsort(np.asarray(cols, dtype=np.int64) * nrows + rows, kind='stable')
    cols_t = rows[perm].astype(np.int32)
    cnt = np.bincount(np.asarray(cols, dtype=np.int64), minlength=ncols)
    rowptr_t = np.concatenate([[0], np.cumsum(cnt)]).astype(np.int32)
    return rowptr_t, cols_t, perm.astype(np.int64)


def boundary_facets(mesh: HostMesh, markers: np.ndarray, marker_id: int):
    """Exterior facets with the given id (a ``ds(id)`` measure): (facet, cell, local facet)."""
    f = np.flatnonzero((np.asarray(markers) == marker_id) & mesh.edge_on_boundary)
    return f, mesh.edge_cells[f, 0].astype(np.int64), mesh.edge_local[f, 0].astype(np.int64)


def dirichlet_dofs_p2(mesh: HostMesh, markers: np.ndarray, marker_id: int) -> np.ndarray:
    """P2 dofs of ``DirichletBC(V, g, markers, id)`` (topological: both vertices + edge dof)."""
    f = np.flatnonzero(np.asarray(markers) == marker_id)
    return np.unique(p2_facet_dofs(mesh, f).ravel()).astype(np.int64)


def dirichlet_dofs_p1(mesh: HostMesh, markers: np.ndarray, marker_id: int) -> np.ndarray:
    f = np.flatnonzero(np.asarray(markers) == marker_id)
    return np.unique(mesh.edges[f].ravel()).astype(np.int64)
