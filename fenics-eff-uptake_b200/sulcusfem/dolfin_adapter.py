"""Adapter between REAL dolfin objects and the B200 solve path.

The reference's callers are written against ``from dolfin import *``: ``simulation.py:128-166`` builds dolfin function
spaces, hands them to the six ``solvers.py`` entry points and then uses the returned ``dolfin.Function`` objects as UFL
coefficients, with ``File(...) << u``, ``c(Point)``, ``c.vector()`` ...  Where dolfin is importable, this module lets
those callers run unmodified on the GPU path:

* :func:`host_mesh_from_dolfin`, :func:`marker_values` -- a ``dolfin.Mesh`` / facet ``MeshFunction`` -> the host mesh and
  marker arrays the device problems are planned from (same vertex and cell numbering; facets matched by vertex pair);
* :func:`dof_map` -- dolfin's DOF numbering is build dependent, so DOFs are matched by coordinate
  (``V.tabulate_dof_coordinates()``; vector / mixed spaces through ``V.sub(i).dofmap().dofs()``);
* :func:`to_dolfin` / :func:`from_dolfin` -- copy a device result into a genuine ``dolfin.Function`` and a dolfin field
  (the velocity handed to ``advdiff_solver``) into this package's ``Function``;
* :class:`DolfinSolvers` -- the six entry points with dolfin objects in and out; :func:`install` registers them as the
  module ``solvers`` so that ``import simulation`` of the reference tree picks them up instead of the dolfin/PETSc ones.

Only duck-typed members are used (``coordinates()``, ``cells()``, ``array()``, ``tabulate_dof_coordinates()``,
``sub(i).dofmap().dofs()``, ``vector().get_local() / set_local()``), so the module imports -- and its matching logic is
tested, ``tests/test_host.py`` -- without dolfin.  dolfin itself is not installable in this image.
"""
from __future__ import annotations

import sys
import types
from typing import Optional

import numpy as np

from . import dofmap as dm
from .fem import Function, FunctionSpace, MixedElement, VectorFunctionSpace
from .hostmesh import HostMesh, MeshMarkers

_ATTR = '_sfem_host'


def host_mesh_from_dolfin(mesh, geometry: Optional[dict] = None) -> HostMesh:
    """HostMesh with the vertex / cell arrays of a ``dolfin.Mesh`` (cached on the dolfin object)."""
    hm = getattr(mesh, _ATTR, None)
    if hm is None:
        coords = np.ascontiguousarray(np.asarray(mesh.coordinates(), dtype=np.float64)[:, :2])
        cells = np.ascontiguousarray(np.asarray(mesh.cells(), dtype=np.int64))
        hm = HostMesh(coords, cells, dict(geometry or {})).check()
        try:
            setattr(mesh, _ATTR, hm)
        except AttributeError:
            pass
    return hm


def _facet_vertices(mesh):
    """[nf, 2] vertex ids of every dolfin facet (edge) in dolfin's facet numbering."""
    if hasattr(mesh, 'init'):
        mesh.init(1)
    conn = mesh.topology()(1, 0)
    flat = np.asarray(conn() if callable(conn) else conn, dtype=np.int64)
    return flat.reshape(-1, 2)


def marker_values(mf, mesh, host: HostMesh) -> np.ndarray:
    """Values of a dolfin facet ``MeshFunction`` re-indexed to the host mesh's edge numbering (facets are matched by
    their vertex pair, which both sides share)."""
    vals = np.asarray(mf.array())
    fv = np.sort(_facet_vertices(mesh), axis=1)
    nv = host.num_vertices
    key_d = fv[:, 0] * nv + fv[:, 1]
    he = np.sort(np.asarray(host.edges, dtype=np.int64), axis=1)
    key_h = he[:, 0] * nv + he[:, 1]
    order = np.argsort(key_d)
    pos = np.searchsorted(key_d[order], key_h)
    if np.any(pos >= len(order)) or not np.array_equal(key_d[order][np.minimum(pos, len(order) - 1)], key_h):
        raise ValueError("dolfin facets and host edges do not describe the same mesh")
    return vals[order[pos]]


def cell_marker_values(mf) -> np.ndarray:
    return np.asarray(mf.array())          # cell numbering is shared


def _match(coords_theirs: np.ndarray, coords_ours: np.ndarray) -> np.ndarray:
    """perm with coords_theirs[perm[i]] == coords_ours[i] (each of ours matched to exactly one of theirs)."""
    from scipy.spatial import cKDTree
    a, b = np.asarray(coords_theirs, dtype=np.float64)[:, :2], np.asarray(coords_ours, dtype=np.float64)
    if len(a) != len(b):
        raise ValueError(f"dof counts differ: dolfin {len(a)}, host {len(b)}")
    scale = max(float(np.ptp(b, axis=0).max()), 1.0)
    dist, idx = cKDTree(a).query(b)
    if dist.max() > 1e-9 * scale or len(np.unique(idx)) != len(idx):
        raise ValueError("dof coordinates of the dolfin space do not match the host space one to one")
    return idx.astype(np.int64)


def dof_map(V, host: HostMesh, kind: str) -> np.ndarray:
    """``theirs`` such that ``dolfin_vector[theirs] == our_vector`` for the space kind 'P1' | 'P2' | 'P2v' | 'TH'
    (our layouts: P2v = [ux | uy] blocked, TH = [ux | uy | p])."""
    X = np.asarray(V.tabulate_dof_coordinates(), dtype=np.float64).reshape(-1, 2)
    P2, P1 = dm.p2_dof_coordinates(host), host.coords

    def sub_dofs(space):
        return np.asarray(space.dofmap().dofs(), dtype=np.int64)
    if kind == 'P2':
        return _match(X, P2)
    if kind == 'P1':
        return _match(X, P1)
    if kind == 'P2v':
        out = []
        for i in range(2):
            d = sub_dofs(V.sub(i))
            out.append(d[_match(X[d], P2)])
        return np.concatenate(out)
    if kind == 'TH':
        out = []
        for i in range(2):
            d = sub_dofs(V.sub(0).sub(i))
            out.append(d[_match(X[d], P2)])
        d = sub_dofs(V.sub(1))
        out.append(d[_match(X[d], P1)])
        return np.concatenate(out)
    raise ValueError(f"unknown space kind {kind!r}")


def _cached_map(V, host, kind):
    cache = getattr(V, '_sfem_dofmap', None)
    if cache is None:
        cache = dof_map(V, host, kind)
        try:
            V._sfem_dofmap = cache
        except AttributeError:
            pass
    return cache


def to_dolfin(f: Function, V, make_function=None):
    """A genuine ``dolfin.Function(V)`` holding the values of ``f`` (DOFs matched by coordinate)."""
    if make_function is None:
        import dolfin
        make_function = dolfin.Function
    host = f.function_space().mesh()
    theirs = _cached_map(V, host, f.function_space().kind)
    out = make_function(V)
    vec = out.vector()
    vals = np.asarray(vec.get_local(), dtype=np.float64).copy()
    vals[theirs] = f.values
    vec.set_local(vals)
    if hasattr(vec, 'apply'):
        vec.apply("insert")
    return out


def from_dolfin(u, host: HostMesh, kind: str) -> Function:
    """This package's ``Function`` with the values of a dolfin Function on the matching space."""
    V = u.function_space()
    theirs = _cached_map(V, host, kind)
    vals = np.asarray(u.vector().get_local(), dtype=np.float64)[theirs]
    space = {'P2': lambda: FunctionSpace(host, 'CG', 2), 'P1': lambda: FunctionSpace(host, 'P', 1),
             'P2v': lambda: VectorFunctionSpace(host, 'P', 2)}[kind]()
    return Function(space, vals)


class DolfinSolvers:
    """The six ``solvers.py`` entry points (reference ``solvers.py:16,59,113,176,237,308``) for callers that hold dolfin
    objects: same signatures, dolfin ``Function`` results, the solves on the device."""

    def __init__(self, make_function=None, geometry: Optional[dict] = None):
        self._make, self._geometry = make_function, geometry

    # ---- conversions
    def _mesh_results(self, mesh_results):
        mesh = mesh_results['mesh']
        host = host_mesh_from_dolfin(mesh, self._geometry)
        cached = getattr(host, '_sfem_mr', None)
        if cached is None:
            cached = {'mesh': host}
            for k, v in mesh_results.items():
                if k.endswith('_markers') and hasattr(v, 'array'):
                    if k == 'domain_markers':
                        cached[k] = MeshMarkers(cell_marker_values(v).astype(np.int32), 2)
                    else:
                        cached[k] = MeshMarkers(marker_values(v, mesh, host).astype(np.int32), 1)
            host._sfem_mr = cached
            host._sfem_markers = {k: v for k, v in cached.items() if k != 'mesh'}
        return cached, host

    def _scalar_space(self, host):
        return FunctionSpace(host, 'CG', 2)

    # ---- entry points
    def stokes_solver(self, mesh_results, W, L_domain, H, mesh_type="sulcus"):
        from . import solvers
        mr, host = self._mesh_results(mesh_results)
        Vh, Qh = VectorFunctionSpace(host, 'P', 2), FunctionSpace(host, 'P', 1)
        u, p = solvers.stokes_solver(mr, FunctionSpace(host, MixedElement([Vh.ufl_element(), Qh.ufl_element()])), L_domain, H,
                                     mesh_type)
        # split(deepcopy=True) of the reference returns functions on the collapsed sub-spaces of W
        Vd, Qd = W.sub(0).collapse(), W.sub(1).collapse()
        return to_dolfin(u, Vd, self._make), to_dolfin(p, Qd, self._make)

    def stokes_solver_no_adv(self, V, Q):
        make = self._make
        if make is None:
            import dolfin
            make = dolfin.Function
        return make(V), make(Q)

    def _velocity(self, u, host):
        if u is None:
            return None
        if hasattr(u, 'function_space'):
            return from_dolfin(u, host, 'P2v')
        return u                                        # Constant((0, 0)) and the like

    def advdiff_solver(self, mesh_results, u, C, D, mu, mesh_type="sulcus"):
        from . import solvers
        mr, host = self._mesh_results(mesh_results)
        c = solvers.advdiff_solver(mr, self._velocity(u, host), self._scalar_space(host), float(D), float(mu), mesh_type)
        return to_dolfin(c, C, self._make)

    def advdiff_solver_variable_mu(self, mesh_results, u, C, D, mu_function, mesh_type="sulcus"):
        from . import solvers
        mr, host = self._mesh_results(mesh_results)
        c = solvers.advdiff_solver_variable_mu(mr, self._velocity(u, host), self._scalar_space(host), float(D), mu_function,
                                               mesh_type)
        return to_dolfin(c, C, self._make)

    def pure_diffusion_solver(self, mesh_results, C, D, mu, mesh_type="sulcus"):
        from . import solvers
        mr, host = self._mesh_results(mesh_results)
        c = solvers.pure_diffusion_solver(mr, self._scalar_space(host), float(D), float(mu), mesh_type)
        return to_dolfin(c, C, self._make)

    def pure_diffusion_solver_variable_mu(self, mesh_results, C, D, mu_function, mesh_type="rectangular", bottom_id=4, u=None):
        from . import solvers
        mr, host = self._mesh_results(mesh_results)
        c = solvers.pure_diffusion_solver_variable_mu(mr, self._scalar_space(host), float(D), mu_function, mesh_type,
                                                      bottom_id=bottom_id, u=self._velocity(u, host))
        return to_dolfin(c, C, self._make)


def install(module_name: str = 'solvers', **kw):
    """Register the dolfin-facing entry points as ``sys.modules[module_name]``: a later ``import simulation`` of the
    reference tree (``from solvers import stokes_solver, ...``, ``simulation.py:31-38``) then runs its Stokes /
    concentration solves on the device and goes on with genuine dolfin Functions."""
    s = DolfinSolvers(**kw)
    mod = types.ModuleType(module_name)
    for name in ('stokes_solver', 'stokes_solver_no_adv', 'pure_diffusion_solver', 'pure_diffusion_solver_variable_mu',
                 'advdiff_solver', 'advdiff_solver_variable_mu'):
        setattr(mod, name, getattr(s, name))
    mod.__doc__ = "sulcusfem.dolfin_adapter: B200 solve path behind the reference's solvers.py signatures"
    sys.modules[module_name] = mod
    return mod
