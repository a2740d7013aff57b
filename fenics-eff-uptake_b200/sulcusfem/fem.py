"""Minimal dolfin-shaped host objects that the reference's callers hand to / expect from solvers.py.

Only the members the reference actually touches are provided:
``Constant`` (``float(c)``), ``UserExpression`` (``eval(values, x)``, ``value_shape()``),
``FunctionSpace`` / ``VectorFunctionSpace`` / ``MixedElement`` (``simulation.py:128-130,146``),
``Function`` with ``vector().get_local()/set_local()/apply()``, ``function_space().mesh()``,
``set_allow_extrapolation`` and point evaluation ``f(Point)`` / ``f((x, y))``
(``solvers.py:87-105``, ``analysis.py:361,371,578``, ``plotting.py:332``).

A ``Function`` owns its nodal values twice: a host numpy array (what ``get_local`` returns) and,
lazily, device tensors used by the CUDA path; ``set_local`` invalidates the device copy.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from . import dofmap as dm
from .hostmesh import HostMesh


class Constant:
    def __init__(self, value):
        self._v = np.asarray(value, dtype=np.float64)

    def __float__(self):
        if self._v.ndim != 0:
            raise TypeError("vector Constant has no float value")
        return float(self._v)

    def values(self):
        return np.atleast_1d(self._v).copy()

    def __call__(self, *args):
        return float(self) if self._v.ndim == 0 else self._v.copy()


class UserExpression:
    """Base class for user coefficients: subclasses implement ``eval(values, x)``."""

    def __init__(self, **kwargs):
        self._degree = kwargs.get('degree', 2)      # dolfin's default for UserExpression (App. A.3)

    def value_shape(self):
        return ()

    def eval(self, values, x):     # pragma: no cover - interface
        raise NotImplementedError

    def __call__(self, x):
        x = np.asarray(getattr(x, 'array', lambda: x)() if hasattr(x, 'array') else x, dtype=np.float64)
        v = np.zeros(max(1, int(np.prod(self.value_shape() or (1,)))))
        self.eval(v, x)
        return float(v[0]) if self.value_shape() == () else v


def evaluate_expression(expr, pts: np.ndarray) -> np.ndarray:
    """Scalar coefficient at points: float / Constant / vectorised ``mu_at`` / per-point ``eval``."""
    pts = np.asarray(pts, dtype=np.float64)
    if np.isscalar(expr) or isinstance(expr, Constant):
        return np.full(len(pts), float(expr))
    if hasattr(expr, 'mu_at'):
        return np.asarray(expr.mu_at(pts[:, 0]), dtype=np.float64)
    out = np.empty(len(pts))
    v = np.zeros(1)
    for i, x in enumerate(pts):
        expr.eval(v, x)
        out[i] = v[0]
    return out


class Point:
    def __init__(self, *xy):
        self._x = np.asarray(xy, dtype=np.float64)

    def array(self):
        return self._x

    def x(self):
        return float(self._x[0])

    def y(self):
        return float(self._x[1])


class Element:
    def __init__(self, family, degree, shape=()):
        self.family_, self.degree_, self.shape = family, int(degree), tuple(shape)

    def family(self):
        return self.family_

    def degree(self):
        return self.degree_

    def value_shape(self):
        return self.shape


class MixedElement:
    def __init__(self, elements):
        self.elements = list(elements)

    def sub_elements(self):
        return self.elements


class _DofMap:
    def __init__(self, cell_dofs):
        self._cd = cell_dofs

    def cell_dofs(self, i):
        return self._cd[int(i)]


class FunctionSpace:
    """P1 / P2 scalar, P2 vector, or Taylor-Hood mixed space on a HostMesh."""

    def __init__(self, mesh: HostMesh, family_or_element, degree: Optional[int] = None, _parent=None, _index=None):
        self._mesh = mesh
        if isinstance(family_or_element, MixedElement):
            self.element = family_or_element
            e = family_or_element.elements
            ok = (len(e) == 2 and e[0].shape == (2,) and e[0].degree() == 2 and e[1].shape == () and e[1].degree() == 1)
            if not ok:
                raise ValueError("only the Taylor-Hood element [P2^2, P1] is supported as a mixed space")
            self.kind = 'TH'
        elif isinstance(family_or_element, Element):
            self.element = family_or_element
            self.kind = self._kind(family_or_element)
        else:
            fam = str(family_or_element)
            if fam not in ('CG', 'P', 'Lagrange'):
                raise ValueError(f"unsupported element family {fam!r}")
            self.element = Element('Lagrange', degree)
            self.kind = self._kind(self.element)
        self._parent, self._index = _parent, _index

    @staticmethod
    def _kind(e: Element):
        if e.shape == () and e.degree() in (1, 2):
            return f'P{e.degree()}'
        if e.shape == (2,) and e.degree() == 2:
            return 'P2v'
        raise ValueError("supported spaces: P1, P2, P2 vector, Taylor-Hood")

    def mesh(self):
        return self._mesh

    def ufl_element(self):
        return self.element

    def dim(self):
        m = self._mesh
        return {'P1': m.num_vertices, 'P2': dm.p2_num_dofs(m), 'P2v': 2 * dm.p2_num_dofs(m),
                'TH': dm.th_num_dofs(m)}[self.kind]

    def num_sub_spaces(self):
        return {'P1': 0, 'P2': 0, 'P2v': 2, 'TH': 2}[self.kind]

    def sub(self, i):
        if self.kind == 'TH':
            el = self.element.elements[i]
            return FunctionSpace(self._mesh, el, _parent=self, _index=i)
        if self.kind == 'P2v':
            return FunctionSpace(self._mesh, Element('Lagrange', 2), _parent=self, _index=i)
        raise ValueError("space has no sub spaces")

    def collapse(self):
        return FunctionSpace(self._mesh, self.element)

    def tabulate_dof_coordinates(self):
        m = self._mesh
        x2 = dm.p2_dof_coordinates(m)
        return {'P1': m.coords, 'P2': x2, 'P2v': np.concatenate([x2, x2]),
                'TH': np.concatenate([x2, x2, m.coords])}[self.kind]

    def dofmap(self):
        m = self._mesh
        if self.kind == 'P1':
            return _DofMap(dm.p1_cell_dofs(m))
        if self.kind == 'P2':
            return _DofMap(dm.p2_cell_dofs(m))
        if self.kind == 'P2v':
            c = dm.p2_cell_dofs(m).astype(np.int64)
            return _DofMap(np.concatenate([c, c + dm.p2_num_dofs(m)], axis=1))
        return _DofMap(dm.th_cell_dofs(m))


def VectorFunctionSpace(mesh, family, degree, dim=2):
    if dim != 2:
        raise ValueError("2-D vector spaces only")
    return FunctionSpace(mesh, Element('Lagrange', degree, (2,)))


class _Vector:
    def __init__(self, owner: "Function"):
        self._f = owner

    def get_local(self):
        return self._f.values.copy()

    def _edited(self):
        f = self._f
        f._dev = None                       # device copy and cached functionals describe the old values
        f._functionals = None
        f._version = getattr(f, '_version', 0) + 1

    def set_local(self, a):
        self._f.values[:] = np.asarray(a, dtype=np.float64)
        self._edited()

    def apply(self, mode):
        return None

    def zero(self):
        self._f.values[:] = 0.0
        self._edited()

    def size(self):
        return len(self._f.values)

    def __setitem__(self, key, value):
        self._f.values[key] = value
        self._edited()

    def __getitem__(self, key):
        return self._f.values[key]

    def min(self):
        return float(self._f.values.min())

    def max(self):
        return float(self._f.values.max())


class Function:
    def __init__(self, V: FunctionSpace, values: Optional[np.ndarray] = None):
        self._V = V
        self.values = np.zeros(V.dim()) if values is None else np.ascontiguousarray(values, dtype=np.float64)
        if len(self.values) != V.dim():
            raise ValueError("value array does not match the space dimension")
        self._dev = None
        self._version = 0
        self._functionals = None
        self._extrapolate = False
        self._locator = None

    def function_space(self):
        return self._V

    def vector(self):
        return _Vector(self)

    def set_allow_extrapolation(self, flag):
        self._extrapolate = bool(flag)

    def split(self, deepcopy=False):
        V = self._V
        m = V.mesh()
        n2 = dm.p2_num_dofs(m)
        if V.kind == 'TH':
            u = Function(VectorFunctionSpace(m, 'P', 2), self.values[:2 * n2].copy())
            p = Function(FunctionSpace(m, 'P', 1), self.values[2 * n2:].copy())
            return u, p
        if V.kind == 'P2v':
            return (Function(FunctionSpace(m, 'P', 2), self.values[:n2].copy()),
                    Function(FunctionSpace(m, 'P', 2), self.values[n2:].copy()))
        raise ValueError("cannot split a scalar function")

    # ------------------------------------------------------------------ device copy
    def device_components(self):
        """Device tensors of the scalar components (P2v -> (ux, uy); scalar -> (c,))."""
        if self._dev is None:
            from .device import Context
            ctx = Context.get()
            n2 = dm.p2_num_dofs(self._V.mesh())
            if self._V.kind == 'P2v':
                self._dev = (ctx.up(self.values[:n2], np.float64), ctx.up(self.values[n2:], np.float64))
            else:
                self._dev = (ctx.up(self.values, np.float64),)
        return self._dev

    # ------------------------------------------------------------------ point evaluation (device, batched)
    def eval_points(self, pts, tol=None):
        """Values at many points in one launch of ``sfem_eval_points`` (uniform-grid locator, sulcusfem/locator.py).

        Returns ``(values, inside)``: values ``[npts]`` (scalar spaces) or ``[npts, 2]`` (P2 vector), ``inside``
        the mask of points that lie in some cell -- what the reference tests with
        ``mesh.bounding_box_tree().compute_first_entity_collision(Point)`` before calling ``f(Point)``
        (``analysis.py:367-372``).  Points outside get value 0."""
        from .locator import locator_for, TOL
        kind = self._V.kind
        if kind == 'TH':
            raise ValueError("point evaluation of a mixed function is not supported")
        loc = locator_for(self._V.mesh())
        vals, cell = loc.eval(pts, self.device_components(), degree=1 if kind == 'P1' else 2,
                              tol=TOL if tol is None else tol)
        return (vals[0] if kind != 'P2v' else vals.T.copy()), cell >= 0

    # ------------------------------------------------------------------ point evaluation (host, one point)
    def __call__(self, *x):
        from .hierarchy import locate_points
        p = x[0] if len(x) == 1 else x
        p = np.asarray(p.array() if hasattr(p, 'array') else p, dtype=np.float64).reshape(1, -1)[:, :2]
        m = self._V.mesh()
        cell, lam = locate_points(m, p)
        lam, cell = lam[0], int(cell[0])
        if lam.min() < -1e-10 and not self._extrapolate:
            raise RuntimeError("point is outside the mesh (set_allow_extrapolation(True) to extrapolate)")
        kind = self._V.kind
        if kind == 'P1':
            return float(self.values[m.cells[cell]] @ lam)
        l0, l1, l2 = lam
        phi = np.array([l0 * (2 * l0 - 1), l1 * (2 * l1 - 1), l2 * (2 * l2 - 1), 4 * l1 * l2, 4 * l0 * l2, 4 * l0 * l1])
        d = dm.p2_cell_dofs(m)[cell]
        if kind == 'P2':
            return float(self.values[d] @ phi)
        if kind == 'P2v':
            n2 = dm.p2_num_dofs(m)
            return np.array([self.values[d] @ phi, self.values[d + n2] @ phi])
        raise ValueError("point evaluation of a mixed function is not supported")
