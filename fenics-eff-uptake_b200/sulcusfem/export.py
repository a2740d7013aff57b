"""ParaView export ``File(path) << f`` (reference ``simulation.py:137-138,165``: velocity.pvd, pressure.pvd,
concentration.pvd) -- SURVEY 8(f)-4, pure host I/O.

dolfin's VTK writer stores a Lagrange function by its vertex values on the linear mesh (one ``.vtu`` per time step
plus a ``.pvd`` collection); the same layout is written here: ``<stem>.pvd`` referencing ``<stem>000000.vtu`` with
Points (z = 0), triangle Cells and one PointData array (scalar, or a 3-component vector with w = 0) named after
the function.  ASCII with 17 significant digits, so values read back bit-exactly.
"""
from __future__ import annotations

import os

import numpy as np


def _fmt(a):
    return ' '.join(repr(float(v)) for v in np.asarray(a, dtype=np.float64).ravel())


def vertex_values(f):
    """[nv] (scalar spaces) or [nv, 2] (P2 vector): P2 vertex dofs are the values at the vertices."""
    V = f.function_space()
    nv = V.mesh().num_vertices
    if V.kind in ('P1', 'P2'):
        return f.values[:nv].copy()
    if V.kind == 'P2v':
        n2 = len(f.values) // 2
        return np.stack([f.values[:nv], f.values[n2:n2 + nv]], axis=1)
    raise ValueError("split a mixed function before exporting it")


class File:
    def __init__(self, path: str):
        if not path.endswith('.pvd'):
            raise ValueError("only .pvd output is supported")
        self.path = path
        self.count = 0
        self._entries = []

    def __lshift__(self, obj):
        f, t = (obj if isinstance(obj, tuple) else (obj, float(self.count)))
        mesh = f.function_space().mesh()
        stem = os.path.splitext(os.path.basename(self.path))[0]
        vtu = f"{stem}{self.count:06d}.vtu"
        d = os.path.dirname(self.path) or '.'
        os.makedirs(d, exist_ok=True)
        name = getattr(f, 'name_', None) or 'f'
        vals = vertex_values(f)
        nv, nc = mesh.num_vertices, mesh.num_cells
        pts = np.concatenate([mesh.coords, np.zeros((nv, 1))], axis=1)
        with open(os.path.join(d, vtu), 'w') as out:
            out.write('<?xml version="1.0"?>\n<VTKFile type="UnstructuredGrid" version="0.1">\n<UnstructuredGrid>\n')
            out.write(f'<Piece NumberOfPoints="{nv}" NumberOfCells="{nc}">\n')
            out.write('<Points>\n<DataArray type="Float64" NumberOfComponents="3" format="ascii">')
            out.write(_fmt(pts) + '</DataArray>\n</Points>\n<Cells>\n')
            out.write('<DataArray type="UInt32" Name="connectivity" format="ascii">'
                      + ' '.join(map(str, mesh.cells.ravel().tolist())) + '</DataArray>\n')
            out.write('<DataArray type="UInt32" Name="offsets" format="ascii">'
                      + ' '.join(map(str, range(3, 3 * nc + 1, 3))) + '</DataArray>\n')
            out.write('<DataArray type="UInt8" Name="types" format="ascii">' + ' '.join(['5'] * nc) + '</DataArray>\n</Cells>\n')
            if vals.ndim == 1:
                out.write(f'<PointData Scalars="{name}">\n<DataArray type="Float64" Name="{name}" format="ascii">')
                out.write(_fmt(vals))
            else:
                v3 = np.concatenate([vals, np.zeros((nv, 1))], axis=1)
                out.write(f'<PointData Vectors="{name}">\n<DataArray type="Float64" Name="{name}" NumberOfComponents="3" format="ascii">')
                out.write(_fmt(v3))
            out.write('</DataArray>\n</PointData>\n</Piece>\n</UnstructuredGrid>\n</VTKFile>\n')
        self._entries.append((t, vtu))
        self.count += 1
        with open(self.path, 'w') as out:
            out.write('<?xml version="1.0"?>\n<VTKFile type="Collection" version="0.1">\n<Collection>\n')
            for tt, fn in self._entries:
                out.write(f'<DataSet timestep="{tt}" part="0" file="{fn}" />\n')
            out.write('</Collection>\n</VTKFile>\n')
        return self
