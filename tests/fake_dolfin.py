"""A minimal stand-in for the dolfin objects the adapter touches (tests only): a mesh with ``coordinates() / cells() /
topology()``, facet / cell MeshFunctions, and function spaces whose DOF numbering is a RANDOM permutation of the host
numbering with interleaved vector components -- like real dolfin, whose numbering is build dependent."""
import numpy as np


class FakeMesh:
    def __init__(self, host, rng):
        self._host = host
        self._facet_perm = rng.permutation(len(host.edges))          # dolfin numbers facets differently

    def coordinates(self):
        return self._host.coords.copy()

    def cells(self):
        return self._host.cells.copy()

    def init(self, dim):
        return None

    def topology(self):
        mesh = self

        class _T:
            def __call__(self, d0, d1):
                assert (d0, d1) == (1, 0)
                fv = mesh._host.edges[mesh._facet_perm][:, ::-1]     # reversed vertex order inside a facet, too
                return lambda: fv.ravel()
        return _T()


class FakeMeshFunction:
    def __init__(self, values):
        self._v = np.asarray(values)

    def array(self):
        return self._v


def facet_function(mesh: FakeMesh, host_values):
    return FakeMeshFunction(np.asarray(host_values)[mesh._facet_perm])


class _DofMap:
    def __init__(self, dofs):
        self._d = np.asarray(dofs)

    def dofs(self):
        return self._d


class FakeSpace:
    """dofs = permuted concatenation of component blocks; ``coords[i]`` = coordinate of dolfin dof i."""

    def __init__(self, mesh, blocks, rng, perm=None, parent=None, index=None):
        self._mesh = mesh
        self.blocks = blocks                        # list of coordinate arrays (one per scalar component)
        n = sum(len(b) for b in blocks)
        self.perm = rng.permutation(n) if perm is None else perm      # our blocked index -> dolfin dof
        self.coords = np.empty((n, 2))
        self.coords[self.perm] = np.concatenate(blocks)
        self._rng = rng

    def mesh(self):
        return self._mesh

    def dim(self):
        return len(self.coords)

    def tabulate_dof_coordinates(self):
        return self.coords.copy()

    def dofmap(self):
        return _DofMap(np.arange(self.dim()))

    def sub(self, i):
        return _SubSpace(self, [i])


class _SubSpace:
    def __init__(self, root, path):
        self._root, self._path = root, path

    def _range(self):
        # TH root: blocks [ux, uy, p]; sub(0) = velocity (blocks 0, 1), sub(1) = pressure (block 2); P2v root: sub(i) = block i
        r = self._root
        sizes = [len(b) for b in r.blocks]
        off = np.concatenate([[0], np.cumsum(sizes)])
        if len(r.blocks) == 3:
            if self._path == [0]:
                return [0, 1], off
            if self._path == [1]:
                return [2], off
            if self._path[0] == 0:
                return [self._path[1]], off
        return [self._path[-1]], off

    def sub(self, i):
        return _SubSpace(self._root, self._path + [i])

    def dofmap(self):
        blocks, off = self._range()
        idx = np.concatenate([self._root.perm[off[b]:off[b + 1]] for b in blocks])
        return _DofMap(np.sort(idx))

    def collapse(self):
        blocks, _ = self._range()
        return FakeSpace(self._root._mesh, [self._root.blocks[b] for b in blocks], self._root._rng)


class FakeVector:
    def __init__(self, n):
        self._v = np.zeros(n)

    def get_local(self):
        return self._v.copy()

    def set_local(self, a):
        self._v[:] = a

    def apply(self, mode):
        return None


class FakeFunction:
    def __init__(self, V):
        self._V = V
        self._vec = FakeVector(V.dim())

    def function_space(self):
        return self._V

    def vector(self):
        return self._vec


class FakeConstant:
    def __init__(self, v):
        self._v = float(v)

    def __float__(self):
        return self._v
