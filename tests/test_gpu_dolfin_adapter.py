"""SURVEY 8(b)(ii) on the GPU: the six solvers.py entry points called with dolfin-LIKE objects (a mesh, facet
MeshFunctions and function spaces whose numberings are random permutations, tests/fake_dolfin.py) must return functions
on THOSE spaces whose values, mapped back by coordinate, equal the oracle's LU solution -- the call sequence is the one
of the reference's simulation.py:122-166 (W.sub(i).collapse() spaces, u handed on to advdiff_solver)."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def test_reference_call_sequence_with_dolfin_like_objects():
    import fake_dolfin as fd
    from oracle import cpu_oracle as co
    from sulcusfem import dofmap as dm, dolfin_adapter as da, hostmesh as hm
    from sulcusfem.unstructured import mesh_domain
    rng = np.random.default_rng(11)
    host0 = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.1, 'sulcus')
    mk = hm.build_markers(host0, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    dmesh = fd.FakeMesh(host0, rng)
    mesh_results = {'mesh': dmesh, 'mesh_info': {}}
    for k in ('bc_markers', 'bottom_segment_markers', 'y0_markers'):
        mesh_results[k] = fd.facet_function(dmesh, mk[k].values)
    mesh_results['domain_markers'] = fd.FakeMeshFunction(mk['domain_markers'].values)
    P2, P1 = dm.p2_dof_coordinates(host0), host0.coords
    W = fd.FakeSpace(dmesh, [P2, P2, P1], rng)
    C = fd.FakeSpace(dmesh, [P2], rng)
    s = da.DolfinSolvers(make_function=fd.FakeFunction, geometry=host0.geometry)
    # simulation.py:133 -> stokes_solver(mesh_results, W, L, H, domain_type)
    u, p = s.stokes_solver(mesh_results, W, 10.0, 1.0, 'sulcus')
    D, mu = 1.0 / 40.0, 1.0
    # simulation.py:160-163 -> advdiff_solver(mesh_results, u, C, Constant(D), Constant(mu), domain_type)
    c = s.advdiff_solver(mesh_results, u, C, fd.FakeConstant(D), fd.FakeConstant(mu), 'sulcus')
    c0 = s.pure_diffusion_solver(mesh_results, C, fd.FakeConstant(1.0), fd.FakeConstant(mu), 'sulcus')
    om = co.Mesh(host0.coords, host0.cells)
    bm = mk['bc_markers'].values
    ux, uy, pr, _, _ = co.solve_stokes(om, bm, 1.0)
    cr, _, _ = co.solve_concentration(om, bm, D, mu=mu, ux=ux, uy=uy)
    c0r, _, _ = co.solve_concentration(om, bm, 1.0, mu=mu)
    Vd = u.function_space()
    assert Vd.dim() == 2 * om.n_p2 and p.function_space().dim() == om.nv
    assert _rel(u.vector().get_local()[Vd.perm], np.concatenate([ux, uy])) < 1e-10
    assert _rel(p.vector().get_local()[p.function_space().perm], pr) < 1e-9
    assert c.function_space() is C
    assert _rel(c.vector().get_local()[C.perm], cr) < 1e-10
    assert _rel(c0.vector().get_local()[C.perm], c0r) < 1e-10
