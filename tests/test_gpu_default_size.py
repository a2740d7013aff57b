"""GPU parity at the REFERENCE'S DEFAULT MESH SIZE (mesh_size = 0.02, parameters.py:108-116: ~120 k P2 /
~270 k Taylor-Hood dofs) against the oracle's sparse LU on the same mesh -- BASELINE configs 0 / 1 / 2 through the
reference-facing front end (sulcusfem.simulation -> solvers / analysis -> C ABI -> CUDA).

This is the size at which the full multigrid depth, the non-nested coarse levels and the sliced-ELL engine at its
natural threshold are in play (the toy-mesh tests force the engines instead).  The LU of the Taylor-Hood system takes
~20-50 s of one host core per case.  Bars (north_star): fields 1e-10 relative L2, functionals 1e-9, mu_eff 1e-8.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
H_DEFAULT = 0.02
FIELD_TOL, P_TOL, FUNC_TOL, MUEFF_TOL = 1e-10, 1e-9, 1e-9, 1e-8


def _params(mode, **kw):
    from sulcusfem.parameters import Parameters
    p = Parameters(mode=mode, mesh_size_dim=H_DEFAULT, sulci_w_dim=0.5, sulci_h_dim=1.0, **kw)
    p.validate()
    p.nondim()
    return p


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def _flat(d, prefix=''):
    out = {}
    for k, v in d.items():
        if str(k).startswith('_'):
            continue
        if isinstance(v, dict):
            out.update(_flat(v, prefix + str(k) + '.'))
        elif isinstance(v, (int, float, np.floating)):
            out[prefix + str(k)] = float(v)
    return out


def _check_dict(got, want, tol):
    got, want = _flat(got), _flat(want)
    scale = max(1e-3, max(abs(v) for v in want.values() if np.isfinite(v)))
    missing = [k for k in want if k not in got]
    assert not missing, missing
    for k, w in want.items():
        if np.isfinite(w):
            assert abs(got[k] - w) <= tol * max(abs(w), 1e-3 * scale), (k, got[k], w)


def _oracle_mesh(res):
    from oracle import cpu_oracle as co
    mr = res['mesh_results']
    mesh = mr['mesh']
    keys = [k for k in ('bc_markers', 'bottom_segment_markers', 'y0_markers', 'domain_markers') if k in mr]
    return co, co.Mesh(mesh.coords, mesh.cells), {k: mr[k].values for k in keys}


def test_config0_no_advection_sulcus_default_size():
    from sulcusfem import simulation
    p = _params('no-adv')
    res = simulation.run_simulation('no-adv', 'Default', 'c0', 'sulcus', p)
    co, om, mk = _oracle_mesh(res)
    assert 100_000 < om.n_p2 < 160_000                       # the reference's default problem size
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], 1.0, mu=p.mu)
    assert _rel(res['c'].vector().get_local(), c) <= FIELD_TOL
    fl = co.flux_metrics(om, mk, 'sulcus', 1.0, c, mu=p.mu)
    _check_dict(res['flux_metrics'], fl, FUNC_TOL)
    _check_dict(res['mass_metrics'], co.mass_metrics(om, c, 'sulcus', mk['domain_markers']), FUNC_TOL)
    me = co.mu_eff_metrics(fl, p.L, p.sulci_h, p.sulci_w, p.mu)
    got = res['mu_eff_comparison']
    for k in ('mu_eff_arc', 'mu_eff_enh', 'mu_eff_sim', 'mu_eff_open'):
        assert abs(got[k] - me[k]) <= MUEFF_TOL * abs(me[k]), (k, got[k], me[k])


def test_config1_advection_diffusion_sulcus_default_size():
    """BASELINE configs[1] exactly as the reference runs it: default mesh, default Pe = 40, mu = 1."""
    from sulcusfem import simulation
    p = _params('adv-diff')
    assert abs(p.Pe - 40.0) < 1e-12
    res = simulation.run_simulation('adv-diff', 'Default', 'c1', 'sulcus', p)
    co, om, mk = _oracle_mesh(res)
    ux, uy, pr, _, _ = co.solve_stokes(om, mk['bc_markers'], p.H)
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], p.D, mu=p.mu, ux=ux, uy=uy)
    n2 = om.n_p2
    u = res['u'].vector().get_local()
    assert _rel(u, np.concatenate([ux, uy])) <= FIELD_TOL
    assert _rel(u[:n2], ux) <= FIELD_TOL
    assert np.linalg.norm(u[n2:] - uy) <= FIELD_TOL * np.linalg.norm(ux)      # |u_y| << |u_x|: same absolute accuracy
    assert _rel(res['p'].vector().get_local(), pr) <= P_TOL
    assert _rel(res['c'].vector().get_local(), c) <= FIELD_TOL
    fl = co.flux_metrics(om, mk, 'sulcus', p.D, c, ux, uy, mu=p.mu)
    _check_dict(res['flux_metrics'], fl, FUNC_TOL)
    _check_dict(res['mass_metrics'], co.mass_metrics(om, c, 'sulcus', mk['domain_markers']), FUNC_TOL)
    me = co.mu_eff_metrics(fl, p.L, p.sulci_h, p.sulci_w, p.mu)
    got = res['mu_eff_comparison']
    for k in ('mu_eff_sim', 'mu_eff_open'):
        assert abs(got[k] - me[k]) <= MUEFF_TOL * abs(me[k]), (k, got[k], me[k])
    assert res['u'].solver_info['converged'] and res['c'].solver_info['converged']


def test_config2_rectangle_with_step_uptake_default_size():
    """adv_diff_analysis.py:144-178 at the reference's mesh size: rectangle surrogate, StepUptakeOpen mu(x), Pe = 10."""
    from sulcusfem import simulation
    from sulcusfem.parameters import StepUptakeOpen
    p = _params('adv-diff', U_ref_dim=0.003)
    step = dict(mu_base=1.0, mu_eff_target=1.7700044654465237, xL=4.75, xR=5.25, L_c=0.05, Gamma=5.0)
    p.mu = p.mu_dim = StepUptakeOpen(mu_base=step['mu_base'], mu_eff_target=step['mu_eff_target'], sulcus_left_x=step['xL'],
                                     sulcus_right_x=step['xR'], L_c=step['L_c'], Gamma=step['Gamma'], degree=2)
    res = simulation.run_simulation('adv-diff', 'Default', 'c2', 'rectangular', p, mu_variable=True)
    co, om, mk = _oracle_mesh(res)
    ux, uy, pr, _, _ = co.solve_stokes(om, mk['bc_markers'], p.H)
    mun = co.interpolate_p2(om, co.StepUptakeOpen(step['mu_base'], step['mu_eff_target'], step['xL'], step['xR'],
                                                   L_c=step['L_c'], Gamma=step['Gamma']))
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], p.D, mu_nodal=mun, ux=ux, uy=uy)
    n2 = om.n_p2
    u = res['u'].vector().get_local()
    assert _rel(u[:n2], ux) <= FIELD_TOL
    assert abs(u[:n2].max() - 1.0) < 1e-10                  # exact Poiseuille (reference CSV: 1.0000000000002331)
    assert _rel(res['p'].vector().get_local(), pr) <= P_TOL
    assert _rel(res['c'].vector().get_local(), c) <= FIELD_TOL
    _check_dict(res['flux_metrics'], co.flux_metrics(om, mk, 'rectangular', p.D, c, ux, uy, mu_nodal=mun), FUNC_TOL)
    _check_dict(res['mass_metrics'], co.mass_metrics(om, c, 'rectangular'), FUNC_TOL)


def test_pure_diffusion_variable_mu_with_advecting_velocity():
    """pure_diffusion_solver_variable_mu(..., u=<Stokes velocity>): the optional advection argument of
    solvers.py:176-231 (`u` defaults to Constant((0,0)), :200-201) with the quadrature-point clamp of mu."""
    from sulcusfem import solvers, hostmesh as hm
    from sulcusfem.fem import Constant, FunctionSpace, MixedElement, UserExpression, VectorFunctionSpace
    from sulcusfem.unstructured import mesh_domain
    from oracle import cpu_oracle as co
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.08, 'rectangular')
    mr = {'mesh': mesh}
    mr.update(hm.build_markers(mesh, 10.0, 1.0, 4.75, 5.25, 'rectangular'))
    V, Q = VectorFunctionSpace(mesh, 'P', 2), FunctionSpace(mesh, 'P', 1)
    W = FunctionSpace(mesh, MixedElement([V.ufl_element(), Q.ufl_element()]))
    u, _ = solvers.stokes_solver(mr, W, 10.0, 1.0, 'rectangular')

    class Mu(UserExpression):
        def eval(self, values, x):
            values[0] = 0.5 + np.cos(3.0 * x[0])

        def value_shape(self):
            return ()
    D = 0.1
    Csp = FunctionSpace(mesh, 'CG', 2)
    c = solvers.pure_diffusion_solver_variable_mu(mr, Csp, Constant(D), Mu(degree=2), 'rectangular', bottom_id=4, u=u)
    om = co.Mesh(mesh.coords, mesh.cells)
    bm = mr['bc_markers'].values
    ux, uy, _, _, _ = co.solve_stokes(om, bm, 1.0)
    X = om.p2_dof_coords()
    want, _, _ = co.solve_concentration(om, bm, D, mu_nodal=0.5 + np.cos(3.0 * X[:, 0]), ux=ux, uy=uy, clamp_mu=True)
    assert _rel(c.vector().get_local(), want) <= FIELD_TOL
    # and the default (u=None) is the no-advection solve, different from the advected one
    c0 = solvers.pure_diffusion_solver_variable_mu(mr, Csp, Constant(D), Mu(degree=2), 'rectangular')
    want0, _, _ = co.solve_concentration(om, bm, D, mu_nodal=0.5 + np.cos(3.0 * X[:, 0]), clamp_mu=True)
    assert _rel(c0.vector().get_local(), want0) <= FIELD_TOL
    assert _rel(c0.vector().get_local(), want) > 1e-3
