"""GPU tests of the study drivers (sulcusfem/studies.py): CSV schemas identical to the reference's checked-in files,
rows consistent with direct oracle solves."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'study_columns.json')))
H = 0.1            # coarse meshes: the schema and the arithmetic are what is tested here


def test_phase_b_rows_and_schema(tmp_path):
    from oracle import cpu_oracle as co
    from sulcusfem import studies, simulation
    from sulcusfem.parameters import Parameters, create_geometry_variations
    geos = create_geometry_variations(Parameters(mode='no-adv'), max_width=1.0)
    pick = {k: geos[k] for k in list(geos)[:2]}
    df = studies.run_no_adv_mu_sweep(str(tmp_path), mu_factors=[0.5, 1.0], geometries=pick, mesh_size_dim=H)
    assert list(df.columns) == GOLD['no_adv_mu_sweep_results.csv']['columns']
    assert len(df) == 4 and os.path.exists(tmp_path / 'no_adv_mu_sweep_results.csv')
    assert np.allclose(df['CR'], df['avg_conc_sulc'] / df['avg_conc_rect'], rtol=1e-14)
    assert np.allclose(df['flux_ratio'], df['flux_rect_bottom'] / df['flux_sulc_y0'], rtol=1e-14)
    # one row against the oracle on the very mesh the driver used
    row = df.iloc[0]
    p = studies._params_no_adv(row['mu_factor'], row['width_mm'], row['depth_mm'], H)
    mr = simulation._simulation_generate_mesh(p, 'rectangular')
    om = co.Mesh(mr['mesh'].coords, mr['mesh'].cells)
    c, _, _ = co.solve_concentration(om, mr['bc_markers'].values, p.D, mu=p.mu)
    mm = co.mass_metrics(om, c, 'rectangular')
    assert abs(mm['average_concentration'] - row['avg_conc_rect']) < 1e-10 * abs(row['avg_conc_rect'])


def test_advdiff_validation_schema_and_step_surrogate(tmp_path):
    from sulcusfem import studies
    df = studies.run_advdiff_step_validation(str(tmp_path), pe_values=[1.0], mu_factors=[1.0], mesh_size_dim=H)
    assert list(df.columns) == GOLD['advdiff_validation_step_pe_x_mu.csv']['columns']
    assert list(df['domain_type']) == ['rectangular', 'sulcus'] and list(df['surrogate_type']) == ['step_open', 'reference']
    rect, sulc = df.iloc[0], df.iloc[1]
    assert rect['mu_eff_open'] == sulc['mu_eff_open'] and np.isfinite(rect['flux_error_pct'])
    assert abs(rect['flux_ratio'] - rect['total_flux'] / sulc['total_flux']) < 1e-14
    assert abs(rect['CR'] - sulc['avg_conc'] / rect['avg_conc']) < 1e-14
    # the step surrogate reproduces the sulcus flux to within a few per cent (reference CSV: -1.2 % ... +0.04 %)
    assert abs(rect['flux_error_pct']) < 5.0
    meta = json.load(open(tmp_path / 'study_metadata.json'))
    assert meta['Pe_values'] == [1.0]


def test_phase_a_schemas(tmp_path):
    from sulcusfem import studies
    df = studies.run_mu_sweep(str(tmp_path), regimes={'small_uptake': [0.25, 1.0], 'high_uptake': [50.0]}, mesh_size_dim=H)
    assert list(df.columns) == GOLD['mu_parameter_sweep_results.csv']['columns']
    assert list(df['Config']) == ['small_uptake_mu_0.2x', 'small_uptake_mu_1.0x', 'high_uptake_mu_50.0x']
    assert np.all(np.diff(df['Mu']) > 0) and np.all(df['Ratio_Sim'] > 0)
    # closed forms are bit-exact against the reference CSV whatever the mesh (mu_eff_arc for w = h = 0.25, mu = 1)
    one = df[df['Mu_Factor'] == 1.0].iloc[0]
    assert one['Mu_Eff_Analytical'] == 1.0326223165338422
    cases = [c for c in studies.aspect_ratio_cases() if c[2] in (0.25, 0.5)][:3]
    da = studies.run_aspect_ratio_analysis(str(tmp_path), cases=cases, mesh_size_dim=H)
    assert list(da.columns) == GOLD['aspect_ratio_analysis_results.csv']['columns']
    assert len(da) == 3 and os.path.exists(tmp_path / 'aspect_ratio_analysis_results.csv')


@pytest.mark.timeout(900)
def test_advdiff_validation_against_reference_csv_at_reference_resolution():
    """The reference's own numbers (dolfin 2019.1 on Gmsh meshes, h = 0.02; checked-in CSV
    Advection-Diffusion/Results Data/advdiff_validation_step_pe_x_mu.csv) against this path on its own h = 0.02
    meshes, rows (Pe, mu) = (1, 1) and (10, 1).  The meshes differ (Gmsh's cannot be regenerated), so agreement is
    at mesh-discretisation tolerance.  Measured on B200 (profiles/r01_reference_csv_agreement.md): average
    concentration 1.4e-6 ... 9e-5, uptake flux 1.8e-6 ... 5.4e-5, y=0 / bottom flux (corner singularities at the
    inlet / floor) 1.8e-3 ... 2.8e-3, mu_eff_sim 1.8e-3 ... 2.8e-3, mu_eff_open 1e-3 ... 8.4e-3, mu_eff_arc bit-exact;
    the bounds below leave a factor 2-4 of margin."""
    from sulcusfem import studies
    ref = {(float(r['Pe']), float(r['mu_factor']), r['domain_type']): r
           for r in GOLD['advdiff_validation_step_pe_x_mu.csv']['rows']}
    df = studies.run_advdiff_step_validation(None, pe_values=[1.0, 10.0], mu_factors=[1.0], mesh_size_dim=0.02)
    report = []
    for _, row in df.iterrows():
        want = ref[(float(row['Pe']), float(row['mu_factor']), row['domain_type'])]
        for key, tol in (('avg_conc', 3e-4), ('uptake_flux', 2e-4), ('total_flux', 6e-3), ('diffusive_flux', 6e-3),
                         ('mu_eff_arc', 1e-15), ('mu_eff_sim', 6e-3), ('mu_eff_open', 2e-2)):
            got, w = float(row[key]), float(want[key])
            rel = abs(got - w) / abs(w)
            report.append((row['Pe'], row['domain_type'], key, got, w, rel))
            assert rel < tol, (row['Pe'], row['domain_type'], key, got, w, rel)
        if row['domain_type'] == 'rectangular':
            # the surrogate's flux error: reference +0.022 % (Pe = 1), +0.010 % (Pe = 10)
            report.append((row['Pe'], 'rect', 'flux_error_pct', float(row['flux_error_pct']), float(want['flux_error_pct']), 0.0))
            assert abs(float(row['flux_error_pct'])) < 0.1
            assert abs(float(row['CR']) - float(want['CR'])) < 5e-3
    for r in report:
        print("REFCSV", *r)


def test_no_uptake_geometry_study_schema_and_physics(tmp_path):
    """run_geometry_study (no_uptake_analysis.py:921-975): schema of geometry_comparison_results.csv, mu = 0 physics
    (no flux leaves through the walls: inlet + outlet flux ~ 0; uniform c = 1 would be the Pe -> 0 limit)."""
    from sulcusfem import studies
    from sulcusfem.parameters import Parameters, create_geometry_variations
    geos = create_geometry_variations(Parameters(mode='no-uptake'), max_width=1.0)
    pick = {k: geos[k] for k in ('largest', 'square_small') if k in geos} or {k: geos[k] for k in list(geos)[:2]}
    df = studies.run_geometry_study(str(tmp_path), peclet_numbers=(1.0, 10.0), geometries=pick, mesh_size_dim=H)
    assert list(df.columns) == GOLD['geometry_comparison_results.csv']['columns']
    assert list(df['Domain']) == ['sulcus'] * (2 * len(pick)) + ['rectangle'] * 2
    assert os.path.exists(tmp_path / 'geometry_comparison_results.csv')
    s = df[df['Domain'] == 'sulcus']
    assert np.all(np.abs(s['Inlet-Outlet Flux'].astype(float)) < 5e-2)        # = minus the (weakly zero) wall flux: O(h^2)
    assert np.all(np.abs(s['Mouth Net Check'].astype(float) - s['Mouth_Flux_Total'].astype(float)) < 1e-10)
    assert np.all(s['Mouth Q_in'].astype(float) > 0) and np.all(s['Mouth Q_out'].astype(float) > 0)
    assert np.allclose(s['Intradomain_Enrichment'].astype(float),
                       s['Sulcus Avg Concentration'].astype(float) / s['Main Channel Avg Concentration'].astype(float))
    r = df[df['Domain'] == 'rectangle']
    # Poiseuille: max |u_x| on the mid-channel line is 1 (reference CSV: 1.0000000000002331)
    assert np.all(np.abs(r['Max_Ux_mid_channel'].astype(float) - 1.0) < 1e-9)
    assert np.all(np.isfinite(s['VR_mid_max'].astype(float)))


@pytest.mark.timeout(900)
def test_concentration_profiles_against_reference_samples(tmp_path):
    """Point values of c along the reference's profile lines: the reference checked in 400-sample line profiles of
    the no-uptake study for two geometries (dolfin on Gmsh meshes, h = 0.02); this path evaluates the same lines on
    its own h = 0.02 meshes through sfem_eval_points.  Field-level agreement at mesh-discretisation tolerance."""
    from sulcusfem import studies
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'profile_samples.json')))
    out = studies.run_profile_export(str(tmp_path), geometry_keys=('largest', 'square_small'), mesh_size_dim=0.02)
    worst = {}
    for gkey, lines in gold['geometries'].items():
        df, dfs = out[gkey]
        assert list(df.columns) == gold['columns']
        assert os.path.exists(tmp_path / f'profiles_samples_{gkey}.csv') and os.path.exists(tmp_path / f'profiles_{gkey}.csv')
        for L in lines:
            sel = df[(df['Peclet'] == L['peclet']) & (df['LineName'] == L['name'])]
            # (the checked-in CSV predates the current analysis.py: its mouth-level line sits at y = 0, today's at 1e-6 H)
            assert abs(float(sel['y'].iloc[0]) - L['y']) <= 1e-6 + 1e-12
            # the in-cavity line is cut by the cavity wall: the two meshes may disagree on one end sample
            assert abs(len(sel) - L['n_valid']) <= (2 if L['name'] == 'sulcus_mid' else 0), (gkey, L['name'], len(sel), L['n_valid'])
            xs, cs = sel['x'].to_numpy(), sel['c'].to_numpy()
            j = np.searchsorted(xs, np.array(L['x']) - 1e-9)
            ok = (j < len(xs)) & (np.abs(xs[np.minimum(j, len(xs) - 1)] - np.array(L['x'])) < 1e-9)
            assert ok.sum() >= len(L['x']) - 2
            d = np.abs(cs[j[ok]] - np.array(L['c'])[ok])
            key = (gkey, L['peclet'], L['name'])
            worst[key] = float(d.max())
    for k, v in sorted(worst.items()):
        print("REFPROFILE", k, v)
    # measured on B200 (profiles/r01_reference_csv_agreement.md): 1e-11 ... 3e-6 on every line except the mouth-level line
    # of the largest cavity (8.8e-5: it runs through the two re-entrant mouth corners)
    assert max(worst.values()) < 5e-4
    assert max(v for k, v in worst.items() if k[2] != 'mouth_level') < 2e-5


def test_mu_eff_analysis_against_reference_csv(tmp_path):
    """run_mu_eff_analysis (no_advection_analysis_A.py:1583-1682) at the reference's resolution: the rows of the
    reference's checked-in mu_eff_analysis_results.csv (BASELINE config 1: 0.5 x 1.0 mm sulcus, mu x {0.1, 1, 10};
    dolfin on a Gmsh mesh that cannot be regenerated) -- closed forms bit-exact, simulated values to mesh tolerance."""
    from sulcusfem import studies
    g = GOLD['mu_eff_analysis_results.csv']
    df = studies.run_mu_eff_analysis(str(tmp_path), mesh_size_dim=0.02)
    assert list(df.columns) == g['columns'] and len(df) == 3
    assert os.path.exists(tmp_path / 'mu_eff_analysis_results.csv')
    for (_, row), ref in zip(df.iterrows(), g['rows']):
        assert row['Config'] == ref['Config']
        assert row['Mu_Eff_Analytical'] == float(ref['Mu_Eff_Analytical'])            # scipy quad closed form: bit-exact
        assert abs(row['Mu_Eff_Enhanced'] - float(ref['Mu_Eff_Enhanced'])) <= 2e-16 * abs(float(ref['Mu_Eff_Enhanced'])) + 1e-18
        assert abs(row['Mu_Eff_Simulation'] / float(ref['Mu_Eff_Simulation']) - 1.0) < 2e-2     # mesh-dependent (corner singularities; 1.1 % at mu = 10)
        assert abs(row['Mu_Eff_Opening'] / float(ref['Mu_Eff_Opening']) - 1.0) < 5e-2
        assert abs(row['Ratio_Sim'] - row['Mu_Eff_Simulation'] / row['Mu_base_nondim']) < 1e-14
        assert row['Mu_X_Array'] == ref['Mu_X_Array']


def test_geometry_analysis_rows(tmp_path):
    """run_geometry_analysis (no_advection_analysis_A.py:1463-1581): geometry x mu cases, device problems of a geometry
    reused across its mu values; one row against the oracle on the mesh the driver used."""
    from oracle import cpu_oracle as co
    from sulcusfem import studies, simulation
    from sulcusfem.parameters import Parameters, create_geometry_variations
    geos = create_geometry_variations(Parameters(mode='no-adv'))
    pick = {k: geos[k] for k in list(geos)[:2]}
    df = studies.run_geometry_analysis(str(tmp_path), mu_factors=[0.1, 10], geometries=pick, mesh_size_dim=H)
    assert len(df) == 4 and os.path.exists(tmp_path / 'geometry_analysis_results.csv')
    assert list(df.columns)[:8] == ['Config', 'Geometry_Name', 'Mu_Value', 'Mu_Factor', 'Sulcus_Width_mm', 'Sulcus_Depth_mm',
                                    'Aspect_Ratio', 'Aspect_Ratio_Category']
    assert list(df['Config']) == [f"{g}_mu_{f}" for g in pick for f in (0.1, 10)]
    row = df.iloc[1]
    p = Parameters(mode='no-adv', mesh_size_dim=H)
    p.sulci_w_dim, p.sulci_h_dim = row['Sulcus_Width_mm'], row['Sulcus_Depth_mm']
    p.mu_dim = row['Mu_Value']
    p.validate()
    p.nondim()
    mr = simulation._simulation_generate_mesh(p, 'sulcus')
    om = co.Mesh(mr['mesh'].coords, mr['mesh'].cells)
    mk = {k: mr[k].values for k in ('bc_markers', 'bottom_segment_markers', 'y0_markers', 'domain_markers')}
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], p.D, mu=p.mu)
    fl = co.flux_metrics(om, mk, 'sulcus', p.D, c, mu=p.mu)
    me = co.mu_eff_metrics(fl, p.L, p.sulci_h, p.sulci_w, p.mu)
    assert abs(row['Mu_Eff_Simulation'] - me['mu_eff_sim']) <= 1e-8 * abs(me['mu_eff_sim'])
    assert abs(row['Mu_Eff_Opening'] - me['mu_eff_open']) <= 1e-8 * abs(me['mu_eff_open'])
    assert abs(row['Total_Mass'] - co.mass_metrics(om, c, 'sulcus', mk['domain_markers'])['total_mass']) < 1e-10


def test_mu_sweep_with_frozen_coarse_levels_and_concurrent_streams():
    """The mu sweep keeps the multigrid levels of a nearby mu (solvers.frozen_coarse_levels) and can keep several cases
    in flight (sweep.run_concurrent): both are throughput devices only -- every row must equal the plain serial sweep to
    solver tolerance, and CG must not need many more iterations behind the frozen levels."""
    from sulcusfem import studies, solvers, simulation
    from sulcusfem.parameters import Parameters
    regimes = {'small_uptake': [0.1, 0.25, 0.5, 1.0, 2.0, 3.0], 'high_uptake': [50.0, 100.0, 150.0]}
    plain = studies.run_mu_sweep(None, regimes=regimes, mesh_size_dim=H, frozen_coarse=False, batch=False)
    frozen = studies.run_mu_sweep(None, regimes=regimes, mesh_size_dim=H, frozen_coarse=True, batch=False)
    conc = studies.run_mu_sweep(None, regimes=regimes, mesh_size_dim=H, frozen_coarse=True, streams=3, batch=False)
    # batched Krylov loops (the default): 9 coefficients = one batch of 8 + one of 1; serial and with 3 cases in flight
    batched = studies.run_mu_sweep(None, regimes=regimes, mesh_size_dim=H)
    batched_conc = studies.run_mu_sweep(None, regimes=regimes, mesh_size_dim=H, streams=3)
    assert list(plain['Config']) == list(frozen['Config']) == list(conc['Config']) == list(batched['Config'])
    for col in ('Mu_Eff_Simulation', 'Mu_Eff_Opening', 'Total_Mass', 'Mouth_Flux_Total'):
        for other in (frozen, conc, batched, batched_conc):
            assert np.allclose(other[col], plain[col], rtol=1e-9, atol=1e-13), col
    # iteration counts behind frozen levels: solve mu = 3 with levels built for mu = 1 (ratio 3 < 4) and compare
    p = Parameters(mode='no-adv', mesh_size_dim=H)
    p.sulci_w_dim = p.sulci_h_dim = 0.25
    base = float(getattr(Parameters, 'MU_DIM_NO_ADV'))
    its = {}
    for frozen_mode in (False, True):
        out = []
        for f in (1.0, 3.0):
            p.mu_dim = base * f
            p.validate()
            p.nondim()
            with solvers.frozen_coarse_levels(4.0 if frozen_mode else 1.0):
                r = simulation.run_simulation('no-adv', 'T', 'c', 'sulcus', p)
            out.append(r['c'].solver_info['iterations'])
        its[frozen_mode] = out
    assert its[True][1] <= its[False][1] + 4, its
