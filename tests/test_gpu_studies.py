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
