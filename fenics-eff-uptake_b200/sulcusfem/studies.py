"""Data paths of the reference's study drivers -- the loops over independent cases, the row extraction and the CSV
files with the reference's exact column names -- without the plotting / menu layers.

    reference driver                                         here
    no_advection_analysis_B.py:86-200  run_no_adv_mu_sweep   run_no_adv_mu_sweep      no_adv_mu_sweep_results.csv
    adv_diff_analysis.py:177-300       run_advdiff_step_...  run_advdiff_step_validation  advdiff_validation_step_pe_x_mu.csv
    no_advection_analysis_A.py:1257-1347  run_mu_sweep       run_mu_sweep             mu_parameter_sweep_results.csv
    no_advection_analysis_A.py:1349-1452  run_aspect_ratio_analysis  run_aspect_ratio_analysis  aspect_ratio_analysis_results.csv
    no_advection_analysis_A.py:1463-1581  run_geometry_analysis  run_geometry_analysis    geometry_analysis_results.csv
    no_advection_analysis_A.py:1583-1682  run_mu_eff_analysis    run_mu_eff_analysis      mu_eff_analysis_results.csv
    no_uptake_analysis.py:50-313,921-975  run_geometry_study  run_geometry_study      geometry_comparison_results.csv

The reference runs every case serially; the cases are independent, so each driver deals them round-robin to the
ranks (one process per GPU, ``sweep.run_sharded``) and gathers only the small row dictionaries -- no data-path
collective (BASELINE config 4).  On one GPU a driver reuses the device problems of a geometry across its mu / Pe
values through the per-mesh caches of ``simulation`` / ``solvers``.

Every driver takes ``mesh_size_dim`` (default: the reference's 0.02) so tests and smoke runs can use coarse meshes,
and ``rank`` / ``world`` (default: read from ``torch.distributed`` when initialised).
"""
from __future__ import annotations

import contextlib
import io
import json
import os
import threading
import time
from typing import Dict, Iterable, List, Optional, Sequence

import numpy as np

from .parameters import Parameters, StepUptakeOpen, create_geometry_variations
from .simulation import run_simulation
from .solvers import frozen_coarse_levels
from .sweep import run_sharded


def _world(rank, world):
    if rank is not None and world is not None:
        return int(rank), int(world)
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(), dist.get_world_size()
    except Exception:
        pass
    return 0, 1


def _run(quiet, **kw):
    from .sweep import current_slot
    if not quiet or current_slot() != 0 or threading.current_thread() is not threading.main_thread():
        return run_simulation(**kw)             # worker threads: stdout is redirected once, around the whole pool
    with contextlib.redirect_stdout(io.StringIO()):
        return run_simulation(**kw)


def _sharded(cases, one, rank, world, streams, quiet):
    """``sweep.run_sharded`` with the study's stdout handling: with several streams the (process-global) redirect is
    set once around the worker pool instead of once per case."""
    if streams > 1 and quiet:
        with contextlib.redirect_stdout(io.StringIO()):
            return run_sharded(cases, one, rank, world, streams=streams)
    return run_sharded(cases, one, rank, world, streams=streams)


def _prefetch(jobs, rank=0, world=1, enable=True):
    """Hand the (params, domain_type) pairs of this rank's cases to ``simulation.prefetch_meshes`` (parallel host
    meshing); ``jobs`` is a list of lists, one per case, in case order."""
    if not enable:
        return 0
    from .simulation import prefetch_meshes
    mine = [j for i, case_jobs in enumerate(jobs) if i % world == rank for j in case_jobs]
    return prefetch_meshes(mine)


def _frame(rows: List[dict], sort: Optional[Sequence[str]] = None):
    import pandas as pd
    df = pd.DataFrame(rows)
    if sort and len(df):
        df = df.sort_values(list(sort)).reset_index(drop=True)
    return df


def _save(df, directory, name, meta=None, rank=0):
    if directory is None or rank != 0:
        return None
    os.makedirs(directory, exist_ok=True)
    path = os.path.join(directory, name)
    df.to_csv(path, index=False)
    if meta is not None:
        with open(os.path.join(directory, "study_metadata.json"), 'w') as f:
            json.dump(meta, f, indent=4)
    return path


# ====================================================================== Phase B: no advection, mu x geometry
MU_FACTORS_PHASE_B = [0.1, 0.5, 1.0]                       # no_advection_analysis_B.py:34


def _params_no_adv(mu_factor, w_dim, h_dim, mesh_size_dim):
    p = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
    p.mu_dim = float(getattr(Parameters, 'MU_DIM_NO_ADV', p.mu_dim)) * float(mu_factor)   # :48-50
    p.sulci_w_dim, p.sulci_h_dim = w_dim, h_dim
    p.validate()
    p.nondim()
    return p


def _flux_of(results, domain_type):                       # no_advection_analysis_B.py:55-66
    fm = results.get('flux_metrics') or {}
    if domain_type == 'sulcus':
        pf = (fm.get('sulcus_specific') or {}).get('physical_flux') or {}
        for key in ('y0_flux', 'y0_combined'):
            if key in pf and isinstance(pf[key], dict):
                return pf[key].get('total', np.nan)
        return np.nan
    return ((fm.get('physical_flux') or {}).get('bottom', {}) or {}).get('total', np.nan)


def _avg_conc_of(results, domain_type):                   # no_advection_analysis_B.py:68-78
    avg = (results.get('mass_metrics') or {}).get('average_concentration', None)
    if domain_type == 'sulcus':
        return avg.get('total', None) if isinstance(avg, dict) else None
    return avg if isinstance(avg, (int, float)) else None


def phase_b_case(case, mesh_size_dim=0.02, quiet=True):
    """One (mu, geometry) pair: sulcus run + rectangle run -> CSV row (no_advection_analysis_B.py:112-172)."""
    mu, gkey, gcfg = case
    name = f"{gkey}_mu{str(mu).replace('.', 'p')}"
    sulc = _run(quiet, mode='no-adv', study_type="mu Sweep", config_name=f"Sulcus_{name}", domain_type='sulcus',
                params=_params_no_adv(mu, gcfg['sulci_w_dim'], gcfg['sulci_h_dim'], mesh_size_dim))
    rect = _run(quiet, mode='no-adv', study_type="mu Sweep", config_name=f"Rect_{name}", domain_type='rectangular',
                params=_params_no_adv(mu, gcfg['sulci_w_dim'], gcfg['sulci_h_dim'], mesh_size_dim))
    return phase_b_row(gkey, gcfg, mu, _avg_conc_of(sulc, 'sulcus'), _avg_conc_of(rect, 'rectangular'),
                       _flux_of(sulc, 'sulcus'), _flux_of(rect, 'rectangular'))


def phase_b_row(gkey, gcfg, mu, conc_s, conc_r, flux_s, flux_r):
    """CSV row of one (mu, geometry) pair from the four extracted numbers (no_advection_analysis_B.py:147-172)."""
    CR = (conc_s / conc_r) if (conc_s is not None and conc_r not in (None, 0)) else np.nan
    if flux_s is None or not np.isfinite(flux_s) or np.isclose(flux_s, 0.0):
        flux_ratio = flux_err = np.nan
    else:
        flux_ratio = flux_r / flux_s
        denom = abs(flux_s) if not np.isclose(abs(flux_s), 0.0) else 1.0
        flux_err = 100.0 * (flux_r - flux_s) / denom
    return {'geometry': gkey, 'width_mm': gcfg['sulci_w_dim'], 'depth_mm': gcfg['sulci_h_dim'],
            'aspect_ratio': gcfg.get('aspect_ratio'), 'mu_factor': mu, 'avg_conc_sulc': conc_s, 'avg_conc_rect': conc_r,
            'flux_sulc_y0': flux_s, 'flux_rect_bottom': flux_r, 'CR': CR, 'flux_ratio': flux_ratio,
            'flux_error_pct': flux_err}


def run_no_adv_mu_sweep(output_dir=None, mu_factors: Iterable[float] = None, geometries: Optional[Dict] = None,
                        mesh_size_dim=0.02, rank=None, world=None, quiet=True, prefetch=True, streams=1, batch=True):
    """Reference ``run_no_adv_mu_sweep`` (23 geometries x 3 mu x {sulcus, rectangle} = 138 solves by default).
    ``batch`` (default): the mu values of every mesh of this rank are solved in one batched Krylov loop first
    (``presolve_no_adv``).  Returns the DataFrame with the reference's columns; rank 0 writes
    ``no_adv_mu_sweep_results.csv``."""
    rank, world = _world(rank, world)
    mu_factors = list(MU_FACTORS_PHASE_B if mu_factors is None else mu_factors)
    configs = geometries if geometries is not None else create_geometry_variations(Parameters(mode='no-adv'), max_width=1.0)
    cases = [(mu, g, cfg) for mu in mu_factors for g, cfg in configs.items()]
    t0 = time.time()
    _prefetch([[(_params_no_adv(mu, cfg['sulci_w_dim'], cfg['sulci_h_dim'], mesh_size_dim), 'sulcus'),
                (_params_no_adv(mu, cfg['sulci_w_dim'], cfg['sulci_h_dim'], mesh_size_dim), 'rectangular')]
               for mu, _, cfg in cases], rank, world, prefetch)
    if batch:
        presolve_no_adv([(_params_no_adv(mu, cfg['sulci_w_dim'], cfg['sulci_h_dim'], mesh_size_dim), dom)
                         for i, (mu, _, cfg) in enumerate(cases) if i % world == rank
                         for dom in ('sulcus', 'rectangular')], quiet)
    done = _sharded(cases, lambda c: phase_b_case(c, mesh_size_dim, quiet), rank, world, streams, quiet)
    df = _frame([row for _, row in done], ['mu_factor', 'geometry'])
    p0 = Parameters(mode='no-adv')
    p0.validate()
    p0.nondim()
    meta = {'study_type': 'No Advection — mu Sweep', 'timestamp': time.strftime("%Y-%m-%dT%H:%M:%S"),
            'mu_factors': mu_factors, 'n_gpus': world, 'wall_s': time.time() - t0,
            'baselines': {'MU_DIM_NO_ADV': getattr(Parameters, 'MU_DIM_NO_ADV', None), 'D_dim': p0.D_dim,
                          'H_dim': p0.H_dim, 'L_dim': p0.L_dim}}
    _save(df, output_dir, "no_adv_mu_sweep_results.csv", meta, rank)
    return df


# ====================================================================== adv-diff validation: Pe x mu, step surrogate
PE_VALUES = [0.1, 1.0, 10]                                 # adv_diff_analysis.py:49-50
MU_FACTORS_ADV = [0.1, 1.0, 10]
REFERENCE_GEOMETRY = {'L_dim': 10.0, 'H_dim': 1.0, 'sulci_w_dim': 0.5, 'sulci_h_dim': 1.0, 'mesh_size_dim': 0.02,
                      'refinement_factor': 1}              # :52-59
D_DIM = 0.0003                                             # :61
MU_DIM_BASE = 0.0003                                       # :62
STEP_PARAMS = {'L_c': None, 'Gamma': 5.0, 'degree': 2}     # :64-68


def create_base_parameters(Pe_target, mu_factor, mesh_size_dim=None):
    """adv_diff_analysis.py:75-87."""
    geo = dict(REFERENCE_GEOMETRY)
    if mesh_size_dim is not None:
        geo['mesh_size_dim'] = mesh_size_dim
    params = Parameters(mode='adv-diff', U_ref_dim=Pe_target * D_DIM / geo['H_dim'], D_dim=D_DIM, **geo)
    params.mu_dim = MU_DIM_BASE * float(mu_factor)
    return params


def extract_flux_data(results, domain_type):
    """adv_diff_analysis.py:89-109."""
    fm = results.get('flux_metrics', {}) or {}
    if domain_type == 'sulcus':
        src = ((fm.get('sulcus_specific') or {}).get('physical_flux') or {}).get('y0_flux', {}) or {}
    else:
        src = (fm.get('physical_flux') or {}).get('bottom', {}) or {}
    return {'total_flux': src.get('total', None), 'diffusive_flux': src.get('diffusive', None),
            'advective_flux': src.get('advective', None), 'uptake_flux': fm.get('uptake_flux', None)}


def advdiff_case(case, mesh_size_dim=None, quiet=True):
    """One (Pe, mu): sulcus reference, then the rectangle with the step mu(x) built from its mu_eff_open
    (adv_diff_analysis.py:115-175, 201-260).  Returns the one or two CSV rows."""
    Pe, mu_factor = case
    params = create_base_parameters(Pe, mu_factor, mesh_size_dim)
    params.validate()
    params.nondim()
    name = f"Pe_{Pe:.1f}_mu_{mu_factor:.1f}".replace('.', 'p')
    sulc = _run(quiet, mode='adv-diff', study_type='AdvDiff Step Validation', config_name="Sulcus_" + name,
                domain_type='sulcus', params=params)
    me = sulc.get('mu_eff_comparison', {}) or {}
    arc, sim, opn = me.get('mu_eff_arc'), me.get('mu_eff_sim'), me.get('mu_eff_open')
    fs = extract_flux_data(sulc, 'sulcus')
    avg_s = sulc.get('mass_metrics', {}).get('average_concentration', {}).get('total')
    rows = [{'Pe': Pe, 'mu_factor': mu_factor, 'domain_type': 'sulcus', 'surrogate_type': 'reference', **fs,
             'mu_eff_arc': arc, 'mu_eff_sim': sim, 'mu_eff_open': opn, 'avg_conc': avg_s, 'CR': np.nan,
             'Mu_base_nondim': sulc['params'].mu, 'Domain_Length_mm': sulc['params'].L_dim,
             'Sulcus_Width_mm': sulc['params'].sulci_w_dim}]
    if opn is None:
        return rows
    pr = create_base_parameters(Pe, mu_factor, mesh_size_dim)
    pr.validate()
    pr.nondim()
    xl, xr = pr.L / 2 - pr.sulci_w / 2, pr.L / 2 + pr.sulci_w / 2
    mu_step = StepUptakeOpen(mu_base=float(mu_factor), mu_eff_target=float(opn), sulcus_left_x=xl, sulcus_right_x=xr,
                             L_c=STEP_PARAMS['L_c'] or (0.1 * pr.sulci_w), Gamma=STEP_PARAMS['Gamma'],
                             degree=STEP_PARAMS['degree'])
    pr.mu = mu_step
    pr.mu_dim = mu_step
    rect = _run(quiet, mode='adv-diff', study_type='AdvDiff Step Validation', config_name="Rect_step_open_" + name,
                domain_type='rectangular', params=pr, mu_variable=True)
    fr = extract_flux_data(rect, 'rectangular')
    avg_r = (rect.get('mass_metrics', {}) or {}).get('average_concentration')
    rows.append({'Pe': Pe, 'mu_factor': mu_factor, 'domain_type': 'rectangular', 'surrogate_type': 'step_open', **fr,
                 'mu_eff_arc': arc, 'mu_eff_sim': sim, 'mu_eff_open': opn, 'avg_conc': avg_r,
                 'CR': (avg_s / avg_r) if (avg_s is not None and avg_r not in (None, 0.0)) else np.nan})
    return rows


def add_surrogate_errors(df, pe_values, mu_factors):
    """flux_error_pct / flux_ratio of the step surrogate against the sulcus row of the same (Pe, mu)
    (adv_diff_analysis.py:266-279)."""
    df['flux_error_pct'] = np.nan
    df['flux_ratio'] = np.nan
    for Pe in pe_values:
        for mu in mu_factors:
            ref = (df['Pe'] == Pe) & (df['mu_factor'] == mu) & (df['domain_type'] == 'sulcus')
            rec = (df['Pe'] == Pe) & (df['mu_factor'] == mu) & (df['domain_type'] == 'rectangular') & \
                  (df['surrogate_type'] == 'step_open')
            if not ref.any() or not rec.any():
                continue
            rf = df.loc[ref, 'total_flux'].iloc[0]
            df.loc[rec, 'flux_ratio'] = df.loc[rec, 'total_flux'] / (rf if rf != 0 else 1.0)
            df.loc[rec, 'flux_error_pct'] = 100.0 * (df.loc[rec, 'total_flux'] - rf) / (abs(rf) if rf != 0 else 1.0)
    return df


def run_advdiff_step_validation(output_dir=None, pe_values=None, mu_factors=None, mesh_size_dim=None, rank=None,
                                world=None, quiet=True):
    """Reference ``run_advdiff_step_validation``: 3 Pe x 3 mu x {sulcus, step rectangle} = 18 solves (+ 1 Stokes solve
    per geometry: the flow depends on the geometry only).  Rank 0 writes ``advdiff_validation_step_pe_x_mu.csv``."""
    rank, world = _world(rank, world)
    pe_values = list(PE_VALUES if pe_values is None else pe_values)
    mu_factors = list(MU_FACTORS_ADV if mu_factors is None else mu_factors)
    cases = [(Pe, mu) for Pe in pe_values for mu in mu_factors]
    done = run_sharded(cases, lambda c: advdiff_case(c, mesh_size_dim, quiet), rank, world)
    df = _frame([r for _, rows in done for r in rows], ['Pe', 'mu_factor', 'domain_type'])
    df = add_surrogate_errors(df, pe_values, mu_factors)
    meta = {'study_type': 'AdvDiff Validation (Pe x mu) - Step mu only', 'timestamp': time.strftime("%Y-%m-%dT%H:%M:%S"),
            'Pe_values': pe_values, 'mu_factors': mu_factors, 'reference_geometry': REFERENCE_GEOMETRY,
            'parameters': {'D_dim': D_DIM, 'mu_dim_base': MU_DIM_BASE}, 'n_gpus': world}
    _save(df, output_dir, "advdiff_validation_step_pe_x_mu.csv", meta, rank)
    return df


# ====================================================================== Phase A: mu sweep and aspect-ratio study
REGIMES = {'small_uptake': [0.1, 0.25, 0.5, 0.75, 1.0, 1.25, 1.5, 2.0, 2.5, 3.0],          # no_advection_analysis_A.py:1276-1292
           'moderate_uptake': [5.0, 7.5, 10.0, 12.5, 15.0],
           'high_uptake': [50.0, 75.0, 100.0, 125.0, 150.0]}


def _mu_eff_columns(result):
    row = {}
    if 'mu_eff_comparison' in result:                      # no_advection_analysis_A.py:64-90
        d = result['mu_eff_comparison']
        r, e = d.get('ratios', {}), d.get('errors_vs_sim', {})
        row.update({'Mu_Eff_Simulation': d.get('mu_eff_sim'), 'Mu_Eff_Analytical': d.get('mu_eff_arc'),
                    'Mu_Eff_Enhanced': d.get('mu_eff_enh'), 'Mu_Eff_Opening': d.get('mu_eff_open'),
                    'Ratio_Sim': r.get('sim'), 'Ratio_Analytical': r.get('arc'), 'Ratio_Enhanced': r.get('enh'),
                    'Ratio_Opening': r.get('open'), 'Relative_Error_Analytical': e.get('arc'),
                    'Relative_Error_Enhanced': e.get('enh'), 'Relative_Error_Opening': e.get('open')})
    if 'mass_metrics' in result:
        row['Total_Mass'] = result['mass_metrics'].get('total_mass')
    if 'flux_metrics' in result:
        fm = result['flux_metrics'] or {}
        mouth = ((fm.get('sulcus_specific') or {}).get('physical_flux') or {}).get('sulcus_opening') or {}
        row['Mouth_Flux_Total'] = mouth.get('total')
    return row


def extract_mu_sweep_data(result, config_name, peclet_num=0):
    """no_advection_analysis_A.py:51-105 (same columns, same order)."""
    row = {'Config': config_name, 'Regime': result.get('regime', 'unknown'), 'Mu_Factor': result.get('mu_factor', 1.0),
           'Mu_dim': result.get('mu_dim_used'), 'Mu': result.get('mu_used'), 'Baseline_Mu_dim': result.get('baseline_mu_dim')}
    row.update(_mu_eff_columns(result))
    return row


def extract_aspect_ratio_data(result, config_name, aspect_ratio_type, width, depth, aspect_ratio):
    """no_advection_analysis_A.py:107-164."""
    row = {'Config': config_name, 'Aspect_Ratio_Type': aspect_ratio_type, 'Width': width, 'Depth': depth,
           'Aspect_Ratio': aspect_ratio}
    if 'parameters' in result or hasattr(result, 'mu'):
        row['Mu'] = result.get('mu_used') or result.get('mu', 0)
    row.update(_mu_eff_columns(result))
    return row


def _slim(result, **extra):
    """What the row extractors read from a run_simulation result (fields stay on the rank that computed them)."""
    out = {k: result[k] for k in ('mu_eff_comparison', 'mass_metrics', 'flux_metrics') if k in result}
    out.update(extra)
    return out


PRESOLVE_IN_BACKGROUND = os.environ.get('SFEM_PRESOLVE_BACKGROUND', '1') != '0'


def presolve_no_adv(jobs, quiet=True, background=None):
    """Batched pre-solve of no-advection cases: ``jobs`` = (params, domain_type) pairs; the pairs that
    share a mesh and a diffusivity are solved together, up to ``solvers.BATCH`` Robin coefficients per Krylov loop
    (``solvers.presolve_pure_diffusion``), and the fields are parked on the mesh for the ``run_simulation`` calls that
    follow.  ``background`` (default ``PRESOLVE_IN_BACKGROUND``): the Krylov loops run on a worker thread / stream and
    overlap the per-case post-processing of the batches already delivered.  Returns the number of fields."""
    background = PRESOLVE_IN_BACKGROUND if background is None else bool(background)
    from .fem import FunctionSpace
    from .simulation import _mesh_key, _simulation_generate_mesh
    from .solvers import presolve_pure_diffusion
    groups = {}
    for p, dom in jobs:
        groups.setdefault((_mesh_key(p, dom)[0], float(p.D)), []).append((p, dom))
    n = 0
    with contextlib.redirect_stdout(io.StringIO()) if quiet else contextlib.nullcontext():
        for (_, D), members in groups.items():
            p0, dom = members[0]
            mr = _simulation_generate_mesh(p0, dom)
            mus = sorted({float(p.mu) for p, _ in members})
            n += presolve_pure_diffusion(mr, FunctionSpace(mr['mesh'], "CG", 2), D, mus, background=background)
    return n


def run_mu_sweep(output_dir=None, regimes: Optional[Dict[str, List[float]]] = None, w_dim=0.25, h_dim=0.25,
                 mesh_size_dim=0.02, rank=None, world=None, quiet=True, streams=1, frozen_coarse=True, batch=True):
    """Reference ``run_mu_sweep``: 20 mu values in three uptake regimes on the 0.25 x 0.25 mm sulcus; the mesh,
    patterns and multigrid hierarchy are built once and reused by every mu.  ``batch`` (default): this rank's mu
    values are solved up to 8 at a time in batched Krylov loops before the per-case drivers run (``presolve_no_adv``);
    ``batch=False`` solves case by case (``frozen_coarse`` then keeps the multigrid levels between nearby mu).
    Rank 0 writes ``mu_parameter_sweep_results.csv``."""
    rank, world = _world(rank, world)
    regimes = REGIMES if regimes is None else regimes
    base = float(getattr(Parameters, 'MU_DIM_NO_ADV'))
    cases = [(reg, f) for reg, fs in regimes.items() for f in fs]

    def params_of(factor):
        p = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
        p.sulci_w_dim, p.sulci_h_dim = w_dim, h_dim
        p.mu_dim = base * factor
        p.validate()
        p.nondim()
        return p
    if batch:
        presolve_no_adv([(params_of(f), 'sulcus') for i, (_, f) in enumerate(cases) if i % world == rank], quiet)

    def one(case):
        reg, factor = case
        p = params_of(factor)
        name = f"{reg}_mu_{factor:.1f}x"
        # one geometry, 20 Robin coefficients: the multigrid levels are kept while mu stays within 4x of the value they
        # were assembled for (solvers.frozen_coarse_levels: preconditioner data only, same residual target)
        with frozen_coarse_levels(4.0 if frozen_coarse else 1.0):
            res = _run(quiet, mode='no-adv', study_type="Phase A/Mu Parameter Sweep Simulations", config_name=name,
                       domain_type='sulcus', params=p)
        return extract_mu_sweep_data(_slim(res, regime=reg, mu_factor=factor, mu_dim_used=p.mu_dim, mu_used=p.mu,
                                           baseline_mu_dim=base), name)
    done = _sharded(cases, one, rank, world, streams, quiet)
    df = _frame([row for _, row in done])
    _save(df, output_dir, "mu_parameter_sweep_results.csv", None, rank)
    return df


def aspect_ratio_cases(max_width=1.0):
    """The (type, AR, depth, width) list of no_advection_analysis_A.py:1355-1392."""
    micro = np.logspace(np.log10(0.01), np.log10(0.10), 10)
    meso = np.array([0.12, 0.15, 0.20, 0.25, 0.35, 0.50, 0.75, 1.00])
    macro = np.array([1.50, 2.00, 2.50, 3.00, 3.50, 4.00, 4.50, 5.00])
    depths = sorted(set(np.round(np.concatenate([micro, meso, macro]), 4)))
    out = []
    for name, ar in {'h_equals_w': 1.0, 'h_equals_2w': 2.0, 'h_equals_half_w': 0.5}.items():
        for h in depths:
            w = h / ar
            if w > max_width:
                continue
            out.append((name, ar, float(h), float(w)))
    return out


def run_aspect_ratio_analysis(output_dir=None, cases=None, mesh_size_dim=0.02, rank=None, world=None, quiet=True,
                              prefetch=True):
    """Reference ``run_aspect_ratio_analysis``.  Rank 0 writes ``aspect_ratio_analysis_results.csv``."""
    rank, world = _world(rank, world)
    cases = aspect_ratio_cases() if cases is None else list(cases)

    def params_of(case):
        _, _, h, w = case
        p = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
        p.sulci_w_dim, p.sulci_h_dim = w, h
        p.validate()
        p.nondim()
        return p
    _prefetch([[(params_of(c), 'sulcus')] for c in cases], rank, world, prefetch)

    def one(case):
        name, ar, h, w = case
        p = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
        p.sulci_w_dim, p.sulci_h_dim = w, h
        p.validate()
        p.nondim()
        cfg = f"{name}_h{h}"
        res = _run(quiet, mode='no-adv', study_type="Phase A/Aspect Ratio Study Simulations", config_name=cfg,
                   domain_type='sulcus', params=p)
        return extract_aspect_ratio_data(_slim(res), cfg, name, w, h, ar)
    done = run_sharded(cases, one, rank, world)
    df = _frame([row for _, row in done])
    _save(df, output_dir, "aspect_ratio_analysis_results.csv", None, rank)
    return df


# ====================================================================== Phase A: geometry analysis and mu_eff spatial analysis
def extract_geometry_analysis_data(result, config_name, geometry_name, mu_value, mu_factor, config_results=None):
    """no_advection_analysis_A.py:165-231 (same columns, same order)."""
    row = {'Config': config_name, 'Geometry_Name': geometry_name, 'Mu_Value': mu_value, 'Mu_Factor': mu_factor}
    if config_results and 'geometry_config' in config_results:
        g = config_results['geometry_config']
        w, h = g.get('sulci_w_dim'), g.get('sulci_h_dim')
        row['Sulcus_Width_mm'], row['Sulcus_Depth_mm'] = w, h
        if w and h and w > 0:
            row['Aspect_Ratio'] = h / w
        row['Aspect_Ratio_Category'] = g.get('aspect_ratio_category', 'unknown')
    row.update(_mu_eff_columns(result))
    return row


def extract_mu_eff_analysis_data(result, config_name, mu_value, mu_factor):
    """no_advection_analysis_A.py:233-296 (same columns, same order; mu(x) sampled at 100 points of the floor)."""
    from .analysis import sample_mu_along_bottom
    row = {'Config': config_name, 'Mu_Value': mu_value, 'Mu_Factor': mu_factor}
    params = result.get('params')
    if params is not None:
        row.update({'Sulcus_Width_mm': getattr(params, 'sulci_w_dim', 0.5), 'Sulcus_Depth_mm': getattr(params, 'sulci_h_dim', 1.0),
                    'Domain_Length_mm': getattr(params, 'L_dim', 10.0),
                    'L_ref': getattr(params, 'L_ref', getattr(params, 'H_dim', 1.0)),
                    'L_nondim': getattr(params, 'L', 10.0), 'H_nondim': getattr(params, 'H', 1.0),
                    'Sulcus_W_nondim': getattr(params, 'sulci_w', 0.5), 'Sulcus_H_nondim': getattr(params, 'sulci_h', 1.0),
                    'Mu_base_nondim': getattr(params, 'mu', mu_value)})
    else:
        row.update({'Sulcus_Width_mm': 0.5, 'Sulcus_Depth_mm': 1.0, 'Domain_Length_mm': 10.0, 'L_ref': 1.0, 'L_nondim': 10.0,
                    'H_nondim': 1.0, 'Sulcus_W_nondim': 0.5, 'Sulcus_H_nondim': 1.0, 'Mu_base_nondim': mu_value})
    if 'mu_eff_comparison' in result:
        d = result['mu_eff_comparison']
        r = d.get('ratios', {})
        row.update({'Mu_Eff_Simulation': d.get('mu_eff_sim'), 'Mu_Eff_Analytical': d.get('mu_eff_arc'),
                    'Mu_Eff_Enhanced': d.get('mu_eff_enh'), 'Mu_Eff_Opening': d.get('mu_eff_open'),
                    'Ratio_Sim': r.get('sim'), 'Ratio_Analytical': r.get('arc'), 'Ratio_Enhanced': r.get('enh'),
                    'Ratio_Opening': r.get('open')})
    try:
        ms = sample_mu_along_bottom(result, n_points=100)
        row.update({'Mu_Mean_Bottom': ms['mu_mean'], 'Mu_Min_Bottom': ms['mu_min'], 'Mu_Max_Bottom': ms['mu_max'],
                    'Mu_X_Array': str(np.asarray(ms['x']).tolist()), 'Mu_Values_Array': str(np.asarray(ms['mu']).tolist())})
    except Exception as e:
        print(f"Warning: Could not sample mu for {config_name}: {e}")
        row.update({'Mu_Mean_Bottom': None, 'Mu_Min_Bottom': None, 'Mu_Max_Bottom': None, 'Mu_X_Array': None,
                    'Mu_Values_Array': None})
    return row


def run_geometry_analysis(output_dir=None, mu_factors=(0.1, 1.0, 10), geometries: Optional[Dict] = None, mesh_size_dim=0.02,
                          rank=None, world=None, quiet=True, prefetch=True, streams=1, batch=True):
    """Reference ``run_geometry_analysis`` (no_advection_analysis_A.py:1463-1581): every geometry of
    ``create_geometry_variations`` x mu factors (23 x 3 = 69 sulcus solves by default); the device problems of a
    geometry are reused by its mu values.  Rank 0 writes ``geometry_analysis_results.csv``."""
    rank, world = _world(rank, world)
    base_params = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
    base_params.sulci_w_dim = base_params.sulci_h_dim = 0.25                 # :1474-1477
    base_params.validate()
    base_params.nondim()
    base = float(getattr(Parameters, 'MU_DIM_NO_ADV', base_params.mu_dim))
    geos = geometries if geometries is not None else create_geometry_variations(base_params)
    mu_factors = list(mu_factors)
    # geometry-major case order (:1506-1507): consecutive cases of a rank share the cached geometry
    cases = [(g, cfg, f) for g, cfg in geos.items() for f in mu_factors]

    def params_of(cfg, f):
        p = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
        p.sulci_w_dim, p.sulci_h_dim = cfg['sulci_w_dim'], cfg['sulci_h_dim']
        p.mu_dim = base * f
        p.validate()
        p.nondim()
        return p
    _prefetch([[(params_of(cfg, f), 'sulcus')] for _, cfg, f in cases], rank, world, prefetch)
    if batch:        # the mu values of a geometry in one batched Krylov loop (presolve_no_adv)
        presolve_no_adv([(params_of(cfg, f), 'sulcus') for i, (_, cfg, f) in enumerate(cases) if i % world == rank], quiet)

    def one(case):
        g, cfg, f = case
        name = f"{g}_mu_{f}"
        res = _run(quiet, mode='no-adv', study_type="Phase A/Geometry Comparison Simulations", config_name=name,
                   domain_type='sulcus', params=params_of(cfg, f))
        return extract_geometry_analysis_data(_slim(res), name, g, base * f, f, {'geometry_config': cfg})
    done = _sharded(cases, one, rank, world, streams, quiet)
    df = _frame([row for _, row in done])
    _save(df, output_dir, "geometry_analysis_results.csv", None, rank)
    return df


def run_mu_eff_analysis(output_dir=None, mu_factors=(0.1, 1.0, 10.0), w_dim=0.5, h_dim=1.0, mesh_size_dim=0.02, rank=None,
                        world=None, quiet=True):
    """Reference ``run_mu_eff_analysis`` (no_advection_analysis_A.py:1583-1682): the 0.5 x 1.0 mm sulcus at three mu
    values, with mu(x) sampled along the floor.  Rank 0 writes ``mu_eff_analysis_results.csv`` (the reference's
    checked-in copy of this file holds the BASELINE config-1 numbers)."""
    rank, world = _world(rank, world)
    base = float(getattr(Parameters, 'MU_DIM_NO_ADV', 0.0003))

    def one(factor):
        p = Parameters(mode='no-adv', mesh_size_dim=mesh_size_dim)
        p.sulci_w_dim, p.sulci_h_dim = w_dim, h_dim
        p.mu_dim = base * factor
        p.validate()
        p.nondim()
        name = f"mu_eff_analysis_mu_{factor}x"
        res = _run(quiet, mode='no-adv', study_type="Phase A/Mu_Eff Spatial Analysis Simulations", config_name=name,
                   domain_type='sulcus', params=p)
        keep = _slim(res, params=res.get('params', p), mesh_results={'mesh': res['mesh_results']['mesh']})
        return extract_mu_eff_analysis_data(keep, name, p.mu_dim, factor)
    done = run_sharded([float(f) for f in mu_factors], one, rank, world)
    df = _frame([row for _, row in done])
    _save(df, output_dir, "mu_eff_analysis_results.csv", None, rank)
    return df


# ====================================================================== no-uptake geometry comparison (mu = 0)
def format_filename_value(value):
    """plotting.py:249-253."""
    if abs(value - round(value)) < 0.001:
        return f"{value:.0f}"
    return f"{value:.1f}".replace('.', 'p') if value >= 1.0 else f"{value:.3f}".replace('.', 'p')


def _no_uptake_params(pe, mesh_size_dim, w_dim=None, h_dim=None):
    """no_uptake_analysis.py:114-127, 945-951 (same order of assignments)."""
    p = Parameters(mode='no-uptake', mesh_size_dim=mesh_size_dim)
    if w_dim is not None:
        p.sulci_w_dim, p.sulci_h_dim = w_dim, h_dim
    else:
        p.mu_dim = 0.0
    p.U_ref_dim = pe * p.D_dim / p.H_dim
    p.validate()
    p.nondim()
    p.D_dim = p.U_ref_dim * p.H_dim / p.Pe
    return p


def _vel_cols(vm, sulcus):
    """The reference reads 'max_ux_sulcus_level' (no_uptake_analysis.py:229-232), a key its current
    compute_velocity_metrics no longer produces (the line is called 'mouth_level' now, analysis.py:748) although the
    checked-in CSV still carries the values; the mouth-level line is used for those two columns."""
    vm = vm if isinstance(vm, dict) else {}
    g = vm.get
    return {'Max_Ux_mid_channel': g('max_ux_mid_channel'), 'Avg_Ux_mid_channel': g('avg_ux_mid_channel'),
            'Max_Ux_sulcus_level': g('max_ux_sulcus_level', g('max_ux_mouth_level')) if sulcus else None,
            'Avg_Ux_sulcus_level': g('avg_ux_sulcus_level', g('avg_ux_mouth_level')) if sulcus else None}


def _common_cols(params, domain, w=None, h=None):
    D_dim = params.U_ref_dim * params.H_dim / params.Pe
    return {'Domain': domain, 'Mode': params.mode, 'Peclet': params.Pe, 'U_ref': getattr(params, 'U_ref', None),
            'Sulcus Width (mm)': w, 'Sulcus Depth (mm)': h,
            'Aspect_Ratio': (h / w) if (w and w > 0) else None, 'U_ref (Dim)': params.U_ref_dim, 'Diff Coef (Dim)': D_dim,
            'Delta (mm)': D_dim / params.U_ref_dim}


def extract_simulation_data(result):
    """Sulcus row of geometry_comparison_results.csv (no_uptake_analysis.py:146-236)."""
    params = result.get('params')
    if not params:
        return None
    row = _common_cols(params, 'sulcus', getattr(params, 'sulci_w_dim', None), getattr(params, 'sulci_h_dim', None))
    mm = result.get('mass_metrics', {})
    if isinstance(mm, dict):
        avg = mm.get('average_concentration', {})
        d = isinstance(avg, dict)
        row.update({'Total Mass': mm.get('total_mass'), 'Sulcus Mass': mm.get('sulcus_mass'),
                    'Main Channel Mass': mm.get('rectangle_mass'), 'Avg Concentration': avg.get('total') if d else avg,
                    'Sulcus Avg Concentration': avg.get('sulcus_region') if d else None,
                    'Main Channel Avg Concentration': avg.get('rectangle_region') if d else None})
    fm = result.get('flux_metrics', {})
    extra = {}
    if isinstance(fm, dict):
        spf = fm.get('sulcus_specific', {}).get('physical_flux', {})
        row['Mouth_Flux_Total'] = spf.get('sulcus_opening', {}).get('total')
        pf = fm.get('physical_flux', {})
        row['Inlet-Outlet Flux'] = pf.get('left', {}).get('total', 0) + pf.get('right', {}).get('total', 0)
        extra = spf.get('sulcus_opening_extra', {})
    if isinstance(extra, dict):
        row.update({'Mouth E_L1': extra.get('E_L1'), 'Mouth E_avg': extra.get('E_avg'), 'Mouth Q_in': extra.get('Q_in'),
                    'Mouth Q_out': extra.get('Q_out'), 'Mouth Net Check': extra.get('net_check'),
                    'Mouth Length': extra.get('length')})
    row.update(_vel_cols(result.get('vel_metrics', {}), True))
    return row


def extract_rectangular_data(result):
    """Rectangle row (no_uptake_analysis.py:50-107), columns in the order of the sulcus rows."""
    params = result.get('params')
    if not params:
        return None
    mm = result.get('mass_metrics', {})
    fm = result.get('flux_metrics', {})
    pf = fm.get('physical_flux', {}) if isinstance(fm, dict) else {}
    inlet = pf.get('left', {}).get('total', 0) if isinstance(pf.get('left'), dict) else 0
    outlet = pf.get('right', {}).get('total', 0) if isinstance(pf.get('right'), dict) else 0
    row = _common_cols(params, 'rectangle')
    row.update({'Total Mass': mm.get('total_mass'), 'Sulcus Mass': None, 'Main Channel Mass': mm.get('total_mass', None),
                'Avg Concentration': mm.get('average_concentration', None), 'Sulcus Avg Concentration': None,
                'Main Channel Avg Concentration': mm.get('average_concentration', None), 'Mouth_Flux_Total': None,
                'Inlet-Outlet Flux': inlet + outlet, 'Mouth E_L1': None, 'Mouth E_avg': None, 'Mouth Q_in': None,
                'Mouth Q_out': None, 'Mouth Net Check': None, 'Mouth Length': None})
    row.update(_vel_cols(result.get('vel_metrics', {}), False))
    return row


def add_ratio_metrics(df):
    """no_uptake_analysis.py:262-313 on the DataFrame (sulcus rows against the rectangle baseline of the same Pe)."""
    rect = df[df['Domain'] == 'rectangle'].groupby('Peclet').agg({'Avg Concentration': 'mean', 'Max_Ux_mid_channel': 'mean',
                                                                  'Avg_Ux_mid_channel': 'mean'})
    for col in ('Concentration_Ratio', 'Channel_Conc_Ratio', 'Intradomain_Enrichment', 'VR_mid_avg', 'VR_mid_max',
                'VR_intradomain_avg', 'VR_intradomain_max'):
        df[col] = np.nan
    num = lambda s: df.loc[mask, s].astype(float)          # noqa: E731
    for pe in rect.index:
        mask = (df['Domain'] == 'sulcus') & (df['Peclet'] == pe)
        if not mask.any():
            continue
        df.loc[mask, 'Concentration_Ratio'] = num('Avg Concentration') / rect.loc[pe, 'Avg Concentration']
        df.loc[mask, 'Channel_Conc_Ratio'] = num('Main Channel Avg Concentration') / rect.loc[pe, 'Avg Concentration']
        df.loc[mask, 'VR_mid_avg'] = num('Avg_Ux_mid_channel') / rect.loc[pe, 'Avg_Ux_mid_channel']
        df.loc[mask, 'VR_mid_max'] = num('Max_Ux_mid_channel') / rect.loc[pe, 'Max_Ux_mid_channel']
        df.loc[mask, 'Intradomain_Enrichment'] = num('Sulcus Avg Concentration') / num('Main Channel Avg Concentration')
        df.loc[mask, 'VR_intradomain_avg'] = num('Avg_Ux_sulcus_level') / num('Avg_Ux_mid_channel')
        df.loc[mask, 'VR_intradomain_max'] = num('Max_Ux_sulcus_level') / num('Max_Ux_mid_channel')
    return df


def run_geometry_study(output_dir=None, peclet_numbers=(0.1, 1.0, 10.0), geometries: Optional[Dict] = None,
                       mesh_size_dim=0.02, rank=None, world=None, quiet=True, prefetch=True):
    """Reference ``run_geometry_study`` (no_uptake_analysis.py:921-975): mu = 0, every geometry x every Pe on the
    sulcus plus one rectangle baseline per Pe (23 x 3 + 3 = 72 rows by default), ratio columns added.  The Stokes flow of
    a geometry is solved once and reused by its Pe values.  Rank 0 writes ``geometry_comparison_results.csv``."""
    rank, world = _world(rank, world)
    peclet_numbers = list(peclet_numbers)
    configs = geometries if geometries is not None else create_geometry_variations(Parameters(mode='no-uptake'), max_width=1.0)
    # a geometry's Pe values stay on one rank (they share the mesh and the Stokes solution)
    cases = [('sulcus', k, cfg) for k, cfg in configs.items()] + [('rectangle', None, None)]
    pe0 = peclet_numbers[0]
    _prefetch([[(_no_uptake_params(pe0, mesh_size_dim, cfg['sulci_w_dim'], cfg['sulci_h_dim']), 'sulcus')] if kind == 'sulcus'
               else [(_no_uptake_params(pe0, mesh_size_dim), 'rectangular')] for kind, _, cfg in cases], rank, world, prefetch)

    def one(case):
        kind, key, cfg = case
        rows = []
        for pe in peclet_numbers:
            if kind == 'sulcus':
                p = _no_uptake_params(pe, mesh_size_dim, cfg['sulci_w_dim'], cfg['sulci_h_dim'])
                res = _run(quiet, mode='no-uptake', study_type="Geometry Comparison",
                           config_name=f"{key}_Pe{format_filename_value(pe)}", domain_type='sulcus', params=p)
                rows.append(extract_simulation_data(res))
            else:
                p = _no_uptake_params(pe, mesh_size_dim)
                res = _run(quiet, mode='no-uptake', study_type='Rectangular Baselines',
                           config_name=f'rect_Pe{format_filename_value(pe)}', domain_type='rectangular', params=p)
                rows.append(extract_rectangular_data(res))
        return rows
    done = run_sharded(cases, one, rank, world)
    df = add_ratio_metrics(_frame([r for _, rows in done for r in rows]))
    _save(df, output_dir, "geometry_comparison_results.csv", None, rank)
    return df


# ---- concentration line profiles of selected geometries (no_uptake_analysis.py:315-436, 977-1015)
def collect_profile_rows(result, geometry_key=None):
    """One tidy row per sample of result['mass_metrics']['profiles_full'] (horizontal lines, like the reference:
    no_uptake_analysis.py:315-359); columns = those of the reference's profiles_samples_<geometry>.csv."""
    rows = []
    if not result:
        return rows
    params = result.get('params', {})
    mm = result.get('mass_metrics') or {}
    full, meta = mm.get('profiles_full') or {}, mm.get('profiles_meta') or {}
    domain = result.get('domain_type', 'unknown')
    config = result.get('config_name', result.get('geometry'))
    pe = getattr(params, 'Pe', None)
    x_rng, y_rng, n_points = meta.get('x_range'), meta.get('y_range'), meta.get('n_points')
    for name, payload in (full.get('horizontal') or {}).items():
        y = float(payload['y'])
        for i, (xx, cc) in enumerate(zip(np.asarray(payload['x']), np.asarray(payload['c']))):
            rows.append({'Domain': domain, 'Geometry': geometry_key or result.get('geometry'), 'Config': config, 'Peclet': pe,
                         'LineType': 'horizontal', 'LineName': name, 'Index': i, 'x': float(xx), 'y': y, 'c': float(cc),
                         'n_points': n_points, 'x_min': None if x_rng is None else float(x_rng[0]),
                         'x_max': None if x_rng is None else float(x_rng[1]),
                         'y_min': None if y_rng is None else float(y_rng[0]),
                         'y_max': None if y_rng is None else float(y_rng[1])})
    return rows


def run_profile_export(output_dir=None, geometry_keys=('largest', 'square_small'), peclet_numbers=(0.1, 1.0, 10.0),
                       mesh_size_dim=0.02, n_points=400, quiet=True):
    """The selective profile step of ``run_geometry_study`` (no_uptake_analysis.py:977-1015): for the chosen
    geometries and every Pe, run the no-uptake case, ``compute_conc_profiles`` (all eight lines in one device launch)
    and write ``profiles_samples_<geometry>.csv`` / ``profiles_<geometry>.csv``.  Returns {geometry: (samples
    DataFrame, statistics DataFrame)}."""
    from .analysis import compute_conc_profiles
    configs = create_geometry_variations(Parameters(mode='no-uptake'), max_width=1.0)
    out = {}
    for gkey in geometry_keys:
        cfg = configs[gkey]
        rows, stats = [], []
        for pe in peclet_numbers:
            p = _no_uptake_params(pe, mesh_size_dim, cfg['sulci_w_dim'], cfg['sulci_h_dim'])
            res = _run(quiet, mode='no-uptake', study_type="Geometry Comparison",
                       config_name=f"{gkey}_Pe{format_filename_value(pe)}", domain_type='sulcus', params=p)
            res.update({'geometry': gkey, 'peclet': pe, 'domain_type': 'sulcus'})
            compute_conc_profiles(res, n_points=n_points)
            rows.extend(collect_profile_rows(res, geometry_key=gkey))
            for name, st in ((res['mass_metrics'].get('profiles') or {}).get('horizontal') or {}).items():
                stats.append({'Geometry': gkey, 'Peclet': pe, 'line_type': 'horizontal', 'name': name, 'x': None,
                              'y': st.get('y'), 'min_c': st.get('min_c'), 'max_c': st.get('max_c'),
                              'avg_c': st.get('avg_c'), 'n_samples': st.get('n_samples')})
        df, dfs = _frame(rows), _frame(stats)
        _save(df, output_dir, f"profiles_samples_{gkey}.csv")
        _save(dfs, output_dir, f"profiles_{gkey}.csv")
        out[gkey] = (df, dfs)
    return out
