"""Flux / mass / mu_eff metrics with the reference's function names and result schema.

Mirrors reference ``analysis.py``: ``compute_flux_metrics`` (:640), ``compute_mass_metrics`` (:677),
``compute_concentration_profiles`` (:884), ``compute_mu_eff_arc/_enh/_sim/_sim_mouth`` (:948-1031),
``compute_mu_eff_metrics`` (:1033).  Where the reference issues ~38 separate dolfin ``assemble``
calls (each a full mesh loop), this module evaluates *all* facet and cell integrals of a run in two
kernel launches (``sfem_facet_functionals`` / ``sfem_cell_functionals``) and then only fills the
nested dictionaries the study drivers index.
"""
from __future__ import annotations

import functools

import numpy as np
from scipy.integrate import quad

from .fem import Constant, evaluate_expression
from . import dofmap as dm
from .hostmesh import MARKERS  # noqa: F401

_G = {'left': 0, 'right': 1, 'top': 2, 'bottom': 3, 'bottom_left': 4, 'sulcus': 5, 'bottom_right': 6,
      'y0_ext': 7, 'mouth': 8}
_DIFF, _ADV, _UPT, _C, _LEN, _EL1, _QIN, _QOUT = range(8)


def _plan(mesh_results, domain_type):
    mesh = mesh_results['mesh']
    cache = getattr(mesh, '_sfem_cache', None)
    if cache is None:
        cache = {}
        setattr(mesh, '_sfem_cache', cache)
    from .solvers import _slot_key
    key = _slot_key(('functionals', domain_type))
    if key not in cache:
        from .device import FunctionalPlan
        cache[key] = FunctionalPlan(mesh, mesh_results, domain_type)
    return cache[key]


def _mu_key(mu_val):
    if mu_val is None:
        return 0.0
    if np.isscalar(mu_val) or isinstance(mu_val, Constant):
        return float(mu_val)
    return ('expr', id(mu_val))


def _field_key(f):
    """Identity + edit counter of a Function (``vector().set_local`` etc. bump the counter, so edited fields are
    never served stale functionals)."""
    return None if f is None else (id(f), getattr(f, '_version', 0))


def evaluate_functionals(c, u, mesh_results, domain_type, D_val, mu_val):
    """All facet + cell integrals of one run (cached on ``c``): returns (F[groups,8], M[markers,2]).  The cache key
    holds every argument, so the functions below stay pure in (c, u, D, mu) like the reference's."""
    key = (_field_key(c), _field_key(u), float(D_val), _mu_key(mu_val), domain_type)
    cached = getattr(c, '_functionals', None)
    if cached is not None and cached[0] == key:
        return cached[1], cached[2]
    plan = _plan(mesh_results, domain_type)
    cd = c.device_components()[0]
    ux = uy = None
    if u is not None and np.any(u.values):
        ux, uy = u.device_components()
    mu_const, mu_nodal = 0.0, None
    if np.isscalar(mu_val) or isinstance(mu_val, Constant):
        mu_const = float(mu_val)
    elif mu_val is not None:
        mesh = mesh_results['mesh']
        X = dm.p2_dof_coordinates(mesh)
        vals = np.zeros(len(X))
        bottom = np.flatnonzero(X[:, 1] <= 0.0)           # only dofs of facets on the floor are read
        vals[bottom] = evaluate_expression(mu_val, X[bottom])
        mu_nodal = plan.ctx.up(vals, np.float64)
    F, M = plan.evaluate(cd, ux, uy, D=float(D_val), mu_const=mu_const, mu_nodal=mu_nodal)
    c._functionals = (key, F, M, u)                      # holding `u` keeps its id() from being recycled
    return F, M


def _mesh_results_of(c):
    """Marker containers of the mesh ``c`` lives on (attached by ``hostmesh.build_markers``), for the reference
    functions whose signature carries only ``measures``."""
    mesh = c.function_space().mesh()
    mk = getattr(mesh, '_sfem_markers', None)
    if mk is None:
        raise ValueError("the mesh carries no marker sets (build them with hostmesh.build_markers / MeshGenerator)")
    mr = {'mesh': mesh}
    mr.update(mk)
    return mr, ('sulcus' if 'domain_markers' in mk else 'rectangular')


def _cached_or_evaluate(c, mu_val=None, need_mu=False):
    """(F, M) of ``c``: the cached evaluation when it still matches the field (and, if ``need_mu``, the given mu);
    otherwise a fresh one (u = None, D = 1: the mu c, c and mass integrals do not depend on them)."""
    cached = getattr(c, '_functionals', None)
    if cached is not None and cached[0][0] == _field_key(c) and (not need_mu or cached[0][3] == _mu_key(mu_val)):
        return cached[1], cached[2]
    mr, domain_type = _mesh_results_of(c)
    u = cached[3] if cached is not None and cached[0][0] == _field_key(c) else None
    D = cached[0][2] if u is not None else 1.0
    return evaluate_functionals(c, u, mr, domain_type, D, mu_val if need_mu else 0.0)


def _pf(F, g):
    d, a = float(F[g, _DIFF]), float(F[g, _ADV])
    return {'diffusive': d, 'advective': a, 'total': float(d + a)}


# ====================================================================== flux metrics
def compute_physical_flux_boundary(c, u, mesh_results, measures, boundary_marker, D_val):
    domain_type = 'sulcus' if 'domain_markers' in mesh_results else 'rectangular'
    F, _ = evaluate_functionals(c, u, mesh_results, domain_type, D_val, 0.0)
    return _pf(F, int(boundary_marker) - 1)


def compute_sulcus_segment_fluxes(c, u, mesh_results, measures, D_val, _F=None):
    """Segment, mouth and y=0 fluxes of the sulcus mesh (reference analysis.py:181-298)."""
    F = _F if _F is not None else evaluate_functionals(c, u, mesh_results, 'sulcus', D_val, 0.0)[0]
    fl = {name: _pf(F, _G[name]) for name in ('bottom_left', 'sulcus', 'bottom_right')}
    m = _G['mouth']
    fl['sulcus_opening'] = _pf(F, m)
    L_sig = float(F[m, _LEN])
    fl['sulcus_opening_extra'] = {
        'E_L1': float(F[m, _EL1]), 'E_avg': float(F[m, _EL1] / L_sig) if L_sig else float('nan'),
        'Q_in': float(F[m, _QIN]), 'Q_out': float(F[m, _QOUT]),
        'net_check': float(F[m, _QIN] - F[m, _QOUT]), 'length': L_sig}
    e = _G['y0_ext']
    jd, ja = float(F[e, _DIFF] + F[m, _DIFF]), float(F[e, _ADV] + F[m, _ADV])
    fl['y0_flux'] = {'diffusive': jd, 'advective': ja, 'total': float(jd + ja)}

    def sum_fields(keys):
        return {nm: float(sum(fl[k][nm] for k in keys)) for nm in ('diffusive', 'advective', 'total')}
    fl['bottom_combined'] = sum_fields(['bottom_left', 'sulcus', 'bottom_right'])
    fl['y0_combined'] = sum_fields(['bottom_left', 'bottom_right', 'sulcus_opening'])
    diff_val = abs(fl['y0_flux']['total'] - fl['y0_combined']['total'])
    if diff_val > 1e-10:
        print(f"⚠️ y0_flux vs y0_combined differ by {diff_val:.3e}")
    return fl


def compute_flux_metrics(c, u, mesh_results, domain_type, measures, D_val, mu_val):
    """Reference analysis.py:640-675: same nested dictionary."""
    F, _ = evaluate_functionals(c, u, mesh_results, domain_type, D_val, mu_val)
    out = {'physical_flux': {nm: _pf(F, _G[nm]) for nm in ('left', 'right', 'top', 'bottom')},
           'uptake_flux': float(F[_G['bottom'], _UPT])}
    if domain_type == 'sulcus':
        bl, su, br = (float(F[_G[k], _UPT]) for k in ('bottom_left', 'sulcus', 'bottom_right'))
        out['sulcus_specific'] = {
            'physical_flux': compute_sulcus_segment_fluxes(c, u, mesh_results, measures, D_val, _F=F),
            'uptake_flux': {'bottom_left': bl, 'sulcus': su, 'bottom_right': br, 'total': bl + su + br}}
    return out


def compute_uptake_flux_bottom(c, measures, mu_val):
    """``assemble(mu c ds(4))`` (reference analysis.py:307-311): pure in (c, mu_val)."""
    F, _ = _cached_or_evaluate(c, mu_val, need_mu=True)
    return float(F[_G['bottom'], _UPT])


def compute_uptake_flux_segments(c, measures, mu_val):
    """Uptake flux mu c over the external bottom segments bottom_left / sulcus / bottom_right and their sum
    (reference analysis.py:313-333): pure in (c, mu_val)."""
    F, _ = _cached_or_evaluate(c, mu_val, need_mu=True)
    bl, su, br = (float(F[_G[k], _UPT]) for k in ('bottom_left', 'sulcus', 'bottom_right'))
    return {'bottom_left': bl, 'sulcus': su, 'bottom_right': br, 'total': bl + su + br}


def sample_mu_along_bottom(results, n_points=500, y_at_bottom=0.0, save_csv_path=None):
    """Sample the Robin coefficient mu(x) along the bottom wall (reference analysis.py:838-882, same keys): mu may be
    a float, a ``Constant`` or a ``UserExpression`` such as ``StepUptakeOpen``.  Host-only (it evaluates the
    coefficient, not a field)."""
    params = results.get('params', None)
    mesh = results.get('mesh_results', {}).get('mesh', None)
    if params is None or mesh is None:
        raise ValueError("results must contain 'params' and 'mesh_results[mesh]'")
    mu_obj = getattr(params, 'mu', None)
    coords = mesh.coordinates()
    xs = np.linspace(float(coords[:, 0].min()), float(coords[:, 0].max()), int(n_points))
    mus = evaluate_expression(mu_obj, np.stack([xs, np.full(len(xs), float(y_at_bottom))], axis=1))
    trapz = getattr(np, 'trapezoid', None) or np.trapz
    out = {'x': xs, 'mu': mus,
           'mu_mean': float(trapz(mus, xs) / (xs[-1] - xs[0]) if len(xs) > 1 else mus.mean()),
           'mu_min': float(np.min(mus)), 'mu_max': float(np.max(mus))}
    if save_csv_path:
        import os
        import pandas as pd
        os.makedirs(os.path.dirname(save_csv_path), exist_ok=True)
        pd.DataFrame({'x': xs, 'mu': mus}).to_csv(save_csv_path, index=False)
    return out


# ====================================================================== mass metrics
def compute_mass_metrics(c, measures, domain_type):
    """Reference analysis.py:677-719."""
    _, M = _cached_or_evaluate(c)
    if domain_type == 'sulcus':
        sm, sa = float(M[1, 0]), float(M[1, 1])
        rm, ra = float(M[2, 0]), float(M[2, 1])
        tm, ta = sm + rm, sa + ra
        return {'total_mass': tm, 'sulcus_mass': sm, 'rectangle_mass': rm, 'total_area': ta, 'sulcus_area': sa,
                'rectangle_area': ra,
                'average_concentration': {'total': tm / ta if ta > 0 else None,
                                          'sulcus_region': sm / sa if sa > 0 else None,
                                          'rectangle_region': rm / ra if ra > 0 else None}}
    tm, ta = float(M[0, 0]), float(M[0, 1])
    return {'total_mass': tm, 'total_area': ta, 'average_concentration': ta and tm / ta or 0.0}


# ====================================================================== line profiles (device point evaluation)
def _line_points(mesh, fixed, axis, rng, n_points):
    k = 1 if axis == 'v' else 0
    if rng is None:
        X = mesh.coordinates()
        lo, hi = X[:, k].min(), X[:, k].max()
    else:
        lo, hi = rng
    s = np.linspace(lo, hi, n_points)
    f = np.full(n_points, float(fixed))
    return s, (np.stack([f, s], axis=1) if axis == 'v' else np.stack([s, f], axis=1))


def _eval_lines(f, mesh, lines, n_points):
    """All sample points of all lines in ONE device launch.  lines: [(fixed, axis, range)] ->
    [(coords_inside, values_inside)] (values [k] or [k, 2])."""
    S, P = zip(*[_line_points(mesh, fx, ax, rg, n_points) for fx, ax, rg in lines]) if lines else ((), ())
    if not lines:
        return []
    vals, ok = f.eval_points(np.concatenate(P, axis=0))
    out = []
    for i, s in enumerate(S):
        sl = slice(i * n_points, (i + 1) * n_points)
        out.append((s[ok[sl]], vals[sl][ok[sl]]))
    return out


def extract_concentration_vertical_line_profile(c, mesh, x_location, y_range=None, n_points=100):
    """Reference analysis.py:341-378 (same keys); the per-point collision test + c(Point) loop is one launch."""
    (y, v), = _eval_lines(c, mesh, [(x_location, 'v', y_range)], n_points)
    return {'y_coords': y, 'c': v}


def extract_concentration_horizontal_line_profile(c, mesh, y_location, x_range=None, n_points=100):
    """Reference analysis.py:380-419."""
    (x, v), = _eval_lines(c, mesh, [(y_location, 'h', x_range)], n_points)
    return {'x_coords': x, 'c': v}


def extract_velocity_vertical_line_profile(u, mesh, x_location, y_range=None, n_points=100):
    """Reference analysis.py:544-586."""
    (y, v), = _eval_lines(u, mesh, [(x_location, 'v', y_range)], n_points)
    return {'y_coords': y, 'u_x': v[:, 0], 'u_y': v[:, 1], 'u_mag': np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2)}


def extract_velocity_horizontal_line_profile(u, mesh, y_location, x_range=None, n_points=100):
    """Reference analysis.py:588-632."""
    (x, v), = _eval_lines(u, mesh, [(y_location, 'h', x_range)], n_points)
    return {'x_coords': x, 'u_x': v[:, 0], 'u_y': v[:, 1], 'u_mag': np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2)}


def compute_conc_profiles(results, *, n_points=400):
    """Horizontal and vertical concentration line profiles, statistics and full samples
    (reference analysis.py:421-542: same lines, same result keys under results['mass_metrics']).  All eight lines
    are located and evaluated in one launch."""
    c = results.get('c')
    mesh = (results.get('mesh_results') or {}).get('mesh')
    params = results.get('params', None)
    if c is None or mesh is None or params is None:
        return results
    L = float(getattr(params, 'L_dim', getattr(params, 'L', 1.0)))
    H = float(getattr(params, 'H_dim', getattr(params, 'H', 1.0)))
    domain_type = results.get('domain_type', None)
    if domain_type is None:
        h_dim = getattr(params, 'sulci_h_dim', 0.0)
        domain_type = 'sulcus' if (h_dim and h_dim > 0) else 'rectangular'
        results['domain_type'] = domain_type
    mass_metrics = results.setdefault('mass_metrics', {})

    def _stats(vals):
        vals = np.asarray(vals)
        if vals.size == 0:
            return {'min_c': None, 'max_c': None, 'avg_c': None, 'n_samples': 0}
        return {'min_c': float(np.min(vals)), 'max_c': float(np.max(vals)), 'avg_c': float(np.mean(vals)),
                'n_samples': int(vals.size)}
    vert_lines = [(0.25 * L, "x_quarter"), (0.50 * L, "x_mid"), (0.75 * L, "x_three_quarters")]
    horiz_lines = [(1e-6 * H, "mouth_level"), (0.25 * H, "lower_channel"), (0.50 * H, "mid_channel"),
                   (0.75 * H, "upper_channel")]
    if domain_type == 'rectangular':
        x_range, y_range = (0.0, float(L)), (0.0, float(H))
    else:
        coords = mesh.coordinates()
        y_min = float(coords[:, 1].min())
        horiz_lines = [(0.5 * (y_min + 0.0), "sulcus_mid")] + horiz_lines
        x_range, y_range = (float(coords[:, 0].min()), float(coords[:, 0].max())), None
    try:
        c.set_allow_extrapolation(True)
    except Exception:
        pass
    lines = [(float(y), 'h', x_range) for y, _ in horiz_lines] + [(float(x), 'v', y_range) for x, _ in vert_lines]
    profs = _eval_lines(c, mesh, lines, n_points)
    profiles_stats = {'horizontal': {}, 'vertical': {}}
    profiles_full = {'horizontal': {}, 'vertical': {}}
    for (y_loc, name), (xs, cs) in zip(horiz_lines, profs[:len(horiz_lines)]):
        s = _stats(cs)
        if s['n_samples'] > 0:
            profiles_stats['horizontal'][name] = {'y': float(y_loc), **s}
            profiles_full['horizontal'][name] = {'y': float(y_loc), 'x': np.asarray(xs).tolist(), 'c': np.asarray(cs).tolist()}
    for (x_loc, name), (ys, cs) in zip(vert_lines, profs[len(horiz_lines):]):
        s = _stats(cs)
        if s['n_samples'] > 0:
            profiles_stats['vertical'][name] = {'x': float(x_loc), **s}
            profiles_full['vertical'][name] = {'x': float(x_loc), 'y': np.asarray(ys).tolist(), 'c': np.asarray(cs).tolist()}
    mass_metrics['profiles'] = profiles_stats
    mass_metrics['profiles_full'] = profiles_full
    mass_metrics['profiles_meta'] = {
        'n_points': int(n_points), 'domain_type': domain_type,
        'x_range': tuple(map(float, x_range)) if x_range is not None else None,
        'y_range': tuple(map(float, y_range)) if y_range is not None else None}
    return results


def compute_velocity_metrics(u, mesh_results, params, rng=None):
    """Key velocity metrics (reference analysis.py:721-830, same keys): statistics of 4 horizontal and 3 vertical
    line profiles (100 samples each) and of u at <= 1000 randomly chosen mesh vertices.  All 700 line samples go
    through one device launch; the vertex sample needs no location (vertex dofs are nodal values).  Like the
    reference the vertex sample is unseeded unless ``rng`` (a numpy Generator) is given; modes without flow
    return {} (:733-734)."""
    if u is None:
        return {}
    mesh = mesh_results['mesh']
    mode = getattr(params, 'mode', 'unknown')
    if mode not in ['adv-diff', 'no-uptake']:
        return {}
    try:
        L, H, sulcus_w = params.L, params.H, params.sulci_w
        xc = L / 2
        horizontal_lines = [(1e-6 * H, "mouth_level"), (0.25 * H, "lower_channel"), (0.50 * H, "mid_channel"),
                            (0.75 * H, "upper_channel")]
        vertical_lines = [(xc - sulcus_w / 2, "sulcus_leading"), (xc, "sulcus_center"), (xc + sulcus_w / 2, "sulcus_trailing")]
        hl = [(y, n) for y, n in horizontal_lines if 0 <= y <= H]
        vl = [(x, n) for x, n in vertical_lines if 0 <= x <= L]
        profs = _eval_lines(u, mesh, [(float(y), 'h', (0, L)) for y, _ in hl] + [(float(x), 'v', (0, H)) for x, _ in vl], 100)
        vm = {}
        for (_, name), (_, v) in zip(hl, profs[:len(hl)]):
            ok = len(v) > 0
            mag = np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2) if ok else None
            vm[f'max_ux_{name}'] = np.max(np.abs(v[:, 0])) if ok else 0
            vm[f'max_umag_{name}'] = np.max(mag) if ok else 0
            vm[f'avg_ux_{name}'] = np.mean(np.abs(v[:, 0])) if ok else 0
            vm[f'avg_umag_{name}'] = np.mean(mag) if ok else 0
        for (_, name), (_, v) in zip(vl, profs[len(hl):]):
            ok = len(v) > 0
            mag = np.sqrt(v[:, 0] ** 2 + v[:, 1] ** 2) if ok else None
            vm[f'max_umag_{name}'] = np.max(mag) if ok else 0
            vm[f'max_uy_{name}'] = np.max(np.abs(v[:, 1])) if ok else 0
            vm[f'avg_umag_{name}'] = np.mean(mag) if ok else 0
            vm[f'avg_uy_{name}'] = np.mean(np.abs(v[:, 1])) if ok else 0
        coords = mesh.coordinates()
        n_sample = min(1000, len(coords))
        idx = (rng.choice(len(coords), n_sample, replace=False) if rng is not None
               else np.random.choice(len(coords), n_sample, replace=False))
        n2 = len(u.values) // 2
        gx, gy = u.values[idx], u.values[n2 + idx]             # P2 vertex dofs = values at the vertices
        gm = np.sqrt(gx ** 2 + gy ** 2)
        if len(gm):
            vm.update({'global_max_umag': np.max(gm), 'global_avg_umag': np.mean(gm),
                       'global_max_ux': np.max(np.abs(gx)), 'global_avg_ux': np.mean(np.abs(gx)),
                       'global_max_uy': np.max(np.abs(gy)), 'global_avg_uy': np.mean(np.abs(gy))})
        else:
            vm.update({k: 0 for k in ('global_max_umag', 'global_avg_umag', 'global_max_ux', 'global_avg_ux',
                                      'global_max_uy', 'global_avg_uy')})
        return vm
    except Exception as e:
        print(f"⚠️ Warning: Could not extract velocity metrics: {e}")
        return {}


# ====================================================================== mu_eff
def compute_concentration_profiles(results):
    """Line integrals of c along y=0 with the channel-side trace on the mouth (analysis.py:884-946)."""
    F, _ = _cached_or_evaluate(results['c'])
    e, m = _G['y0_ext'], _G['mouth']
    C_ext, C_m, L_ext, L_m = (float(F[e, _C]), float(F[m, _C]), float(F[e, _LEN]), float(F[m, _LEN]))
    tot_L = L_ext + L_m
    return {'C_y0_ext': C_ext, 'C_mouth': C_m, 'C_y0_total': C_ext + C_m,
            'lengths': {'L_y0_ext': L_ext, 'L_mouth': L_m, 'L_y0_total': tot_L},
            'means': {'mean_y0_ext': C_ext / L_ext if L_ext > 0 else np.nan,
                      'mean_mouth': C_m / L_m if L_m > 0 else np.nan,
                      'mean_y0_total': (C_ext + C_m) / tot_L if tot_L > 0 else np.nan}}


@functools.lru_cache(maxsize=256)
def _arc_integral(h, w):
    """int_0^1 sqrt(1 + (pi h / w cos(pi s))^2) ds by the reference's adaptive quadrature (same call, same tolerances);
    memoised per geometry -- a mu sweep asks for it once per case."""
    val, _ = quad(lambda s: np.sqrt(1.0 + (np.pi * h / w * np.cos(np.pi * s)) ** 2), 0.0, 1.0,
                  epsabs=1e-10, epsrel=1e-10, limit=200)
    return float(val)


def compute_mu_eff_arc(results):
    """mu (1 + (L_sulcus - w)/L) with the arc length by adaptive quadrature (analysis.py:948-970)."""
    p = results['params']
    L, h, w, mu = float(p.L), float(p.sulci_h), float(p.sulci_w), float(p.mu)
    if w <= 0 or h <= 0 or L <= 0:
        return None
    return float(mu * (1.0 + (w * _arc_integral(h, w) - w) / L))


def compute_mu_eff_enh(results, kappa=10.0):
    """analysis.py:972-985."""
    p = results['params']
    L, h, w, mu = float(p.L), float(p.sulci_h), float(p.sulci_w), float(p.mu)
    if L <= 0 or mu < 0 or w <= 0:
        return None
    f = 1.0 / np.sqrt(1.0 + kappa * mu * (h ** 2) / w)
    return float(mu * ((L - w) / L + (w / L) * f))


def _flux_total(results, keys):
    pf = results.get('flux_metrics', {}).get('sulcus_specific', {}).get('physical_flux', {})
    for k in keys:
        if k in pf and 'total' in pf[k]:
            return float(pf[k]['total'])
    return None


def compute_mu_eff_sim(results, conc=None):
    conc = conc if conc is not None else compute_concentration_profiles(results)
    C_y0 = conc['C_y0_total']
    J = _flux_total(results, ('y0_flux', 'y0_combined'))
    if not np.isfinite(C_y0) or C_y0 <= 0.0 or J is None:
        return None
    return float(J / C_y0)


def compute_mu_eff_sim_mouth(results, conc=None):
    conc = conc if conc is not None else compute_concentration_profiles(results)
    C_s = conc['C_mouth']
    J = _flux_total(results, ('opening', 'mouth', 'y0_opening', 'y0_mouth', 'sulcus_opening'))
    if not np.isfinite(C_s) or C_s <= 0.0 or J is None:
        return None
    return float(J / C_s)


def compute_mu_eff_metrics(results, kappa=10.0):
    """Reference analysis.py:1033-1097: same report dictionary."""
    mu = float(results['params'].mu)
    conc = compute_concentration_profiles(results)
    arc, enh = compute_mu_eff_arc(results), compute_mu_eff_enh(results, kappa=kappa)
    sim, opn = compute_mu_eff_sim(results, conc=conc), compute_mu_eff_sim_mouth(results, conc=conc)

    def ratio(x, y):
        return float(x / y) if (x is not None and y not in (None, 0.0)) else None

    def pct(approx, truth):
        return None if (truth in (None, 0.0) or approx is None) else float(abs(approx - truth) / abs(truth) * 100.0)
    return {'mu_eff_arc': arc, 'mu_eff_enh': enh, 'mu_eff_sim': sim, 'mu_eff_open': opn,
            'ratios': {'arc': ratio(arc, mu), 'enh': ratio(enh, mu), 'sim': ratio(sim, mu), 'open': ratio(opn, mu)},
            'errors_vs_sim': {'arc': pct(arc, sim), 'enh': pct(enh, sim), 'open': pct(opn, sim)},
            'audit': {'concentrations': {k: conc[k] for k in ('C_y0_ext', 'C_mouth', 'C_y0_total')},
                      'lengths': conc.get('lengths', {}), 'means': conc.get('means', {}),
                      'fluxes': {'J_y0_total': _flux_total(results, ('y0_flux', 'y0_combined')),
                                 'J_sigma_mouth': _flux_total(results, ('opening', 'mouth', 'y0_opening', 'y0_mouth', 'sulcus_opening'))}}}
