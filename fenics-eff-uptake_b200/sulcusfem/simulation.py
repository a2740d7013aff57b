"""``run_simulation`` with the reference's signature and ``results`` schema (simulation.py:270-349).

Flow: mesh -> velocity (Stokes or zero) -> concentration -> post-processing.  The reference's
plotting, ParaView export and JSON dump (simulation.py:91-92,137-138,165,235-268) are presentation
layers outside the hot path and are not reproduced; everything the study drivers read from the
returned dictionary is there.  Meshes (and with them patterns, gather maps and the
multigrid hierarchy on the device) are cached per geometry so a mu-sweep on one geometry reuses
them, where the reference re-runs Gmsh for every case.
"""
from __future__ import annotations

import time

from .analysis import (compute_flux_metrics, compute_mass_metrics, compute_mu_eff_metrics,
                       compute_velocity_metrics)
from .fem import Constant, FunctionSpace, MixedElement, VectorFunctionSpace
from .mesh import MeshGenerator, setup_rectangular_measures, setup_sulcus_measures
from .solvers import (advdiff_solver, advdiff_solver_variable_mu, pure_diffusion_solver,
                      pure_diffusion_solver_variable_mu, stokes_solver, stokes_solver_no_adv)

_MESH_CACHE = {}
MESH_OPTIONS = {'mesher': 'delaunay', 'uniform_refinements': 0}
# ParaView export (reference simulation.py:137-138,165,300-311): off by default -- set to a directory (the
# reference's top-level "Results") to get <dir>/<Mode> Simulations/<study>/<config>/ParaView Files/{velocity,
# pressure,concentration}.pvd exactly where the reference puts them
EXPORT_BASE_DIR = None
_MODE_DIRS = {'adv-diff': 'Adv-Diff', 'no-adv': 'No Advection', 'no-uptake': 'No Uptake'}


def _paraview_dir(mode, study_type, config_name):
    if EXPORT_BASE_DIR is None:
        return None
    import os
    d = os.path.join(EXPORT_BASE_DIR, f"{_MODE_DIRS.get(mode, mode.replace('-', ' ').title())} Simulations",
                     study_type, config_name, "ParaView Files")
    os.makedirs(d, exist_ok=True)
    return d


def _export(paraview_dir, name, f):
    if paraview_dir is not None and f is not None:
        import os
        from .export import File
        f.name_ = name
        File(os.path.join(paraview_dir, name + ".pvd")) << f


def _mesh_key(params, domain_type):
    mp = params.get_mesh_generator_params()
    key = (domain_type, mp['width'], mp['height'], mp['sulcus_depth'], mp['sulcus_width'], mp['mesh_size'],
           mp['refinement_factor'], MESH_OPTIONS['mesher'], MESH_OPTIONS['uniform_refinements'])
    return key, mp


def prefetch_meshes(jobs, workers=None, with_hierarchy=True, with_plans=True):
    """Build the meshes (Delaunay + smoothing), markers and multigrid hierarchies of many geometries in parallel
    worker processes and put them into the per-geometry cache ``run_simulation`` reads.

    ``jobs``: iterable of ``(params, domain_type)``.  The host work per geometry (seconds at h = 0.02) dwarfs the
    milliseconds a case costs on the GPU once its geometry is cached, and it is embarrassingly parallel over
    geometries, so a study hands all its geometries here first.  Returns the number of meshes built.  The results
    are the ones the in-process generator produces (same code, deterministic), so a sweep gives bit-identical fields
    with and without prefetching.  With ``with_plans`` the workers also build the CSR patterns and gather maps the
    device problems will need (memoised on the mesh objects, ``dofmap.memo_pattern``), leaving only uploads and the
    sliced-ELL plans to the solving process."""
    import os
    todo, plans = {}, {}
    for params, domain_type in jobs:
        key, mp = _mesh_key(params, domain_type)
        if key not in _MESH_CACHE and key not in todo:
            mp['output_dir'] = None
            mp['domain_type'] = domain_type
            todo[key] = dict(mp, **MESH_OPTIONS)
            plans[key] = 'scalar' if getattr(params, 'mode', None) == 'no-adv' else 'stokes'
        elif key in todo and getattr(params, 'mode', None) != 'no-adv':
            plans[key] = 'stokes'
    if not todo:
        return 0
    if workers is None:
        world = int(os.environ.get('WORLD_SIZE', '1') or 1)
        workers = max(1, min(len(todo), (os.cpu_count() or 1) // max(world, 1), 16))
    from .mesh import generate_mesh_job
    if workers <= 1 or len(todo) == 1:
        results = {k: generate_mesh_job(kw, with_hierarchy, plans.get(k) if with_plans else None) for k, kw in todo.items()}
    else:
        import multiprocessing as mp_
        import warnings
        from concurrent.futures import ProcessPoolExecutor
        results = {}
        try:
            # spawn: the parent may hold a CUDA context, which must not be forked
            with ProcessPoolExecutor(max_workers=workers, mp_context=mp_.get_context('spawn')) as ex:
                futs = {k: ex.submit(generate_mesh_job, kw, with_hierarchy, plans.get(k) if with_plans else None)
                        for k, kw in todo.items()}
                for k, f in futs.items():
                    results[k] = f.result()
        except Exception as e:          # e.g. no importable __main__ (python -c / stdin): spawned workers cannot start
            warnings.warn(f"parallel mesh prefetch unavailable ({type(e).__name__}: {e}); meshing in-process")
            for k, kw in todo.items():
                if k not in results:
                    results[k] = generate_mesh_job(kw, with_hierarchy, plans.get(k) if with_plans else None)
    for k, (mesh_results, hier) in results.items():
        _MESH_CACHE[k] = mesh_results
        if hier is not None and mesh_results:
            cache = getattr(mesh_results['mesh'], '_sfem_cache', None)
            if cache is None:
                cache = {}
                mesh_results['mesh']._sfem_cache = cache
            cache.setdefault('hierarchy', hier)
    return len(results)


def _simulation_generate_mesh(params, domain_type, mesh_dir=None, paraview_dir=None):
    print("\n Generating mesh...")
    key, mp = _mesh_key(params, domain_type)
    if key not in _MESH_CACHE:
        mp['output_dir'] = mesh_dir
        mp['domain_type'] = domain_type
        _MESH_CACHE[key] = MeshGenerator(**mp, **MESH_OPTIONS).generate_mesh()
    mesh_results = _MESH_CACHE[key]
    if mesh_results:
        info = mesh_results['mesh_info']
        print("✓ Mesh generated successfully!")
        print(f"      • Vertices: {info['num_vertices']:,}")
        print(f"      • Elements: {info['num_cells']:,}")
        print(f"      • h_min: {info['hmin']:.6f}")
        print(f"      • h_max: {info['hmax']:.6f}")
        return mesh_results
    print("✗ Mesh generation failed!")
    return None


def _simulation_generate_vel(mode, domain_type, params, mesh_results, paraview_dir=None):
    print("\n Generating velocity field...")
    mesh = mesh_results['mesh']
    V = VectorFunctionSpace(mesh, "P", 2)
    Q = FunctionSpace(mesh, "P", 1)
    W = FunctionSpace(mesh, MixedElement([V.ufl_element(), Q.ufl_element()]))
    if mode == 'no-adv':
        return stokes_solver_no_adv(V, Q)
    cache = mesh_results.setdefault('_stokes_solution', {})
    key = (float(params.L), float(params.H))
    if key not in cache:                       # Stokes depends on the geometry only, not on Pe or mu
        cache[key] = stokes_solver(mesh_results, W, params.L, params.H, domain_type)
    return cache[key]


def _simulation_generate_conc(u, mode, domain_type, params, mesh_results, paraview_dir=None, mu_variable=False):
    print("\n Generating concentration field...")
    mesh = mesh_results['mesh']
    C = FunctionSpace(mesh, "CG", 2)
    D_const = Constant(params.D)
    mu_val = params.mu
    if isinstance(mu_val, (int, float)):
        mu_val = Constant(mu_val)
    if mode == 'no-adv':
        if mu_variable:
            return pure_diffusion_solver_variable_mu(mesh_results, C, D_const, mu_val, domain_type)
        return pure_diffusion_solver(mesh_results, C, D_const, mu_val, domain_type)
    if mu_variable:
        return advdiff_solver_variable_mu(mesh_results, u, C, D_const, mu_val, domain_type)
    return advdiff_solver(mesh_results, u, C, D_const, mu_val, domain_type)


def _simulation_post_process(domain_type, params, mesh_results, c, u, p):
    print("\n Analysing results...")
    mesh = mesh_results['mesh']
    bc_markers = mesh_results['bc_markers']
    if domain_type == 'sulcus':
        ds_bc, ds_bottom, dS_bottom, ds_y0, dS_y0, dx_dom = setup_sulcus_measures(
            mesh, bc_markers, mesh_results['bottom_segment_markers'], mesh_results['y0_markers'],
            mesh_results['domain_markers'])
        measures = {'ds_bc': ds_bc, 'ds_bottom': ds_bottom, 'dS_bottom': dS_bottom, 'ds_y0': ds_y0, 'dS_y0': dS_y0,
                    'dx_domain_sulc': dx_dom}
    else:
        ds_bc, dx_dom = setup_rectangular_measures(mesh, bc_markers)
        measures = {'ds_bc': ds_bc, 'dx_domain_rect': dx_dom}
    flux_metrics = compute_flux_metrics(c, u, mesh_results, domain_type, measures, params.D, params.mu)
    mass_metrics = compute_mass_metrics(c, measures, domain_type)
    vel_metrics = compute_velocity_metrics(u, mesh_results, params)
    results = {'c': c, 'u': u, 'p': p, 'mass_metrics': mass_metrics, 'flux_metrics': flux_metrics,
               'vel_metrics': vel_metrics, 'params': params, 'mesh_results': mesh_results, 'measures': measures}
    if domain_type == 'sulcus':
        results['mu_eff_comparison'] = compute_mu_eff_metrics(results)
    return results


def _simulation_save_results(results, filename):
    """JSON summary of one run (reference simulation.py:235-262: same keys -- params, mass_metrics, flux_metrics,
    mesh_info, mu_eff_comparison).  In the reference this call always ends in its ``except`` branch because
    ``Parameters.to_dict`` raises; here ``to_dict`` works, so the file is written.  numpy scalars and
    arrays inside the metric dictionaries are converted to plain Python."""
    import json

    def plain(o):
        import numpy as np
        if isinstance(o, dict):
            return {str(k): plain(v) for k, v in o.items()}
        if isinstance(o, (list, tuple)):
            return [plain(v) for v in o]
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, np.generic):
            return o.item()
        return o
    try:
        mesh = (results.get('mesh_results', {}) or {}).get('mesh')
        mesh_info = {}
        if mesh:
            mesh_info = {'num_vertices': int(mesh.num_vertices), 'num_cells': int(mesh.num_cells), 'hmin': mesh.hmin(),
                         'hmax': mesh.hmax()}
        out = {'params': plain(results['params'].to_dict()), 'mass_metrics': plain(results['mass_metrics']),
               'flux_metrics': plain(results['flux_metrics']), 'mesh_info': mesh_info,
               'mu_eff_comparison': plain(results.get('mu_eff_comparison', None))}
        with open(filename, 'w') as f:
            json.dump(out, f, indent=4)
        print(f"✓ Results saved to {filename}")
    except Exception as e:
        print(f"Error saving results: {e}")


def run_simulation(mode, study_type, config_name, domain_type, params, mu_variable=False):
    """Run one case; same arguments and result dictionary as the reference's ``run_simulation``."""
    start_time = time.time()
    valid_modes = ['adv-diff', 'no-adv', 'no-uptake']
    if mode not in valid_modes:
        raise ValueError(f"Invalid mode '{mode}'. Must be one of: {valid_modes}")
    valid_domain_types = ['sulcus', 'rectangular']
    if domain_type not in valid_domain_types:
        raise ValueError(f"Invalid domain type '{domain_type}'. Must be one of: {valid_domain_types}")
    paraview_dir = _paraview_dir(mode, study_type, config_name)
    mesh_results = _simulation_generate_mesh(params, domain_type)
    u, p = _simulation_generate_vel(mode, domain_type, params, mesh_results)
    _export(paraview_dir, "velocity", u)
    _export(paraview_dir, "pressure", p)
    c = _simulation_generate_conc(u, mode, domain_type, params, mesh_results, mu_variable=mu_variable)
    _export(paraview_dir, "concentration", c)
    results = _simulation_post_process(domain_type, params, mesh_results, c, u, p)
    print(f"\n✓ Simulation completed in {time.time() - start_time:.1f}s")
    return results
