"""Sweep throughput (BASELINE configs[3]: independent mu / Pe cases): solves per second through the study drivers.

    python tools/sweep_bench.py [--h 0.02]            (under torchrun: cases are dealt round-robin to the ranks)

Two reference studies at the reference's resolution:
  * Phase A mu sweep -- 20 mu values on ONE geometry (0.25 x 0.25 mm sulcus): mesh, patterns, hierarchy and device
    problems are built once, every further mu is assemble + CG solve + functionals;
  * adv-diff Pe x mu validation -- 9 sulcus + 9 step-rectangle solves on two meshes, one Stokes solve per mesh.
Wall clock around the driver calls (host mesh generation and set-up included -- that is what a user waits for), plus the
marginal time per case once the geometry is cached.
"""
import argparse
import contextlib
import io
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h', type=float, default=0.02)
    ap.add_argument('--streams', default='1,2,4,8')     # concurrent cases per GPU (worker threads, one CUDA stream each)
    args = ap.parse_args()
    import torch
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f"cuda:{int(os.environ.get('LOCAL_RANK', 0))}"))
    from sulcusfem import studies, simulation
    out = {'n_gpus': world, 'h': args.h}

    def timed(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, time.perf_counter() - t
    # Phase A mu sweep: first call builds the geometry, second call (same cases) runs on the cached one
    df, t_cold = timed(lambda: studies.run_mu_sweep(None, mesh_size_dim=args.h))
    df, t_warm = timed(lambda: studies.run_mu_sweep(None, mesh_size_dim=args.h))
    out['mu_sweep'] = {'cases': len(df), 'wall_s_first': t_cold, 'wall_s_cached_geometry': t_warm,
                       'solves_per_s_first': len(df) / t_cold, 'solves_per_s_cached_geometry': len(df) / t_warm,
                       'ms_per_case_cached_geometry': 1e3 * t_warm / max(len(df) / world, 1)}
    # the same 20 cases with several cases in flight per GPU (sweep.run_concurrent): first call per stream count builds
    # that many sets of device problems, the second one is the steady state
    out['mu_sweep_concurrent'] = {}
    for k in [int(v) for v in str(args.streams).split(',') if v.strip() and int(v) > 1]:
        dfk, tk0 = timed(lambda: studies.run_mu_sweep(None, mesh_size_dim=args.h, streams=k))
        dfk, tk = timed(lambda: studies.run_mu_sweep(None, mesh_size_dim=args.h, streams=k))
        same = bool((abs(dfk['Mu_Eff_Simulation'] - df['Mu_Eff_Simulation']) <= 1e-10 * abs(df['Mu_Eff_Simulation'])).all())
        out['mu_sweep_concurrent'][str(k)] = {'wall_s_first': tk0, 'wall_s_cached_geometry': tk,
                                              'solves_per_s_cached_geometry': len(dfk) / tk, 'rows_match_serial': same}
    # a longer sweep (100 mu values) so that thread start-up does not dominate; 3 repetitions per stream count (worker
    # threads share the GIL: the concurrent numbers scatter from run to run -- median and best are reported)
    many = {'dense': [float(v) for v in __import__('numpy').geomspace(0.1, 150.0, 100)]}
    # batched = the default (up to 8 mu per Krylov loop, sfem_krylov_cg_batch); per_case = one CG solve per mu
    ref_rows = None
    for mode, batch in (('mu_sweep_100', True), ('mu_sweep_100_per_case', False)):
        out[mode] = {}
        for k in [1] + [int(v) for v in str(args.streams).split(',') if v.strip() and int(v) > 1]:
            studies.run_mu_sweep(None, regimes={'dense': many['dense'][:max(k, 2)]}, mesh_size_dim=args.h, streams=k, batch=batch)
            ts = []
            for _ in range(3):
                dfm, tm = timed(lambda: studies.run_mu_sweep(None, regimes=many, mesh_size_dim=args.h, streams=k, batch=batch))
                ts.append(tm)
            ts.sort()
            out[mode][str(k)] = {'wall_s_median': ts[1], 'solves_per_s_median': len(dfm) / ts[1],
                                 'solves_per_s_best': len(dfm) / ts[0], 'solves_per_s_worst': len(dfm) / ts[2]}
            if ref_rows is None:
                ref_rows = dfm
            out[mode][str(k)]['rows_match_first_run'] = bool(
                (abs(dfm['Mu_Eff_Simulation'] - ref_rows['Mu_Eff_Simulation']) <= 1e-9 * abs(ref_rows['Mu_Eff_Simulation'])).all())
    dfp, tp = timed(lambda: studies.run_mu_sweep(None, regimes=many, mesh_size_dim=args.h, streams=1, frozen_coarse=False, batch=False))
    out['mu_sweep_100_per_case']['1_full_reassembly_every_case'] = {'solves_per_s': len(dfp) / tp}
    # the batched solves alone (device time of the Krylov loops, no per-case post-processing): 100 mu = 13 batches
    from sulcusfem import solvers
    from sulcusfem.fem import FunctionSpace
    p0 = studies.Parameters(mode='no-adv', mesh_size_dim=args.h)
    p0.sulci_w_dim = p0.sulci_h_dim = 0.25
    p0.validate()
    p0.nondim()
    with contextlib.redirect_stdout(io.StringIO()):
        mr = simulation._simulation_generate_mesh(p0, 'sulcus')
        Cs = FunctionSpace(mr['mesh'], "CG", 2)
        mus100 = [p0.mu * f for f in many['dense']]
        solvers.pure_diffusion_solver_batch(mr, Cs, p0.D, mus100[:8])
        fs, tb = timed(lambda: solvers.pure_diffusion_solver_batch(mr, Cs, p0.D, mus100))
    out['batched_solver_only'] = {'solves': len(fs), 'wall_s': tb, 'solves_per_s': len(fs) / tb,
                                  'iterations': sorted({f.solver_info['iterations'] for f in fs})}
    df2, t2 = timed(lambda: studies.run_advdiff_step_validation(None, mesh_size_dim=args.h))
    df2, t2w = timed(lambda: studies.run_advdiff_step_validation(None, mesh_size_dim=args.h))
    out['advdiff_validation'] = {'solves': len(df2), 'wall_s_first': t2, 'wall_s_cached_geometry': t2w,
                                 'solves_per_s_first': len(df2) / t2, 'solves_per_s_cached_geometry': len(df2) / t2w}
    if rank == 0:
        print(json.dumps(out))
        os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
        with open(os.path.join(ROOT, 'gpurun_out', f'sweep_bench_n{world}.json'), 'w') as f:
            json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
