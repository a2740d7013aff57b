"""Device time of the batched Robin sweep (sfem_krylov_cg_batch) at the reference's mesh size, per batch width.

    python tools/batch_profile.py [--h 0.02] [--nbs 1,2,4,8]            CUDA-event timings (median of --reps)
    SFEM_GRAPHS=0 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/batch_launches.csv \
        python tools/batch_profile.py --ncu                            one un-graphed nb = 8 solve for a launch list

The geometry is the one of the reference's mu sweep (0.25 x 0.25 mm sulcus, no_advection_analysis_A.py:1264-1266).
"""
import argparse
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h', type=float, default=0.02)
    ap.add_argument('--nbs', default='1,2,4,8,16')
    ap.add_argument('--reps', type=int, default=7)
    ap.add_argument('--ncu', action='store_true')
    args = ap.parse_args()
    import numpy as np
    import torch
    torch.cuda.set_device(0)
    from sulcusfem import simulation, solvers, studies
    p0 = studies.Parameters(mode='no-adv', mesh_size_dim=args.h)
    p0.sulci_w_dim = p0.sulci_h_dim = 0.25
    p0.validate()
    p0.nondim()
    with contextlib.redirect_stdout(io.StringIO()):
        mr = simulation._simulation_generate_mesh(p0, 'sulcus')
    prob = solvers.scalar_problem(mr['mesh'], mr['bc_markers'], 4)
    bcv = {1: 1.0, 2: 0.0}
    out = {'h': args.h, 'n_p2': prob.n, 'levels': [l.n for l in prob.levels], 'batches': {}}
    if args.ncu:
        mus = [p0.mu * f for f in np.geomspace(0.5, 4.0, 8)]
        prob.solve_batch(p0.D, mus, bcv)
        torch.cuda.synchronize()
        print('batch_profile ncu pass done')
        return
    for nb in [int(v) for v in args.nbs.split(',')]:
        mus = [p0.mu * f for f in np.geomspace(0.5, 4.0, nb)]
        prob.solve_batch(p0.D, mus, bcv)                          # warm-up: graph capture, work space
        ts, its = [], 0
        for _ in range(args.reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _, infos = prob.solve_batch(p0.D, mus, bcv)
            e1.record()
            e1.synchronize()
            ts.append(e0.elapsed_time(e1))
            its = infos[0]['iterations']
        ts.sort()
        ms = ts[len(ts) // 2]
        out['batches'][str(nb)] = {'ms_per_batch': ms, 'ms_per_solve': ms / nb, 'iterations': its,
                                   'ms_per_iteration_upper_bound': ms / max(its + 1, 1)}
        if nb > 1:                                                # first version: coarse levels of mu_ref for all columns
            _, infos = prob.solve_batch(p0.D, mus, bcv, shared_coarse=True)
            out['batches'][str(nb)]['iterations_shared_coarse_levels'] = infos[0]['iterations']
            wide = [p0.mu * f for f in np.geomspace(0.1, 6.4, nb)]    # a batch at the span limit of solvers.BATCH_SPAN
            _, infos = prob.solve_batch(p0.D, wide, bcv)
            out['batches'][str(nb)]['iterations_span_64'] = infos[0]['iterations']
    # the single-solve path on the same problem for comparison
    prob.assemble(p0.D, mu_const=p0.mu, bc_values=bcv)
    prob.solve('cg', rtol=1e-13)
    ts = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        prob.assemble(p0.D, mu_const=p0.mu, bc_values=bcv)
        prob.solve('cg', rtol=1e-13)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    out['single'] = {'ms_per_solve': ts[len(ts) // 2], 'iterations': prob.last_info['iterations']}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'batch_profile.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
