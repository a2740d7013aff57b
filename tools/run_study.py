"""Run one of the reference's studies on the CUDA path and write its CSV (the data path of the study scripts; no plots).

    python tools/run_study.py STUDY --out DIR [--h 0.02]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/run_study.py STUDY --out DIR

STUDY: phase_a_mu | phase_a_aspect | phase_b | advdiff | no_uptake | profiles   (sulcusfem/studies.py).
Under torchrun the independent cases are dealt round-robin to the ranks (one process per GPU, no data-path collective) and
rank 0 writes the files.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('study', choices=['phase_a_mu', 'phase_a_aspect', 'phase_b', 'advdiff', 'no_uptake', 'profiles'])
    ap.add_argument('--out', required=True)
    ap.add_argument('--h', type=float, default=0.02, help='mesh_size_dim (the reference uses 0.02)')
    args = ap.parse_args()
    import torch
    rank, world = int(os.environ.get('RANK', 0)), int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    from sulcusfem import studies
    t0 = time.perf_counter()
    if args.study == 'phase_a_mu':
        df = studies.run_mu_sweep(args.out, mesh_size_dim=args.h)
    elif args.study == 'phase_a_aspect':
        df = studies.run_aspect_ratio_analysis(args.out, mesh_size_dim=args.h)
    elif args.study == 'phase_b':
        df = studies.run_no_adv_mu_sweep(args.out, mesh_size_dim=args.h)
    elif args.study == 'advdiff':
        df = studies.run_advdiff_step_validation(args.out, mesh_size_dim=args.h)
    elif args.study == 'no_uptake':
        df = studies.run_geometry_study(args.out, mesh_size_dim=args.h)
    else:
        out = studies.run_profile_export(args.out, mesh_size_dim=args.h) if rank == 0 else {}
        df = None
        rows = sum(len(v[0]) for v in out.values())
    torch.cuda.synchronize()
    if rank == 0:
        n = len(df) if df is not None else rows
        print(json.dumps({'study': args.study, 'rows': n, 'n_gpus': world, 'wall_s': time.perf_counter() - t0, 'out': args.out}))


if __name__ == '__main__':
    main()
