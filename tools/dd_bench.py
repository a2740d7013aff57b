"""Domain-decomposed solve of the no-advection sulcus problem on a refined mesh (BASELINE config 5).

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/dd_bench.py --refine 2

Every rank builds the mesh / hierarchy and assembles the global operators (replicated set-up), then
 (a) solves the system alone on its GPU (the single-GPU reference time t_1 and field), and
 (b) solves it row-partitioned over the N ranks with peer-memory halo exchange / all-reduce (t_N).
Rank 0 prints one JSON line: DOFs, iterations, t_1, t_N, strong-scaling efficiency t_1 / (N t_N), and the
relative L2 difference between the distributed and the single-GPU field (parity bar: 1e-10).
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
    ap.add_argument('--refine', type=int, default=2)
    ap.add_argument('--h', type=float, default=0.02)
    ap.add_argument('--reps', type=int, default=5)
    ap.add_argument('--mu', type=float, default=1.0)
    ap.add_argument('--replicate-below', type=int, default=20000)
    args = ap.parse_args()
    import numpy as np
    import torch
    import torch.distributed as dist
    rank = int(os.environ.get('RANK', 0)); world = int(os.environ.get('WORLD_SIZE', 1))
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    from bench import build_mesh
    from sulcusfem.device import Context, ScalarProblem
    from sulcusfem.dist import DistScalarProblem
    from sulcusfem.hierarchy import build_hierarchy
    ctx = Context.get()
    t0 = time.perf_counter()
    mr = build_mesh(args.h, args.refine)
    mesh, bm = mr['mesh'], mr['bc_markers'].values
    prob = ScalarProblem(mesh, bm, hierarchy=build_hierarchy(mesh), ctx=ctx)
    prob.assemble(1.0, mu_const=args.mu, bc_values={1: 1.0, 2: 0.0})
    torch.cuda.synchronize()
    t_setup = time.perf_counter() - t0

    def timed(fn):
        ts = []
        for _ in range(args.reps):
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=ctx.device)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    prob.solve('cg', rtol=1e-13)                       # warm-up (graph capture)
    t1 = timed(lambda: prob.solve('cg', rtol=1e-13))
    x1 = prob.x.clone()
    info1 = dict(prob.last_info)
    out = {"n_gpus": world, "dofs": int(prob.n), "levels": [int(l.n) for l in prob.levels], "refine": args.refine,
           "single_gpu": {"ms": t1, **info1}, "setup_s": t_setup}
    if world > 1:
        dp = DistScalarProblem(prob, rank, world, replicate_below=args.replicate_below)
        dp.refresh()
        dp.solve(rtol=1e-13)
        tN = timed(lambda: dp.solve(rtol=1e-13))
        xl = dp.solve(rtol=1e-13)
        ref = x1[dp.owned]
        num = torch.tensor([float(((xl - ref) ** 2).sum()), float((ref ** 2).sum())], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(num)
        out["distributed"] = {"ms": tN, **dp.last_info, "distributed_levels": dp.nd,
                              "halo_dofs_level0": int(len(dp.parts[0].ghost)), "owned_level0": int(dp.parts[0].n_own)}
        out["rel_l2_vs_single_gpu"] = float((num[0] / num[1]).sqrt())
        out["strong_scaling_efficiency"] = t1 / (world * tN)
        out["dofs_per_s"] = prob.n / (tN / 1e3)
        dp.close()
    else:
        out["dofs_per_s"] = prob.n / (t1 / 1e3)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
