"""Time the multigrid V-cycle (the dominant cost of every Krylov iteration) on the velocity hierarchy.

    python tools/vcycle_bench.py --refine 2 --nb 2 [--reps 20]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--refine', type=int, default=2)
    ap.add_argument('--nb', type=int, default=2)
    ap.add_argument('--reps', type=int, default=20)
    args = ap.parse_args()
    import numpy as np
    import torch
    from bench import build_mesh
    from sulcusfem.device import Context, ScalarProblem
    ctx = Context.get()
    mr = build_mesh(0.02, args.refine)
    prob = ScalarProblem(mr['mesh'], mr['bc_markers'].values, dirichlet_ids=(1, 4, 3), robin_id=None, ctx=ctx, nb=args.nb)
    prob.assemble(1.0, robin=False)
    torch.manual_seed(0)
    r = torch.rand(prob.n * args.nb, dtype=torch.float64, device=ctx.device)
    z = torch.empty_like(r)
    for _ in range(3):
        prob.mg.vcycle(r, z)
    torch.cuda.synchronize()
    ms = []
    for _ in range(args.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prob.mg.vcycle(r, z); e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    A = prob.fine.A
    fine_bytes = 4 * (12.0 * A.nnz + 12.0 * prob.n) + (48 + 32 + 24 + 24) * 8.0 / 8 * args.nb * prob.n
    print(json.dumps({"lib": os.environ.get('SFEM_LIB', 'default'), "refine": args.refine, "nb": args.nb, "n": prob.n,
                      "nnz": A.nnz, "vcycle_ms_med": float(np.median(ms)), "vcycle_ms_min": float(min(ms)),
                      "fine_level_GBs_if_all_time_were_fine_level": fine_bytes / 1e9 / (float(np.median(ms)) / 1e3)}))


if __name__ == '__main__':
    main()
