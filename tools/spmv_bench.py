"""SpMV micro-benchmark: FP64 CSR SpMV on P2 / Taylor-Hood patterns of rect(r) meshes.

    python tools/spmv_bench.py [--nx 2000 --ny 200] [--space p2|th] [--variants vector,staged:128:3 ...]
                               [--iters 20] [--only NAME] [--out gpurun_out/spmv_bench.json]

Every timed launch is preceded by an L2 flush (512 MiB memset) unless --no-flush; time = CUDA events
on the launching stream; achieved = (12 nnz + 20 N) / t  (SURVEY 8(d) algorithmic bytes).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))

import numpy as np  # noqa: E402


def build_matrix(nx, ny, space, sulcus=None):
    from sulcusfem import hostmesh as hm, dofmap as dm
    if sulcus:                                   # "H,R": the bench.py mesh (unstructured sulcus, h = H, R refinements)
        from sulcusfem.unstructured import mesh_domain
        h, r = sulcus.split(',')
        mesh = hm.refine_n(mesh_domain(10.0, 1.0, 0.5, 1.0, float(h), 'sulcus'), int(r))
    else:
        mesh = hm.rectangle_mesh(10.0, 1.0, nx, ny)
    if space == 'p2':
        cd = dm.p2_cell_dofs(mesh)
        n = dm.p2_num_dofs(mesh)
    else:
        cd = dm.th_cell_dofs(mesh)
        n = dm.th_num_dofs(mesh)
    pat = dm.build_pattern(n, n, [(cd, cd)])
    return n, pat.rowptr, pat.cols


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--nx', type=int, default=2000)
    ap.add_argument('--ny', type=int, default=200)
    ap.add_argument('--space', default='p2')
    ap.add_argument('--sulcus', default=None, help='H,R: use the bench.py sulcus mesh instead of rect(nx, ny)')
    ap.add_argument('--variants', default='vector,staged:1536,staged:1024,staged:2048')
    ap.add_argument('--iters', type=int, default=20)
    ap.add_argument('--nb', type=int, default=1)
    ap.add_argument('--no-flush', action='store_true')
    ap.add_argument('--flush', default='write', help='write: 512 MiB memset (leaves dirty lines); read: memset then read another 512 MiB buffer (clean L2)')
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'spmv_bench.json'))
    args = ap.parse_args()
    import torch
    from sulcusfem.device import Context, DeviceCsr
    ctx = Context.get()
    n, rowptr, cols = build_matrix(args.nx, args.ny, args.space, args.sulcus)
    nnz = len(cols)
    rng = np.random.default_rng(0)
    vals = rng.random(nnz)
    torch.manual_seed(0)
    x = torch.rand(n, dtype=torch.float64, device=ctx.device)
    y = torch.empty(n, dtype=torch.float64, device=ctx.device)
    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)
    flush2 = torch.ones(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)
    bytes_alg = 12.0 * nnz + 4.0 * n + 16.0 * n * args.nb
    peak = 6449.4
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peak = float(json.load(f)['hbm_gbs'])
    except Exception:
        pass
    ref = None
    results = []
    for v in args.variants.split(','):
        parts = v.split(':')
        staged = parts[0] == 'staged'
        sell = parts[0] == 'sell'                      # sell[:sigma]
        if staged and len(parts) > 1:
            DeviceCsr.STAGED_CAP = int(parts[1])
        if staged and len(parts) > 2:
            DeviceCsr.STAGED_ROWS = int(parts[2])
        if sell and len(parts) > 1:
            DeviceCsr.SELL_SIGMA = int(parts[1])
        ctx.lib.sfem_staged_set_min_tiles(1 if staged else 1 << 30)
        ctx.lib.sfem_sell_set_min_rows(1 if sell else 1 << 30)
        A = DeviceCsr(ctx, n, n, rowptr, cols, vals, staged=staged, sell=sell)
        nbv = args.nb
        if nbv == 2 and x.numel() == n:
            x = torch.rand(2 * n, dtype=torch.float64, device=ctx.device)
            y = torch.empty(2 * n, dtype=torch.float64, device=ctx.device)
        for _ in range(3):
            A.spmv(x, y, staged=staged, nb=nbv, sell=sell)
        torch.cuda.synchronize()
        got = y.clone()
        if ref is None:
            import scipy.sparse as sp
            M = sp.csr_matrix((vals, cols, rowptr), shape=(n, n))
            ref = torch.from_numpy(np.ascontiguousarray(M @ x.cpu().numpy().reshape(n, nbv)).ravel()).to(ctx.device)
        err = float((got - ref).norm() / ref.norm())
        ms = []
        for _ in range(args.iters):
            if not args.no_flush:
                flush.zero_()
                if args.flush == 'read':
                    flush2.sum()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            A.spmv(x, y, staged=staged, nb=nbv, sell=sell)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms = np.array(ms)
        r = {"variant": v, "nb": nbv, "ntiles": A.ntiles, "sell_fill": (A.sell['fill'] if A.sell else None), "space": args.space, "mesh": (args.sulcus or f"rect {args.nx}x{args.ny}"), "n": n, "nnz": nnz, "ms_med": float(np.median(ms)), "ms_min": float(ms.min()),
             "gbs_med": bytes_alg / 1e9 / (float(np.median(ms)) / 1e3), "gbs_best": bytes_alg / 1e9 / (float(ms.min()) / 1e3),
             "rel_err": err, "flush": ("none" if args.no_flush else args.flush)}
        r["frac_of_measured_peak"] = r["gbs_med"] / peak
        results.append(r)
        print(json.dumps(r), flush=True)
        del A
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, 'a') as f:
        for r in results:
            f.write(json.dumps(r) + "\n")


if __name__ == '__main__':
    main()
