"""Krylov tolerance vs field error (which tolerance meets the 1e-10 field-parity bar with margin?).

    python tools/tolerance_study.py [--h-lu 0.04] [--refine 2]

Part 1 (mesh small enough for the oracle's sparse LU): Stokes MINRES and adv-diff FGMRES at a range of tolerances,
relative L2 error of u, p, c against the LU fields.  Part 2 (the bench.py mesh): iterations, time and the distance
to the rtol = 1e-14 / 1e-13 solution of the same solver (self-convergence; no LU at that size).
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))

import numpy as np  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h-lu', type=float, default=0.04)
    ap.add_argument('--refine', type=int, default=2)
    args = ap.parse_args()
    import torch
    from bench import build_mesh
    from oracle import cpu_oracle as co
    from sulcusfem import dofmap as dm
    from sulcusfem.device import Context, ScalarProblem, StokesProblem
    from sulcusfem.hierarchy import build_hierarchy
    ctx = Context.get()
    D, mu = 1.0 / 40.0, 1.0

    def problems(h, r):
        mr = build_mesh(h, r)
        mesh, bm = mr['mesh'], mr['bc_markers'].values
        hier = build_hierarchy(mesh)
        st = StokesProblem(mesh, bm, hierarchy=hier, ctx=ctx)
        sc = ScalarProblem(mesh, bm, hierarchy=hier, ctx=ctx)
        X = dm.p2_dof_coordinates(mesh)
        d1 = dm.dirichlet_dofs_p2(mesh, bm, 1)
        st.set_bcs({1: (4.0 * X[d1, 1] * (1.0 - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})
        return mesh, bm, st, sc

    def rel(a, b):
        return float(np.linalg.norm(a - b) / np.linalg.norm(b))

    def timed(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        out = fn()
        torch.cuda.synchronize()
        return out, 1e3 * (time.perf_counter() - t)

    out = {'lu_mesh': [], 'bench_mesh': []}
    # ---- part 1: against the LU oracle
    mesh, bm, st, sc = problems(args.h_lu, 0)
    om = co.Mesh(mesh.coords, mesh.cells)
    rx, ry, rp, _, _ = co.solve_stokes(om, bm, 1.0)
    rc, _, _ = co.solve_concentration(om, bm, D, mu=mu, ux=rx, uy=ry)
    un = np.concatenate([rx, ry])
    st.assemble(bc_mode=1)
    st.solve(rtol=1e-14)
    for rtol in (1e-14, 1e-13, 1e-12, 1e-11, 1e-10, 1e-9):
        (ux, uy, p), ms = timed(lambda: st.solve(rtol=rtol))
        u = np.concatenate([ux.cpu().numpy(), uy.cpu().numpy()])
        row = {'solver': 'stokes_minres', 'dofs': st.n, 'rtol': rtol, 'iterations': st.last_info['iterations'],
               'true_relres': st.last_info['relres'], 'err_u': rel(u, un), 'err_p': rel(p.cpu().numpy(), rp), 'ms': ms}
        out['lu_mesh'].append(row)
        print(json.dumps(row), flush=True)
    ux, uy, p = st.solve(rtol=1e-14)
    sc.assemble(D, ux, uy, mu_const=mu, bc_values={1: 1.0, 2: 0.0})
    sc.solve('fgmres', rtol=1e-13)
    for rtol in (1e-13, 1e-12, 1e-11, 1e-10, 1e-9):
        c, ms = timed(lambda: sc.solve('fgmres', rtol=rtol))
        row = {'solver': 'advdiff_fgmres', 'dofs': sc.n, 'rtol': rtol, 'iterations': sc.last_info['iterations'],
               'true_relres': sc.last_info['relres'], 'err_c': rel(c.cpu().numpy(), rc), 'ms': ms}
        out['lu_mesh'].append(row)
        print(json.dumps(row), flush=True)
    # sensitivity of c to the Stokes tolerance (c solved to 1e-13 every time), against the LU concentration
    for rtol in (1e-13, 1e-12, 1e-11, 1e-10):
        ux, uy, p = st.solve(rtol=rtol)
        sc.assemble(D, ux, uy, mu_const=mu, bc_values={1: 1.0, 2: 0.0})
        c = sc.solve('fgmres', rtol=1e-13)
        row = {'solver': 'chain', 'dofs': sc.n, 'stokes_rtol': rtol, 'err_c_vs_lu': rel(c.cpu().numpy(), rc)}
        out['lu_mesh'].append(row)
        print(json.dumps(row), flush=True)
    del st, sc
    # ---- part 2: the bench mesh, self-convergence
    mesh, bm, st, sc = problems(0.02, args.refine)
    st.assemble(bc_mode=1)
    st.solve(rtol=1e-14)
    (ux, uy, p), _ = timed(lambda: st.solve(rtol=1e-14))
    u0, p0 = torch.cat([ux, uy]).clone(), p.clone()
    for rtol in (1e-14, 1e-13, 1e-12, 1e-11, 1e-10):
        (ux, uy, p), ms = timed(lambda: st.solve(rtol=rtol))
        u = torch.cat([ux, uy])
        row = {'solver': 'stokes_minres', 'dofs': st.n, 'rtol': rtol, 'iterations': st.last_info['iterations'],
               'true_relres': st.last_info['relres'], 'dist_u_to_1e-14': float((u - u0).norm() / u0.norm()),
               'dist_p_to_1e-14': float((p - p0).norm() / p0.norm()), 'ms': ms}
        out['bench_mesh'].append(row)
        print(json.dumps(row), flush=True)
    ux, uy, p = st.solve(rtol=1e-14)
    sc.assemble(D, ux, uy, mu_const=mu, bc_values={1: 1.0, 2: 0.0})
    c0 = sc.solve('fgmres', rtol=1e-13).clone()
    for rtol in (1e-13, 1e-12, 1e-11, 1e-10):
        c, ms = timed(lambda: sc.solve('fgmres', rtol=rtol))
        row = {'solver': 'advdiff_fgmres', 'dofs': sc.n, 'rtol': rtol, 'iterations': sc.last_info['iterations'],
               'true_relres': sc.last_info['relres'], 'dist_c_to_1e-13': float((c - c0).norm() / c0.norm()), 'ms': ms}
        out['bench_mesh'].append(row)
        print(json.dumps(row), flush=True)
    for rtol in (1e-13, 1e-12, 1e-11, 1e-10):
        ux, uy, p = st.solve(rtol=rtol)
        sc.assemble(D, ux, uy, mu_const=mu, bc_values={1: 1.0, 2: 0.0})
        c = sc.solve('fgmres', rtol=1e-13)
        row = {'solver': 'chain', 'dofs': sc.n, 'stokes_rtol': rtol, 'dist_c_to_c(stokes 1e-14)': float((c - c0).norm() / c0.norm())}
        out['bench_mesh'].append(row)
        print(json.dumps(row), flush=True)
    os.makedirs(os.path.join(ROOT, 'gpurun_out'), exist_ok=True)
    with open(os.path.join(ROOT, 'gpurun_out', 'tolerance_study.json'), 'w') as f:
        json.dump(out, f, indent=1)


if __name__ == '__main__':
    main()
