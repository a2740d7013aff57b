"""A short un-graphed pass over the bench workload for `ncu --set full` captures of the hot kernels.

    SFEM_GRAPHS=0 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
        -k regex:'EpiChebPtr<2>|EpiResidD0Ptr<2>|k_elem_p2|k_gather|k_facet_functionals|k_cell_functionals' -c 40 \
        -o gpurun_out/r02_hot python tools/ncu_targets.py --refine 2 --iters 2

Builds the bench case (bench.Case), then runs: Stokes assembly + `iters` MINRES iterations, adv-diff assembly +
`iters` FGMRES iterations, functionals -- every kernel family of the step appears at its real size, system-level
launches first in every iteration.  Nothing here is a timing.
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--refine', type=int, default=2)
    ap.add_argument('--h', type=float, default=0.02)
    ap.add_argument('--iters', type=int, default=2)
    args = ap.parse_args()
    import torch
    import bench
    from sulcusfem.device import Context
    torch.cuda.set_device(0)
    ctx = Context.get()
    case = bench.Case(ctx, args.h, args.refine, bench.MU)
    st, sc = case.stokes, case.scalar
    st.assemble(bc_mode=1)
    ux, uy, p = st.solve(rtol=1e-12, maxit=args.iters)
    sc.assemble(case.D, ux, uy, mu_const=case.mu, bc_values={1: 1.0, 2: 0.0})
    c = sc.solve('fgmres', rtol=1e-13, maxit=args.iters)
    case.plan.evaluate(c, ux, uy, D=case.D, mu_const=case.mu)
    torch.cuda.synchronize()
    print("ncu_targets done", st.last_info, sc.last_info)


if __name__ == '__main__':
    main()
