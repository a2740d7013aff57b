"""cProfile of the host set-up of the bench case (mesh, markers, hierarchy, device problems): where do the seconds go?
    python tools/setup_profile.py [--h 0.02] [--refine 2]"""
import argparse
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h', type=float, default=0.02)
    ap.add_argument('--refine', type=int, default=2)
    args = ap.parse_args()
    import torch
    torch.cuda.set_device(0)
    import bench
    from sulcusfem.device import Context
    ctx = Context.get()
    torch.zeros(1, device=ctx.device)
    torch.cuda.synchronize()
    pr = cProfile.Profile()
    pr.enable()
    case = bench.Case(ctx, args.h, args.refine, bench.MU)
    torch.cuda.synchronize()
    pr.disable()
    print(case.setup)
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(28)
    st.sort_stats('cumulative').print_stats(45)


if __name__ == '__main__':
    main()
