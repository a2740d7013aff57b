"""cProfile of the host side of a cached-geometry mu sweep (where do the milliseconds per case go?).
    python tools/sweep_profile.py [--h 0.02] [--n 60]"""
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
    ap.add_argument('--n', type=int, default=60)
    args = ap.parse_args()
    import numpy as np
    import torch
    torch.cuda.set_device(0)
    from sulcusfem import studies
    many = {'dense': [float(v) for v in np.geomspace(0.1, 150.0, args.n)]}
    studies.run_mu_sweep(None, regimes={'dense': many['dense'][:2]}, mesh_size_dim=args.h)
    pr = cProfile.Profile()
    pr.enable()
    studies.run_mu_sweep(None, regimes=many, mesh_size_dim=args.h)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats('tottime').print_stats(35)
    st.sort_stats('cumulative').print_stats(45)


if __name__ == '__main__':
    main()
