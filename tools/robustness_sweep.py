"""Robustness sweep: every geometry of the reference's catalogue (parameters.py:365-402, 23 sulci from 0.05 x 0.05 mm to
1.0 x 3.0 mm) through the three solver families on the GPU path, at a moderate mesh size -- did every case mesh, converge
and produce finite metrics?  Prints one JSON summary; exits non-zero on any failure.

    python tools/robustness_sweep.py [--h 0.04] [--streams 4]
"""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--h', type=float, default=0.04)
    ap.add_argument('--streams', type=int, default=4)
    args = ap.parse_args()
    import numpy as np
    import torch
    torch.cuda.set_device(0)
    from sulcusfem import studies
    out = {'h': args.h}
    bad = []
    with warnings.catch_warnings(record=True) as wlist:
        warnings.simplefilter('always')
        t = time.perf_counter()
        dfb = studies.run_no_adv_mu_sweep(None, mesh_size_dim=args.h, streams=args.streams, prefetch=False)     # 23 x 3 x 2 solves
        out['phase_b'] = {'rows': len(dfb), 'wall_s': time.perf_counter() - t,
                          'finite': bool(np.isfinite(dfb[['avg_conc_sulc', 'avg_conc_rect', 'flux_sulc_y0', 'flux_rect_bottom', 'CR']].to_numpy()).all()),
                          'CR_range': [float(dfb['CR'].min()), float(dfb['CR'].max())],
                          'flux_error_pct_range': [float(dfb['flux_error_pct'].min()), float(dfb['flux_error_pct'].max())]}
        if len(dfb) != 69 or not out['phase_b']['finite']:
            bad.append('phase_b')
        t = time.perf_counter()
        dfg = studies.run_geometry_analysis(None, mesh_size_dim=args.h, prefetch=False)                         # 23 x 3 solves
        cols = ['Mu_Eff_Simulation', 'Mu_Eff_Opening', 'Total_Mass', 'Mouth_Flux_Total']
        out['geometry_analysis'] = {'rows': len(dfg), 'wall_s': time.perf_counter() - t,
                                    'finite': bool(np.isfinite(dfg[cols].to_numpy(dtype=float)).all()),
                                    'ratio_sim_range': [float(dfg['Ratio_Sim'].min()), float(dfg['Ratio_Sim'].max())]}
        if len(dfg) != 69 or not out['geometry_analysis']['finite']:
            bad.append('geometry_analysis')
        t = time.perf_counter()
        dfu = studies.run_geometry_study(None, mesh_size_dim=args.h, prefetch=False)                            # Stokes + adv-diff, mu = 0: 72 rows
        num = dfu.select_dtypes(include=[np.number])
        out['no_uptake_geometry_study'] = {'rows': len(dfu), 'wall_s': time.perf_counter() - t,
                                           'finite_fraction': float(np.isfinite(num.to_numpy(dtype=float)).mean())}
        if len(dfu) != 72:
            bad.append('no_uptake_geometry_study')
    out['warnings'] = sorted(set(str(w.message)[:160] for w in wlist))[:10]
    out['failed'] = bad
    print(json.dumps(out))
    sys.exit(1 if bad else 0)


if __name__ == '__main__':
    main()
