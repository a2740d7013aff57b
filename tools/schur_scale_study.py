"""Experiment: MINRES iterations of the Taylor-Hood solve against a relative scaling of the Schur block of the
block-diagonal preconditioner (env SFEM_SCHUR_SCALE, read once per process -> one process per value).
    python tools/schur_scale_study.py --refine 1"""
import argparse, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))

def one(refine):
    import torch, bench
    from sulcusfem.device import Context
    ctx = Context.get()
    case = bench.Case(ctx, 0.02, refine, 1.0)
    st = case.stokes
    st.assemble(bc_mode=1)
    st.solve(rtol=1e-12)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); st.solve(rtol=1e-12); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"scale": float(os.environ.get('SFEM_SCHUR_SCALE', 1.0)), "refine": refine, "ms": e0.elapsed_time(e1), **st.last_info}))

if __name__ == '__main__':
    ap = argparse.ArgumentParser(); ap.add_argument('--refine', type=int, default=1); ap.add_argument('--one', action='store_true')
    ap.add_argument('--scales', default='0.25,0.5,0.7,1.0,1.4,2.0,4.0')
    a = ap.parse_args()
    if a.one:
        one(a.refine)
    else:
        for s in a.scales.split(','):
            env = dict(os.environ, SFEM_SCHUR_SCALE=s)
            r = subprocess.run([sys.executable, __file__, '--one', '--refine', str(a.refine)], env=env, capture_output=True, text=True)
            print([l for l in r.stdout.splitlines() if l.startswith('{')][-1] if r.returncode == 0 else r.stderr[-500:], flush=True)
