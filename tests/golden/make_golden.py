"""Generate tests/golden/reference_vectors.json by importing the reference's OWN python modules.

Run once in the build container (``python tests/golden/make_golden.py``); the GPU box has no
/root/reference, so only the committed JSON travels.  dolfin / meshio / matplotlib are absent here,
so they are replaced by inert stubs that provide just the names the reference modules touch at
import time (``UserExpression``, ``near``, ``DOLFIN_EPS`` ...).  Everything written to the JSON is
computed by reference code (parameters.py, mesh.py, analysis.py) or read from the reference's
checked-in result CSVs -- nothing from this repo is involved.
"""
import csv
import json
import os
import sys
import types

import numpy as np

REF = '/root/reference'
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'reference_vectors.json')


def _install_stubs():
    d = types.ModuleType('dolfin')

    class UserExpression:                      # dolfin.UserExpression: only __init__(**kwargs) is used
        def __init__(self, **kwargs):
            self._kwargs = kwargs
    d.UserExpression = UserExpression
    d.DOLFIN_EPS = 3.0e-16
    d.near = lambda x, x0, eps=3.0e-16: (x0 - eps) <= x <= (x0 + eps)
    for name in ('SubDomain', 'Expression', 'Constant', 'Function', 'MeshFunction', 'Mesh', 'Measure',
                 'FacetNormal', 'FunctionSpace', 'File', 'Point'):
        setattr(d, name, type(name, (), {}))
    d.__all__ = [k for k in d.__dict__ if not k.startswith('_')]
    sys.modules['dolfin'] = d
    for name in ('meshio', 'matplotlib', 'matplotlib.pyplot', 'ufl'):
        sys.modules[name] = types.ModuleType(name)
    sys.modules['matplotlib'].pyplot = sys.modules['matplotlib.pyplot']


def main():
    _install_stubs()
    sys.path.insert(0, REF)
    import parameters as P
    import mesh as M
    import analysis as A

    out = {}

    # 1. StepUptakeOpen.eval (parameters.py:24-84)
    cases = [dict(mu_base=1.0, mu_eff_target=1.7700044654465237, sulcus_left_x=4.75, sulcus_right_x=5.25, L_c=0.05, Gamma=5.0),
             dict(mu_base=0.1, mu_eff_target=0.40360903095592376, sulcus_left_x=4.75, sulcus_right_x=5.25, L_c=None, Gamma=5.0),
             dict(mu_base=10.0, mu_eff_target=3.2, sulcus_left_x=4.0, sulcus_right_x=6.0, L_c=5.0, Gamma=8.0),
             dict(mu_base=2.0, mu_eff_target=0.5, sulcus_left_x=1.0, sulcus_right_x=1.5, L_c=0.0, Gamma=5.0)]
    step = []
    for kw in cases:
        f = P.StepUptakeOpen(degree=2, **kw)
        xs = np.unique(np.concatenate([np.linspace(f.xL - 0.3, f.xR + 0.3, 61),
                                       [f.xL, f.xR, f.xL + f.L_c, f.xR - f.L_c, 0.5 * (f.xL + f.xR)]]))
        vals = []
        for x in xs:
            v = np.zeros(1)
            f.eval(v, np.array([x, 0.0]))
            vals.append(float(v[0]))
        step.append({'kwargs': kw, 'L_c_effective': f.L_c, 'x': [float(x) for x in xs], 'mu': vals})
    out['step_uptake_open'] = step

    # 2. Parameters.nondim (parameters.py:200-226)
    nd = []
    for mode, kw in (('adv-diff', {}), ('no-adv', {}), ('no-uptake', {}),
                     ('adv-diff', dict(H_dim=2.0, L_dim=10.0, sulci_w_dim=0.25, sulci_h_dim=0.25, U_ref_dim=0.001, D_dim=0.0005)),
                     ('no-adv', dict(H_dim=0.5, mesh_size_dim=0.01))):
        p = P.Parameters(mode=mode, **kw)
        p.validate()
        p.nondim()
        nd.append({'mode': mode, 'kwargs': kw,
                   'values': {k: getattr(p, k) for k in ('L', 'H', 'sulci_h', 'sulci_w', 'mesh_size', 'D', 'mu', 'U_ref', 'Pe', 'Re', 'mu_dim')},
                   'mesh_generator_params': p.get_mesh_generator_params()})
    out['nondim'] = nd

    # 3. geometry catalogue (parameters.py:342-447)
    base = P.Parameters(mode='no-adv')
    out['geometry_variations'] = {k: {'w': v['sulci_w_dim'], 'h': v['sulci_h_dim'], 'is_small': bool(v['is_small'])}
                                  for k, v in P.create_geometry_variations(base).items()}
    out['geometry_variations_small'] = sorted(P.create_geometry_variations(base, include_small=True).keys())

    # 4. closed forms (analysis.py:948-985)
    cf = []
    for (w, h, mu) in ((0.5, 1.0, 1.0), (0.25, 0.25, 1.0), (0.5, 1.0, 0.1), (1.0, 2.0, 10.0), (0.05, 1.0, 0.5), (0.01, 0.01, 1.0)):
        prm = types.SimpleNamespace(L=10.0, sulci_h=h, sulci_w=w, mu=mu)
        r = {'params': prm}
        cf.append({'L': 10.0, 'w': w, 'h': h, 'mu': mu, 'mu_eff_arc': A.compute_mu_eff_arc(r), 'mu_eff_enh': A.compute_mu_eff_enh(r)})
    out['closed_forms'] = cf

    # 5. sulcus floor samples + marker predicates (mesh.py:139-155, 196-214)
    mg = M.MeshGenerator.__new__(M.MeshGenerator)
    mg._store_parameters(10.0, 1.0, 1.0, 0.5, 0.02, 1, 'sulcus')
    pts = []
    sec = mg._generate_sulcus_points()['points_section'].splitlines()
    for ln in sec:
        nums = ln[ln.index('{') + 1:ln.index('}')].split(',')
        pts.append([float(nums[0]), float(nums[1])])
    out['sulcus_points'] = pts
    mg._create_boundary_functions()
    samples = []
    e = 3.0e-16
    for x in (0.0, e, 2 * e, 1.0, 4.75 - 1e-9, 4.75, 4.75 + e, 4.75 + 1e-9, 5.0, 5.25 - 1e-9, 5.25, 5.25 + 1e-9, 10.0 - e, 10.0, 7.3):
        for y in (1.0, 1.0 - e, 0.5, 5e-16, 7e-16, 0.0, -1e-16, -5e-16, -1e-9, -0.5):
            for ob in (True, False):
                samples.append({'x': x, 'y': y, 'on_boundary': ob,
                                'inside': {k: bool(fn(np.array([x, y]), ob)) for k, fn in mg.boundary_functions.items()}})
    out['boundary_predicates'] = {'width': 10.0, 'height': 1.0, 'xL': mg.sulcus_left_x, 'xR': mg.sulcus_right_x, 'samples': samples}
    out['marker_ids'] = dict(M.MeshGenerator.MARKERS)
    out['mark_order'] = {'bc': ["left", "right", "top", "bottom"],
                         'bottom_segment': ["bottom_left", "bottom_right", "sulcus", "sulcus_opening"], 'y0': ["y0_line"]}

    # 6. rows of the reference's checked-in result CSVs (the only recorded outputs of past runs)
    def rows(path):
        with open(os.path.join(REF, path), newline='') as f:
            return list(csv.DictReader(f))
    pb = rows('No Advection - Phase B/no_adv_mu_sweep_results.csv')
    out['phaseB_reference_rows'] = [r for r in pb if r['geometry'] == 'reference']
    out['phaseB_rect_avg_conc'] = {mu: sorted({float(r['avg_conc_rect']) for r in pb if float(r['mu_factor']) == float(mu)})
                                   for mu in ('0.1', '0.5', '1.0')}
    ar = rows('No Advection - Phase A/Aspect Ratio Study Analysis/aspect_ratio_analysis_results.csv')
    key0 = list(ar[0].keys())[0]
    out['phaseA_aspect_rows'] = [r for r in ar if r[key0] in ('h_equals_2w_h1.0', 'h_equals_w_h0.25')]
    out['advdiff_validation_rows'] = rows('Advection-Diffusion/Results Data/advdiff_validation_step_pe_x_mu.csv')
    gc = rows('No Uptake Simulations/Geometry Comparison Analysis/geometry_comparison_results.csv')
    k0 = list(gc[0].keys())[0]
    out['no_uptake_rectangle_rows'] = [r for r in gc if 'rect' in str(r[k0]).lower() or 'rect' in str(list(r.values())[:3]).lower()]
    with open(OUT, 'w') as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print('wrote', OUT, os.path.getsize(OUT), 'bytes')


if __name__ == '__main__':
    main()
