"""Golden: header rows and case keys of the reference's checked-in study CSVs (run here, where /root/reference exists).

    python tests/golden/make_study_columns.py  ->  tests/golden/study_columns.json
"""
import csv
import json
import os

REF = '/root/reference'
FILES = {
    'no_adv_mu_sweep_results.csv': 'No Advection - Phase B/no_adv_mu_sweep_results.csv',
    'advdiff_validation_step_pe_x_mu.csv': 'Advection-Diffusion/Results Data/advdiff_validation_step_pe_x_mu.csv',
    'mu_parameter_sweep_results.csv': 'No Advection - Phase A/Mu Parameter Sweep Analysis/mu_parameter_sweep_results.csv',
    'aspect_ratio_analysis_results.csv': 'No Advection - Phase A/Aspect Ratio Study Analysis/aspect_ratio_analysis_results.csv',
    'geometry_comparison_results.csv': 'No Uptake Simulations/Geometry Comparison Analysis/geometry_comparison_results.csv',
    'mu_eff_analysis_results.csv': 'No Advection - Phase A/Mu_Eff Spatial Analysis Analysis/mu_eff_analysis_results.csv',
}
csv.field_size_limit(10 ** 9)
out = {}
for name, rel in FILES.items():
    with open(os.path.join(REF, rel), newline='') as f:
        rows = list(csv.reader(f))
    hdr = rows[0]
    entry = {'columns': hdr, 'n_rows': len(rows) - 1}
    if 'Config' in hdr:
        entry['configs'] = [r[hdr.index('Config')] for r in rows[1:]]
    if name.startswith('no_adv_mu_sweep'):
        entry['geometries'] = sorted(set(r[hdr.index('geometry')] for r in rows[1:]))
        entry['mu_factors'] = sorted(set(float(r[hdr.index('mu_factor')]) for r in rows[1:]))
    entry['rows'] = [dict(zip(hdr, r)) for r in rows[1:]]          # the reference's own results (dolfin on Gmsh meshes)
    if name.startswith('advdiff'):
        entry['cases'] = sorted(set((float(r[hdr.index('Pe')]), float(r[hdr.index('mu_factor')])) for r in rows[1:]))
    out[name] = entry
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'study_columns.json'), 'w') as f:
    json.dump(out, f, indent=1)
print({k: (len(v['columns']), v['n_rows']) for k, v in out.items()})
