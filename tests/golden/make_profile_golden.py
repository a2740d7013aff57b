"""Golden: concentration line-profile samples the reference checked in (no-uptake geometry study, dolfin on Gmsh meshes).

    python tests/golden/make_profile_golden.py  ->  tests/golden/profile_samples.json

Source: /root/reference/No Uptake Simulations/Geometry Comparison Analysis/Profiles/profiles_samples_{largest,square_small}.csv
(5 horizontal lines x 3 Peclet numbers, 400 samples per line; every 20th sample of the channel lines and every sample of
the short in-cavity line are kept).
"""
import csv
import json
import os

REF = '/root/reference/No Uptake Simulations/Geometry Comparison Analysis/Profiles'
out = {'columns': None, 'geometries': {}}
for g in ('largest', 'square_small'):
    with open(os.path.join(REF, f'profiles_samples_{g}.csv'), newline='') as f:
        rd = csv.DictReader(f)
        out['columns'] = rd.fieldnames
        lines = {}
        for r in rd:
            key = f"{r['Peclet']}|{r['LineName']}"
            L = lines.setdefault(key, {'peclet': float(r['Peclet']), 'name': r['LineName'], 'y': float(r['y']), 'n_valid': 0,
                                       'index': [], 'x': [], 'c': []})
            L['n_valid'] += 1
            i = int(r['Index'])
            if r['LineName'] == 'sulcus_mid' or i % 20 == 0:
                L['index'].append(i)
                L['x'].append(float(r['x']))
                L['c'].append(float(r['c']))
        out['geometries'][g] = list(lines.values())
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'profile_samples.json'), 'w') as f:
    json.dump(out, f)
print({g: (len(v), sum(len(l['x']) for l in v)) for g, v in out['geometries'].items()})
