#!/usr/bin/env python
"""Generate tests/golden/solution_golden.json: oracle (CPU, sparse LU) results of the BASELINE configs on
small synthetic meshes.  Run here (no GPU needed):  python tests/golden/make_solution_golden.py

The GPU tests (tests/test_gpu_golden.py) run the same cases through the reference-facing front end
(sulcusfem.simulation.run_simulation / sulcusfem.solvers) and compare with these committed numbers,
so parity is checked against fixed fixtures and not only against a live oracle run.

Cases (mesh size 0.1 unless stated; w = 0.5, d = 1.0):
  c0_noadv_sulcus        BASELINE configs[0]: no-adv sulcus, constant Robin mu = 1
  c1_advdiff_sulcus      BASELINE configs[1]: Stokes + advection-diffusion at the default Pe = 40, mu = 1
  c2_step_rectangle      BASELINE configs[2]: rectangle surrogate with StepUptakeOpen mu(x), Pe = 10
  c2b_noadv_variable_mu  pure_diffusion_solver_variable_mu with a mu(x) that changes sign (clamp at quadrature points)
  c3_no_uptake           no-uptake mode (mu = 0), Pe = 1
"""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))

import numpy as np  # noqa: E402

SAMPLE = 16


def fingerprint(v):
    v = np.asarray(v, dtype=np.float64)
    idx = (np.arange(SAMPLE) * max(1, len(v) // SAMPLE)) % len(v)
    return {"n": int(len(v)), "l2": float(np.linalg.norm(v)), "sum": float(v.sum()), "min": float(v.min()),
            "max": float(v.max()), "idx": idx.tolist(), "vals": v[idx].tolist()}


def flat(d, prefix=''):
    out = {}
    for k, v in d.items():
        if k.startswith('_'):
            continue
        if isinstance(v, dict):
            out.update(flat(v, prefix + k + '.'))
        elif isinstance(v, (int, float, np.floating)):
            out[prefix + k] = float(v)
    return out


def mesh_hash(mesh):
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(mesh.coords).tobytes())
    h.update(np.ascontiguousarray(mesh.cells.astype(np.int64)).tobytes())
    return h.hexdigest()


def build(domain, h=0.1):
    from sulcusfem import hostmesh as hm
    from sulcusfem.unstructured import mesh_domain
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, h, domain)
    mk = hm.build_markers(mesh, 10.0, 1.0, 4.75, 5.25, domain)
    return mesh, {k: v.values for k, v in mk.items()}


def main():
    from oracle import cpu_oracle as co
    out = {"_about": "oracle results; regenerate with tests/golden/make_solution_golden.py", "cases": {}}
    L, H, w, d = 10.0, 1.0, 0.5, 1.0

    # ---- c0
    mesh, mk = build('sulcus')
    om = co.Mesh(mesh.coords, mesh.cells)
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], 1.0, mu=1.0)
    fl = co.flux_metrics(om, mk, 'sulcus', 1.0, c, mu=1.0)
    ms = co.mass_metrics(om, c, 'sulcus', mk['domain_markers'])
    me = co.mu_eff_metrics(fl, L, d, w, 1.0)
    out["cases"]["c0_noadv_sulcus"] = {"mesh_sha256": mesh_hash(mesh), "c": fingerprint(c), "flux": flat(fl), "mass": flat(ms),
                                       "mu_eff": flat(me)}
    # ---- c1
    ux, uy, p, _, _ = co.solve_stokes(om, mk['bc_markers'], H)
    D = 1.0 / 40.0
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], D, mu=1.0, ux=ux, uy=uy)
    fl = co.flux_metrics(om, mk, 'sulcus', D, c, ux, uy, mu=1.0)
    ms = co.mass_metrics(om, c, 'sulcus', mk['domain_markers'])
    me = co.mu_eff_metrics(fl, L, d, w, 1.0)
    out["cases"]["c1_advdiff_sulcus"] = {"mesh_sha256": mesh_hash(mesh), "ux": fingerprint(ux), "uy": fingerprint(uy),
                                         "p": fingerprint(p), "c": fingerprint(c), "flux": flat(fl), "mass": flat(ms),
                                         "mu_eff": flat(me)}
    # ---- c3 no-uptake, Pe = 1
    c, _, _ = co.solve_concentration(om, mk['bc_markers'], 1.0, mu=0.0, ux=ux, uy=uy)
    fl = co.flux_metrics(om, mk, 'sulcus', 1.0, c, ux, uy, mu=0.0)
    out["cases"]["c3_no_uptake"] = {"mesh_sha256": mesh_hash(mesh), "c": fingerprint(c), "flux": flat(fl)}

    # ---- c2 rectangle + step mu
    meshr, mkr = build('rectangular')
    omr = co.Mesh(meshr.coords, meshr.cells)
    uxr, uyr, pr, _, _ = co.solve_stokes(omr, mkr['bc_markers'], H)
    step = co.StepUptakeOpen(1.0, 1.7700044654465237, 4.75, 5.25, L_c=0.05, Gamma=5.0)
    mun = co.interpolate_p2(omr, step)
    D = 0.1
    c, _, _ = co.solve_concentration(omr, mkr['bc_markers'], D, mu_nodal=mun, ux=uxr, uy=uyr)
    fl = co.flux_metrics(omr, mkr, 'rectangular', D, c, uxr, uyr, mu_nodal=mun)
    ms = co.mass_metrics(omr, c, 'rectangular')
    out["cases"]["c2_step_rectangle"] = {"mesh_sha256": mesh_hash(meshr), "ux": fingerprint(uxr), "p": fingerprint(pr),
                                         "c": fingerprint(c), "flux": flat(fl), "mass": flat(ms),
                                         "step": {"mu_base": 1.0, "mu_eff_target": 1.7700044654465237, "xL": 4.75, "xR": 5.25,
                                                  "L_c": 0.05, "Gamma": 5.0}, "D": D}
    # ---- c2b sign-changing mu, clamp at quadrature points, no advection
    X = omr.p2_dof_coords()
    mun2 = 0.5 + np.cos(3.0 * X[:, 0])
    c, _, _ = co.solve_concentration(omr, mkr['bc_markers'], 1.0, mu_nodal=mun2, clamp_mu=True)
    out["cases"]["c2b_noadv_variable_mu"] = {"mesh_sha256": mesh_hash(meshr), "c": fingerprint(c)}

    path = os.path.join(ROOT, 'tests', 'golden', 'solution_golden.json')
    with open(path, 'w') as f:
        json.dump(out, f, indent=1)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == '__main__':
    main()
