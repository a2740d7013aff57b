#!/usr/bin/env python
"""bench.py -- throughput of the sulcus FEM hot path on B200 (see DESIGN.md "Measurement").

    python bench.py --gpus N --steps K --warmup W [--refine R] [--h H] [--impl reference]

One *step* = one full advection-diffusion sulcus run (BASELINE.json configs[1]): Taylor-Hood Stokes
assembly + MINRES solve, P2 advection-diffusion assembly (Pe = 40, Robin mu = 1) + FGMRES solve, and
all flux / mass functionals, on a synthetic sulcus mesh (w = 0.5, d = 1.0, h = 0.02, R uniform
refinements; default R = 2 -> 6.1 M DOFs per case, matrices larger than L2).  Metric: DOFs/s = (Taylor-Hood dofs + P2 dofs) solved per second, whole job.
N > 1: every rank solves its own independent sweep case (no data-path collective) -> weak scaling.

`value`    device-resident inputs, CUDA-event timed per step, L2 flushed between steps (512 MiB written, then
           512 MiB of another buffer read, so the cache holds clean unrelated lines).
`e2e`      the same run through the reference-facing API (sulcusfem.solvers / analysis) with host
           inputs and host results, wall clock around each call (H2D / D2H inside).
`roofline` dominant kernel family, from a per-launch CUDA-event profile of one extra step.
`cpu_baseline` the CPU oracle (numpy assembly + SuperLU) on a bounded sample of the workload.
`same_mesh` GPU and CPU oracle on the SAME unrefined mesh with the parity of every field (exit code 3 on a miss).
`mu_sweep` BASELINE configs[3]: the reference's 20-coefficient mu sweep through studies.run_mu_sweep -- batched Krylov
           loops against one solve per coefficient, rows compared, two coefficients against the oracle's LU.
`dd_strong` (N > 1) BASELINE config 5: the same step, one case row-partitioned over all ranks.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))

import numpy as np  # noqa: E402

W_SULCUS, D_SULCUS, L_CH, H_CH = 0.5, 1.0, 10.0, 1.0
PE, MU = 40.0, 1.0
RTOL, STOKES_RTOL = 1e-13, 1e-12          # = sulcusfem.solvers.RTOL / STOKES_RTOL (the API path's defaults; checked in main)
CAT_NAMES = ['spmv', 'spmv_dot', 'cheb_step', 'resid_d0', 'elem', 'gather', 'vec', 'other', 'spmv_staged', 'halo']


def build_mesh(h, refine):
    from sulcusfem import hostmesh as hm
    from sulcusfem.unstructured import mesh_domain
    mesh = mesh_domain(L_CH, H_CH, W_SULCUS, D_SULCUS, h, 'sulcus')
    mesh = hm.refine_n(mesh, refine)
    mk = hm.build_markers(mesh, L_CH, H_CH, L_CH / 2 - W_SULCUS / 2, L_CH / 2 + W_SULCUS / 2, 'sulcus')
    res = {'mesh': mesh, 'mesh_info': {}}
    res.update(mk)
    return res


# ------------------------------------------------------------------------------------ CPU arm
def oracle_step(h, mu=MU, mr=None, fields=False):
    """One step of the same workload on the CPU oracle (numpy assembly + SuperLU), timed."""
    from oracle import cpu_oracle as co
    mr = build_mesh(h, 0) if mr is None else mr
    mesh = mr['mesh']
    om = co.Mesh(mesh.coords, mesh.cells)
    bm = mr['bc_markers'].values
    mk = {k: mr[k].values for k in ('bc_markers', 'bottom_segment_markers', 'y0_markers', 'domain_markers')}
    t0 = time.perf_counter()
    ux, uy, p, _, _ = co.solve_stokes(om, bm, H_CH)
    c, _, _ = co.solve_concentration(om, bm, 1.0 / PE, mu=mu, ux=ux, uy=uy)
    fl = co.flux_metrics(om, mk, 'sulcus', 1.0 / PE, c, ux, uy, mu=mu)
    co.mass_metrics(om, c, 'sulcus', mk['domain_markers'])
    dt = time.perf_counter() - t0
    ndof = 2 * om.n_p2 + om.nv + om.n_p2
    if fields:
        me = co.mu_eff_metrics(fl, L_CH, D_SULCUS, W_SULCUS, mu)
        return dt, ndof, {'ux': ux, 'uy': uy, 'p': p, 'c': c, 'flux': fl, 'mu_eff': me}
    return dt, ndof, fl


def run_reference(args, rank):
    """The CPU arm: K timed (+ W warm-up) steps of the SAME step (Stokes + adv-diff + functionals) on the oracle
    port, on a BOUNDED mesh of the same workload family (h = --cpu-h, no refinement) -- a sparse LU of the GPU arm's
    6 M-dof mesh does not end within minutes.  `config` states the mesh this arm really solved; the GPU arm's
    `same_mesh` records time the GPU on exactly this mesh (and check parity there)."""
    if rank != 0:
        return
    hs = args.cpu_h
    mr = build_mesh(hs, 0)
    for _ in range(args.warmup):
        oracle_step(hs, mr=mr)
    times, ndof = [], 0
    for _ in range(args.steps):
        dt, ndof, _ = oracle_step(hs, mr=mr)
        times.append(dt)
    val = ndof * len(times) / sum(times)
    sample = (f"same adv-diff sulcus step (Stokes + adv-diff + functionals) on the h={hs} unrefined mesh, {ndof} dofs, "
              f"numpy assembly + scipy SuperLU, 1 thread")
    cfg = workload_config(args, h=hs, refine=0, dofs=ndof)
    cfg["l2"] = "n/a (CPU)"
    cfg["parallelism"] = "1 host thread (serial reference path)"
    cfg["bounded_sample_of"] = {"h": args.h, "refine": args.refine,
                                "why": "per-DOF throughput across UNEQUAL meshes when set against the GPU arm's headline value; "
                                       "the same-mesh ratio is the GPU arm's same_mesh[] record for this h"}
    line = {
        "impl": "reference", "metric": "fem_dofs_per_s", "value": val, "unit": "DOFs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * sum(times) / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": "DOFs/s", "cores": 1, "kind": "port", "sample": sample,
                         "host_cores": os.cpu_count()},
        "e2e": {"value": val, "unit": "DOFs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "dolfin/PETSc is not installable in this image; the CPU arm is the oracle port of the reference path",
    }
    print(json.dumps(line))


def workload_config(args, h=None, refine=None, dofs=None):
    h = args.h if h is None else h
    refine = args.refine if refine is None else refine
    return {"workload": f"adv-diff sulcus (BASELINE configs[1]): Stokes TH + adv-diff P2 + functionals, "
                        f"w={W_SULCUS} d={D_SULCUS} Pe={PE:g} mu={MU:g}, synthetic Delaunay mesh h={h} "
                        f"+ {refine} uniform refinements",
            "h": h, "refine": refine, "dofs": dofs, "l2": "flushed between steps (512 MiB memset + 512 MiB read of a second buffer)",
            "krylov_rtol": {"scalar_true_residual": RTOL, "stokes_preconditioned_residual": STOKES_RTOL,
                            "basis": "fields within 1e-10 relative L2 of the LU solution with >= 2.5x (scalar) / 100x (Stokes) margin, profiles/r01_tolerance_study.md"}, "parallelism": f"case-sharded x{args.gpus} (no collectives)"}


# ------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(',')]
            try:
                sm.append(float(f[0])); mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# ------------------------------------------------------------------------------------ GPU arm
class Case:
    """One sulcus case on this rank's GPU: mesh, device problems, the device-resident step and the same step through
    the reference-facing API (host in, host out)."""

    def __init__(self, ctx, h, refine, mu):
        import torch
        from sulcusfem import dofmap as dm
        from sulcusfem.device import FunctionalPlan, ScalarProblem, StokesProblem
        from sulcusfem.hierarchy import build_hierarchy
        self.ctx, self.h, self.refine, self.mu, self.D = ctx, h, refine, mu, 1.0 / PE
        self.setup = {}
        t0 = time.perf_counter()
        self.mr = mr = build_mesh(h, refine)
        self.mesh, self.bm = mr['mesh'], mr['bc_markers'].values
        t1 = time.perf_counter()
        self.hier = build_hierarchy(self.mesh)
        t2 = time.perf_counter()
        self.stokes = StokesProblem(self.mesh, self.bm, hierarchy=self.hier, ctx=ctx)
        t3 = time.perf_counter()
        self.scalar = ScalarProblem(self.mesh, self.bm, hierarchy=self.hier, ctx=ctx)
        t4 = time.perf_counter()
        self.plan = FunctionalPlan(self.mesh, mr, 'sulcus', ctx=ctx)
        # the reference-facing API (e2e leg) finds the same device objects through the per-mesh cache
        self.mesh._sfem_cache = {'hierarchy': self.hier, 'stokes': self.stokes, ('scalar', 4): self.scalar,
                                 ('functionals', 'sulcus'): self.plan}
        X = dm.p2_dof_coordinates(self.mesh)
        d1 = dm.dirichlet_dofs_p2(self.mesh, self.bm, 1)
        self.stokes.set_bcs({1: (4.0 * X[d1, 1] * (H_CH - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})
        self.stokes.set_channel_flow_guess(L_CH, H_CH)       # what solvers.stokes_solver does: analytic channel flow as x0
        self.stokes._inflow_H = float(H_CH)
        torch.cuda.synchronize()
        t5 = time.perf_counter()
        self.setup = {"mesh_markers_s": t1 - t0, "hierarchy_s": t2 - t1, "stokes_problem_s": t3 - t2,
                      "scalar_problem_s": t4 - t3, "functionals_bcs_s": t5 - t4, "total_s": t5 - t0}
        self.ndof = self.stokes.n + self.scalar.n
        self.info = {}
        self.api_ms = {}
        self._spaces = None

    def step(self):
        st, sc, mu, D = self.stokes, self.scalar, self.mu, self.D
        st.assemble(bc_mode=1)
        ux, uy, p = st.solve(rtol=STOKES_RTOL)
        sc.assemble(D, ux, uy, mu_const=mu, bc_values={1: 1.0, 2: 0.0})
        c = sc.solve('fgmres', rtol=RTOL)
        F, M = self.plan.evaluate(c, ux, uy, D=D, mu_const=mu)
        self.info['stokes'], self.info['advdiff'] = dict(st.last_info), dict(sc.last_info)
        return F, M

    def api_step(self):
        """Stokes -> adv-diff -> functionals through sulcusfem.solvers / analysis; returns (u, p, c, flux, mass)."""
        import contextlib
        import io
        from sulcusfem import solvers, analysis
        from sulcusfem.fem import Constant, FunctionSpace, MixedElement, VectorFunctionSpace
        if self._spaces is None:
            V = VectorFunctionSpace(self.mesh, "P", 2)
            Q = FunctionSpace(self.mesh, "P", 1)
            self._spaces = (FunctionSpace(self.mesh, MixedElement([V.ufl_element(), Q.ufl_element()])),
                            FunctionSpace(self.mesh, "CG", 2))
        Wsp, Csp = self._spaces
        t = [time.perf_counter()]
        with contextlib.redirect_stdout(io.StringIO()):
            u, p = solvers.stokes_solver(self.mr, Wsp, L_CH, H_CH, 'sulcus')
            t.append(time.perf_counter())
            u._dev = None                                  # velocity re-enters from its host array (H2D)
            c = solvers.advdiff_solver(self.mr, u, Csp, Constant(self.D), Constant(self.mu), 'sulcus')
            t.append(time.perf_counter())
            c._dev = None
            fm = analysis.compute_flux_metrics(c, u, self.mr, 'sulcus', {}, self.D, self.mu)
            mm = analysis.compute_mass_metrics(c, {}, 'sulcus')
            t.append(time.perf_counter())
        for name, a, b in (('stokes_solver', 0, 1), ('advdiff_solver', 1, 2), ('functionals', 2, 3)):
            self.api_ms[name] = 1e3 * (t[b] - t[a])
        return u, p, c, fm, mm

    def bytes_per_step(self):
        n2, nv = self.stokes.n2, self.stokes.nv
        h2d = 8 * (2 * n2) + 8 * n2          # velocity components + concentration re-uploaded for the functionals
        d2h = 8 * (2 * n2 + nv) + 8 * n2 + 8 * (self.plan.ngroups * 8 + self.plan.nmarkers * 2)
        return h2d, d2h


def _rel(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / np.linalg.norm(b))


def same_mesh_leg(ctx, h, flush_l2, reps=5):
    """GPU and CPU oracle on the SAME unrefined mesh (h = 0.02 is the reference's default size, BASELINE configs[1]
    as the reference itself runs it): device-timed and API-timed GPU step, the CPU step, their ratio, and the parity
    of the API results against the oracle's LU fields (bars: fields 1e-10 relative L2, mu_eff 1e-8)."""
    import torch
    from types import SimpleNamespace
    from sulcusfem import analysis
    case = Case(ctx, h, 0, MU)
    for _ in range(3):
        case.step()
    ts = []
    for _ in range(reps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); case.step(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    case.api_step()
    te = []
    for _ in range(reps):
        flush_l2()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        u, p, c, fm, mm = case.api_step()
        torch.cuda.synchronize()
        te.append(1e3 * (time.perf_counter() - t0))
    res = {'params': SimpleNamespace(L=L_CH, sulci_h=D_SULCUS, sulci_w=W_SULCUS, mu=MU), 'c': c, 'flux_metrics': fm}
    me = analysis.compute_mu_eff_metrics(res)
    cpu_s, ndof, o = oracle_step(h, mr=case.mr, fields=True)
    n2 = case.stokes.n2
    par = {"rel_l2_u": _rel(u.values, np.concatenate([o['ux'], o['uy']])), "rel_l2_p": _rel(p.values, o['p']),
           "rel_l2_c": _rel(c.values, o['c']),
           "mu_eff_sim_rel": abs(me['mu_eff_sim'] - o['mu_eff']['mu_eff_sim']) / abs(o['mu_eff']['mu_eff_sim']),
           "mu_eff_open_rel": abs(me['mu_eff_open'] - o['mu_eff']['mu_eff_open']) / abs(o['mu_eff']['mu_eff_open']),
           "uptake_rel": abs(fm['uptake_flux'] - o['flux']['uptake_flux']) / abs(o['flux']['uptake_flux']),
           "bars": {"fields": 1e-10, "mu_eff": 1e-8, "uptake": 1e-9}}
    par["ok"] = bool(par["rel_l2_u"] <= 1e-10 and par["rel_l2_c"] <= 1e-10 and par["mu_eff_sim_rel"] <= 1e-8
                     and par["mu_eff_open_rel"] <= 1e-8 and par["uptake_rel"] <= 1e-9 and par["rel_l2_p"] <= 1e-9)
    gpu_ms, e2e_ms = float(np.median(ts)), float(np.median(te))
    rec = {"h": h, "refine": 0, "dofs": int(ndof), "gpu_ms": gpu_ms, "gpu_e2e_ms": e2e_ms, "cpu_ms": 1e3 * cpu_s,
           "ratio": 1e3 * cpu_s / e2e_ms, "ratio_device_resident": 1e3 * cpu_s / gpu_ms,
           "gpu_dofs_per_s_e2e": ndof / (e2e_ms / 1e3), "cpu_dofs_per_s": ndof / cpu_s,
           "iterations": {"stokes_minres": case.info['stokes']['iterations'], "advdiff_fgmres": case.info['advdiff']['iterations']},
           "setup_s": case.setup, "parity": par,
           "note": "GPU = this repo's API path (host in / host out) and device-resident step; CPU = oracle port "
                   "(numpy assembly + SuperLU + one extended-precision refinement), 1 thread; same mesh, same parameters"}
    del case
    return rec


def sweep_leg(with_oracle, h=0.02):
    """BASELINE configs[3] / SURVEY 8(e) on the driver's clock: the reference's Phase A mu sweep
    (no_advection_analysis_A.py:1257-1347: 20 Robin coefficients on the 0.25 x 0.25 mm sulcus, mesh size 0.02) through
    ``studies.run_mu_sweep`` -- batched Krylov loops (sfem_krylov_cg_batch, the default) against one CG solve per
    coefficient, wall clock around the driver call on a cached geometry, rows compared; with the CPU legs enabled two
    coefficients are also solved by the oracle's sparse LU on the same mesh (field parity bar 1e-10, and the CPU time
    per solve the reference's serial loop would need)."""
    import contextlib
    import io
    import torch
    from sulcusfem import simulation, solvers, studies
    from sulcusfem.fem import FunctionSpace

    def timed(fn):
        torch.cuda.synchronize()
        t = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, time.perf_counter() - t
    many = {'dense': [float(v) for v in np.geomspace(0.1, 150.0, 100)]}
    df0, t_first = timed(lambda: studies.run_mu_sweep(None, mesh_size_dim=h))
    rec = {"h": h, "cases_reference_sweep": len(df0), "wall_s_first_call_with_geometry_build": t_first}
    for name, kw in (("batched", dict(batch=True)), ("per_case", dict(batch=False))):
        studies.run_mu_sweep(None, mesh_size_dim=h, **kw)
        ts20, ts100 = [], []
        for _ in range(3):
            df20, t = timed(lambda: studies.run_mu_sweep(None, mesh_size_dim=h, **kw))
            ts20.append(t)
            df100, t = timed(lambda: studies.run_mu_sweep(None, regimes=many, mesh_size_dim=h, **kw))
            ts100.append(t)
        rec[name] = {"solves_per_s_20_cases": len(df20) / float(np.median(ts20)),
                     "solves_per_s_100_cases": len(df100) / float(np.median(ts100))}
        rec[name + "_rows"] = df20
    a, b = rec.pop("batched_rows"), rec.pop("per_case_rows")
    cols = ('Mu_Eff_Simulation', 'Mu_Eff_Opening', 'Total_Mass', 'Mouth_Flux_Total')
    rec["rows_batched_vs_per_case_max_rel"] = float(max(np.max(np.abs(a[c] - b[c]) / np.abs(b[c])) for c in cols))
    rec["speedup_batched_over_per_case_100_cases"] = rec["batched"]["solves_per_s_100_cases"] / rec["per_case"]["solves_per_s_100_cases"]
    ok = rec["rows_batched_vs_per_case_max_rel"] <= 1e-9
    if with_oracle:
        from oracle import cpu_oracle as co
        p = studies.Parameters(mode='no-adv', mesh_size_dim=h)
        p.sulci_w_dim = p.sulci_h_dim = 0.25
        p.validate()
        p.nondim()
        with contextlib.redirect_stdout(io.StringIO()):
            mr = simulation._simulation_generate_mesh(p, 'sulcus')
            mus = [p.mu * f for f in (0.1, 1.0, 12.5, 150.0)]
            fs = solvers.pure_diffusion_solver_batch(mr, FunctionSpace(mr['mesh'], "CG", 2), p.D, mus)
        om = co.Mesh(mr['mesh'].coords, mr['mesh'].cells)
        errs, cpu_s = [], []
        for k in (1, 3):
            t0 = time.perf_counter()
            ref, _, _ = co.solve_concentration(om, mr['bc_markers'].values, p.D, mu=mus[k])
            cpu_s.append(time.perf_counter() - t0)
            errs.append(_rel(fs[k].values, ref))
        rec["parity_vs_oracle_lu"] = {"mu_factors": [1.0, 150.0], "rel_l2_c": errs, "bar": 1e-10, "p2_dofs": int(om.n_p2),
                                      "batch_iterations": fs[0].solver_info['iterations']}
        rec["cpu_oracle_s_per_solve"] = float(np.mean(cpu_s))
        rec["ratio_batched_over_cpu_oracle"] = rec["batched"]["solves_per_s_100_cases"] * rec["cpu_oracle_s_per_solve"]
        ok = ok and max(errs) <= 1e-10
    rec["ok"] = bool(ok)
    return rec


def run_dd_strong(args, case, rank, world, dist, flush_l2):
    """BASELINE config 5 on the driver's clock: ONE case (the bench step: Stokes + adv-diff + functionals, mu = 1)
    row-partitioned over all ranks -- peer-memory halo exchange + in-kernel all-reduces (csrc/sfem_dist.cu), NCCL only
    for the all-gather of the fields between the replicated stages -- against the same step on one GPU.
    Returns one record per refinement level: strong-scaling efficiency t_1 / (N t_N) and the relative L2 distance of
    the distributed fields from the single-GPU ones."""
    import torch
    from sulcusfem.dist import DistScalarProblem, DistStokesProblem, DistWorld, num_distributed_levels, plan_partitions
    ctx = case.ctx
    out = []
    refines = [args.refine] + [int(r) for r in str(args.dd_refine).split(',') if r.strip() != '' and int(r) != args.refine]
    for refine in refines:
        cs = case if refine == args.refine else Case(ctx, args.h, refine, MU)
        mu_saved, cs.mu = cs.mu, MU                       # every rank: the SAME case
        st, sc, plan, D = cs.stokes, cs.scalar, cs.plan, cs.D
        rb = args.dd_replicate_below

        def timed(fn, reps):
            ts = []
            for _ in range(reps):
                flush_l2()
                dist.barrier()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); fn(); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            t = torch.tensor([float(np.median(ts))], dtype=torch.float64, device=ctx.device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())

        # single-GPU time of the same case (every rank runs it; no communicator is active yet)
        reps = max(3, min(args.steps, 5))
        for _ in range(2):
            cs.step()
        t1 = timed(cs.step, reps)
        info1 = {k: dict(v) for k, v in cs.info.items()}
        ref = [t.clone() for t in (st.x[:st.n2], st.x[st.n2:2 * st.n2], sc.x)]

        def asm_only():                                   # assembly share of the single-GPU step (for the solve-only ratio)
            st.assemble(bc_mode=1)
            sc.assemble(D, ref[0], ref[1], mu_const=MU, bc_values={1: 1.0, 2: 0.0})
        t1_asm = timed(asm_only, 3)
        t0 = time.perf_counter()
        nd_v = num_distributed_levels(st.vel, rb)
        nd_c = num_distributed_levels(sc, rb)
        vec_cap = max(2 * st.vel.levels[nd_v].n, sc.levels[nd_c].n)
        wd = DistWorld(ctx, rank, world, vec_cap)
        ds = DistStokesProblem(st, wd, replicate_below=rb)
        dc = DistScalarProblem(sc, rank, world, world=wd, nb=1, plan=plan_partitions(sc, world, rb))
        wd.commit()
        ds.finalize()
        dc.finalize()
        torch.cuda.synchronize()
        t_plan = time.perf_counter() - t0
        marks = {}

        def dd_step(ev=None):
            def mark(name):
                if ev is not None:
                    e = torch.cuda.Event(enable_timing=True)
                    e.record()
                    ev.append((name, e))
            mark('start')
            ds.assemble_local()
            mark('stokes_assemble_local')
            ds.solve(rtol=STOKES_RTOL)
            mark('stokes_minres')
            ux, uy, p = ds.gather()
            mark('allgather_u')
            dc.assemble_local(D, ux, uy, mu_const=MU, bc_values={1: 1.0, 2: 0.0})
            mark('advdiff_assemble_local')
            dc.solve('fgmres', rtol=RTOL)
            mark('advdiff_fgmres')
            c = dc.gather()
            F, M = plan.evaluate(c, ux, uy, D=D, mu_const=MU)
            mark('allgather_c+functionals')
            return ux, uy, p, c, F

        # the distributed assembly must produce exactly the rows of the replicated (global) assembly
        st.assemble(bc_mode=1)
        ds.refresh()
        keep = [t.clone() for t in (ds.vel.A[0].csr.vals, ds.B.csr.vals, ds.BT.csr.vals, ds.Mp.csr.vals, ds.b[:ds.n_own])]
        ds.assemble_local()
        now = (ds.vel.A[0].csr.vals, ds.B.csr.vals, ds.BT.csr.vals, ds.Mp.csr.vals, ds.b[:ds.n_own])
        # operators: bit for bit (same element arithmetic, same gather order); right-hand side: the lifting SpMV sums a
        # row in the order of the engine's chunking, which differs between the global and the local mirror -> rounding level
        local_diff = {nm: float((a - b).abs().max()) for nm, a, b in zip(('K', 'B', 'BT', 'Mp', 'rhs'), keep, now)}
        nrm = float(keep[4].norm())                       # zero on ranks whose slab carries no inhomogeneous Dirichlet data
        local_diff['rhs_rel'] = float((keep[4] - now[4]).norm()) / nrm if nrm > 0 else float((keep[4] - now[4]).norm())
        local_equals = bool(all(torch.equal(a, b) for a, b in zip(keep[:4], now[:4])) and local_diff['rhs_rel'] < 1e-13)
        eq = torch.tensor([1.0 if local_equals else 0.0], dtype=torch.float64, device=ctx.device)
        dist.all_reduce(eq, op=dist.ReduceOp.MIN)
        local_equals = bool(eq.item() > 0.5)
        del keep
        for _ in range(2):
            dd_step()
        tN = timed(dd_step, reps)
        ev = []
        flush_l2(); dist.barrier(); torch.cuda.synchronize()
        ux, uy, p, c, F = dd_step(ev)
        torch.cuda.synchronize()
        phases = {ev[i][0]: ev[i - 1][1].elapsed_time(ev[i][1]) for i in range(1, len(ev))}

        # one un-graphed dd step with a CUDA event pair around every launch (all ranks run it; rank 0 reports): where
        # the distributed step spends its kernel time, and what one peer-memory exchange costs
        import ctypes as C
        cap = 400000
        lib = ctx.lib
        dist.barrier()
        lib.sfem_profile_start(cap)
        dd_step()
        cats, byts, mss = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_float * cap)()
        nrec = lib.sfem_profile_stop(cap, cats, byts, mss)
        cats, mss = np.array(cats[:nrec]), np.array(mss[:nrec], dtype=np.float64)
        prof = {name: {"launches": int((cats == k).sum()), "ms": float(mss[cats == k].sum())}
                for k, name in enumerate(CAT_NAMES) if (cats == k).any()}
        if 'halo' in prof and prof['halo']['launches']:
            prof['halo']['note'] = "peer-memory halo exchanges and vector all-reduces (send + wait + unpack): un-graphed"
            prof['halo']['avg_us'] = 1e3 * prof['halo']['ms'] / prof['halo']['launches']

        def rel(a, b):
            return float(((a - b).norm() / b.norm()).item())
        lu = ds.vel.parts[0]
        rec = {"refine": refine, "dofs": int(cs.ndof), "n_gpus": world, "ms_1gpu": t1, "ms_ngpu": tN,
               "efficiency": t1 / (world * tN), "speedup": t1 / tN, "dofs_per_s": cs.ndof / (tN / 1e3),
               "rel_l2_vs_single": {"ux": rel(ux, ref[0]), "uy_abs_over_ux": float(((uy - ref[1]).norm() / ref[0].norm()).item()),
                                    "c": rel(c, ref[2])},
               "iterations": {"single": {k: v['iterations'] for k, v in info1.items()},
                              "distributed": {"stokes": ds.last_info['iterations'], "advdiff": dc.last_info['iterations']}},
               "phases_ms_last_step": phases, "profiled_step_by_category": prof,
               "solve_only": {"ms_ngpu": phases['stokes_minres'] + phases['advdiff_fgmres'],
                              "ms_1gpu_estimate": t1 - (t1_asm if t1_asm is not None else 0.0)},
               "distributed_levels": {"velocity": ds.vel.nd, "concentration": dc.nd, "replicate_below": rb},
               "exchanges_per_vcycle": {"velocity": 6 * ds.vel.nd - 2, "concentration": 6 * dc.nd - 2,
                                        "note": "per row-partitioned level: Chebyshev step, residual, restriction, prolongation, "
                                                "residual + Chebyshev step of the post-smoother (the last level restricts / prolongs "
                                                "through one vector all-reduce instead)"},
               "halo_bytes": {"velocity_level0_per_exchange": 16 * int(len(lu.ghost)), "owned_velocity_dofs": int(lu.n_own),
                              "concentration_level0_per_exchange": 8 * int(len(dc.parts[0].ghost))},
               "plan_s": t_plan, "setup_s": cs.setup.get('total_s'),
               "assembly": "distributed: every rank assembles the rows it owns from the cells that touch them (system level "
                           "and row-partitioned multigrid levels); only the coarse levels below replicate_below are assembled "
                           "on every rank", "local_assembly_equals_replicated": local_equals,
               "local_assembly_max_abs_diff": local_diff}
        rec["parity_ok"] = bool(local_equals and rec["rel_l2_vs_single"]["ux"] <= 1e-10 and rec["rel_l2_vs_single"]["c"] <= 1e-10
                                and rec["rel_l2_vs_single"]["uy_abs_over_ux"] <= 1e-10)
        out.append(rec)
        ds.close()
        dc.close()
        wd.close()
        cs.mu = mu_saved
        if cs is not case:
            del cs
        dist.barrier()
    return out


def run_gpu(args, rank, world):
    import torch
    import ctypes as C
    from sulcusfem import capi
    from sulcusfem.device import Context
    local = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device(f'cuda:{local}'))
    ctx = Context.get()
    lib = ctx.lib
    from sulcusfem import solvers as _sv
    assert (_sv.RTOL, _sv.STOKES_RTOL) == (RTOL, STOKES_RTOL), "bench.py tolerances differ from the API defaults"
    mu = MU + 0.05 * rank                      # every rank: its own sweep case
    case = Case(ctx, args.h, args.refine, mu)
    mesh, stokes, scalar, plan = case.mesh, case.stokes, case.scalar, case.plan
    t_setup = case.setup['total_s']
    ndof = case.ndof
    info = case.info
    step = case.step

    flush = torch.empty(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)
    flush2 = torch.ones(512 * 1024 * 1024 // 8, dtype=torch.float64, device=ctx.device)

    def flush_l2():
        flush.zero_()          # evicts everything ...
        flush2.sum()           # ... and replaces the dirty lines of the memset by clean ones

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    lib.sfem_launch_count_reset()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    mark = torch.full((1,), 0.5, device=ctx.device)
    mark.erfinv_()             # marker kernel: tools/ncu_summary.py cuts the launch list of the timed region between two of these
    for e0, e1 in ev:
        flush_l2()
        e0.record()
        F, M = step()
        e1.record()
    mark.erfinv_()
    barrier()
    launches = int(lib.sfem_launch_count())
    clocks = sampler.stop() if sampler else None
    ms = sum(e0.elapsed_time(e1) for e0, e1 in ev)
    t = torch.tensor([ms], dtype=torch.float64, device=ctx.device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())

    # ---- e2e through the reference-facing API (host in, host out)
    api_ms = case.api_ms
    case.api_step()
    barrier()
    t0 = time.perf_counter()
    n_e2e = max(1, min(args.steps, 3))
    for _ in range(n_e2e):
        flush_l2()
        _, _, _, fm, mm = case.api_step()
    torch.cuda.synchronize()
    e2e_s = (time.perf_counter() - t0) / n_e2e
    te = torch.tensor([e2e_s], dtype=torch.float64, device=ctx.device)
    if dist is not None:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_s = float(te.item())
    h2d, d2h = case.bytes_per_step()

    # ---- config 5: the SAME step, one case domain-decomposed over all ranks (strong scaling), on the driver's clock
    dd = None
    if world > 1 and not args.no_dd:
        dd = run_dd_strong(args, case, rank, world, dist, flush_l2)

    if rank != 0:
        return
    # ---- per-launch profile of one extra step (roofline of the dominant kernel family)
    cap = 400000
    lib.sfem_profile_start(cap)
    step()
    cats, byts, mss = (C.c_int * cap)(), (C.c_double * cap)(), (C.c_float * cap)()
    n = lib.sfem_profile_stop(cap, cats, byts, mss)
    cats, byts, mss = np.array(cats[:n]), np.array(byts[:n]), np.array(mss[:n], dtype=np.float64)
    per_cat = {}
    for k, name in enumerate(CAT_NAMES):
        sel = cats == k
        if sel.any():
            per_cat[name] = {"launches": int(sel.sum()), "ms": float(mss[sel].sum()), "gbytes": float(byts[sel].sum() / 1e9)}
    # dominant kernel family: SpMV-type launches (plain / +dot / Chebyshev step / residual) on the system-level
    # matrices, i.e. every launch whose algorithmic bytes are >= half of the largest such launch
    spmv_family = np.isin(cats, [0, 1, 2, 3, 8])
    big = spmv_family & (byts >= 0.5 * byts[spmv_family].max())
    ach = float(byts[big].sum() / 1e9 / (mss[big].sum() / 1e3))
    fine_lvl = {}
    for k, name in enumerate(CAT_NAMES):
        sel = big & (cats == k)
        if sel.any():
            fine_lvl[name] = {"launches": int(sel.sum()), "avg_ms": float(mss[sel].mean()),
                              "gbs": float(byts[sel].sum() / 1e9 / (mss[sel].sum() / 1e3))}
    peaks = {}
    try:
        with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
            peaks = json.load(f)
    except Exception:
        pass
    peak = float(peaks.get('hbm_gbs', 6650.0))
    # DRAM traffic of the dominant kernel: not measurable here (needs ncu); taken from the committed `ncu --set full`
    # capture of this same command (profiles/r02_sell_traffic.json: measured dram bytes / algorithmic bytes of the
    # system-level launches) and scaled to this run's algorithmic bytes per launch
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, 'profiles', 'r02_sell_traffic.json')) as f:
            tj = json.load(f)
        traffic = float(tj['dram_over_algorithmic']) * float(byts[big].mean())
        traffic_src = tj.get('source')
    except Exception:
        pass
    roofline = {"bound": "hbm", "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak, "traffic": traffic,
                "traffic_source": traffic_src,
                "kernel": "FP64 sliced-ELL SpMV family on the system-level matrices (k_sell_stream with the store / dot / "
                          "Chebyshev-step / residual epilogues; CSR lane-group kernels serve the small multigrid levels)",
                "launches": int(big.sum()), "avg_launch_ms": float(mss[big].mean()),
                "algorithmic_bytes_per_launch": float(byts[big].mean()),
                "peak_source": "measured (MEASURED_PEAKS.json hbm_gbs)" if 'hbm_gbs' in peaks else "fallback 6650 GB/s",
                "share_of_profiled_step": float(mss[big].sum() / mss.sum()),
                "system_level_by_kernel": fine_lvl,
                "by_category": per_cat, "profiled_step_kernel_ms": float(mss.sum()),
                "note": "profiled step runs un-graphed with a CUDA event pair around every launch; traffic: see profiles/"}

    # ---- BASELINE configs[3]: the reference's mu sweep (batched Krylov loops), N = 1 only; wall-clock numbers of a
    # Python driver loop, so it runs before the CPU legs fill the heap with the oracle's factorisations
    sweep = None
    if world == 1 and not args.no_sweep:
        import gc
        gc.collect()
        try:
            sweep = sweep_leg(with_oracle=not args.no_cpu)
        except Exception as e:                   # an auxiliary leg must not cost the headline line; reported and exit code 3
            sweep = {"ok": False, "error": f"{type(e).__name__}: {e}"[:500]}

    # ---- CPU baseline (oracle port) + same-mesh GPU legs with parity, rank 0 at N = 1 only
    cpu, same = None, []
    if not args.no_cpu and world == 1:
        try:
            for hh in dict.fromkeys([args.same_h, args.cpu_h]):
                same.append(same_mesh_leg(ctx, hh, flush_l2))
            r0 = same[0]
            cpu = {"value": r0["cpu_dofs_per_s"], "unit": "DOFs/s", "cores": 1, "kind": "port",
                   "sample": f"one step of the same workload on the unrefined h={r0['h']} mesh ({r0['dofs']} dofs; h=0.02 is the "
                             f"reference's default mesh size): numpy assembly + SuperLU, {r0['cpu_ms'] / 1e3:.1f} s; the GPU on this "
                             f"same mesh: same_mesh[0]",
                   "host_cores": os.cpu_count()}
        except Exception as e:                   # reported in the line (and exit code 3) instead of losing the headline numbers
            same.append({"error": f"{type(e).__name__}: {e}"[:500], "parity": {"ok": False}})

    value = world * ndof * args.steps / (ms_total / 1e3)
    line = {
        "metric": "fem_dofs_per_s", "value": value, "unit": "DOFs/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_total / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args),
        "solves_per_s": world * args.steps / (ms_total / 1e3), "dofs_per_case": ndof,
        "dofs": {"taylor_hood": stokes.n, "p2": scalar.n, "cells": int(mesh.num_cells)},
        "iterations": {"stokes_minres": info['stokes']['iterations'], "advdiff_fgmres": info['advdiff']['iterations'],
                       "stokes_relres": info['stokes']['relres'], "advdiff_relres": info['advdiff']['relres']},
        "e2e": {"value": world * ndof / e2e_s, "unit": "DOFs/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "ms_per_step": 1e3 * e2e_s, "last_step_breakdown_ms": api_ms,
                "api": "sulcusfem.solvers.stokes_solver + advdiff_solver + analysis.compute_*_metrics"},
        "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        "setup_s": t_setup, "setup_breakdown_s": case.setup,
        "same_mesh": same, "dd_strong": dd, "mu_sweep": sweep,
        "functionals": {"uptake_flux": fm['uptake_flux'], "total_mass": mm['total_mass']},
    }
    line["config"]["dofs"] = ndof
    print(json.dumps(line), flush=True)
    bad = [r for r in same if not r["parity"]["ok"]]
    if bad:
        print("PARITY FAILURE in the same-mesh leg: " + json.dumps([r["parity"] for r in bad]), file=sys.stderr)
        sys.exit(3)
    if sweep is not None and not sweep["ok"]:
        print("PARITY FAILURE in the mu-sweep leg: " + json.dumps(sweep), file=sys.stderr)
        sys.exit(3)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200')
    ap.add_argument('--h', type=float, default=0.02)
    ap.add_argument('--refine', type=int, default=int(os.environ.get('SFEM_BENCH_REFINE', 2)))
    ap.add_argument('--cpu-h', type=float, default=0.03)    # reference arm: ~4-9 s of single-thread CPU work per step (168 k dofs)
    ap.add_argument('--same-h', type=float, default=0.02)   # same-mesh leg / cpu_baseline: the reference's default mesh size (385 k dofs)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-sweep', action='store_true')      # skip the mu-sweep leg (BASELINE configs[3])
    ap.add_argument('--no-dd', action='store_true')         # N > 1: skip the domain-decomposed (config 5) leg
    ap.add_argument('--dd-replicate-below', type=int, default=200000)   # multigrid levels below this many unknowns stay replicated
    # N > 1: the dd leg also runs these refinement levels (the bench mesh, 6.1 M dofs, is too small for 8 GPUs; r = 3 =
    # 24.4 M dofs adds ~50 s of set-up per rank); '' = the bench mesh only
    ap.add_argument('--dd-refine', default=os.environ.get('SFEM_BENCH_DD_REFINE', '3'))
    args = ap.parse_args()
    rank = int(os.environ.get('RANK', 0))
    world = int(os.environ.get('WORLD_SIZE', 1))
    if args.impl == 'reference':
        run_reference(args, rank)
        return
    run_gpu(args, rank, world)
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
