"""The six solver entry points of the reference's ``solvers.py``, served by the B200 path.

Signatures, argument meaning, printed diagnostics and error behaviour follow the reference
(``solvers.py:16,59,113,176,237,308``) so ``simulation.run_simulation`` and the study drivers call
them unchanged.  Inputs are duck-typed: ``mesh_results['mesh']`` is a :class:`HostMesh`,
``mesh_results['bc_markers']`` anything with ``.array()``; ``C`` / ``W`` / ``V`` / ``Q`` are
``fem.FunctionSpace`` objects (only the mesh and element are consulted); ``D`` / ``mu`` are
``Constant`` or floats; ``mu_function`` is any object with ``eval(values, x)`` (e.g.
``StepUptakeOpen``); ``u`` is a P2-vector ``Function``.

What the reference hands to dolfin's ``solve(a == L, u, bcs)`` (assemble -> ``DirichletBC.apply`` ->
sparse LU) runs here as: element/facet kernels + gather into the CSR pattern, Dirichlet kernel,
then multigrid-preconditioned CG (pure diffusion), FGMRES (advection-diffusion) or MINRES
(Taylor-Hood Stokes) to a relative residual of ``RTOL`` -- all on the device through the C ABI of
``libsulcusfem.so``.  There is no CPU fallback.
"""
from __future__ import annotations

import contextlib
import threading

import numpy as np

from . import dofmap as dm
from .fem import Constant, Function, FunctionSpace, VectorFunctionSpace, evaluate_expression

# Krylov tolerances, chosen from measurements against the LU oracle (tools/tolerance_study.py,
# profiles/r01_tolerance_study.md) so that every field meets the 1e-10 relative-L2 parity bar with margin:
#  * scalar CG / FGMRES: true relative residual 1e-13 -- the concentration error is ~40x (30 k dofs) to ~1200x
#    (1.9 M dofs) the relative residual, i.e. <= 4e-11 on the largest mesh measured; do not loosen;
#  * Stokes MINRES: preconditioned residual 1e-12 -- the u / p errors track it one to one (3e-13 at 66 k dofs, 9e-13
#    at 4.2 M dofs), 100x below the bar, and the concentration computed from that velocity moves by 3e-13
RTOL = 1e-13
STOKES_RTOL = 1e-12
_CACHE_ATTR = '_sfem_cache'


def _cache(mesh):
    c = getattr(mesh, _CACHE_ATTR, None)
    if c is None:
        c = {}
        setattr(mesh, _CACHE_ATTR, c)
    return c


def _slot_key(key):
    """Cache key of a device object: unchanged on the main thread, tagged with the worker's slot inside
    ``sweep.run_concurrent`` (concurrent cases own separate device buffers, graphs and streams)."""
    from .sweep import current_slot
    slot = current_slot()
    return key if slot == 0 else (key, 'slot', slot)


def _markers_array(m):
    return np.asarray(m.array() if hasattr(m, 'array') else m)


def _hierarchy(mesh):
    c = _cache(mesh)
    if 'hierarchy' not in c:
        from .hierarchy import build_hierarchy
        c['hierarchy'] = build_hierarchy(mesh)
    return c['hierarchy']


def scalar_problem(mesh, bc_markers, robin_id=4):
    """Cached device problem (patterns, gather maps, multigrid hierarchy) of one mesh."""
    from .device import ScalarProblem
    c = _cache(mesh)
    key = _slot_key(('scalar', int(robin_id)))
    if key not in c:
        c[key] = ScalarProblem(mesh, _markers_array(bc_markers), dirichlet_ids=(1, 2), robin_id=robin_id,
                               hierarchy=_hierarchy(mesh))
    return c[key]


def stokes_problem(mesh, bc_markers):
    from .device import StokesProblem
    c = _cache(mesh)
    key = _slot_key('stokes')
    if key not in c:
        c[key] = StokesProblem(mesh, _markers_array(bc_markers), hierarchy=_hierarchy(mesh))
    return c[key]


def _check_space(C, kind):
    if not isinstance(C, FunctionSpace) or C.kind != kind:
        raise ValueError(f"expected a {kind} function space")
    return C.mesh()


def _velocity_components(u, mesh):
    if u is None:
        return None, None
    if isinstance(u, Constant):
        v = u.values()
        if np.all(v == 0.0):
            return None, None
        from .device import Context
        ctx = Context.get()
        n2 = dm.p2_num_dofs(mesh)
        return ctx.up(np.full(n2, v[0]), np.float64), ctx.up(np.full(n2, v[1]), np.float64)
    if u.function_space().kind != 'P2v':
        raise ValueError("velocity must be a P2 vector Function")
    if not np.any(u.values):
        return None, None
    return u.device_components()


def _mu_nodal(mu_function, prob):
    """P2 interpolant of mu at the dofs of the Robin boundary cells' facets (App. A.3)."""
    lev = prob.fine
    X = dm.p2_dof_coordinates(prob.mesh)
    facet_dofs = np.unique(dm.p2_facet_dofs(prob.mesh, lev.robin_facets).ravel()) if lev.nf else np.zeros(0, dtype=np.int64)
    vals = np.zeros(lev.n)
    if len(facet_dofs):
        vals[facet_dofs] = evaluate_expression(mu_function, X[facet_dofs])
    mean = float(vals[facet_dofs].mean()) if len(facet_dofs) else 0.0
    return prob.ctx.up(vals, np.float64), vals, mean


# ---------------------------------------------------------------------- mu sweeps: frozen coarse levels
# In a Robin-coefficient sweep on one geometry (no_advection_analysis_A.py:1306-1347: 20 mu values, one mesh) only the
# boundary rows of A(mu) = D K + mu M_Gamma change.  Inside ``frozen_coarse_levels()`` a constant-mu pure-diffusion solve
# re-assembles the system level only and keeps the multigrid levels of the last full assembly as long as mu stays within
# ``factor`` of the mu they were built for: the coarse levels are preconditioner data, CG still iterates on the exact
# system to the same true residual (RTOL), so the fields are unchanged to solver tolerance (tests/test_gpu_studies.py) --
# it saves ~30 coarse-level launches, the dense coarse inverse and ~30 host calls per case.
_sweep = threading.local()


@contextlib.contextmanager
def frozen_coarse_levels(factor=4.0):
    prev = getattr(_sweep, 'factor', None)
    _sweep.factor = float(factor)
    try:
        yield
    finally:
        _sweep.factor = prev


def _reuse_coarse(prob, mu):
    factor = getattr(_sweep, 'factor', None)
    ref = getattr(prob, '_coarse_mu', None)
    if factor is None or ref is None or mu is None or not (mu > 0.0 and ref > 0.0):
        return False
    return max(mu / ref, ref / mu) <= factor


def _solve_scalar(mesh_results, C, D, u, mu=None, mu_function=None, clamp=False, bottom_id=4):
    mesh = _check_space(C, 'P2')
    prob = scalar_problem(mesh, mesh_results['bc_markers'], bottom_id)
    ux, uy = _velocity_components(u, mesh)
    kw = dict(mu_const=0.0)
    mu_host = None
    if mu_function is not None:
        mu_dev, mu_host, mean = _mu_nodal(mu_function, prob)
        kw = dict(mu_nodal=mu_dev, clamp=clamp, coarse_mu=max(mean, 0.0))
    elif mu is not None:
        kw = dict(mu_const=float(mu))
    reuse = ux is None and mu_function is None and mu is not None and _reuse_coarse(prob, float(mu))
    prob.assemble(float(D), ux, uy, bc_values={1: 1.0, 2: 0.0}, reuse_coarse=reuse, **kw)
    method = 'cg' if ux is None else 'fgmres'
    x = prob.solve(method, rtol=RTOL)
    _accept_scalar(prob.last_info)
    return prob, x


# A TRUE relative residual of 10 * RTOL = 1e-12 sits close to the FP64 floor eps * ||A|| ||x|| / ||b|| of large refined
# (or row-partitioned) systems; the reference's LU never "fails" there.  A recurrence that stagnates below this
# parity-safe level (concentration error <= ~1200x the residual, profiles/r01_tolerance_study.md) is accepted with a
# warning; anything above it is an error.
PARITY_SAFE_RELRES = 5e-12


def _accept_scalar(info):
    if info['converged']:
        return
    if np.isfinite(info['relres']) and info['relres'] <= PARITY_SAFE_RELRES:
        import warnings
        warnings.warn(f"Krylov recurrence stagnated at true relative residual {info['relres']:.2e} "
                      f"(target {10 * RTOL:.0e}); accepted, parity-safe: {info}")
        return
    raise RuntimeError(f"Krylov solver did not converge: {info}")


def _post(prob, x, fix_nonfinite):
    import ctypes as C
    from . import capi
    stats = (C.c_double * 6)()
    capi.check(prob.ctx.lib.sfem_postprocess_concentration(prob.n, capi.ptr(x), int(fix_nonfinite), stats, prob.ctx.stream),
               'sfem_postprocess_concentration')
    return {'nonfinite': int(stats[0]), 'negative': int(stats[1]), 'min': stats[2], 'max': stats[3], 'mean': stats[4],
            'clamped': bool(stats[5])}


def _to_function(C, prob, x):
    f = Function(C, prob.ctx.down(x))
    f._dev = (x.clone(),)
    f.solver_info = dict(prob.last_info)
    return f


# ====================================================================== advection-diffusion
def advdiff_solver(mesh_results, u, C, D, mu, mesh_type="sulcus"):
    """Steady advection-diffusion with constant Robin uptake (reference solvers.py:16-57)."""
    prob, x = _solve_scalar(mesh_results, C, D, u, mu=mu)
    return _to_function(C, prob, x)


def advdiff_solver_variable_mu(mesh_results, u, C, D, mu_function, mesh_type="sulcus"):
    """Advection-diffusion with a spatially varying Robin coefficient mu(x) (solvers.py:59-107)."""
    prob, x = _solve_scalar(mesh_results, C, D, u, mu_function=mu_function, clamp=False)
    st = _post(prob, x, fix_nonfinite=True)
    if st['nonfinite']:
        print(f"WARNING: {st['nonfinite']} non-finite concentration entries; clamping to 0.")
    if st['negative']:
        if st['clamped']:
            print("✓ Clamped tiny negative values to 0 (numerical noise).")
            st = _post(prob, x, fix_nonfinite=False)
        else:
            print(f"WARNING: {st['negative']} negative values; most negative {st['min']:.3e}")
    print(f"Solution stats: min={st['min']:.6e}, max={st['max']:.6e}, mean={st['mean']:.6e}")
    return _to_function(C, prob, x)


# ====================================================================== diffusion only
def _pure_diffusion_stats(prob, x):
    """The validation block of the reference's pure_diffusion_solver (solvers.py:154-173) as (stats, printed lines)."""
    st = _post(prob, x, fix_nonfinite=False)
    lines = []
    if st['negative'] > 0:
        if st['clamped']:
            st = _post(prob, x, fix_nonfinite=False)
        else:
            lines += [f"WARNING: {st['negative']} negative concentration values found!",
                      f"  Most negative: {st['min']:.6e}",
                      f"  Min: {st['min']:.6e}, Max: {st['max']:.6e}",
                      "  Check: mesh quality, boundary conditions, solver settings"]
    else:
        lines.append("✓ All concentration values are non-negative")
    lines.append(f"Solution stats: min={st['min']:.6e}, max={st['max']:.6e}, mean={st['mean']:.6e}")
    return st, lines


def pure_diffusion_solver(mesh_results, C, D, mu, mesh_type="sulcus"):
    """Steady diffusion with constant Robin uptake (reference solvers.py:113-174).  A field that
    :func:`presolve_pure_diffusion` already computed for this (mesh, D, mu) is handed out instead of solving again."""
    hit = _take_presolved(_check_space(C, 'P2'), _presolve_key(D, mu))
    if hit is not None:
        f, lines = hit
        print("\n".join(lines))
        return f
    prob, x = _solve_scalar(mesh_results, C, D, None, mu=mu)
    _, lines = _pure_diffusion_stats(prob, x)
    print("\n".join(lines))
    return _to_function(C, prob, x)


# ---------------------------------------------------------------------- batched Robin sweeps
# The reference's mu sweeps (no_advection_analysis_A.py:1306-1347: 20 mu on one mesh; no_advection_analysis_B.py:110-141:
# 3 mu per geometry) call pure_diffusion_solver once per mu.  A(mu) = D K + mu M_Gamma differs between the cases in its
# boundary rows only, so up to BATCH coefficients are solved in ONE Krylov loop (sfem_krylov_cg_batch: interleaved
# right-hand sides, per-column CG scalars; the multigrid hierarchy assembled for the batch's geometric-mean mu supplies
# patterns, transfers and the dense coarsest inverse, every column smooths with its own operator on every level).
BATCH = int(__import__('os').environ.get('SFEM_BATCH', 16))     # coefficients per Krylov loop (library limit: 16)
BATCH_SPAN = 64.0         # largest mu / smallest mu inside one batch: bounds the Chebyshev steps of the coarsest-level solve
                          # (every column runs on its own operators; iteration counts do not depend on the span -- measured)


def _presolve_key(D, mu):
    return (float(D), float(mu))


# Parked fields live in the per-mesh cache: 'presolved' {key: (Function, printed lines)}; 'presolve_pending' = keys a
# background pre-solve (presolve_pure_diffusion(background=True)) has announced but not delivered yet.
_presolve_cv = threading.Condition()
_presolve_stream = None
_presolve_pool = None          # one persistent worker thread: its library work spaces and captured graphs survive between sweeps
PRESOLVE_SLOT = 1000           # device-problem slot of the worker (never shared with the per-case path of any other thread)


def _take_presolved(mesh, key, timeout=120.0):
    """The parked field of ``key`` (None if there is none); waits while a background pre-solve still owes it."""
    c = _cache(mesh)
    with _presolve_cv:
        while True:
            pre = c.get('presolved')
            if pre and key in pre:
                return pre.pop(key)
            pend = c.get('presolve_pending')
            if not pend or key not in pend:
                return None
            if not _presolve_cv.wait(timeout):
                pend.discard(key)          # the worker is stuck or gone: the caller solves this case itself
                return None


def _publish_presolved(mesh, keys, hits):
    c = _cache(mesh)
    with _presolve_cv:
        pre = c.setdefault('presolved', {})
        pend = c.get('presolve_pending')
        for k, h in zip(keys, hits):
            if h is not None:
                pre[k] = h
            if pend is not None:
                pend.discard(k)
        _presolve_cv.notify_all()


def pure_diffusion_solver_batch(mesh_results, C, D, mus, mesh_type="sulcus"):
    """``[pure_diffusion_solver(mesh_results, C, D, mu) for mu in mus]`` in batched solves of up to ``BATCH`` coefficients
    (nearby coefficients share a batch: the list is processed in ascending order of mu).  Returns the Functions in the
    order of ``mus``; each carries ``solver_info`` and prints the reference's validation lines."""
    out = _solve_batches(mesh_results, C, D, mus)
    fs = []
    for f, lines in out:
        print("\n".join(lines))
        fs.append(f)
    return fs


def presolve_pure_diffusion(mesh_results, C, D, mus, background=False):
    """Solve all ``mus`` of one geometry in batches and park the fields on the mesh; the next
    ``pure_diffusion_solver(mesh_results, C, D, mu)`` call for each of them (e.g. from ``run_simulation``) returns the
    parked field.  ``background=True``: the batches are solved by a worker thread on its own CUDA stream and device
    problem while the caller goes on (a consumer that asks for a field not delivered yet waits for it), so the
    per-case post-processing of one batch overlaps the Krylov loops of the next.  Returns the number of fields."""
    mesh = _check_space(C, 'P2')
    mus = [float(m) for m in mus]
    keys = [_presolve_key(D, mu) for mu in mus]
    if not background:
        _solve_batches(mesh_results, C, D, mus, publish=lambda idx, hits: _publish_presolved(mesh, [keys[i] for i in idx], hits))
        import torch
        torch.cuda.current_stream().synchronize()  # the parked fields may be consumed on other streams (sweep workers)
        return len(mus)
    global _presolve_pool
    import torch
    from concurrent.futures import ThreadPoolExecutor
    if _presolve_pool is None:
        _presolve_pool = ThreadPoolExecutor(max_workers=1, thread_name_prefix='sfem-presolve')
    with _presolve_cv:
        _cache(mesh).setdefault('presolve_pending', set()).update(keys)
    device = torch.cuda.current_device()
    torch.cuda.current_stream().synchronize()      # whatever the caller queued (mesh uploads) is visible to the worker

    def work():
        from . import sweep
        try:
            torch.cuda.set_device(device)
            sweep._tls.slot = PRESOLVE_SLOT
            global _presolve_stream
            if _presolve_stream is None:
                _presolve_stream = torch.cuda.Stream()     # one stream for the worker's lifetime (captured graphs are keyed on it)
            st = _presolve_stream
            with torch.cuda.stream(st):
                _solve_batches(mesh_results, C, D, mus,
                               publish=lambda idx, hits: (st.synchronize(), _publish_presolved(mesh, [keys[i] for i in idx], hits)))
        except BaseException as e:               # the consumers fall back to their own solves
            import warnings
            warnings.warn(f"background pre-solve failed ({type(e).__name__}: {e}); cases are solved one by one")
        finally:
            _publish_presolved(mesh, keys, [None] * len(keys))       # nothing stays pending
    _presolve_pool.submit(work)
    return len(mus)


def _solve_batches(mesh_results, C, D, mus, publish=None):
    """Batched solves of ``mus``; returns [(Function, printed lines)] in the caller's order.  ``publish(indices, hits)``
    is called after every batch with the positions (in ``mus``) and results it produced."""
    mesh = _check_space(C, 'P2')
    prob = scalar_problem(mesh, mesh_results['bc_markers'], 4)
    mus = [float(m) for m in mus]
    order = sorted(range(len(mus)), key=lambda i: mus[i])
    out = [None] * len(mus)
    chunks, cur = [], []
    for i in order:              # a batch holds <= BATCH coefficients within a factor BATCH_SPAN (one shared preconditioner)
        if cur and (len(cur) >= BATCH or (mus[cur[0]] > 0.0 and mus[i] > BATCH_SPAN * mus[cur[0]])):
            chunks.append(cur)
            cur = []
        cur.append(i)
    if cur:
        chunks.append(cur)
    for idx in chunks:
        X, infos = prob.solve_batch(float(D), [mus[i] for i in idx], {1: 1.0, 2: 0.0}, rtol=RTOL)
        for c, i in enumerate(idx):
            _accept_scalar(infos[c])
            x = prob.batch_column(X, len(idx), c)
            _, lines = _pure_diffusion_stats(prob, x)
            f = Function(C, prob.ctx.down(x))
            f._dev = (x,)
            f.solver_info = dict(infos[c])
            out[i] = (f, lines)
        if publish is not None:
            publish(list(idx), [out[i] for i in idx])
    return out


def pure_diffusion_solver_variable_mu(mesh_results, C, D, mu_function, mesh_type="rectangular", bottom_id=4, u=None):
    """Diffusion (optionally advected by ``u``) with mu(x) clamped to >= 0 at the quadrature points
    (reference solvers.py:176-231)."""
    prob, x = _solve_scalar(mesh_results, C, D, u, mu_function=mu_function, clamp=True, bottom_id=bottom_id)
    st = _post(prob, x, fix_nonfinite=False)
    if st['negative'] > 0:
        if st['clamped']:
            st = _post(prob, x, fix_nonfinite=False)
        else:
            print(f"WARNING: {st['negative']} negative concentration values found!")
            print(f"  Most negative: {st['min']:.6e}")
            print("  Check: mesh quality, BCs, solver settings")
    print(f"Solution stats: min={st['min']:.6e}, max={st['max']:.6e}, mean={st['mean']:.6e}")
    return _to_function(C, prob, x)


# ====================================================================== Stokes
def stokes_solver(mesh_results, W, L_domain, H, mesh_type="sulcus"):
    """Taylor-Hood Stokes flow driven by a Poiseuille inlet (reference solvers.py:237-306).

    Inlet ``(4y(H-y), 0)`` on id 1, no-slip on ids 4 and 3 (applied in that order), natural outflow
    on id 2.  The reference's "pointwise" pressure pin matches no dof, so the
    pressure level is fixed by the outflow condition alone -- same here.
    """
    bm_obj = mesh_results['bc_markers']
    unique_vals = getattr(bm_obj, '_unique_ids', None)
    if unique_vals is None:
        unique_vals = np.unique(_markers_array(bm_obj))
        try:
            bm_obj._unique_ids = unique_vals
        except AttributeError:
            pass
    print(f"Boundary markers present: {unique_vals}")
    try:
        mesh = _check_space(W, 'TH')
        bm = _markers_array(mesh_results['bc_markers'])
        prob = stokes_problem(mesh, bm)
        print(f"Trying pressure constraint at outlet center: ({L_domain}, {H/2})")
        if getattr(prob, '_inflow_H', None) != float(H):    # Dirichlet data depend on the mesh and H only
            X = dm.p2_dof_coordinates(mesh)
            d1 = dm.dirichlet_dofs_p2(mesh, bm, 1)
            prob.set_bcs({1: (4.0 * X[d1, 1] * (H - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})
            prob.set_channel_flow_guess(float(L_domain), float(H))
            prob._inflow_H = float(H)
        prob.assemble(bc_mode=1)
        ux, uy, p = prob.solve(rtol=STOKES_RTOL)
        info = prob.last_info
        if not info['converged']:
            # MINRES stopped on maxit / stagnation above STOKES_RTOL (preconditioned residual).  The u / p errors track
            # that residual one to one, so only a true residual within 10x of the target is still parity-safe.
            if not np.isfinite(info['relres']) or info['relres'] > 10.0 * STOKES_RTOL:
                raise RuntimeError(f"MINRES stalled: {info}")
            import warnings
            warnings.warn(f"MINRES stopped before the preconditioned residual reached {STOKES_RTOL:.0e}; true relative "
                          f"residual {info['relres']:.2e} accepted: {info}")
        n2 = prob.n2
        u = Function(VectorFunctionSpace(mesh, 'P', 2), prob.ctx.down(prob.x[:2 * n2]))
        u._dev = (ux.clone(), uy.clone())
        pf = Function(FunctionSpace(mesh, 'P', 1), prob.ctx.down(p))
        u.solver_info = pf.solver_info = dict(prob.last_info)
        print(f"✓ Stokes solver completed using outlet point constraint for {mesh_type} mesh")
        return u, pf
    except Exception as e:
        print(f"Outlet point constraint failed: {e}")
        raise RuntimeError("Unable to solve Stokes system.")


def stokes_solver_no_adv(V, Q):
    """Zero velocity / pressure fields for the no-advection mode (reference solvers.py:308-316)."""
    return Function(V), Function(Q)
