"""CPU checks of the two spectral facts the batched Robin sweep (csrc/sfem_batch.cu) relies on, and of the host-side
batching logic of ``solvers._solve_batches`` -- on the oracle's own matrices (test infrastructure only)."""
import numpy as np
import pytest


@pytest.fixture(scope='module')
def robin_pair():
    """A0 = D K with identity Dirichlet rows and M = M_Gamma with zero Dirichlet rows / columns on a coarse sulcus mesh."""
    import scipy.sparse as sp
    from oracle import cpu_oracle as co
    from sulcusfem import hostmesh as hm
    from sulcusfem.unstructured import mesh_domain
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.2, 'sulcus')
    mk = hm.build_markers(mesh, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    om = co.Mesh(mesh.coords, mesh.cells)
    bm = mk['bc_markers'].values
    K = co.assemble_p2_stiffness(om).tocsr()
    f4 = np.flatnonzero((bm == 4) & om.on_boundary)
    M = co.assemble_p2_robin(om, f4, mu_const=1.0).tocsr()
    dofs, _ = co.concentration_bcs(om, bm)
    keep = np.ones(om.n_p2)
    keep[dofs] = 0.0
    Dk = sp.diags(keep)
    A0 = (Dk @ K @ Dk + sp.diags(1.0 - keep)).toarray()
    Mm = (Dk @ M @ Dk).toarray()
    return A0, Mm


def test_spectrum_of_reference_preconditioned_operator_is_the_mu_ratio_window(robin_pair):
    """Coarsest level of a batch: the dense inverse belongs to A(mu_ref); column c runs a Chebyshev iteration on
    A(mu_ref)^-1 A(mu_c) with the window [min(1, mu_c / mu_ref), max(1, mu_c / mu_ref)].  A(mu) = A0 + mu M with M
    positive semi-definite is monotone in mu (Loewner order), so the window must contain the whole spectrum."""
    import scipy.linalg as sl
    A0, M = robin_pair
    assert np.linalg.eigvalsh(M).min() > -1e-12
    for mu_ref, mu_c in ((1.0, 0.05), (1.0, 1.0), (2.0, 60.0), (0.3, 0.0)):
        ev = sl.eigh(A0 + mu_c * M, A0 + mu_ref * M, eigvals_only=True)
        ratio = mu_c / mu_ref
        assert ev.min() >= min(1.0, ratio) - 1e-10 and ev.max() <= max(1.0, ratio) + 1e-10, (mu_ref, mu_c, ev.min(), ev.max())
        if mu_c == 0.0:
            assert ev.min() > 0.0          # mu_c = 0: in (0, 1], the lower end of the window is only a guess there


def test_gershgorin_bound_of_a_batch_is_attained_at_an_end_point(robin_pair):
    """One set of Chebyshev coefficients serves every column of a level: the Gershgorin row bound of D^-1 (A0 + mu M) is
    convex-over-linear in mu, so its maximum over [mu_lo, mu_hi] sits at mu_lo or mu_hi (kb_gershgorin evaluates both)."""
    A0, M = robin_pair

    def bound(mu):
        A = A0 + mu * M
        return (np.abs(A).sum(axis=1) / np.abs(np.diag(A))).max()
    for lo, hi in ((0.0, 1.0), (0.1, 6.4), (2.0, 300.0)):
        ends = max(bound(lo), bound(hi))
        for mu in np.geomspace(max(lo, 1e-3), hi, 23):
            assert bound(mu) <= ends * (1.0 + 1e-12)
        # and the bound really bounds the spectrum of the Jacobi-preconditioned operator of every column
        for mu in (lo, 0.5 * (lo + hi), hi):
            A = A0 + mu * M
            d = 1.0 / np.sqrt(np.diag(A))
            assert np.linalg.eigvalsh(d[:, None] * A * d[None, :]).max() <= ends * (1.0 + 1e-12)


def test_chebyshev_coarse_solve_reaches_the_promised_reduction(robin_pair):
    """The coefficient recurrence of the coarsest-level solve (host code of sfem_krylov_cg_batch, restated): m steps chosen
    from the widest window give at least the promised (20-fold) reduction of the energy-norm error for every column."""
    A0, M = robin_pair
    mu_ref, mus, eps = 1.0, [0.125, 0.7, 1.0, 8.0], 0.01
    inv = np.linalg.inv(A0 + mu_ref * M)
    lo = [min(1.0, m / mu_ref) * (1 - eps) for m in mus]
    hi = [max(1.0, m / mu_ref) * (1 + eps) for m in mus]
    kmax = max(h / l for h, l in zip(hi, lo))
    rr = (np.sqrt(kmax) - 1) / (np.sqrt(kmax) + 1)
    m = 1
    while m < 8 and 2 * rr ** m / (1 + rr ** (2 * m)) > 0.05:
        m += 1
    rng = np.random.default_rng(3)
    b = rng.standard_normal(A0.shape[0])
    for c, mu in enumerate(mus):
        A = A0 + mu * M
        theta, delta = 0.5 * (hi[c] + lo[c]), 0.5 * (hi[c] - lo[c])
        sigma = theta / delta
        rho = 1.0 / sigma
        d = inv @ b / theta
        x = d.copy()
        for _ in range(1, m):
            z = inv @ (b - A @ x)
            rho_new = 1.0 / (2.0 * sigma - rho)
            d = rho_new * rho * d + (2.0 * rho_new / delta) * z
            x += d
            rho = rho_new
        exact = np.linalg.solve(A, b)
        e, e0 = x - exact, exact
        assert np.sqrt(e @ A @ e) <= 0.05 * np.sqrt(e0 @ A @ e0), (mu, m)


def test_batches_are_formed_by_width_and_span_and_results_keep_the_callers_order(monkeypatch):
    """solvers._solve_batches: ascending mu, at most BATCH coefficients and a factor BATCH_SPAN per batch, results handed
    back in the caller's order (host logic; the device problem is a stand-in)."""
    import sulcusfem.solvers as sv
    calls = []

    class FakeCtx:
        def down(self, x):
            return np.asarray(x)

    class FakeProb:
        ctx = FakeCtx()
        n = 3

        def solve_batch(self, D, mus, bc, rtol):
            calls.append(list(mus))
            return ('X', list(mus)), [{'iterations': 1, 'relres': 0.0, 'converged': True, 'estimate': 0.0, 'method': 'cg_batch'}
                                      for _ in mus]

        def batch_column(self, X, nb, c):
            return np.full(3, X[1][c])
    monkeypatch.setattr(sv, 'scalar_problem', lambda mesh, bm, rid: FakeProb())
    monkeypatch.setattr(sv, '_check_space', lambda C, kind: object())
    monkeypatch.setattr(sv, '_pure_diffusion_stats', lambda prob, x: ({}, ['ok']))

    class FakeFunction:
        def __init__(self, C, values):
            self.values = values
    monkeypatch.setattr(sv, 'Function', FakeFunction)
    monkeypatch.setattr(sv, 'BATCH', 4)
    mus = [50.0, 0.1, 3.0, 0.2, 1000.0, 0.4, 0.8, 1.6, 7.0]
    out = sv._solve_batches({'bc_markers': None}, None, 1.0, mus)
    assert [f.values[0] for f, _ in out] == mus                       # caller's order
    assert calls == [[0.1, 0.2, 0.4, 0.8], [1.6, 3.0, 7.0, 50.0], [1000.0]]   # width 4, then the span limit (1000 > 64 * 1.6 would
    assert all(max(c) <= sv.BATCH_SPAN * min(c) for c in calls)               # also have split a wider batch)


def test_parked_fields_wait_for_a_background_presolve_and_fall_back_when_it_never_delivers():
    """solvers._take_presolved / _publish_presolved: a consumer asking for a field a background pre-solve has announced
    waits for it; one that is not announced returns at once; one that is never delivered gives up (the caller then
    solves the case itself)."""
    import threading
    import time
    import sulcusfem.solvers as sv

    class M:
        pass
    mesh = M()
    key, other, lost = (1.0, 2.0), (1.0, 3.0), (1.0, 4.0)
    assert sv._take_presolved(mesh, key) is None                       # nothing parked, nothing pending
    with sv._presolve_cv:
        sv._cache(mesh).setdefault('presolve_pending', set()).update([key, lost])
    t = threading.Timer(0.2, lambda: sv._publish_presolved(mesh, [key], [('field', ['line'])]))
    t0 = time.perf_counter()
    t.start()
    assert sv._take_presolved(mesh, other) is None                     # not announced: no waiting
    assert time.perf_counter() - t0 < 0.15
    assert sv._take_presolved(mesh, key) == ('field', ['line'])        # announced: waited for the delivery
    assert time.perf_counter() - t0 >= 0.19
    assert sv._take_presolved(mesh, key) is None                       # handed out exactly once
    assert sv._take_presolved(mesh, lost, timeout=0.1) is None         # never delivered: give up, and stop waiting for it
    assert lost not in sv._cache(mesh)['presolve_pending']
    t.join()
