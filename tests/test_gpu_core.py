"""GPU parity tests of the device kernels against the CPU oracle (through the C ABI)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope='module')
def ctx():
    from sulcusfem.device import Context
    return Context.get()


@pytest.fixture(scope='module')
def small():
    """Small unstructured sulcus mesh + markers + oracle mesh."""
    from sulcusfem import hostmesh as hm
    from sulcusfem.unstructured import mesh_domain
    from oracle import cpu_oracle as co
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.08, 'sulcus')
    mk = hm.build_markers(mesh, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    om = co.Mesh(mesh.coords, mesh.cells)
    return mesh, mk, om


@pytest.fixture(params=['vector', 'staged', 'sell'])
def engine(request, ctx):
    """Run a test once with the vector SpMV engine only, once with the TMA-staged engine forced on every matrix
    that has a tile plan and once with the sliced-ELL engine forced on every matrix that has a mirror (by
    default only large matrices use the latter two)."""
    old = ctx.lib.sfem_staged_set_min_tiles(1 if request.param == 'staged' else 1 << 30)
    old_sell = ctx.lib.sfem_sell_set_min_rows(1 if request.param == 'sell' else 1 << 30)
    yield request.param
    ctx.lib.sfem_staged_set_min_tiles(old)
    ctx.lib.sfem_sell_set_min_rows(old_sell)


def _rel(a, b):
    return np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(b), 1e-300)


def test_spmv_variants(ctx):
    import torch
    import scipy.sparse as sp
    from sulcusfem.device import DeviceCsr
    rng = np.random.default_rng(0)
    ctx.lib.sfem_sell_set_min_rows(1 << 30)          # this test: lane-group and staged engines only
    for n, dens in ((1, 1.0), (37, 0.3), (1000, 0.02), (20011, 0.0006)):
        A = sp.random(n, n, dens, random_state=1, format='csr') + sp.eye(n, format='csr')
        A.sort_indices()
        # an empty row and a long row
        if n > 100:
            A = A.tolil(); A[5, :] = 0; A[7, :200] = 1.5; A = A.tocsr(); A.eliminate_zeros(); A.sort_indices()
        dA = DeviceCsr(ctx, n, n, A.indptr, A.indices, A.data)
        x = rng.random(n); b = rng.random(n)
        dx = torch.from_numpy(x).cuda(); db = torch.from_numpy(b).cuda()
        ref = A @ x
        for min_tiles, staged in ((1 << 30, False), (1, True)):
            ctx.lib.sfem_staged_set_min_tiles(min_tiles)
            y = dA.spmv(dx, staged=staged)
            assert _rel(y.cpu().numpy(), ref) < 1e-14, (n, staged)
            r = dA.spmv(dx, b=db, mode=1, staged=staged)
            assert _rel(r.cpu().numpy(), b - ref) < 1e-13, (n, staged)
            y2 = torch.from_numpy(b.copy()).cuda()
            dA.spmv(dx, y=y2, mode=2, staged=staged)
            assert _rel(y2.cpu().numpy(), b + ref) < 1e-14
        ctx.lib.sfem_staged_set_min_tiles(0)
    # a row longer than one tile: no plan -> the vector engine serves the matrix
    n = 5000
    A = sp.random(n, n, 0.001, random_state=2, format='lil')
    A[11, :3000] = 0.25
    A = A.tocsr(); A.sort_indices()
    dA = DeviceCsr(ctx, n, n, A.indptr, A.indices, A.data)
    assert dA.tile_row is None
    x = rng.random(n)
    assert _rel(dA.spmv(torch.from_numpy(x).cuda()).cpu().numpy(), A @ x) < 1e-14
    ctx.lib.sfem_sell_set_min_rows(0)


def test_spmv_sell(ctx):
    """Sliced-ELL mirror: all modes, 1 and 2 right-hand sides, ragged / empty / long rows, rectangular operators,
    every sorting window; and the mirror follows the CSR values (dirty -> re-packed, never stale)."""
    import torch
    import scipy.sparse as sp
    from sulcusfem.device import DeviceCsr
    rng = np.random.default_rng(7)
    old = ctx.lib.sfem_sell_set_min_rows(1)
    sig0, min0, fill0 = DeviceCsr.SELL_SIGMA, DeviceCsr.SELL_MIN_ROWS, DeviceCsr.SELL_MAX_FILL
    try:
        DeviceCsr.SELL_MIN_ROWS = 1
        DeviceCsr.SELL_MAX_FILL = 64.0          # tiny random matrices pad a lot
        for (m, n, dens, sigma) in ((1, 1, 1.0, 256), (37, 37, 0.3, 32), (53, 31, 0.2, 64), (1000, 1000, 0.02, 256),
                                    (3000, 4100, 0.004, 128), (20011, 20011, 0.0006, 1024)):
            DeviceCsr.SELL_SIGMA = sigma
            A = sp.random(m, n, dens, random_state=3, format='csr')
            if m == n:
                A = A + sp.eye(n, format='csr')
            if m > 100:
                A = A.tolil(); A[5, :] = 0; A[7, :25] = 1.5; A = A.tocsr(); A.eliminate_zeros()
            A.sort_indices()
            dA = DeviceCsr(ctx, m, n, A.indptr, A.indices, A.data)
            assert dA.sell is not None, (m, n)
            for nb in (1, 2):
                X = rng.random((n, nb)); B = rng.random((m, nb))
                dX = torch.from_numpy(X.ravel().copy()).cuda(); dB = torch.from_numpy(B.ravel().copy()).cuda()
                ref = A @ X
                y = dA.spmv(dX, nb=nb, sell=True).cpu().numpy().reshape(m, nb)
                assert _rel(y, ref) < 1e-14, (m, n, nb)
                r = dA.spmv(dX, b=dB, mode=1, nb=nb, sell=True).cpu().numpy().reshape(m, nb)
                assert _rel(r, B - ref) < 1e-13, (m, n, nb)
                y2 = dB.clone()
                dA.spmv(dX, y=y2, mode=2, nb=nb, sell=True)
                assert _rel(y2.cpu().numpy().reshape(m, nb), B + ref) < 1e-14
                # the default entry point picks the same engine and gives the same bits
                y3 = dA.spmv(dX, nb=nb).cpu().numpy().reshape(m, nb)
                assert np.array_equal(y3, y)
            # values rewritten behind the library's back: mark_dirty -> the next product uses the new values
            dA.vals.mul_(2.0)
            dA.mark_dirty()
            x = rng.random(n)
            assert _rel(dA.spmv(torch.from_numpy(x).cuda(), sell=True).cpu().numpy(), 2.0 * (A @ x)) < 1e-14
            # values rewritten by a library entry (sfem_csr_extract): the mirror is refreshed without being told
            new = torch.from_numpy(rng.random(A.nnz)).cuda()
            slot = torch.arange(A.nnz, dtype=torch.int32, device='cuda')
            from sulcusfem import capi
            capi.check(ctx.lib.sfem_csr_extract(A.nnz, capi.ptr(slot), capi.ptr(new), capi.ptr(dA.vals), ctx.stream))
            A2 = sp.csr_matrix((new.cpu().numpy(), A.indices, A.indptr), shape=A.shape)
            assert _rel(dA.spmv(torch.from_numpy(x).cuda()).cpu().numpy(), A2 @ x) < 1e-14
        # a matrix with wildly irregular rows gets no mirror (padding would exceed SELL_MAX_FILL)
        DeviceCsr.SELL_SIGMA = 32
        DeviceCsr.SELL_MAX_FILL = fill0
        A = sp.random(5000, 5000, 0.001, random_state=2, format='lil')
        A[11, :3000] = 0.25
        A = A.tocsr(); A.sort_indices()
        dA = DeviceCsr(ctx, 5000, 5000, A.indptr, A.indices, A.data)
        assert dA.sell is None
        x = rng.random(5000)
        assert _rel(dA.spmv(torch.from_numpy(x).cuda()).cpu().numpy(), A @ x) < 1e-14
    finally:
        DeviceCsr.SELL_SIGMA, DeviceCsr.SELL_MIN_ROWS, DeviceCsr.SELL_MAX_FILL = sig0, min0, fill0
        ctx.lib.sfem_sell_set_min_rows(old)


def test_spmv_two_rhs_and_rectangular(ctx):
    """nb = 2 interleaved right-hand sides (matrix read once) and rectangular operators."""
    import torch
    import scipy.sparse as sp
    from sulcusfem.device import DeviceCsr
    rng = np.random.default_rng(2)
    for (m, n, dens) in ((1, 1, 1.0), (53, 31, 0.2), (3000, 4100, 0.004), (20011, 20011, 0.0005)):
        A = sp.random(m, n, dens, random_state=5, format='csr')
        if m > 100:
            A = A.tolil(); A[3, :] = 0; A[9, :300] = -0.5; A = A.tocsr(); A.eliminate_zeros()
        A.sort_indices()
        dA = DeviceCsr(ctx, m, n, A.indptr, A.indices, A.data)
        X = rng.random((n, 2)); B = rng.random((m, 2))
        dX = torch.from_numpy(X.ravel().copy()).cuda(); dB = torch.from_numpy(B.ravel().copy()).cuda()
        ref = A @ X
        for min_tiles, staged in ((1 << 30, False), (1, True)):
            ctx.lib.sfem_staged_set_min_tiles(min_tiles)
            y = dA.spmv(dX, nb=2, staged=staged).cpu().numpy().reshape(m, 2)
            assert _rel(y, ref) < 1e-14 or np.linalg.norm(ref) == 0
            r = dA.spmv(dX, b=dB, mode=1, nb=2, staged=staged).cpu().numpy().reshape(m, 2)
            assert _rel(r, B - ref) < 1e-13
            y2 = dB.clone()
            dA.spmv(dX, y=y2, mode=2, nb=2, staged=staged)
            assert _rel(y2.cpu().numpy().reshape(m, 2), B + ref) < 1e-14
        ctx.lib.sfem_staged_set_min_tiles(0)
        if m != n:
            y1 = dA.spmv(torch.from_numpy(X[:, 0].copy()).cuda(), nb=1).cpu().numpy()
            assert _rel(y1, ref[:, 0]) < 1e-14 or np.linalg.norm(ref) == 0


def test_vcycle_two_rhs_equals_two_cycles(ctx, small, engine):
    """One nb = 2 V-cycle = two scalar V-cycles (up to the summation order of rows reduced by a
    different number of lanes)."""
    import torch
    from sulcusfem.device import ScalarProblem
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    p1 = ScalarProblem(mesh, bm, dirichlet_ids=(1, 4, 3), robin_id=None, ctx=ctx, nb=1)
    p2 = ScalarProblem(mesh, bm, dirichlet_ids=(1, 4, 3), robin_id=None, hierarchy=p1.hierarchy, ctx=ctx, nb=2)
    p1.assemble(1.0, robin=False)
    p2.assemble(1.0, robin=False)
    torch.manual_seed(1)
    r = torch.rand(p1.n, 2, dtype=torch.float64, device=ctx.device)
    za = p1.mg.vcycle(r[:, 0].contiguous()).cpu().numpy()
    zb = p1.mg.vcycle(r[:, 1].contiguous()).cpu().numpy()
    z2 = p2.mg.vcycle(r.reshape(-1).contiguous()).cpu().numpy().reshape(-1, 2)
    assert _rel(z2[:, 0], za) < 1e-13 and _rel(z2[:, 1], zb) < 1e-13


def test_fused_tail_vcycle_matches_separate_launches(ctx, small):
    """sfem_mg_tail.cu: the small levels of a V-cycle in one thread-block-cluster kernel (phases separated by the
    cluster barrier) against the same cycle as separate launches -- whole hierarchy fused (level 0 has < 10 k rows on
    this mesh), fused from an inner level, one and two right-hand sides; then the Krylov solves on top of it."""
    import torch
    from sulcusfem.device import ScalarProblem, StokesProblem
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    lib = ctx.lib
    prob = ScalarProblem(mesh, bm, ctx=ctx)
    prob.assemble(1.0, mu_const=2.0, bc_values={1: 1.0, 2: 0.0})
    sp_ = StokesProblem(mesh, bm, ctx=ctx)
    sp_.vel.assemble(1.0, robin=False)
    gen = torch.Generator(device='cpu').manual_seed(7)
    old = lib.sfem_mg_set_tail_rows(0)
    try:
        for mg, nb in ((prob.mg, 1), (sp_.vel.mg, 2)):
            n = mg.levels[0].n
            sizes = [l.n for l in mg.levels]
            b = torch.randn(n * nb, generator=gen, dtype=torch.float64).to(ctx.device)
            lib.sfem_mg_set_tail_rows(0)
            ref = mg.vcycle(b).clone()
            for rows in (1 << 20, sizes[1], sizes[-2]):               # everything / from level 1 / last sparse level + dense
                lib.sfem_mg_set_tail_rows(int(rows))
                x = mg.vcycle(b).clone()
                err = float((x - ref).norm() / ref.norm())
                print('fused tail', nb, sizes, rows, err)
                assert err < 1e-13, (nb, rows, err)
                x2 = mg.vcycle(b)
                assert torch.equal(x2, x)                              # run-to-run reproducible
        # solves with the fused tail: same iteration counts (+-1) and fields as with separate launches
        res = {}
        for rows in (0, 1 << 20):
            lib.sfem_mg_set_tail_rows(rows)
            prob.assemble(1.0, mu_const=2.0, bc_values={1: 1.0, 2: 0.0})
            c = prob.solve('cg', rtol=1e-13).clone()
            res[rows] = (c, prob.last_info['iterations'])
            assert prob.last_info['converged']
        assert abs(res[0][1] - res[1 << 20][1]) <= 1
        assert float((res[0][0] - res[1 << 20][0]).norm() / res[0][0].norm()) < 1e-11
    finally:
        lib.sfem_mg_set_tail_rows(old)


def test_p2_assembly_matches_oracle(ctx, small):
    import torch
    from oracle import cpu_oracle as co
    from sulcusfem.device import ScalarLevel
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    lev = ScalarLevel(ctx, mesh, bm, 2, (1, 2), 4)
    # bit-exact pattern vs the oracle's clique pattern
    ip, ix = co.clique_pattern(om.p2_cell_dofs(), om.n_p2)
    assert np.array_equal(ip, lev.pattern.rowptr) and np.array_equal(ix, lev.pattern.cols)
    X = om.p2_dof_coords()
    ux = 4 * np.clip(X[:, 1], 0, 1) * (1 - np.clip(X[:, 1], 0, 1)) + 0.1 * X[:, 0]
    uy = 0.3 * np.sin(X[:, 0]) * X[:, 1]
    f4 = np.flatnonzero((bm == 4) & om.on_boundary)
    K = co.assemble_p2_stiffness(om)
    Cm = co.assemble_p2_advection(om, ux, uy)
    mun = 0.5 + np.cos(3 * X[:, 0])          # changes sign -> exercises the clamp
    cases = [
        (dict(D=1.0, mu_const=1.0), 1.0 * K + co.assemble_p2_robin(om, f4, mu_const=1.0)),
        (dict(D=0.025, ux=ux, uy=uy, mu_const=0.3), 0.025 * K + Cm + co.assemble_p2_robin(om, f4, mu_const=0.3)),
        (dict(D=0.1, ux=ux, uy=uy, mu_nodal=mun), 0.1 * K + Cm + co.assemble_p2_robin(om, f4, mu_nodal=mun)),
        (dict(D=1.0, mu_nodal=mun, clamp=True), K + co.assemble_p2_robin(om, f4, mu_nodal=mun, clamp=True)),
    ]
    for kw, ref in cases:
        kw = dict(kw)
        for k in ('ux', 'uy', 'mu_nodal'):
            if k in kw:
                kw[k] = torch.from_numpy(np.ascontiguousarray(kw[k])).cuda()
        lev.assemble(**kw)
        got = lev.A.to_scipy()
        d = abs(got - ref)
        assert d.max() <= 1e-13 * abs(ref).max(), kw.keys()


def test_stokes_assembly_and_dirichlet(ctx, small):
    import torch
    from oracle import cpu_oracle as co
    from sulcusfem.device import StokesProblem
    from sulcusfem import dofmap as dm
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    sp_ = StokesProblem(mesh, bm, ctx=ctx)
    ip, ix = co.clique_pattern(om.th_cell_dofs(), 2 * om.n_p2 + om.nv)
    assert np.array_equal(ip, sp_.pattern.rowptr) and np.array_equal(ix, sp_.pattern.cols)
    X = dm.p2_dof_coordinates(mesh)
    d1 = dm.dirichlet_dofs_p2(mesh, bm, 1)
    zeros = lambda i: (0.0, 0.0)
    sp_.set_bcs({1: (4 * X[d1, 1] * (1 - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})
    sp_.assemble(bc_mode=0)
    A0 = co.assemble_stokes(om)
    dofs, vals = co.stokes_bcs(om, bm, 1.0)
    assert np.array_equal(np.flatnonzero(sp_.bc_flag_host), dofs)
    Aref, bref = co.apply_dirichlet_rows(A0, np.zeros(A0.shape[0]), dofs, vals)
    got = sp_.A.to_scipy()
    assert abs(got - Aref).max() <= 1e-13 * abs(Aref).max()
    assert np.abs(sp_.rhs.cpu().numpy() - bref).max() < 1e-15


def test_device_sorted_patterns_are_bit_identical_to_the_host_path(ctx):
    """dofmap.build_pattern sorts large key lists on the device (set-up wall of refined meshes); same stable sort of the
    same keys -> every array of the pattern / gather map must equal the numpy path bit for bit."""
    from sulcusfem import dofmap as dm, hostmesh as hm
    mesh = hm.refine_n(hm.rectangle_mesh(10.0, 1.0, 60, 12), 1)
    cd = dm.p2_cell_dofs(mesh)
    n = dm.p2_num_dofs(mesh)
    c1 = dm.p1_cell_dofs(mesh)
    il = np.concatenate([2 * cd.astype(np.int64), 2 * cd.astype(np.int64) + 1], axis=1)
    for nrows, ncols, fam in ((n, n, [(cd, cd), (cd[:70, :3], cd[:70, :3])]), (mesh.num_vertices, 2 * n, [(c1, il)]),
                              (2 * n, mesh.num_vertices, [(il, c1)])):
        saved, dm.DEVICE_SORT = dm.DEVICE_SORT, None
        try:
            a = dm.build_pattern(nrows, ncols, fam)
        finally:
            dm.DEVICE_SORT = saved
        b = dm._build_pattern_device(nrows, ncols, fam, ctx.device)
        for k in ('rowptr', 'cols', 'contrib_ptr', 'contrib_code'):
            x, y = getattr(a, k), getattr(b, k)
            assert x.dtype == y.dtype and np.array_equal(x, y), k
        assert a.family_base == b.family_base and a.buffer_len == b.buffer_len


def test_device_sell_plan_is_bit_identical_to_the_host_plan(ctx):
    import torch
    from sulcusfem import dofmap as dm, hostmesh as hm, sell as sl
    mesh = hm.refine_n(hm.rectangle_mesh(10.0, 1.0, 50, 9), 1)
    cd = dm.p2_cell_dofs(mesh)
    n = dm.p2_num_dofs(mesh)
    saved, dm.DEVICE_SORT = dm.DEVICE_SORT, None
    try:
        pat = dm.build_pattern(n, n, [(cd, cd)])
    finally:
        dm.DEVICE_SORT = saved
    for sigma in (32, 256):
        a = sl.build_plan(pat.rowptr, pat.cols, sigma)
        b = sl.build_plan_device(torch.from_numpy(pat.rowptr).to(ctx.device), torch.from_numpy(pat.cols).to(ctx.device), sigma)
        assert (a.nrows, a.nslices, a.padded) == (b['nrows'], b['nslices'], b['padded'])
        for k in ('slice_ptr', 'perm', 'scols', 'src'):
            x, y = getattr(a, k), b[k].cpu().numpy()
            assert x.dtype == y.dtype and np.array_equal(x, y), k


def test_dense_inverse(ctx):
    """Blocked (32 x 32) multi-CTA Gauss-Jordan: sizes below / at / across block boundaries, called back to back on one
    stream (the workspace is shared) and with a growing workspace."""
    import torch
    import scipy.sparse as sp
    from sulcusfem import capi
    from sulcusfem.device import DeviceCsr, P
    for n in (1, 5, 32, 33, 150, 64, 449, 1000):
        A = (sp.random(n, n, min(1.0, 8.0 / n), random_state=3 + n) + sp.eye(n) * 3).tocsr()
        A.sort_indices()
        dA = DeviceCsr(ctx, n, n, A.indptr, A.indices, A.data, staged=False, sell=False)
        out = ctx.zeros(n * n)
        capi.check(ctx.lib.sfem_dense_inverse_csr(n, P(dA.rowptr), P(dA.cols), P(dA.vals), P(out), ctx.stream))
        inv = out.cpu().numpy().reshape(n, n)
        assert np.abs(inv @ A.toarray() - np.eye(n)).max() < 1e-11, n


def test_diffusion_solve_matches_lu(ctx, small, engine):
    from oracle import cpu_oracle as co
    from sulcusfem.device import ScalarProblem
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    prob = ScalarProblem(mesh, bm, ctx=ctx)
    for mu in (0.1, 1.0, 10.0):
        prob.assemble(1.0, mu_const=mu, bc_values={1: 1.0, 2: 0.0})
        c = prob.solve('cg', rtol=1e-13).cpu().numpy()
        ref, _, _ = co.solve_concentration(om, bm, 1.0, mu=mu)
        print('cg', mu, prob.last_info, prob.mg.lambda_max())
        assert prob.last_info['converged']
        assert prob.last_info['iterations'] < 60
        assert _rel(c, ref) < 1e-10      # north_star: fields within 1e-10 relative L2


@pytest.mark.parametrize('mus', [[0.05, 0.3, 1.0, 4.0, 25.0, 150.0, 600.0, 0.7], [2.0], [0.1, 10.0, 1.0],
                                 [0.2, 0.4, 0.8, 1.6, 3.2], [5.0, 5.0, 0.0, 40.0, 7.0, 9.0, 11.0],
                                 [0.3 * 1.25 ** k for k in range(16)]])
def test_batched_robin_sweep_matches_lu(ctx, small, mus):
    """sfem_krylov_cg_batch (the mu sweep of one geometry in one Krylov loop, no_advection_analysis_A.py:
    1306-1347): every column against the oracle's sparse LU of its own A(mu) -- batch sizes 1 .. 16 including the
    non-power-of-two block shapes, coefficients over four decades behind ONE multigrid hierarchy, a repeated and a zero
    coefficient."""
    from oracle import cpu_oracle as co
    from sulcusfem.device import ScalarProblem
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    prob = ScalarProblem(mesh, bm, ctx=ctx)
    nb = len(mus)
    Xd, infos = prob.solve_batch(1.0, mus, {1: 1.0, 2: 0.0}, rtol=1e-13)
    X = Xd.cpu().numpy().reshape(-1, nb)
    print('cg_batch', nb, [(i['iterations'], i['relres']) for i in infos])
    for c, mu in enumerate(mus):
        ref, _, _ = co.solve_concentration(om, bm, 1.0, mu=mu)
        assert infos[c]['converged'], infos[c]
        assert infos[c]['iterations'] < (40 if max(mus) <= 64 * min(m for m in mus if m > 0) else 400), infos[c]
        assert _rel(X[:, c], ref) < 1e-10, (mu, _rel(X[:, c], ref))
        assert np.array_equal(prob.batch_column(Xd, nb, c).cpu().numpy(), X[:, c])
    # the same problem object still serves single solves (the batch leaves a fully assembled hierarchy behind)
    prob.assemble(1.0, mu_const=mus[0], bc_values={1: 1.0, 2: 0.0})
    c1 = prob.solve('cg', rtol=1e-13).cpu().numpy()
    assert _rel(c1, X[:, 0]) < 1e-10
    if nb == 5:       # sharing the reference coefficient's coarse levels between the columns converges to the same fields
        Xs, infos_s = prob.solve_batch(1.0, mus, {1: 1.0, 2: 0.0}, rtol=1e-13, shared_coarse=True)
        assert all(i['converged'] for i in infos_s) and infos_s[0]['iterations'] >= infos[0]['iterations']
        assert _rel(Xs.cpu().numpy().reshape(-1, nb), X) < 1e-10
    # run-to-run reproducibility of the batched path (fixed summation orders)
    Xd2, _ = prob.solve_batch(1.0, mus, {1: 1.0, 2: 0.0}, rtol=1e-13)
    assert Xd2.data_ptr() == Xd.data_ptr()                 # the result buffer of a width is reused (stable graph key)
    assert np.array_equal(Xd2.cpu().numpy().reshape(-1, nb), X)


def test_batched_sweep_through_the_solver_entry_points(ctx, small):
    """pure_diffusion_solver_batch / presolve_pure_diffusion: Functions equal to the per-case pure_diffusion_solver
    (reference solvers.py:113-174) to solver tolerance, parked fields are handed out exactly once."""
    from sulcusfem import solvers
    from sulcusfem.fem import Constant, FunctionSpace
    mesh, mk, om = small
    mr = {'mesh': mesh}
    mr.update(mk)
    C = FunctionSpace(mesh, "CG", 2)
    mus = [30.0, 0.1, 1.0, 3.0, 0.3, 10.0, 100.0, 0.03, 300.0, 2.0]          # 10 coefficients: two batches
    single = [solvers.pure_diffusion_solver(mr, C, Constant(0.7), Constant(mu)) for mu in mus]
    batch = solvers.pure_diffusion_solver_batch(mr, C, Constant(0.7), mus)
    for f1, fb, mu in zip(single, batch, mus):
        assert _rel(fb.values, f1.values) < 1e-10, mu
        assert fb.solver_info['method'] == 'cg_batch' and fb.solver_info['converged']
        assert np.array_equal(fb.device_components()[0].cpu().numpy(), fb.values)
    assert solvers.presolve_pure_diffusion(mr, C, Constant(0.7), mus[:3]) == 3
    f = solvers.pure_diffusion_solver(mr, C, Constant(0.7), Constant(mus[1]))
    assert f.solver_info['method'] == 'cg_batch' and _rel(f.values, single[1].values) < 1e-10
    f = solvers.pure_diffusion_solver(mr, C, Constant(0.7), Constant(mus[1]))          # parked field already taken
    assert f.solver_info['method'] == 'cg'


def test_advdiff_solve_matches_lu(ctx, small, engine):
    import torch
    from oracle import cpu_oracle as co
    from sulcusfem.device import ScalarProblem
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    X = om.p2_dof_coords()
    y = np.clip(X[:, 1], 0, 1)
    ux, uy = 4 * y * (1 - y), np.zeros(len(y))
    prob = ScalarProblem(mesh, bm, ctx=ctx)
    dux, duy = torch.from_numpy(ux).cuda(), torch.from_numpy(uy).cuda()
    for D in (0.025, 1.0):
        prob.assemble(D, dux, duy, mu_const=1.0, bc_values={1: 1.0, 2: 0.0})
        c = prob.solve('fgmres', rtol=1e-13).cpu().numpy()
        ref, _, _ = co.solve_concentration(om, bm, D, mu=1.0, ux=ux, uy=uy)
        print('fgmres', D, prob.last_info)
        assert prob.last_info['converged']
        assert _rel(c, ref) < 1e-10


def test_stokes_solve_matches_lu(ctx, small, engine):
    from oracle import cpu_oracle as co
    from sulcusfem.device import StokesProblem
    from sulcusfem import dofmap as dm
    mesh, mk, om = small
    bm = mk['bc_markers'].values
    sp_ = StokesProblem(mesh, bm, ctx=ctx)
    X = dm.p2_dof_coordinates(mesh)
    d1 = dm.dirichlet_dofs_p2(mesh, bm, 1)
    sp_.set_bcs({1: (4 * X[d1, 1] * (1 - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})
    sp_.assemble(bc_mode=1)
    ux, uy, p = [t.cpu().numpy() for t in sp_.solve(rtol=1e-14)]
    rx, ry, rp, _, _ = co.solve_stokes(om, bm, 1.0)
    print('minres', sp_.last_info)
    assert sp_.last_info['converged']
    un = np.linalg.norm(np.concatenate([rx, ry]))
    assert np.linalg.norm(np.concatenate([ux - rx, uy - ry])) / un < 1e-10
    assert _rel(p, rp) < 1e-9
    # the solver's block views were assembled DIRECTLY (72 element entries per cell); the full path -- dolfin's
    # mixed-space matrix, Dirichlet applied there, blocks extracted -- gives the same numbers (to the last bit or two)
    import scipy.sparse as sps
    Kd, Bd, BTd = sp_.vel.fine.A.vals.clone(), sp_.B.vals.clone(), sp_.BT.vals.clone()
    rhs_d, rhs_il_d = sp_.rhs.clone(), sp_.rhs_il.clone()
    assert sp_._full is None                      # the mixed-space pattern was never needed so far
    sp_.assemble(bc_mode=1, full=True)
    dK = float((sp_.vel.fine.A.vals - Kd).abs().max()) / float(Kd.abs().max())
    dB = float((sp_.B.vals - Bd).abs().max()) / float(Bd.abs().max())
    dBT = float((sp_.BT.vals - BTd).abs().max()) / float(BTd.abs().max())
    print('direct vs extracted blocks, relative max difference: K', dK, 'B', dB, 'BT', dBT)
    # same element arithmetic (explicitly rounded helpers) and same gather order: identical up to the multiply-add
    # contraction choices the compiler makes per kernel -- 0 or 1 ulp
    assert dK <= 4e-16 and dB <= 4e-16 and dBT <= 4e-16
    scale = float(sp_.rhs.abs().max())
    assert float((sp_.rhs - rhs_d).abs().max()) <= 1e-14 * scale       # lifting sums run in a different order
    assert float((sp_.rhs_il - rhs_il_d).abs().max()) <= 1e-14 * scale
    # block views = the assembled matrix minus structural zeros: K, B, B^T reproduce A exactly
    A = sp_.A.to_scipy()
    n2, nv = sp_.n2, sp_.nv
    K = sp_.vel.fine.A.to_scipy()
    assert abs(K - A[:n2, :n2]).max() == 0.0 and abs(K - A[n2:2 * n2, n2:2 * n2]).max() == 0.0
    perm = np.concatenate([2 * np.arange(n2), 2 * np.arange(n2) + 1])        # blocked -> interleaved position
    B = sp_.B.to_scipy()
    assert abs(B[:, perm] - A[2 * n2:, :2 * n2]).max() == 0.0
    assert abs(sp_.BT.to_scipy() - B.T).max() == 0.0
    # the graph-replayed iteration and the plain launch sequence give the same iterates (both on the system as it
    # stands now, i.e. assembled through the full path: its right-hand side differs from the direct one in the last bits)
    sp_.solve(rtol=1e-14)
    it_graph = sp_.last_info['iterations']
    import os
    x_graph = sp_.x.clone()
    ctx.lib.sfem_profile_start(1)          # profiling on -> graphs bypassed
    sp_.solve(rtol=1e-14)
    ctx.lib.sfem_profile_stop(0, None, None, None)
    assert sp_.last_info['iterations'] == it_graph
    assert float((sp_.x - x_graph).abs().max()) == 0.0
    # analytic channel flow as the starting vector, stopping level anchored to the plain guess (sfem_stokes_solve_from):
    # same fields against the LU, fewer iterations
    it_plain = sp_.last_info['iterations']
    sp_.set_channel_flow_guess(10.0, 1.0)
    sp_.assemble(bc_mode=1)
    uxg, uyg, pg = [t.cpu().numpy() for t in sp_.solve(rtol=1e-14)]
    print('minres from the channel-flow guess', sp_.last_info, 'plain guess:', it_plain)
    assert sp_.last_info['converged'] and sp_.last_info['iterations'] < it_plain
    assert np.linalg.norm(np.concatenate([uxg - rx, uyg - ry])) / un < 1e-10
    assert _rel(pg, rp) < 1e-9
    sp_.set_bcs({1: (4 * X[d1, 1] * (1 - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})     # drops the guess again
    assert sp_.x0_il is None
    # without the lubrication coarse correction: same solution, more iterations
    sp0 = StokesProblem(mesh, bm, ctx=ctx, schur_correction=False)
    sp0.set_bcs({1: (4 * X[d1, 1] * (1 - X[d1, 1]), 0.0), 4: (0.0, 0.0), 3: (0.0, 0.0)})
    sp0.assemble(bc_mode=1)
    ux0, uy0, p0 = [t.cpu().numpy() for t in sp0.solve(rtol=1e-14)]
    print('minres (mass-matrix Schur only)', sp0.last_info)
    assert np.linalg.norm(np.concatenate([ux0 - rx, uy0 - ry])) / un < 1e-10
    assert sp0.last_info['iterations'] > it_plain


def test_functionals_match_oracle(ctx, small):
    import torch
    from oracle import cpu_oracle as co
    from sulcusfem.device import FunctionalPlan
    mesh, mk, om = small
    rng = np.random.default_rng(1)
    X = om.p2_dof_coords()
    c = 1.0 - X[:, 0] / 10 + 0.1 * np.sin(5 * X[:, 0]) * np.cos(3 * X[:, 1])
    ux = 4 * np.clip(X[:, 1], 0, 1) * (1 - np.clip(X[:, 1], 0, 1))
    uy = 0.2 * np.sin(X[:, 0] * 7) * (X[:, 1] - 0.3)
    D, mu = 0.025, 0.7
    omk = {k: v.values for k, v in mk.items()}
    ref = co.flux_metrics(om, omk, 'sulcus', D, c, ux, uy, mu=mu)
    refm = co.mass_metrics(om, c, 'sulcus', omk['domain_markers'])
    plan = FunctionalPlan(mesh, mk, 'sulcus', ctx=ctx)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    F, Cc = plan.evaluate(t(c), t(ux), t(uy), D=D, mu_const=mu)
    names = plan.GROUPS

    def close(a, b, tol=1e-11):
        assert abs(a - b) <= tol * max(1.0, abs(b)), (a, b)
    for i, nm in enumerate(names[:4]):
        close(F[i, 0], ref['physical_flux'][nm]['diffusive'])
        close(F[i, 1], ref['physical_flux'][nm]['advective'])
    close(F[3, 2], ref['uptake_flux'])
    seg = ref['sulcus_specific']['physical_flux']
    for i, nm in ((4, 'bottom_left'), (5, 'sulcus'), (6, 'bottom_right')):
        close(F[i, 0], seg[nm]['diffusive'])
        close(F[i, 1], seg[nm]['advective'])
        close(F[i, 2], ref['sulcus_specific']['uptake_flux'][nm])
    close(F[8, 0], seg['sulcus_opening']['diffusive'])
    close(F[8, 1], seg['sulcus_opening']['advective'])
    ex = seg['sulcus_opening_extra']
    close(F[8, 5], ex['E_L1']); close(F[8, 6], ex['Q_in']); close(F[8, 7], ex['Q_out']); close(F[8, 4], ex['length'])
    close(F[7, 3], ref['_conc']['C_y0_ext']); close(F[8, 3], ref['_conc']['C_mouth'])
    close(Cc[1, 0], refm['sulcus_mass']); close(Cc[2, 0], refm['rectangle_mass'])
    close(Cc[1, 1], refm['sulcus_area']); close(Cc[2, 1], refm['rectangle_area'])


def test_analysis_functions_are_pure_in_their_arguments(ctx, small):
    """compute_uptake_flux_* take (c, measures, mu_val) like analysis.py:307-333: the result must not depend on which
    evaluation ran last on ``c``, and editing ``c`` through ``vector()`` must not serve stale functionals."""
    from oracle import cpu_oracle as co
    from sulcusfem import analysis
    from sulcusfem.fem import Function, FunctionSpace
    mesh, mk, om = small
    X = om.p2_dof_coords()
    cv = 1.0 - X[:, 0] / 10 + 0.1 * np.sin(5 * X[:, 0]) * np.cos(3 * X[:, 1])
    omk = {k: v.values for k, v in mk.items()}
    mr = {'mesh': mesh}
    mr.update(mk)
    c = Function(FunctionSpace(mesh, 'CG', 2), cv.copy())
    ref = co.flux_metrics(om, omk, 'sulcus', 1.0, cv, mu=0.7)
    # no prior evaluation at all
    assert abs(analysis.compute_uptake_flux_bottom(c, {}, 0.7) - ref['uptake_flux']) < 1e-11
    # an evaluation with mu = 0 in between (physical fluxes) must not leak into the uptake integrals
    analysis.compute_physical_flux_boundary(c, None, mr, {}, 4, 1.0)
    assert abs(analysis.compute_uptake_flux_bottom(c, {}, 0.7) - ref['uptake_flux']) < 1e-11
    seg = analysis.compute_uptake_flux_segments(c, {}, 0.7)
    for k in ('bottom_left', 'sulcus', 'bottom_right'):
        assert abs(seg[k] - ref['sulcus_specific']['uptake_flux'][k]) < 1e-11
    assert abs(analysis.compute_uptake_flux_bottom(c, {}, 1.4) - 2 * ref['uptake_flux']) < 1e-11
    m0 = analysis.compute_mass_metrics(c, {}, 'sulcus')['total_mass']
    # edit the field: everything must follow
    c.vector().set_local(2.0 * cv)
    assert abs(analysis.compute_uptake_flux_bottom(c, {}, 0.7) - 2 * ref['uptake_flux']) < 1e-11
    assert abs(analysis.compute_mass_metrics(c, {}, 'sulcus')['total_mass'] - 2 * m0) < 1e-11


def test_mailbox_halo_exchange_and_vector_allreduce_emulated(ctx):
    """Two 'ranks' inside one process on one GPU: the send halves of all ranks run first, then the
    wait + unpack halves (the production kernels do both in one launch, one rank per GPU)."""
    import torch
    import ctypes as C
    from sulcusfem import hostmesh as hm, dofmap as dm, partition as pt, capi
    from sulcusfem.device import P
    from sulcusfem.dist import DistContext, DeviceHalo
    lib = ctx.lib
    mesh = hm.rectangle_mesh(10.0, 1.0, 30, 6)
    cd = dm.p2_cell_dofs(mesh); n = dm.p2_num_dofs(mesh)
    pat = dm.build_pattern(n, n, [(cd, cd)])
    X = dm.p2_dof_coordinates(mesh)
    R = 3
    owner = pt.slab_owner(X[:, 0], R, X[:, 1])
    gh = pt.ghost_sets(owner, R, [(pat.rowptr, pat.cols, owner)])
    vec_cap = 500
    header = int(lib.sfem_dist_header_words(R, vec_cap))
    parts = [pt.partition_level(owner, R, r, gh, [header] * R, nb_max=2) for r in range(R)]
    words = max(p.mailbox_end for p in parts) + 8
    dists = [DistContext(ctx, r, R, words, vec_cap, emulate=True) for r in range(R)]
    try:
        for r in range(R):
            for q in range(R):
                if q != r:
                    capi.check(lib.sfem_dist_set_peer_pointer(dists[r].handle, q, lib.sfem_dist_mailbox(dists[q].handle)))
        halos = [DeviceHalo(ctx, p) for p in parts]
        rng = np.random.default_rng(0)
        for nb in (1, 2):
            for trial in range(3):                       # several rounds: both slot parities get reused
                xg = rng.random((n, nb))
                xs = []
                for p in parts:
                    x = torch.zeros(p.n_loc * nb, dtype=torch.float64, device=ctx.device)
                    x[:p.n_own * nb] = torch.from_numpy(xg[p.owned].ravel()).to(ctx.device)
                    xs.append(x)
                for phase in (1, 2):
                    for r in range(R):
                        dists[r].activate()
                        halos[r].exchange(xs[r], nb=nb, phase=phase)
                torch.cuda.synchronize()
                for p, x in zip(parts, xs):
                    assert np.array_equal(x.cpu().numpy().reshape(-1, nb)[p.n_own:], xg[p.ghost]), (nb, trial, p.rank)
        # vector all-reduce
        for trial in range(3):
            vs = [torch.from_numpy(rng.random(vec_cap - 7)).to(ctx.device) for _ in range(R)]
            want = sum(v.cpu().numpy() for v in vs)
            for phase in (1, 2):
                for r in range(R):
                    capi.check(lib.sfem_dist_allreduce_vec(dists[r].handle, P(vs[r]), vec_cap - 7, phase, ctx.stream))
            torch.cuda.synchronize()
            for v in vs:
                assert np.allclose(v.cpu().numpy(), want, rtol=1e-15)
                assert np.array_equal(v.cpu().numpy(), vs[0].cpu().numpy())       # bit-identical on every rank
        # production vector all-reduce (phase 0): reduce-scatter + all-gather over flagged words in ONE launch per rank;
        # the emulated ranks run concurrently on their own streams (every kernel polls for its peers' words)
        streams = [torch.cuda.Stream() for _ in range(R)]
        torch.cuda.synchronize()
        for nvals in (vec_cap - 7, 1, R, 2 * R + 1):
            for trial in range(3):                       # both parities of the slots get reused
                vs = [torch.from_numpy(rng.random(nvals)).to(ctx.device) for _ in range(R)]
                want = vs[0].cpu().numpy().copy()
                for v in vs[1:]:
                    want = want + v.cpu().numpy()        # rank order, like the kernel
                torch.cuda.synchronize()
                for r in range(R):
                    with torch.cuda.stream(streams[r]):
                        capi.check(lib.sfem_dist_allreduce_vec(dists[r].handle, P(vs[r]), nvals, 0, ctx.stream))
                torch.cuda.synchronize()
                for v in vs:
                    assert np.array_equal(v.cpu().numpy(), want), (nvals, trial)      # bit-identical on every rank
        assert not any(d.error() for d in dists)
        for h in halos:
            lib.sfem_halo_destroy(h.handle)
    finally:
        for d in dists:
            d.close()


def test_point_evaluation_and_line_profiles_match_oracle(ctx, small):
    """sfem_eval_points (uniform-grid locator + P1/P2 evaluation, one launch) against the oracle's brute-force
    location: same inside/outside decisions, values to 1e-13; then the reference-shaped profile extractors,
    compute_conc_profiles and compute_velocity_metrics (analysis.py:341-632, 721-830) against the oracle."""
    from oracle import cpu_oracle as co
    from sulcusfem import analysis, fem
    from sulcusfem.parameters import Parameters
    mesh, mk, om = small
    X = om.p2_dof_coords()
    rng = np.random.default_rng(3)
    c_vals = np.sin(0.7 * X[:, 0]) * np.cos(1.3 * X[:, 1]) + 0.1 * X[:, 0]
    ux_vals = 4 * X[:, 1] * (1 - X[:, 1]) + 0.05 * np.sin(X[:, 0])
    uy_vals = 0.02 * np.cos(2 * X[:, 0]) * X[:, 1]
    c = fem.Function(fem.FunctionSpace(mesh, 'CG', 2), c_vals)
    u = fem.Function(fem.VectorFunctionSpace(mesh, 'CG', 2), np.concatenate([ux_vals, uy_vals]))
    p1 = fem.Function(fem.FunctionSpace(mesh, 'CG', 1), c_vals[:mesh.num_vertices])
    pts = np.concatenate([
        np.stack([rng.uniform(-0.2, 10.2, 3000), rng.uniform(-1.1, 1.1, 3000)], axis=1),
        mesh.coords[rng.integers(0, mesh.num_vertices, 200)],
        mesh.edge_midpoints()[rng.integers(0, mesh.num_edges, 200)],
        np.array([[0.0, 0.0], [10.0, 1.0], [0.0, 1.0], [10.0, 0.0], [5.0, -1.0], [4.75, 0.0], [5.25, 0.0]]),
        np.zeros((0, 2))])
    ref_c, ok = co.eval_points(om, c_vals, pts)
    got_c, inside = c.eval_points(pts)
    assert np.array_equal(inside, ok)
    assert np.abs(got_c - ref_c).max() < 1e-13
    got_u, inside_u = u.eval_points(pts)
    assert np.array_equal(inside_u, ok)
    assert np.abs(got_u[:, 0] - co.eval_points(om, ux_vals, pts)[0]).max() < 1e-13
    assert np.abs(got_u[:, 1] - co.eval_points(om, uy_vals, pts)[0]).max() < 1e-13
    got_p1, _ = p1.eval_points(pts)
    assert np.abs(got_p1 - co.eval_points(om, c_vals[:mesh.num_vertices], pts, degree=1)[0]).max() < 1e-13
    e, ins = c.eval_points(np.zeros((0, 2)))
    assert e.shape == (0,) and ins.shape == (0,)
    # reference-shaped extractors
    prof = analysis.extract_concentration_vertical_line_profile(c, mesh, 5.0, n_points=101)
    s, v = co.line_profile(om, c_vals, 5.0, 'v', None, 101)
    assert np.array_equal(prof['y_coords'], s) and np.abs(prof['c'] - v).max() < 1e-13
    prof = analysis.extract_velocity_horizontal_line_profile(u, mesh, 0.25, x_range=(0, 10.0), n_points=100)
    s, a = co.line_profile(om, ux_vals, 0.25, 'h', (0, 10.0), 100)
    assert np.array_equal(prof['x_coords'], s) and np.abs(prof['u_x'] - a).max() < 1e-13
    params = Parameters(mode='adv-diff', mesh_size_dim=0.08, sulci_w_dim=0.5, sulci_h_dim=1.0)
    params.validate(); params.nondim()
    res = analysis.compute_conc_profiles({'c': c, 'mesh_results': {'mesh': mesh}, 'params': params}, n_points=400)
    want = co.conc_profiles(om, c_vals, float(params.L_dim), float(params.H_dim), 'sulcus', 400)
    got = res['mass_metrics']['profiles']
    assert set(got['horizontal']) == set(want['horizontal']) and set(got['vertical']) == set(want['vertical'])
    for grp in ('horizontal', 'vertical'):
        for name, w in want[grp].items():
            g = got[grp][name]
            assert g['n_samples'] == w['n_samples']
            for k in ('min_c', 'max_c', 'avg_c'):
                assert abs(g[k] - w[k]) < 1e-12, (grp, name, k)
    assert len(res['mass_metrics']['profiles_full']['vertical']['x_mid']['c']) == got['vertical']['x_mid']['n_samples']
    vm = analysis.compute_velocity_metrics(u, {'mesh': mesh}, params, rng=np.random.default_rng(1))
    wantv = co.velocity_line_metrics(om, ux_vals, uy_vals, params.L, params.H, params.sulci_w, 100)
    for k, w in wantv.items():
        assert abs(vm[k] - w) < 1e-12, k
    idx = np.random.default_rng(1).choice(mesh.num_vertices, min(1000, mesh.num_vertices), replace=False)
    assert abs(vm['global_max_umag'] - np.sqrt(ux_vals[idx] ** 2 + uy_vals[idx] ** 2).max()) < 1e-14
    params.mode = 'no-adv'
    assert analysis.compute_velocity_metrics(u, {'mesh': mesh}, params) == {}
