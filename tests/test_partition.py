"""Host-side domain decomposition: partition plans, local operators, and a world_size-2 gloo
run that moves the halos with torch.distributed exactly as the device mailbox plan prescribes."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _problem(nx=40, ny=8):
    import scipy.sparse as sp
    from sulcusfem import hostmesh as hm, dofmap as dm
    mesh = hm.rectangle_mesh(10.0, 1.0, nx, ny)
    cd = dm.p2_cell_dofs(mesh)
    n = dm.p2_num_dofs(mesh)
    pat = dm.build_pattern(n, n, [(cd, cd)])
    X = dm.p2_dof_coordinates(mesh)
    rng = np.random.default_rng(0)
    vals = rng.random(pat.nnz)
    A = sp.csr_matrix((vals, pat.cols, pat.rowptr), shape=(n, n))
    return mesh, pat, X, vals, A


@pytest.mark.parametrize('nranks', [2, 3, 4, 8])
def test_partition_plan_and_local_spmv(nranks):
    import scipy.sparse as sp
    from sulcusfem import partition as pt
    mesh, pat, X, vals, A = _problem()
    n = A.shape[0]
    owner = pt.slab_owner(X[:, 0], nranks, X[:, 1])
    counts = np.bincount(owner, minlength=nranks)
    assert counts.max() - counts.min() <= 1                     # balanced
    gh = pt.ghost_sets(owner, nranks, [(pat.rowptr, pat.cols, owner)])
    parts = [pt.partition_level(owner, nranks, r, gh, [17] * nranks, nb_max=2) for r in range(nranks)]
    rng = np.random.default_rng(1)
    xg = rng.random((n, 2))
    yg = A @ xg
    xs = []
    for lp in parts:
        assert np.array_equal(np.sort(np.concatenate([lp.owned, lp.ghost])), np.flatnonzero(lp.g2l >= 0))
        x = np.zeros((lp.n_loc, 2))
        x[:lp.n_own] = xg[lp.owned]
        xs.append(x.reshape(-1))
    pt.emulate_exchange(parts, xs, nb=2)
    for lp, x in zip(parts, xs):
        x = x.reshape(-1, 2)
        assert np.array_equal(x[lp.n_own:], xg[lp.ghost])
        rp, c, slot = pt.localize_csr(pat.rowptr, pat.cols, lp.owned, lp.g2l)
        Al = sp.csr_matrix((vals[slot], c, rp), shape=(lp.n_own, lp.n_loc))
        assert np.allclose(Al @ x, yg[lp.owned], rtol=1e-13, atol=0)
        # x-slabs: at most two neighbours, symmetric relation, channels do not overlap
        assert len(lp.neighbors) <= 2
        for k, q in enumerate(lp.neighbors):
            assert lp.rank in parts[q].neighbors
            kk = parts[q].neighbors.index(lp.rank)
            assert lp.peer_data_off[k] == parts[q].my_data_off[kk] and lp.peer_flag_off[k] == parts[q].my_flag_off[kk]
            assert lp.cap[k] == parts[q].cap[kk] and lp.cap[k] % 2 == 0 and lp.my_data_off[k] % 2 == 0
            assert len(lp.send_idx[k]) == parts[q].recv_cnt[kk] and 2 * lp.recv_cnt[k] <= lp.cap[k]
        spans = sorted((lp.my_data_off[k], lp.my_flag_off[k] + 2) for k in range(len(lp.neighbors)))
        assert all(spans[i][1] <= spans[i + 1][0] for i in range(len(spans) - 1))
        assert spans[0][0] >= 17 and spans[-1][1] <= lp.mailbox_end


def test_restriction_onto_replicated_level_sums_to_global():
    """R restricted to owned columns: the partial products of all ranks add up to the global product."""
    import scipy.sparse as sp
    from sulcusfem import hostmesh as hm, partition as pt
    from sulcusfem.hierarchy import midpoint_transfer
    mesh = hm.rectangle_mesh(10.0, 1.0, 20, 4)
    T = midpoint_transfer(mesh)                                  # P1(mesh) -> P2(mesh)
    R = sp.csr_matrix((T.t_vals, T.t_cols, T.t_rowptr), shape=(T.n_coarse, T.n_fine))
    from sulcusfem import dofmap as dm
    X = dm.p2_dof_coordinates(mesh)
    nranks = 3
    owner = pt.slab_owner(X[:, 0], nranks, X[:, 1])
    gh = pt.ghost_sets(owner, nranks, [])
    rng = np.random.default_rng(3)
    r = rng.random(T.n_fine)
    total = np.zeros(T.n_coarse)
    for q in range(nranks):
        lp = pt.partition_level(owner, nranks, q, gh, [0] * nranks)
        own_only = np.where(lp.g2l < lp.n_own, lp.g2l, -1)
        rp, c, slot = pt.localize_csr(T.t_rowptr, T.t_cols, np.arange(T.n_coarse), own_only, drop_missing=True)
        Rl = sp.csr_matrix((T.t_vals[slot], c, rp), shape=(T.n_coarse, lp.n_own))
        total += Rl @ r[lp.owned]
    assert np.allclose(total, R @ r, rtol=1e-13)


def _gloo_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, os.path.join(ROOT, 'fenics-eff-uptake_b200'))
    import torch
    import torch.distributed as dist
    import scipy.sparse as sp
    from sulcusfem import partition as pt
    dist.init_process_group('gloo', rank=rank, world_size=world)
    mesh, pat, X, vals, A = _problem(24, 6)
    n = A.shape[0]
    owner = pt.slab_owner(X[:, 0], world, X[:, 1])
    gh = pt.ghost_sets(owner, world, [(pat.rowptr, pat.cols, owner)])
    lp = pt.partition_level(owner, world, rank, gh, [0] * world)
    xg = np.random.default_rng(5).random(n)
    x = np.zeros(lp.n_loc)
    x[:lp.n_own] = xg[lp.owned]
    # halo exchange over gloo following the plan (the device path stores into peer mailboxes instead)
    reqs, bufs = [], []
    for k, q in enumerate(lp.neighbors):
        reqs.append(dist.isend(torch.from_numpy(np.ascontiguousarray(x[lp.send_idx[k]])), q))
        b = torch.empty(lp.recv_cnt[k], dtype=torch.float64)
        bufs.append(b)
        reqs.append(dist.irecv(b, q))
    for r in reqs:
        r.wait()
    for k, b in enumerate(bufs):
        x[lp.recv_off[k]:lp.recv_off[k] + lp.recv_cnt[k]] = b.numpy()
    rp, c, slot = pt.localize_csr(pat.rowptr, pat.cols, lp.owned, lp.g2l)
    y = sp.csr_matrix((vals[slot], c, rp), shape=(lp.n_own, lp.n_loc)) @ x
    # Krylov-style global dot through an all-reduce
    d = torch.tensor([float(y @ y)], dtype=torch.float64)
    dist.all_reduce(d)
    yg = A @ xg
    ok = np.allclose(y, yg[lp.owned], rtol=1e-13) and abs(d.item() - yg @ yg) <= 1e-12 * (yg @ yg)
    # sweep sharding (study drivers): every rank runs its round-robin share, all ranks receive all rows in case order
    from sulcusfem.sweep import run_sharded
    cases = [('g%d' % i, 0.1 * i) for i in range(7)]
    ran = []

    def run_case(c):
        ran.append(c)
        return {'geometry': c[0], 'value': c[1] ** 2, 'rank': rank}
    got = run_sharded(cases, run_case, rank, world)
    ok = ok and ran == [c for i, c in enumerate(cases) if i % world == rank]
    ok = ok and [i for i, _ in got] == list(range(len(cases)))
    ok = ok and all(r['geometry'] == cases[i][0] and r['rank'] == i % world for i, r in got)
    out[rank] = bool(ok)
    dist.destroy_process_group()


def test_world_size_2_gloo_halo_and_allreduce():
    import torch.multiprocessing as mp
    ctxm = mp.get_context('spawn')
    mgr = ctxm.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctxm.Process(target=_gloo_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert out.get(0) is True and out.get(1) is True


def test_case_sharding_round_robin():
    """Sweep cases (independent (geometry, mu) solves) are dealt to ranks without communication."""
    from sulcusfem.sweep import shard_cases
    cases = list(range(23 * 3))
    got = [shard_cases(cases, r, 8) for r in range(8)]
    assert sorted(sum(got, [])) == cases
    assert max(map(len, got)) - min(map(len, got)) <= 1


@pytest.mark.parametrize('nranks', [2, 3])
def test_stokes_vector_layout_with_contiguous_owned_part(nranks):
    """Row-partitioned Taylor-Hood apply in the layout of csrc/sfem_stokes.cu / sulcusfem/dist.py::DistStokesProblem:
    [u owned (interleaved) | p owned | pad | u ghosts | p ghosts].  numpy model: local K (2 right-hand sides), B, B^T
    with the planner's column numbering + the halo exchanges must reproduce the global saddle-point apply."""
    import scipy.sparse as sp
    from sulcusfem import hostmesh as hm, dofmap as dm, partition as pt
    mesh = hm.rectangle_mesh(10.0, 1.0, 24, 5)
    n2, nv = dm.p2_num_dofs(mesh), mesh.num_vertices
    cd = dm.p2_cell_dofs(mesh)
    patk = dm.build_pattern(n2, n2, [(cd, cd)])
    pb, pbt, _, _ = dm.stokes_block_plans(mesh)
    rng = np.random.default_rng(3)
    K = sp.csr_matrix((rng.random(patk.nnz), patk.cols, patk.rowptr), shape=(n2, n2))
    B = sp.csr_matrix((rng.random(pb.nnz), pb.cols[:pb.nnz], pb.rowptr), shape=(nv, 2 * n2))
    BT = sp.csr_matrix((rng.random(pbt.nnz), pbt.cols[:pbt.nnz], pbt.rowptr), shape=(2 * n2, nv))
    X = dm.p2_dof_coordinates(mesh)
    owner_u = pt.slab_owner(X[:, 0], nranks, X[:, 1])
    owner_p = owner_u[:nv]
    gh_u = pt.ghost_sets(owner_u, nranks, [(patk.rowptr, patk.cols, owner_u),
                                           (pb.rowptr, pb.cols.astype(np.int64) // 2, owner_p)])
    gh_p = pt.ghost_sets(owner_p, nranks, [(pbt.rowptr, pbt.cols, np.repeat(owner_u, 2))])
    base = [5] * nranks
    lus, lps = [], []
    for r in range(nranks):
        nv_own = int((owner_p == r).sum())
        hole, _ = pt.stokes_gaps(nv_own, 0)
        lus.append(pt.partition_level(owner_u, nranks, r, gh_u, base, nb_max=2, ghost_gap=hole))
    base = lus[0].all_mailbox_ends
    for r in range(nranks):
        nv_own = int((owner_p == r).sum())
        _, gap_p = pt.stokes_gaps(nv_own, len(lus[r].ghost))
        lps.append(pt.partition_level(owner_p, nranks, r, gh_p, base, nb_max=1, ghost_gap=gap_p))
    # global vectors (solver layout: interleaved velocity | pressure) and the global apply
    zu, zp = rng.random((n2, 2)), rng.random(nv)
    yu = K @ zu + (BT @ zp).reshape(n2, 2)
    yp = B @ zu.ravel()
    # local vectors
    zs = []
    for lu, lp in zip(lus, lps):
        n_alloc = 2 * lu.n_loc + len(lp.ghost)
        z = np.full(n_alloc, np.nan)
        z[:2 * lu.n_own] = zu[lu.owned].ravel()
        z[2 * lu.n_own:2 * lu.n_own + lp.n_own] = zp[lp.owned]
        assert 2 * lu.n_own + lp.n_own <= 2 * (lu.n_own + lu.ghost_gap)          # owned pressure fits in the hole
        assert 2 * lu.n_own + lp.n_loc == n_alloc                                # pressure ghosts end the vector
        zs.append(z)
    pt.emulate_exchange(lus, [z[:2 * lu.n_loc] for z, lu in zip(zs, lus)], nb=2)
    pt.emulate_exchange(lps, [z[2 * lu.n_own:] for z, lu in zip(zs, lus)], nb=1)
    for r, (lu, lp, z) in enumerate(zip(lus, lps, zs)):
        rp, lc, slot = pt.localize_csr(patk.rowptr, patk.cols, lu.owned, lu.g2l)
        Kl = sp.csr_matrix((K.data[slot], lc, rp), shape=(lu.n_own, lu.n_loc))
        col_b = np.full(2 * n2, -1, dtype=np.int64)
        ok = lu.g2l >= 0
        col_b[0::2] = np.where(ok, 2 * lu.g2l, -1)
        col_b[1::2] = np.where(ok, 2 * lu.g2l + 1, -1)
        rp, lc, slot = pt.localize_csr(pb.rowptr, pb.cols, lp.owned, col_b)
        Bl = sp.csr_matrix((B.data[slot], lc, rp), shape=(lp.n_own, 2 * lu.n_loc))
        rows_bt = np.stack([2 * lu.owned, 2 * lu.owned + 1], axis=1).ravel()
        rp, lc, slot = pt.localize_csr(pbt.rowptr, pbt.cols, rows_bt, lp.g2l)
        BTl = sp.csr_matrix((BT.data[slot], lc, rp), shape=(2 * lu.n_own, lp.n_loc))
        zu_l = np.nan_to_num(z[:2 * lu.n_loc], nan=0.0).reshape(lu.n_loc, 2)     # the hole holds pressure / pad: K never reads it
        assert not np.isnan(z[:2 * lu.n_loc].reshape(lu.n_loc, 2)[np.unique(Kl.indices)]).any()
        zp_l = z[2 * lu.n_own:]
        assert not np.isnan(zp_l[np.unique(BTl.indices)]).any()
        got_u = Kl @ zu_l + (BTl @ np.nan_to_num(zp_l, nan=0.0)).reshape(lu.n_own, 2)
        got_p = Bl @ np.nan_to_num(z[:2 * lu.n_loc], nan=0.0)
        assert np.allclose(got_u, yu[lu.owned], rtol=1e-13, atol=1e-13)
        assert np.allclose(got_p, yp[lp.owned], rtol=1e-13, atol=1e-13)


@pytest.mark.parametrize('nranks', [2, 3])
def test_local_gather_plan_reproduces_the_owned_rows_bit_for_bit(nranks):
    """Distributed assembly (dist.plan_local_gather): a rank's gather map = the global one restricted to its rows and
    re-addressed to an element buffer that holds only the cells those rows touch (+ the whole Robin facet family).  Applied
    to the same element matrices it must give exactly the owned rows of the globally gathered matrix (same contributions
    in the same order -> identical floating-point sums)."""
    from sulcusfem import dist, dofmap as dm, hostmesh as hm, partition as pt
    from sulcusfem.unstructured import mesh_domain
    m = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.25, 'sulcus')
    mk = hm.build_markers(m, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    bm = mk['bc_markers'].values
    cd = dm.p2_cell_dofs(m)
    n = dm.p2_num_dofs(m)
    f, _, _ = dm.boundary_facets(m, bm, 4)
    fd = dm.p2_facet_dofs(m, f)
    pat = dm.build_pattern(n, n, [(cd, cd), (fd, fd)])
    rng = np.random.default_rng(5)
    Ecell = rng.random((m.num_cells, 36))
    Efac = rng.random((len(f), 9))
    E = np.concatenate([Ecell.ravel(), Efac.ravel()])
    fb = pat.family_base[1]
    vals = np.array([E[pat.contrib_code[pat.contrib_ptr[s]:pat.contrib_ptr[s + 1]]].sum() for s in range(pat.nnz)])
    # numpy's pairwise .sum() is not the device's sequential order; use an explicit left-to-right sum on both sides
    def seq(codes_of_slot, buf):
        acc = 0.0
        for c in codes_of_slot:
            acc += buf[c]
        return acc
    X = dm.p2_dof_coordinates(m)
    owner = pt.slab_owner(X[:, 0], nranks, X[:, 1])
    gh = pt.ghost_sets(owner, nranks, [(pat.rowptr, pat.cols, owner)])
    seen_cells = 0
    for r in range(nranks):
        lp = pt.partition_level(owner, nranks, r, gh, [9] * nranks, nb_max=1)
        rp, lc, slot = pt.localize_csr(pat.rowptr, pat.cols, lp.owned, lp.g2l)
        lens, codes = dist._flatten_contribs(pat.contrib_ptr, pat.contrib_code, slot)
        cells, maps = dist.plan_local_gather([(lens, codes)], 36, fb)
        ptr, code = maps[0]
        assert len(cells) < m.num_cells and ptr[-1] == len(code) == len(codes)
        seen_cells += len(cells)
        Eloc = np.concatenate([Ecell[cells].ravel(), Efac.ravel()])
        for k in range(0, len(slot), 7):
            g = seq(pat.contrib_code[pat.contrib_ptr[slot[k]]:pat.contrib_ptr[slot[k] + 1]], E)
            l = seq(code[ptr[k]:ptr[k + 1]], Eloc)
            assert g == l                                   # bit for bit
        # every cell that holds an owned dof is present, no cell without one
        touches = np.isin(cd, lp.owned).any(axis=1)
        assert np.array_equal(cells, np.flatnonzero(touches))
    assert seen_cells < 1.5 * m.num_cells                   # one cell layer of overlap, not replication
    del vals
