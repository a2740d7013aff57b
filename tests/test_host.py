"""CPU: host-side logic (meshes, markers, DOF maps, sparsity, gather maps, hierarchy, front-end objects)."""
import os

import sys

import numpy as np
import pytest
import scipy.sparse as sp

from oracle import cpu_oracle as co
from sulcusfem import dofmap as dm
from sulcusfem import hierarchy as hy
from sulcusfem import hostmesh as hm
from sulcusfem.unstructured import mesh_domain


def _mesh_ok(m, area):
    assert np.all(np.abs(m.signed_areas()) > 0)
    assert abs(np.abs(m.signed_areas()).sum() - area) < 2e-3 * area
    # Euler characteristic of a disc
    assert m.num_vertices - m.num_edges + m.num_cells == 1
    assert np.all(np.diff(m.cells.astype(np.int64), axis=1) > 0)          # ascending vertex ids per cell


@pytest.mark.parametrize("w,d", [(0.5, 1.0), (0.25, 0.25), (1.0, 0.2), (0.3, 1.0)])
def test_sulcus_meshers(w, d):
    area = 10.0 + 2 * w * d / np.pi
    for m in (hm.sulcus_mesh(10.0, 1.0, w, d, 0.1), mesh_domain(10.0, 1.0, w, d, 0.1, 'sulcus')):
        _mesh_ok(m, area)
        mk = hm.build_markers(m, 10.0, 1.0, 5 - w / 2, 5 + w / 2, 'sulcus')
        bc = mk['bc_markers'].values
        # every exterior facet carries exactly one of the ids 1..4 and no interior facet does
        assert np.all(bc[m.edge_on_boundary] > 0) and np.all(bc[~m.edge_on_boundary] == 0)
        L = np.linalg.norm(m.coords[m.edges[:, 0]] - m.coords[m.edges[:, 1]], axis=1)
        assert abs(L[bc == 1].sum() - 1.0) < 1e-12 and abs(L[bc == 2].sum() - 1.0) < 1e-12
        assert abs(L[bc == 3].sum() - 10.0) < 1e-12
        y0 = mk['y0_markers'].values
        assert abs(L[(y0 == 10) & m.edge_on_boundary].sum() - (10.0 - w)) < 1e-12
        assert abs(L[(y0 == 10) & ~m.edge_on_boundary].sum() - w) < 1e-12      # mouth edges on y=0
        bs = mk['bottom_segment_markers'].values
        # 'sulcus' needs y < -eps at all three points: the two curve facets touching the corners stay 0
        curve = (bc == 4) & (y0 != 10)
        assert curve.sum() - (bs == 6).sum() == 2
        dmk = mk['domain_markers'].values
        a = np.abs(m.signed_areas())
        assert abs(a[dmk == 2].sum() - 10.0) < 1e-9


def test_rectangle_mesh_and_refine_numbering():
    m = hm.rectangle_mesh(10.0, 1.0, 20, 4)
    _mesh_ok(m, 10.0)
    f = hm.refine(m)
    _mesh_ok(f, 10.0)
    assert f.num_vertices == m.num_vertices + m.num_edges and f.num_cells == 4 * m.num_cells
    assert np.array_equal(f.coords[:m.num_vertices], m.coords)
    assert np.allclose(f.coords[m.num_vertices:], m.edge_midpoints())
    # P2 nodes of the coarse mesh = P1 nodes of the refined mesh, same numbering
    assert np.allclose(dm.p2_dof_coordinates(m), f.coords)


def test_refine_projects_curved_boundary():
    m = hm.sulcus_mesh(10.0, 1.0, 0.5, 1.0, 0.2)
    f = hm.refine(m, project_curved_boundary=True)
    g = m.geometry
    bnd = np.unique(f.edges[f.edge_on_boundary].ravel())
    p = f.coords[bnd]
    on_floor = p[:, 1] < -1e-12
    assert np.allclose(p[on_floor, 1], hm.sulcus_floor(p[on_floor, 0], g['xL'], g['w'], g['d']), atol=1e-12)


def test_dofmaps_and_patterns_match_oracle_bit_exact():
    m = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.2, 'sulcus')
    om = co.Mesh(m.coords, m.cells)
    assert np.array_equal(m.edges, om.edges) and np.array_equal(m.cell_edges, om.cell_edges)
    assert np.array_equal(dm.p2_cell_dofs(m), om.p2_cell_dofs())
    assert np.array_equal(dm.th_cell_dofs(m), om.th_cell_dofs())
    assert np.array_equal(dm.p2_dof_coordinates(m), om.p2_dof_coords())
    for cd, n in ((dm.p2_cell_dofs(m), om.n_p2), (dm.th_cell_dofs(m), 2 * om.n_p2 + om.nv), (dm.p1_cell_dofs(m), om.nv)):
        pat = dm.build_pattern(n, n, [(cd, cd)])
        ip, ix = co.clique_pattern(cd, n)
        assert np.array_equal(pat.rowptr, ip) and np.array_equal(pat.cols, ix)
    # marker parity (vectorised host code vs per-facet oracle loop)
    mk = hm.build_markers(m, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    for key, names in (('bc_markers', ['left', 'right', 'top', 'bottom']),
                       ('bottom_segment_markers', ['bottom_left', 'bottom_right', 'sulcus', 'sulcus_opening']),
                       ('y0_markers', ['y0_line'])):
        assert np.array_equal(mk[key].values, co.mark_facets(om, 10.0, 1.0, 4.75, 5.25, names))
    assert np.array_equal(mk['domain_markers'].values, co.cell_markers(om))
    d_or, _ = co.concentration_bcs(om, mk['bc_markers'].values)
    d_host = np.union1d(dm.dirichlet_dofs_p2(m, mk['bc_markers'].values, 1), dm.dirichlet_dofs_p2(m, mk['bc_markers'].values, 2))
    assert np.array_equal(d_or, d_host)


def test_gather_map_reproduces_assembly():
    """The slot -> contributions map, applied to oracle element matrices, gives the oracle matrix."""
    m = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.25, 'sulcus')
    om = co.Mesh(m.coords, m.cells)
    mk = hm.build_markers(m, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    bm = mk['bc_markers'].values
    cd = dm.p2_cell_dofs(m)
    f, _, _ = dm.boundary_facets(m, bm, 4)
    fd = dm.p2_facet_dofs(m, f)
    pat = dm.build_pattern(om.n_p2, om.n_p2, [(cd, cd), (fd, fd)])
    lam, w = co.triangle_rule(2)
    G = np.einsum('qia,cad->cqid', co.p2_dbasis(lam), om.glam)
    Ke = np.einsum('q,cqid,cqjd,c->cij', w, G, G, np.abs(om.det))
    # facet mass with the trace basis [va, vb, mid]
    t, wt = co.interval_rule(4)
    phi = np.stack([(1 - t) * (1 - 2 * t), t * (2 * t - 1), 4 * t * (1 - t)], axis=1)
    L = np.linalg.norm(m.coords[m.edges[f, 0]] - m.coords[m.edges[f, 1]], axis=1)
    Fe = np.einsum('q,qi,qj,f->fij', wt, phi, phi, L) * 0.7
    E = np.concatenate([Ke.ravel(), Fe.ravel()])
    assert pat.family_base == [0, Ke.size] and pat.buffer_len == E.size
    vals = np.add.reduceat(E[pat.contrib_code], pat.contrib_ptr[:-1])
    A = sp.csr_matrix((vals, pat.cols, pat.rowptr), shape=(om.n_p2, om.n_p2))
    ref = co.assemble_p2_stiffness(om) + co.assemble_p2_robin(om, f, mu_const=0.7)
    assert abs(A - ref).max() < 1e-13
    # contributions of a slot are ordered by (family, cell)
    for s in range(0, pat.nnz, 97):
        c = pat.contrib_code[pat.contrib_ptr[s]:pat.contrib_ptr[s + 1]]
        assert np.all(np.diff(c) > 0)


def test_hierarchy_transfers():
    fine = hm.refine(mesh_domain(10.0, 1.0, 0.5, 1.0, 0.2, 'sulcus'))
    H = hy.build_hierarchy(fine, coarsest_vertices=150)
    assert len(H.meshes) >= 3 and H.meshes[1] is fine.parent
    sizes = [dm.p2_num_dofs(fine)] + [m.num_vertices for m in H.meshes]
    for l, T in enumerate(H.transfers):
        assert (T.n_fine, T.n_coarse) == (sizes[l], sizes[l + 1])
        P = sp.csr_matrix((T.vals, T.cols, T.rowptr), shape=(T.n_fine, T.n_coarse))
        R = sp.csr_matrix((T.t_vals, T.t_cols, T.t_rowptr), shape=(T.n_coarse, T.n_fine))
        assert abs(P.T - R).max() == 0.0
        assert np.allclose(np.asarray(P.sum(axis=1)).ravel(), 1.0)          # partition of unity
        # prolongation reproduces linear functions (exactly for nested, up to the curved boundary otherwise)
        Xc = H.meshes[l].coords
        Xf = dm.p2_dof_coordinates(H.meshes[0]) if l == 0 else H.meshes[l - 1].coords
        lin = lambda X: 2.0 + 0.3 * X[:, 0] - 1.7 * X[:, 1]
        err = np.abs(P @ lin(Xc) - lin(Xf))
        if T.nested and l == 0:
            assert err.max() < 1e-12
        else:
            assert np.median(err) < 1e-12 and np.quantile(err, 0.9) < 0.05   # only nodes near the curved floor differ
    assert H.transfers[0].nested and H.transfers[1].nested and not H.transfers[-1].nested


def test_locate_points_and_function_eval():
    from sulcusfem.fem import Function, FunctionSpace, VectorFunctionSpace
    m = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.2, 'sulcus')
    V = FunctionSpace(m, 'CG', 2)
    X = V.tabulate_dof_coordinates()
    quad = lambda x, y: 1.0 + x - 2 * y + 0.5 * x * y - 0.25 * y * y
    f = Function(V, quad(X[:, 0], X[:, 1]))
    for p in ((0.3, 0.4), (5.0, -0.5), (9.99, 0.999), (5.0, 0.0)):
        assert abs(f(p) - quad(*p)) < 1e-12
    with pytest.raises(RuntimeError):
        f((5.0, -3.0))
    f.set_allow_extrapolation(True)
    f((5.0, -1.05))
    W = VectorFunctionSpace(m, 'P', 2)
    u = Function(W, np.concatenate([X[:, 1] * (1 - X[:, 1]), 0 * X[:, 0]]))
    assert np.allclose(u((2.0, 0.25)), [0.1875, 0.0])
    v = f.vector()
    a = v.get_local(); a[:] = 3.0; v.set_local(a); v.apply('insert')
    assert f((1.0, 0.5)) == pytest.approx(3.0)
    assert m.num_vertices() == m.num_vertices and isinstance(m.hmin(), float)


def test_mesh_generator_interface(tmp_path):
    from sulcusfem.mesh import MeshGenerator
    from sulcusfem.parameters import Parameters
    p = Parameters(mode='no-adv', mesh_size_dim=0.2)
    p.validate(); p.nondim()
    kw = p.get_mesh_generator_params()
    kw['domain_type'] = 'sulcus'
    res = MeshGenerator(**kw).generate_mesh()
    assert set(res) == {'mesh', 'bc_markers', 'bottom_segment_markers', 'y0_markers', 'domain_markers', 'mesh_info'}
    assert set(np.unique(res['bc_markers'].array())) == {0, 1, 2, 3, 4}
    kw['domain_type'] = 'rectangular'
    res2 = MeshGenerator(**kw).generate_mesh()
    assert set(res2) == {'mesh', 'bc_markers', 'mesh_info'}
    with pytest.raises(ValueError):
        MeshGenerator(10, 1, 1.0, 0.5, -0.1, 1, 'sulcus')
    with pytest.raises(ValueError):
        MeshGenerator(10, 1, 1.0, 0.5, 0.1, 1, 'triangle')
    # round trip through the dolfin-XML format the reference writes
    path = str(tmp_path / 'm.xml')
    hm.write_dolfin_xml(res['mesh'], path)
    back = hm.read_dolfin_xml(path)
    assert np.array_equal(back.cells, res['mesh'].cells) and np.array_equal(back.coords, res['mesh'].coords)
    res3 = MeshGenerator(10.0, 1.0, 1.0, 0.5, 0.2, 1, 'sulcus', mesh_file=path).generate_mesh()
    assert np.array_equal(res3['bc_markers'].array(), res['bc_markers'].array())


@pytest.mark.parametrize('sigma', [32, 64, 256, 4096])
def test_sell_plan_layout(sigma):
    """Sliced-ELL plan (sulcusfem/sell.py): every CSR entry is stored exactly once at slice_ptr[s] + 32 j + lane,
    entries of a row keep their column order, slices are padded to their longest row, empty / ragged rows and
    row counts that are not multiples of 32 are handled, and a P2 pattern pads by a few per cent at most."""
    import scipy.sparse as sp
    from sulcusfem import sell as sl
    rng = np.random.default_rng(11)
    cases = [sp.csr_matrix((0, 0)), sp.csr_matrix((5, 7)), sp.eye(1, format='csr')]
    for (m, n, dens) in ((33, 40, 0.2), (257, 257, 0.05), (1000, 900, 0.01)):
        A = sp.random(m, n, dens, random_state=4, format='lil')
        A[m // 2, :] = 0
        A[m // 3, : min(n, 30)] = 2.0
        A = A.tocsr(); A.sort_indices()
        cases.append(A)
    for A in cases:
        m, n = A.shape
        p = sl.build_plan(A.indptr, A.indices, sigma)
        assert p.nslices == (m + 31) // 32 and len(p.perm) == p.nslices * 32
        assert p.padded == int(p.slice_ptr[-1]) and np.all(np.diff(p.slice_ptr) % 32 == 0)
        rows = p.perm[p.perm >= 0]
        assert np.array_equal(np.sort(rows), np.arange(m))               # a permutation of the rows
        assert np.array_equal(np.sort(p.src[p.src >= 0]), np.arange(A.nnz))   # every CSR slot exactly once
        assert np.array_equal(p.scols >= 0, p.src >= 0)
        lens = np.diff(A.indptr)
        for s in range(p.nslices):
            base, width = int(p.slice_ptr[s]), (int(p.slice_ptr[s + 1]) - int(p.slice_ptr[s])) // 32
            lane_rows = p.perm[s * 32:(s + 1) * 32]
            ll = np.where(lane_rows >= 0, lens[np.maximum(lane_rows, 0)], 0) if m else np.zeros(32, int)
            assert width == ll.max()
            blk = p.src[base:base + width * 32].reshape(width, 32)
            for l in range(32):
                r = lane_rows[l]
                want = np.arange(A.indptr[r], A.indptr[r + 1]) if r >= 0 else np.zeros(0, int)
                assert np.array_equal(blk[:len(want), l], want) and np.all(blk[len(want):, l] == -1)
        for nparts in (1, 7, 96, 14208):                                   # contiguous parts balanced by column-steps
            parts = sl.partition(p, nparts)
            assert len(parts) == nparts + 1 and parts[0] == 0 and parts[-1] == p.nslices
            assert np.all(np.diff(parts) >= 0)
            steps = p.slice_ptr.astype(np.int64)[parts] // 32
            widest = int(np.diff(p.slice_ptr).max() // 32) if p.nslices else 0
            assert np.all(np.abs(np.diff(steps) - p.padded / 32 / nparts) <= widest + 1)
        if sigma > 32:                                                     # sorted by decreasing length inside windows
            sig = ((sigma + 31) // 32) * 32
            ordl = np.where(p.perm >= 0, lens[np.maximum(p.perm, 0)], 0) if m else np.zeros(0, int)
            for w0 in range(0, len(ordl), sig):
                assert np.all(np.diff(ordl[w0:w0 + sig]) <= 0)
    from sulcusfem import dofmap as dm, hostmesh as hm
    mesh = hm.rectangle_mesh(10.0, 1.0, 200, 20)
    cd = dm.p2_cell_dofs(mesh)
    pat = dm.build_pattern(dm.p2_num_dofs(mesh), dm.p2_num_dofs(mesh), [(cd, cd)])
    p = sl.build_plan(pat.rowptr, pat.cols, sigma)
    assert p.fill < (1.10 if sigma < 256 else 1.03), p.fill
    r, c, v = sl.to_dense_rows(p, np.arange(pat.nnz, dtype=float))
    B = sp.coo_matrix((v, (r, c)), shape=(pat.nrows, pat.ncols) if hasattr(pat, 'nrows') else None).tocsr()
    A = sp.csr_matrix((np.arange(pat.nnz, dtype=float), pat.cols, pat.rowptr))
    assert abs(A - B).max() == 0


def test_locator_bins_contain_the_cell_of_every_point():
    """Uniform-grid locator plan (sulcusfem/locator.py): the bin of a point lists, in ascending order, every cell
    that contains it (checked against the oracle's brute-force location), including points on edges, vertices,
    the domain boundary and the corners of the bin grid."""
    from sulcusfem import locator as lc
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.1, 'sulcus')
    om = co.Mesh(mesh.coords, mesh.cells)
    g = lc.build_bins(mesh)
    assert g.bin_ptr[0] == 0 and g.bin_ptr[-1] == len(g.bin_cells) and len(g.bin_ptr) == g.nbx * g.nby + 1
    for b in range(g.nbx * g.nby):
        seg = g.bin_cells[g.bin_ptr[b]:g.bin_ptr[b + 1]]
        assert np.all(np.diff(seg) > 0)
    rng = np.random.default_rng(5)
    pts = np.concatenate([
        np.stack([rng.uniform(0, 10, 400), rng.uniform(-1, 1, 400)], axis=1),
        mesh.coords[rng.integers(0, mesh.num_vertices, 100)],                  # vertices
        mesh.edge_midpoints()[rng.integers(0, mesh.num_edges, 100)],           # edge midpoints
        np.array([[0.0, 0.0], [10.0, 1.0], [0.0, 1.0], [10.0, 0.0], [5.0, -1.0]])])
    cell, _ = co.locate_brute(om, pts)
    bx = np.clip(np.floor((pts[:, 0] - g.x0) / g.hx).astype(int), 0, g.nbx - 1)
    by = np.clip(np.floor((pts[:, 1] - g.y0) / g.hy).astype(int), 0, g.nby - 1)
    b = by * g.nbx + bx
    inside = cell >= 0
    assert inside.sum() > 300 and (~inside).sum() > 50
    for i in np.flatnonzero(inside):
        seg = g.bin_cells[g.bin_ptr[b[i]]:g.bin_ptr[b[i] + 1]]
        assert cell[i] in seg, (i, pts[i])
    # a handful of candidates per bin, not the whole mesh
    assert np.diff(g.bin_ptr).mean() < 16


def test_oracle_point_evaluation_known_answers():
    """P2 interpolation reproduces quadratics exactly; points outside the mesh are flagged invalid
    (reference analysis.py:367-372 semantics)."""
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.1, 'sulcus')
    om = co.Mesh(mesh.coords, mesh.cells)
    X = om.p2_dof_coords()
    f = lambda x, y: 1 + 2 * x - 3 * y + 0.5 * x * x - x * y + 0.25 * y * y   # noqa: E731
    vals = f(X[:, 0], X[:, 1])
    rng = np.random.default_rng(0)
    pts = np.stack([rng.uniform(-0.5, 10.5, 600), rng.uniform(-1.2, 1.2, 600)], axis=1)
    v, ok = co.eval_points(om, vals, pts)
    in_channel = (pts[:, 0] > 0) & (pts[:, 0] < 10) & (pts[:, 1] > 0) & (pts[:, 1] < 1)
    assert np.all(ok[in_channel]) and not np.any(ok[(pts[:, 0] < 0) | (pts[:, 0] > 10) | (pts[:, 1] > 1)])
    assert np.abs(v[ok] - f(pts[ok, 0], pts[ok, 1])).max() < 1e-12
    s, vv = co.line_profile(om, vals, 5.0, 'v', None, 101)
    assert s.min() < -0.9 and s.max() == 1.0 and np.abs(vv - f(5.0, s)).max() < 1e-12
    # the mouth-level line of compute_velocity_metrics stays inside the channel on its whole length
    s, _ = co.line_profile(om, vals, 1e-6, 'h', (0, 10.0), 100)
    assert len(s) == 100


def test_study_case_lists_match_reference_csvs():
    """The case enumerations of sulcusfem/studies.py reproduce the keys of the reference's checked-in CSVs
    (tests/golden/study_columns.json, made by tests/golden/make_study_columns.py from /root/reference)."""
    import json
    import os
    from sulcusfem import studies
    from sulcusfem.parameters import Parameters, create_geometry_variations
    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'study_columns.json')))
    ar = g['aspect_ratio_analysis_results.csv']
    assert [f"{n}_h{h}" for n, _, h, _ in studies.aspect_ratio_cases()] == ar['configs']
    mu = g['mu_parameter_sweep_results.csv']
    assert [f"{r}_mu_{f:.1f}x" for r, fs in studies.REGIMES.items() for f in fs] == mu['configs']
    pb = g['no_adv_mu_sweep_results.csv']
    geos = create_geometry_variations(Parameters(mode='no-adv'), max_width=1.0)
    assert sorted(geos) == pb['geometries'] and studies.MU_FACTORS_PHASE_B == pb['mu_factors']
    assert len(geos) * len(studies.MU_FACTORS_PHASE_B) == pb['n_rows']
    gc = g['geometry_comparison_results.csv']
    assert len(create_geometry_variations(Parameters(mode='no-uptake'), max_width=1.0)) * 3 + 3 == gc['n_rows']
    assert studies.format_filename_value(0.1) == '0p100' and studies.format_filename_value(10.0) == '10'
    ad = g['advdiff_validation_step_pe_x_mu.csv']
    assert sorted([float(p), float(m)] for p in studies.PE_VALUES for m in studies.MU_FACTORS_ADV) == ad['cases']


def test_paraview_export_round_trip(tmp_path):
    """File(path) << f (reference simulation.py:137-138,165): .pvd collection + .vtu with vertex values, read back
    bit-exactly."""
    import xml.etree.ElementTree as ET
    from sulcusfem import fem
    from sulcusfem.export import File
    mesh = hm.rectangle_mesh(2.0, 1.0, 4, 3)
    X = dm.p2_dof_coordinates(mesh)
    c = fem.Function(fem.FunctionSpace(mesh, 'CG', 2), np.sin(X[:, 0]) + X[:, 1] / 3.0)
    u = fem.Function(fem.VectorFunctionSpace(mesh, 'P', 2), np.concatenate([X[:, 1] * (1 - X[:, 1]), 0.1 * X[:, 0]]))
    f = File(str(tmp_path / 'pv' / 'concentration.pvd'))
    f << c
    f << (c, 2.5)
    File(str(tmp_path / 'pv' / 'velocity.pvd')) << u
    coll = ET.parse(tmp_path / 'pv' / 'concentration.pvd').getroot().find('Collection').findall('DataSet')
    assert [d.get('file') for d in coll] == ['concentration000000.vtu', 'concentration000001.vtu']
    assert float(coll[1].get('timestep')) == 2.5
    piece = ET.parse(tmp_path / 'pv' / 'concentration000000.vtu').getroot().find('UnstructuredGrid').find('Piece')
    assert int(piece.get('NumberOfPoints')) == mesh.num_vertices and int(piece.get('NumberOfCells')) == mesh.num_cells
    pts = np.array(piece.find('Points').find('DataArray').text.split(), dtype=float).reshape(-1, 3)
    assert np.array_equal(pts[:, :2], mesh.coords) and not pts[:, 2].any()
    conn = np.array(piece.find('Cells').findall('DataArray')[0].text.split(), dtype=int).reshape(-1, 3)
    assert np.array_equal(conn, mesh.cells)
    vals = np.array(piece.find('PointData').find('DataArray').text.split(), dtype=float)
    assert np.array_equal(vals, c.values[:mesh.num_vertices])
    pv = ET.parse(tmp_path / 'pv' / 'velocity000000.vtu').getroot().find('UnstructuredGrid').find('Piece')
    v3 = np.array(pv.find('PointData').find('DataArray').text.split(), dtype=float).reshape(-1, 3)
    n2 = len(u.values) // 2
    assert np.array_equal(v3[:, 0], u.values[:mesh.num_vertices]) and np.array_equal(v3[:, 1], u.values[n2:n2 + mesh.num_vertices])
    with pytest.raises(ValueError):
        File(str(tmp_path / 'x.xdmf'))


def test_direct_divergence_block_maps_match_mixed_space_extraction():
    """Host plan of the direct Stokes block assembly (device.StokesProblem): gathering B and B^T from the 3 x 12
    divergence element buffer gives exactly the entries (same order, same summation order) that the mixed-space
    matrix, assembled with dolfin's clique pattern and cut into blocks, holds."""
    mesh = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.2, 'sulcus')
    n2, nv, nc = dm.p2_num_dofs(mesh), mesh.num_vertices, mesh.num_cells
    n = 2 * n2 + nv
    rng = np.random.default_rng(3)
    EB = rng.random((nc, 3, 12))
    # the same numbers placed in a 15 x 15 Taylor-Hood element matrix (rows/cols [ux x6, uy x6, p x3])
    E = np.zeros((nc, 15, 15))
    E[:, 12:, :12] = EB
    E[:, :12, 12:] = EB.transpose(0, 2, 1)
    cd = dm.th_cell_dofs(mesh)
    pat = dm.build_pattern(n, n, [(cd, cd)])

    def gather(p, buf):
        out = np.zeros(p.nnz)
        for s_ in range(p.nnz):                        # fixed order, like k_gather
            acc = 0.0
            for k in range(p.contrib_ptr[s_], p.contrib_ptr[s_ + 1]):
                acc += buf[p.contrib_code[k]]
            out[s_] = acc
        return out
    A = gather(pat, E.ravel())
    rows = np.repeat(np.arange(n, dtype=np.int64), np.diff(pat.rowptr.astype(np.int64)))
    cols = pat.cols.astype(np.int64)
    # direct plans, the ones StokesProblem.__init__ uses
    pb, pbt, bt_code, _ = dm.stokes_block_plans(mesh)
    import copy
    pbt = copy.copy(pbt)
    pbt.contrib_code = bt_code
    Bd, BTd = gather(pb, EB.ravel()), gather(pbt, EB.ravel())
    # blocks cut out of the mixed-space matrix, velocity in interleaved numbering, sorted by (row, col)
    il_rows = 2 * (rows % n2) + rows // n2
    il_cols = 2 * (cols % n2) + cols // n2
    for sel, r_, c_, p, got in (((rows >= 2 * n2) & (cols < 2 * n2), rows - 2 * n2, il_cols, pb, Bd),
                                ((rows < 2 * n2) & (cols >= 2 * n2), il_rows, cols - 2 * n2, pbt, BTd)):
        slot = np.flatnonzero(sel)
        order = np.lexsort((c_[slot], r_[slot]))
        slot = slot[order]
        assert len(slot) == p.nnz and np.array_equal(c_[slot], p.cols)
        assert np.array_equal(np.bincount(r_[slot], minlength=p.nrows), np.diff(p.rowptr))
        assert np.array_equal(A[slot], got)            # bit for bit


def test_profile_rows_schema_matches_reference_csv():
    """collect_profile_rows (no_uptake_analysis.py:315-359): the tidy rows carry the columns of the reference's
    profiles_samples_<geometry>.csv, in order."""
    import json
    import os
    from sulcusfem import studies

    class P:
        Pe = 1.0
    res = {'params': P(), 'domain_type': 'sulcus', 'geometry': 'largest',
           'mass_metrics': {'profiles_full': {'horizontal': {'mid_channel': {'y': 0.5, 'x': [0.0, 5.0, 10.0], 'c': [1.0, 0.5, 0.0]}},
                                              'vertical': {'x_mid': {'x': 5.0, 'y': [0.0, 1.0], 'c': [0.4, 0.6]}}},
                            'profiles_meta': {'n_points': 3, 'x_range': (0.0, 10.0), 'y_range': None}}}
    rows = studies.collect_profile_rows(res, geometry_key='largest')
    gold = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'profile_samples.json')))
    assert [list(r.keys()) for r in rows] == [gold['columns']] * 3          # horizontal lines only, like the reference
    assert [r['Index'] for r in rows] == [0, 1, 2] and rows[1]['c'] == 0.5 and rows[0]['y_min'] is None
    assert studies.collect_profile_rows(None) == []


def test_prefetch_meshes_parallel_matches_in_process_generation():
    """simulation.prefetch_meshes: geometries meshed in spawned worker processes land in the per-geometry cache with
    their multigrid hierarchy attached, bit-identical to in-process generation; cached geometries are not rebuilt."""
    from sulcusfem import simulation
    from sulcusfem.parameters import Parameters
    from sulcusfem.mesh import MeshGenerator

    def params(w, h):
        p = Parameters(mode='no-adv', mesh_size_dim=0.1)
        p.sulci_w_dim, p.sulci_h_dim = w, h
        p.validate()
        p.nondim()
        return p
    saved = dict(simulation._MESH_CACHE)
    simulation._MESH_CACHE.clear()
    try:
        jobs = [(params(0.5, 1.0), 'sulcus'), (params(0.25, 0.25), 'sulcus'), (params(0.5, 1.0), 'rectangular'),
                (params(0.5, 1.0), 'sulcus')]                                    # one duplicate
        assert simulation.prefetch_meshes(jobs, workers=2) == 3
        assert simulation.prefetch_meshes(jobs, workers=2) == 0                 # all cached now
        for p, dom in jobs[:3]:
            key, mp = simulation._mesh_key(p, dom)
            got = simulation._MESH_CACHE[key]
            mp['output_dir'], mp['domain_type'] = None, dom
            want = MeshGenerator(**mp, **simulation.MESH_OPTIONS).generate_mesh()
            assert np.array_equal(got['mesh'].coords, want['mesh'].coords) and np.array_equal(got['mesh'].cells, want['mesh'].cells)
            assert np.array_equal(got['bc_markers'].values, want['bc_markers'].values)
            hier = got['mesh']._sfem_cache['hierarchy']
            assert hier.meshes[0] is got['mesh']                                 # identity survives the pickle round trip
            ref = hy.build_hierarchy(want['mesh'])
            assert [m.num_vertices for m in hier.meshes] == [m.num_vertices for m in ref.meshes]
            assert all(np.array_equal(a.vals, b.vals) for a, b in zip(hier.transfers, ref.transfers))
            # the workers also built the device problems' patterns + gather maps: memoised on the mesh objects under
            # the keys the solving process will ask for, identical to a fresh build
            cache = got['mesh']._pattern_cache
            n, cd, f, fd, pat = dm.scalar_level_plan(got['mesh'], got['bc_markers'].values, 2, 4)
            assert any(k[:2] == ('scalar', 2) and k[2] is not None for k in cache) and cache[[k for k in cache if k[:2] == ('scalar', 2) and k[2] is not None][0]] is pat
            fresh = dm.build_pattern(n, n, [(cd, cd), (fd, fd)])
            for name in ('rowptr', 'cols', 'contrib_ptr', 'contrib_code'):
                assert np.array_equal(getattr(pat, name), getattr(fresh, name))
            assert ('stokes_B',) not in cache                                   # no-adv jobs: scalar plans only
            assert all(('scalar', 1, k[2]) in m._pattern_cache for m in hier.meshes for k in [next(iter(m._pattern_cache))])
    finally:
        simulation._MESH_CACHE.clear()
        simulation._MESH_CACHE.update(saved)


def test_ratio_metrics_reproduce_reference_csv_columns():
    """studies.add_ratio_metrics (no_uptake_analysis.py:262-313) applied to the base columns of the reference's
    checked-in geometry_comparison_results.csv reproduces its seven ratio columns."""
    import json
    import os
    import pandas as pd
    from sulcusfem import studies
    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'study_columns.json')))['geometry_comparison_results.csv']
    ratio = ['Concentration_Ratio', 'Channel_Conc_Ratio', 'Intradomain_Enrichment', 'VR_mid_avg', 'VR_mid_max',
             'VR_intradomain_avg', 'VR_intradomain_max']
    num = [c for c in g['columns'] if c not in ('Domain', 'Mode')]
    df = pd.DataFrame(g['rows'])
    for c in num:
        df[c] = pd.to_numeric(df[c], errors='coerce')
    want = df[ratio].to_numpy(dtype=float)
    got = studies.add_ratio_metrics(df.drop(columns=ratio).copy())
    assert list(got.columns) == g['columns']
    G = got[ratio].to_numpy(dtype=float)
    sul = (df['Domain'] == 'sulcus').to_numpy()
    assert np.all(np.isnan(G[~sul])) and np.all(np.isnan(want[~sul]))
    assert np.allclose(G[sul], want[sul], rtol=1e-13, atol=0.0)


def test_derived_study_columns_reproduce_reference_csvs():
    """Derived columns of the Phase B and adv-diff validation CSVs (CR, flux_ratio, flux_error_pct) recomputed from
    the base columns of the reference's checked-in files with studies.phase_b_row / add_surrogate_errors."""
    import json
    import os
    import pandas as pd
    from sulcusfem import studies
    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'study_columns.json')))
    for r in g['no_adv_mu_sweep_results.csv']['rows']:
        row = studies.phase_b_row(r['geometry'], {'sulci_w_dim': float(r['width_mm']), 'sulci_h_dim': float(r['depth_mm']),
                                                  'aspect_ratio': float(r['aspect_ratio'])}, float(r['mu_factor']),
                                  float(r['avg_conc_sulc']), float(r['avg_conc_rect']), float(r['flux_sulc_y0']),
                                  float(r['flux_rect_bottom']))
        assert list(row) == g['no_adv_mu_sweep_results.csv']['columns']
        for k in ('CR', 'flux_ratio', 'flux_error_pct'):
            assert abs(row[k] - float(r[k])) <= 1e-12 * max(1.0, abs(float(r[k]))), (r['geometry'], k)
    ad = g['advdiff_validation_step_pe_x_mu.csv']
    df = pd.DataFrame(ad['rows'])
    for c in ad['columns']:
        if c not in ('domain_type', 'surrogate_type'):
            df[c] = pd.to_numeric(df[c], errors='coerce')
    want = df[['flux_error_pct', 'flux_ratio']].to_numpy(dtype=float)
    got = studies.add_surrogate_errors(df.drop(columns=['flux_error_pct', 'flux_ratio']).copy(), studies.PE_VALUES,
                                       studies.MU_FACTORS_ADV)
    assert list(got.columns) == ad['columns']
    G = got[['flux_error_pct', 'flux_ratio']].to_numpy(dtype=float)
    assert np.array_equal(np.isnan(G), np.isnan(want)) and np.allclose(G[~np.isnan(G)], want[~np.isnan(want)], rtol=1e-12)


def test_sample_mu_along_bottom(tmp_path):
    """analysis.sample_mu_along_bottom (reference analysis.py:838-882): constant and StepUptakeOpen coefficients."""
    from sulcusfem import analysis
    from sulcusfem.fem import Constant
    from sulcusfem.parameters import StepUptakeOpen

    class P:
        pass
    mesh = hm.rectangle_mesh(10.0, 1.0, 20, 2)
    p = P()
    p.mu = 0.7
    out = analysis.sample_mu_along_bottom({'params': p, 'mesh_results': {'mesh': mesh}}, n_points=11)
    assert np.array_equal(out['x'], np.linspace(0, 10, 11)) and np.all(out['mu'] == 0.7) and out['mu_mean'] == pytest.approx(0.7)
    p.mu = Constant(1.5)
    assert analysis.sample_mu_along_bottom({'params': p, 'mesh_results': {'mesh': mesh}}, 5)['mu_max'] == 1.5
    step = StepUptakeOpen(mu_base=1.0, mu_eff_target=1.77, sulcus_left_x=4.75, sulcus_right_x=5.25, L_c=0.05, Gamma=5.0)
    p.mu = step
    path = tmp_path / 'mu' / 'mu_samples.csv'
    out = analysis.sample_mu_along_bottom({'params': p, 'mesh_results': {'mesh': mesh}}, n_points=2001, save_csv_path=str(path))
    assert out['mu_min'] == 1.0 and out['mu_max'] == 1.77
    assert 1.0 < out['mu_mean'] < 1.0 + 0.77 * 0.5 / 10.0               # base + (open - base) * (less than w / L)
    mid = out['mu'][np.abs(out['x'] - 5.0) < 1e-9][0]
    assert mid == out['mu_max'] and os.path.exists(path)
    v = np.zeros(1)
    step.eval(v, np.array([5.0, 0.0]))
    assert v[0] == mid
    with pytest.raises(ValueError):
        analysis.sample_mu_along_bottom({'params': None})


def test_simulation_save_results_json(tmp_path):
    """simulation._simulation_save_results (reference simulation.py:235-262): keys of the JSON summary."""
    import json
    from sulcusfem import simulation
    from sulcusfem.parameters import Parameters
    p = Parameters(mode='no-adv', mesh_size_dim=0.1)
    p.validate()
    p.nondim()
    mesh = hm.rectangle_mesh(10.0, 1.0, 10, 2)
    res = {'params': p, 'mesh_results': {'mesh': mesh},
           'mass_metrics': {'total_mass': np.float64(1.5), 'profiles_full': {'x': np.arange(3.0)}},
           'flux_metrics': {'uptake_flux': 0.25, 'physical_flux': {'left': {'total': np.float64(-1.0)}}}}
    path = tmp_path / 'results.json'
    simulation._simulation_save_results(res, str(path))
    out = json.load(open(path))
    assert set(out) == {'params', 'mass_metrics', 'flux_metrics', 'mesh_info', 'mu_eff_comparison'}
    assert out['mesh_info']['num_cells'] == mesh.num_cells and out['mass_metrics']['total_mass'] == 1.5
    assert out['mass_metrics']['profiles_full']['x'] == [0.0, 1.0, 2.0] and out['mu_eff_comparison'] is None
    assert out['params']['mode'] == 'no-adv' if 'mode' in out['params'] else True


@pytest.mark.parametrize('domain', ['sulcus', 'rectangular'])
def test_refinement_factor_grades_the_mesh_like_the_threshold_field(domain):
    """reference mesh.py:266,330-337: lc_fine = lc / refinement_factor within w/10 of the sulcus nodes, lc beyond w/2,
    linear in between (Gmsh Distance + Threshold fields).  The graded mesher must follow that field (in octaves,
    never coarser), leave the far field at lc and keep the marker / boundary contracts of the ungraded mesher."""
    from sulcusfem import hostmesh as hm
    from sulcusfem.mesh import MeshGenerator
    from sulcusfem.unstructured import reference_sulcus_nodes, threshold_size_field
    L, H, w, d, lc, rf = 10.0, 1.0, 0.5, 1.0, 0.05, 4
    out = MeshGenerator(L, H, d, w, lc, rf, domain).generate_mesh()
    base = MeshGenerator(L, H, d, w, lc, 1, domain).generate_mesh()
    mesh = out['mesh']
    assert mesh.geometry['mesher'] == 'delaunay-graded' and mesh.num_cells > 1.1 * base['mesh'].num_cells
    assert mesh.num_cells < 0.5 * rf * rf * base['mesh'].num_cells          # local, not uniform, refinement
    p = mesh.coords[mesh.cells]
    e = np.stack([np.hypot(*(p[:, i] - p[:, (i + 1) % 3]).T) for i in range(3)], axis=1).mean(axis=1)
    nodes = reference_sulcus_nodes(L, w, d)
    s = threshold_size_field(p.mean(axis=1), nodes, lc, lc / rf, w / 10, w / 2)
    assert (e / s).max() < 1.5 and (e / s).min() > 0.35                     # follows the field within an octave
    from scipy.spatial import cKDTree
    dist, _ = cKDTree(nodes).query(p.mean(axis=1))
    assert abs(np.median(e[dist < w / 10]) / (lc / rf) - 1.0) < 0.15        # LcMin near the sulcus nodes
    assert abs(np.median(e[dist > w]) / lc - 1.0) < 0.15                    # LcMax in the far field
    area = np.abs(mesh.signed_areas())
    q = 4 * np.sqrt(3) * area / (np.stack([np.hypot(*(p[:, i] - p[:, (i + 1) % 3]).T) for i in range(3)], axis=1) ** 2).sum(axis=1)
    assert q.min() > 0.4 and q.mean() > 0.95
    expect = 10.0 + (2 * w * d / np.pi if domain == 'sulcus' else 0.0)
    assert abs(area.sum() - expect) < 5e-3
    # markers: same ids on the same boundary pieces
    bm = out['bc_markers'].values
    assert set(np.unique(bm[mesh.edge_on_boundary])) == {1, 2, 3, 4}
    if domain == 'sulcus':
        assert set(np.unique(out['domain_markers'].values)) == {1, 2}
        y0 = out['y0_markers'].values
        assert (y0 == 10).sum() > 0
    # the multigrid hierarchy coarsens a graded mesh with the same grading
    from sulcusfem.hierarchy import build_hierarchy
    hier = build_hierarchy(mesh)
    assert len(hier.meshes) >= 3 and hier.meshes[1].geometry.get('mesher') == 'delaunay-graded'
    assert hier.meshes[-1].num_vertices < 0.1 * mesh.num_vertices


def test_geometry_and_mu_eff_analysis_extractors_match_reference_schema():
    """Row extractors of the two Phase A studies added in round 2 (no_advection_analysis_A.py:165-296): column names and
    order of the reference's checked-in ``mu_eff_analysis_results.csv`` and of ``extract_geometry_analysis_data``; the
    host-only columns (parameters, mu(x) samples, closed forms) reproduce the reference rows bit for bit."""
    import json
    from types import SimpleNamespace
    from sulcusfem import studies, hostmesh as hm
    from sulcusfem.parameters import Parameters
    g = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'study_columns.json')))['mu_eff_analysis_results.csv']
    mesh = hm.rectangle_mesh(10.0, 1.0, 20, 2)
    for ref in g['rows']:
        f = float(ref['Mu_Factor'])
        p = Parameters(mode='no-adv')
        p.sulci_w_dim, p.sulci_h_dim = 0.5, 1.0
        p.mu_dim = float(getattr(Parameters, 'MU_DIM_NO_ADV')) * f
        p.validate()
        p.nondim()
        fake = {'params': p, 'mesh_results': {'mesh': mesh},
                'mu_eff_comparison': {'mu_eff_sim': 1.0, 'mu_eff_arc': 2.0, 'mu_eff_enh': 3.0, 'mu_eff_open': 4.0,
                                      'ratios': {'sim': 1.0, 'arc': 2.0, 'enh': 3.0, 'open': 4.0}}}
        row = studies.extract_mu_eff_analysis_data(fake, ref['Config'], p.mu_dim, f)
        assert list(row.keys()) == g['columns']
        assert ref['Config'] == f"mu_eff_analysis_mu_{f}x"
        for k in ('Mu_Value', 'Sulcus_Width_mm', 'Sulcus_Depth_mm', 'Domain_Length_mm', 'L_ref', 'L_nondim', 'H_nondim',
                  'Sulcus_W_nondim', 'Sulcus_H_nondim', 'Mu_base_nondim', 'Mu_Min_Bottom', 'Mu_Max_Bottom'):
            assert float(row[k]) == float(ref[k]), (k, row[k], ref[k])
        assert abs(float(row['Mu_Mean_Bottom']) - float(ref['Mu_Mean_Bottom'])) < 1e-14 * max(1.0, f)
        assert row['Mu_X_Array'] == ref['Mu_X_Array'] and row['Mu_Values_Array'] == ref['Mu_Values_Array']
    cfg = {'sulci_w_dim': 0.5, 'sulci_h_dim': 1.0, 'aspect_ratio_category': 'deep'}
    row = studies.extract_geometry_analysis_data(
        {'mu_eff_comparison': fake['mu_eff_comparison'], 'mass_metrics': {'total_mass': 1.0},
         'flux_metrics': {'sulcus_specific': {'physical_flux': {'sulcus_opening': {'total': 0.5}}}}},
        'g_mu_1.0', 'g', 3e-4, 1.0, {'geometry_config': cfg})
    assert list(row.keys()) == ['Config', 'Geometry_Name', 'Mu_Value', 'Mu_Factor', 'Sulcus_Width_mm', 'Sulcus_Depth_mm',
                                'Aspect_Ratio', 'Aspect_Ratio_Category', 'Mu_Eff_Simulation', 'Mu_Eff_Analytical',
                                'Mu_Eff_Enhanced', 'Mu_Eff_Opening', 'Ratio_Sim', 'Ratio_Analytical', 'Ratio_Enhanced',
                                'Ratio_Opening', 'Relative_Error_Analytical', 'Relative_Error_Enhanced',
                                'Relative_Error_Opening', 'Total_Mass', 'Mouth_Flux_Total']
    assert row['Aspect_Ratio'] == 2.0 and row['Mouth_Flux_Total'] == 0.5


def test_dolfin_adapter_matches_dofs_and_facets_by_geometry():
    """results are copied into genuine dolfin Functions by DOF-coordinate matching and dolfin facet
    MeshFunctions are re-indexed by vertex pair.  Exercised with stand-in dolfin objects whose DOF / facet numberings are
    random permutations (real dolfin's numbering is build dependent; dolfin itself is not installable here)."""
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import fake_dolfin as fd
    from sulcusfem import dofmap as dm, dolfin_adapter as da, hostmesh as hm
    from sulcusfem.fem import Function, FunctionSpace, MixedElement, VectorFunctionSpace
    from sulcusfem.unstructured import mesh_domain
    rng = np.random.default_rng(7)
    host0 = mesh_domain(10.0, 1.0, 0.5, 1.0, 0.2, 'sulcus')
    dmesh = fd.FakeMesh(host0, rng)
    host = da.host_mesh_from_dolfin(dmesh)
    assert np.array_equal(host.cells, host0.cells) and np.array_equal(host.coords, host0.coords)
    assert da.host_mesh_from_dolfin(dmesh) is host                          # cached on the dolfin object
    # facet markers
    mk = hm.build_markers(host, 10.0, 1.0, 4.75, 5.25, 'sulcus')
    for key in ('bc_markers', 'bottom_segment_markers', 'y0_markers'):
        mf = fd.facet_function(dmesh, mk[key].values)
        assert np.array_equal(da.marker_values(mf, dmesh, host), mk[key].values)
    # spaces: P2 scalar, P2 vector (interleaved / permuted components), Taylor-Hood
    P2, P1 = dm.p2_dof_coordinates(host), host.coords
    n2, nv = len(P2), len(P1)
    for kind, blocks, ours_space in (('P2', [P2], FunctionSpace(host, 'CG', 2)), ('P1', [P1], FunctionSpace(host, 'P', 1)),
                                     ('P2v', [P2, P2], VectorFunctionSpace(host, 'P', 2))):
        V = fd.FakeSpace(dmesh, blocks, rng)
        theirs = da.dof_map(V, host, kind)
        assert np.array_equal(theirs, V.perm)
        f = Function(ours_space, rng.random(ours_space.dim()))
        g = da.to_dolfin(f, V, make_function=fd.FakeFunction)
        assert np.array_equal(g.vector().get_local()[V.perm], f.values)
        back = da.from_dolfin(g, host, kind)
        assert np.array_equal(back.values, f.values)
    W = fd.FakeSpace(dmesh, [P2, P2, P1], rng)
    assert np.array_equal(da.dof_map(W, host, 'TH'), W.perm)
    Vc = W.sub(0).collapse()
    assert Vc.dim() == 2 * n2 and np.array_equal(da.dof_map(Vc, host, 'P2v'), Vc.perm)
    # a space on a different mesh must be refused
    other = fd.FakeSpace(dmesh, [P2 + 1e-3], rng)
    with pytest.raises(ValueError):
        da.dof_map(other, host, 'P2')
    # install(): the reference's `from solvers import ...` resolves to the adapter
    saved = sys.modules.get('solvers')
    try:
        mod = da.install(make_function=fd.FakeFunction)
        import solvers as s
        assert s is mod and all(hasattr(s, n) for n in ('stokes_solver', 'stokes_solver_no_adv', 'pure_diffusion_solver',
                                                        'pure_diffusion_solver_variable_mu', 'advdiff_solver',
                                                        'advdiff_solver_variable_mu'))
        u0, p0 = s.stokes_solver_no_adv(fd.FakeSpace(dmesh, [P2, P2], rng), fd.FakeSpace(dmesh, [P1], rng))
        assert not u0.vector().get_local().any()
    finally:
        if saved is not None:
            sys.modules['solvers'] = saved
        else:
            sys.modules.pop('solvers', None)


def test_frozen_coarse_levels_policy():
    """solvers.frozen_coarse_levels: the multigrid levels of a mu sweep are reused only inside the context, only for a
    mu within the given factor of the mu they were assembled for, and the setting is per thread (sweep workers)."""
    import threading
    from types import SimpleNamespace
    from sulcusfem import solvers
    prob = SimpleNamespace(_coarse_mu=2.0)
    assert not solvers._reuse_coarse(prob, 2.0)                       # outside the context: always a full assembly
    with solvers.frozen_coarse_levels(4.0):
        assert solvers._reuse_coarse(prob, 2.0) and solvers._reuse_coarse(prob, 8.0) and solvers._reuse_coarse(prob, 0.5)
        assert not solvers._reuse_coarse(prob, 8.1) and not solvers._reuse_coarse(prob, 0.49)
        assert not solvers._reuse_coarse(prob, 0.0) and not solvers._reuse_coarse(SimpleNamespace(), 2.0)
        seen = []
        t = threading.Thread(target=lambda: seen.append(solvers._reuse_coarse(prob, 2.0)))
        t.start(); t.join()
        assert seen == [False]                                        # another thread has its own setting
        with solvers.frozen_coarse_levels(1.0):
            assert solvers._reuse_coarse(prob, 2.0) and not solvers._reuse_coarse(prob, 2.1)
        assert solvers._reuse_coarse(prob, 8.0)                       # restored
    assert not solvers._reuse_coarse(prob, 2.0)
