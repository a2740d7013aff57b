"""CPU emulation of the control flow of k_sell_stream (csrc/sfem_spmv_sell.cu) on the host plan of sulcusfem/sell.py.

The CUDA kernel gives every warp a run of consecutive fine parts (one contiguous span of the mirror), walks it in chunks
of U column-steps that ignore slice boundaries, detects the end of a slice per step and swallows slices without entries
(leading, in the middle, trailing).  This test replays exactly that traversal in Python -- same variables, same order of
operations -- and checks y = A x, so the slice / part / chunk logic is covered without a GPU, including the corner cases
real finite-element matrices never produce (fully empty slices, more warps than slices, a single part)."""
import numpy as np
import pytest
import scipy.sparse as sp

from sulcusfem import sell as sl


def emulate(plan, parts, per, vals, x, U=4):
    """One 'warp' per run of `per` fine parts; returns y (rows without a lane stay NaN-free: every row is owned)."""
    y = np.full(plan.nrows, np.nan)
    written = np.zeros(plan.nrows, dtype=int)
    sp_, perm, scols, src = plan.slice_ptr.astype(np.int64), plan.perm, plan.scols, plan.src
    svals = np.where(src >= 0, np.asarray(vals)[np.maximum(src, 0)], 0.0)          # k_sell_pack
    nwarps = (len(parts) - 1) // per
    for gw in range(nwarps):
        s, s_end = int(parts[gw * per]), int(parts[(gw + 1) * per])
        if not s < s_end:
            continue
        g = sp_[s] >> 5
        g_end = sp_[s_end] >> 5
        e = sp_[s + 1] >> 5
        row = perm[s * 32:(s + 1) * 32].copy()
        ne, nrow = e, np.full(32, -1)
        if s + 1 < s_end:
            ne, nrow = sp_[s + 2] >> 5, perm[(s + 1) * 32:(s + 2) * 32].copy()
        acc = np.zeros(32)
        st = dict(s=s, e=e, row=row, ne=ne, nrow=nrow, acc=acc)

        def advance(gnext):
            while True:
                ok = st['row'] >= 0
                y[st['row'][ok]] = st['acc'][ok]
                written[st['row'][ok]] += 1
                st['s'] += 1
                st['row'], st['e'] = st['nrow'], st['ne']
                st['acc'] = np.zeros(32)
                if st['s'] >= s_end:
                    break
                if st['s'] + 1 < s_end:
                    st['ne'] = sp_[st['s'] + 2] >> 5
                    st['nrow'] = perm[(st['s'] + 1) * 32:(st['s'] + 2) * 32].copy()
                if st['e'] != gnext:
                    break
        while st['s'] < s_end and st['e'] == g:
            advance(g)
        while g < g_end:
            for u in range(U):
                if g + u < g_end:
                    k = (g + u) * 32 + np.arange(32)
                    c = scols[k]
                    st['acc'] = st['acc'] + np.where(c >= 0, svals[k] * x[np.maximum(c, 0)], 0.0)
                    if g + u + 1 == st['e']:
                        advance(g + u + 1)
            g += U
        while st['s'] < s_end:
            advance(g_end)
    return y, written


def _check(A, sigma, nparts, per, U=4):
    A = sp.csr_matrix(A)
    A.sort_indices()
    plan = sl.build_plan(A.indptr, A.indices, sigma)
    parts = sl.partition(plan, nparts)
    rng = np.random.default_rng(0)
    x = rng.random(A.shape[1])
    y, written = emulate(plan, parts, per, A.data, x, U)
    assert np.all(written == 1), "every row must be finished exactly once"
    ref = A @ x
    assert np.allclose(y, ref, rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize('sigma', [32, 256])
@pytest.mark.parametrize('nparts,per', [(12, 3), (12, 4), (96, 3), (1, 1), (600, 3)])
def test_stream_traversal_matches_matvec(sigma, nparts, per):
    rng = np.random.default_rng(1)
    # FEM-like: short, nearly uniform rows
    n = 1000
    A = sp.random(n, n, 0.012, random_state=3, format='lil')
    A.setdiag(1.0)
    _check(A, sigma, nparts, per)
    # ragged with fully empty slices at the start, in the middle and at the end (only visible with sigma = 32;
    # with sorting the empty rows collect at the end of every window)
    B = sp.random(700, 900, 0.01, random_state=4, format='lil')
    B[0:96, :] = 0
    B[300:364, :] = 0
    B[636:700, :] = 0
    B[200, :40] = 1.0                                     # one long row
    _check(B, sigma, nparts, per)
    # fewer rows than one slice, and a matrix without any entry
    _check(sp.random(7, 5, 0.5, random_state=5), sigma, nparts, per)
    _check(sp.csr_matrix((50, 50)), sigma, nparts, per)
    # other chunk lengths
    _check(A, sigma, nparts, per, U=3)
    _check(B, sigma, nparts, per, U=2)
