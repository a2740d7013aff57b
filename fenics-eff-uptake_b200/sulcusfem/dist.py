"""Multi-GPU solve path for large refined meshes (BASELINE config 5): one process per GPU.

The reference is serial.  Here the P2 concentration system and its multigrid hierarchy are
row-partitioned in x-slabs over the ranks; halo exchange and the Krylov all-reduces run as kernels that
store straight into the peers' memory over NVLink (``csrc/sfem_dist.cu``) -- ``torch.distributed`` is
used once, to all-gather the 64-byte CUDA-IPC handles of the mailboxes.

Scope: all three solves of the path -- multigrid-preconditioned CG (pure diffusion / Robin), FGMRES
(advection-diffusion, ``solvers.py:16-57``) and the Taylor-Hood MINRES (``solvers.py:237-306``) with its
2-right-hand-side velocity multigrid, pressure-mass Chebyshev sweep and lubrication coarse correction.
Assembly is replicated (every rank assembles the global matrices with the single-GPU kernels and extracts
its rows with ``sfem_csr_extract``); the solves are distributed; fields are all-gathered (NCCL, the one
real exchange of full vectors) where the next replicated stage needs them.  Levels with fewer than
``replicate_below`` unknowns are kept on every rank and solved redundantly after a vector all-reduce of
the restricted residual.  One mailbox (:class:`DistWorld`) serves every problem of the process.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np

from . import capi
from . import dofmap as dm
from . import partition as pt
from .device import Context, DeviceCsr, Multigrid, P, ScalarProblem, _torch


class DistContext:
    """Mailbox communicator of this rank (``sfem_dist_t``)."""

    def __init__(self, ctx: Context, rank: int, nranks: int, mailbox_words: int, vec_cap: int, group=None, emulate=False):
        self.ctx, self.rank, self.nranks = ctx, int(rank), int(nranks)
        lib = ctx.lib
        self.handle = lib.sfem_dist_create(self.rank, self.nranks, int(mailbox_words), int(vec_cap))
        if not self.handle:
            raise capi.SulcusFemError("sfem_dist_create failed: " + lib.sfem_last_error().decode())
        if self.nranks > 1 and not emulate:
            import torch.distributed as dist
            torch = _torch()
            buf = (C.c_ubyte * 64)()
            capi.check(lib.sfem_dist_ipc_handle(self.handle, buf), 'sfem_dist_ipc_handle')
            mine = torch.tensor(list(buf), dtype=torch.uint8, device=ctx.device)
            allh = [torch.zeros(64, dtype=torch.uint8, device=ctx.device) for _ in range(self.nranks)]
            dist.all_gather(allh, mine, group=group)
            raw = np.concatenate([t.cpu().numpy() for t in allh]).astype(np.uint8)
            self._handles = raw.tobytes()
            capi.check(lib.sfem_dist_open_peers(self.handle, self._handles), 'sfem_dist_open_peers')
            dist.barrier(group=group)
        lib.sfem_dist_activate(self.handle)

    def activate(self):
        self.ctx.lib.sfem_dist_activate(self.handle)

    def error(self) -> bool:
        return bool(self.ctx.lib.sfem_dist_error(self.handle))

    def close(self):
        if self.handle:
            self.ctx.lib.sfem_dist_activate(None)
            self.ctx.lib.sfem_dist_destroy(self.handle)
            self.handle = None


class DeviceHalo:
    def __init__(self, ctx: Context, lp: pt.LevelPartition):
        self.ctx, self.lp = ctx, lp
        nn = len(lp.neighbors)
        i32 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int32))
        i64 = lambda a: np.ascontiguousarray(np.asarray(a, dtype=np.int64))
        send_ptr = np.concatenate([[0], np.cumsum([len(s) for s in lp.send_idx])]).astype(np.int32) if nn else np.zeros(1, np.int32)
        send_idx = np.concatenate(lp.send_idx).astype(np.int32) if nn else np.zeros(1, np.int32)
        arrs = [i32(lp.neighbors), i32(send_ptr), i32(send_idx), i32(lp.recv_cnt), i32(lp.recv_off), i64(lp.peer_data_off),
                i64(lp.peer_flag_off), i64(lp.my_data_off), i64(lp.my_flag_off), i64(lp.cap)]
        ptrs = [a.ctypes.data_as(C.c_void_p) for a in arrs]
        self.handle = ctx.lib.sfem_halo_create(nn, lp.n_own, lp.n_loc, *ptrs)
        if not self.handle:
            raise capi.SulcusFemError("sfem_halo_create failed: " + ctx.lib.sfem_last_error().decode())

    def attach(self, csr: DeviceCsr):
        capi.check(self.ctx.lib.sfem_halo_attach(P(csr.rowptr), self.handle), 'sfem_halo_attach')

    def exchange(self, x, nb=1, phase=0):
        capi.check(self.ctx.lib.sfem_halo_exchange(self.handle, P(x), nb, phase, self.ctx.stream), 'sfem_halo_exchange')


class _LocalOp:
    """Rows of a global CSR operator owned by this rank, columns in local numbering, values extracted on the device."""

    def __init__(self, ctx, global_csr: DeviceCsr, g_rowptr, g_cols, rows, col_g2l, ncols_loc, drop_missing=False):
        rp, lc, slot = pt.localize_csr(g_rowptr, g_cols, rows, col_g2l, drop_missing=drop_missing)
        self.csr = DeviceCsr(ctx, len(rows), ncols_loc, rp, lc)
        self.slot_host = slot
        self.slot = ctx.up(slot, np.int32)
        self.src = global_csr
        self.ctx = ctx

    def refresh(self):
        capi.check(self.ctx.lib.sfem_csr_extract(self.csr.nnz, P(self.slot), P(self.src.vals), P(self.csr.vals), self.ctx.stream),
                   'sfem_csr_extract')


class DistWorld:
    """The communicator of this process: ONE mailbox shared by every row-partitioned problem.  Partitions are
    planned first (host only; every rank computes the channel layout of all ranks), then :meth:`commit` allocates
    the mailbox, exchanges the IPC handles and activates the communicator."""

    def __init__(self, ctx: Context, rank: int, nranks: int, vec_cap: int, group=None, emulate=False):
        self.ctx, self.rank, self.nranks, self.group, self.emulate = ctx, int(rank), int(nranks), group, emulate
        self.vec_cap = int(vec_cap)
        header = int(ctx.lib.sfem_dist_header_words(self.nranks, self.vec_cap))
        self.base = [header] * self.nranks
        self.dist: Optional[DistContext] = None

    def plan(self, owner, ghosts, nb_max=1, gap=0) -> pt.LevelPartition:
        if self.dist is not None:
            raise capi.SulcusFemError("DistWorld.plan after commit")
        lp = pt.partition_level(owner, self.nranks, self.rank, ghosts, self.base, nb_max=nb_max, ghost_gap=gap)
        self.base = list(lp.all_mailbox_ends)
        return lp

    def commit(self):
        if self.dist is None:
            self.dist = DistContext(self.ctx, self.rank, self.nranks, int(max(self.base)) + 8, self.vec_cap,
                                    group=self.group, emulate=self.emulate)
        return self.dist

    def error(self) -> bool:
        return self.dist is not None and self.dist.error()

    def close(self):
        if self.dist is not None:
            self.dist.close()
            self.dist = None


def level_coordinates(prob: ScalarProblem, l: int) -> np.ndarray:
    lev = prob.levels[l]
    return dm.p2_dof_coordinates(lev.mesh) if lev.degree == 2 else lev.mesh.coords


def num_distributed_levels(prob: ScalarProblem, replicate_below: int) -> int:
    nl = len(prob.levels)
    nd = 0
    while nd < nl - 1 and prob.levels[nd].n >= replicate_below:
        nd += 1
    return max(nd, 1)                                       # at least the system level is distributed


def plan_partitions(prob: ScalarProblem, nranks: int, replicate_below: int = 20000, extra_ops0=None):
    """Owner arrays + ghost sets of the row-partitioned levels (identical on every rank).  ``extra_ops0(owner0)``:
    further operators whose columns live on the system level (the Stokes divergence block)."""
    nd = num_distributed_levels(prob, replicate_below)
    owners = []
    for l in range(nd):
        X = level_coordinates(prob, l)
        owners.append(pt.slab_owner(X[:, 0], nranks, X[:, 1]))
    ghosts = []
    for l in range(nd):
        lev = prob.levels[l]
        ops = [(lev.pattern.rowptr, lev.pattern.cols, owners[l])]
        if l + 1 < nd:                                      # restriction l -> l+1: rows on l+1, columns on l
            T = prob.hierarchy.transfers[l]
            ops.append((T.t_rowptr, T.t_cols, owners[l + 1]))
        if l > 0:                                           # prolongation l -> l-1: rows on l-1, columns on l
            T = prob.hierarchy.transfers[l - 1]
            ops.append((T.rowptr, T.cols, owners[l - 1]))
        if l == 0 and extra_ops0 is not None:
            ops.extend(extra_ops0(owners[0]))
        ghosts.append(pt.ghost_sets(owners[l], nranks, ops))
    return nd, owners, ghosts


def mailbox_layout(prob: ScalarProblem, nranks: int, nd: int, owners, ghosts, header_words: int, nb_max=1):
    """LevelPartition of every (level, rank) + the mailbox size (same on every rank)."""
    base = [header_words] * nranks
    parts = []
    for l in range(nd):
        row = [pt.partition_level(owners[l], nranks, r, ghosts[l], base, nb_max=nb_max) for r in range(nranks)]
        base = row[0].all_mailbox_ends
        parts.append(row)
    return parts, int(max(base)) + 8


def _gather_map(ctx, idx):
    return ctx.up(np.ascontiguousarray(idx, dtype=np.int32), np.int32)


def _gather(ctx, n, slot, src, dst):
    """dst[i] = src[slot[i]] (device)."""
    capi.check(ctx.lib.sfem_csr_extract(int(n), P(slot), P(src), P(dst), ctx.stream), 'sfem_csr_extract')


class AllGather:
    """All-gather of the ranks' owned entries into full (replicated) vectors: one NCCL all-gather of equally padded
    segments, then gathers by precomputed position maps.  ``owned_lists[q]``: global ids in the order rank q stores
    its owned entries; ``fields``: {name: (offset of the field in a rank's segment as a function of q, global ids per
    rank)} -- built by the problems below."""

    def __init__(self, ctx, nranks, seg_len, group=None):
        torch = _torch()
        self.ctx, self.nranks, self.group = ctx, nranks, group
        self.seg = int(max(seg_len))
        self.send = torch.zeros(self.seg, dtype=torch.float64, device=ctx.device)
        self.recv = torch.zeros(self.seg * nranks, dtype=torch.float64, device=ctx.device)
        self.maps = {}

    def add_field(self, name, n_global, pos_of_global):
        """pos_of_global[g] = index into the concatenated receive buffer of global entry g."""
        assert len(pos_of_global) == n_global and pos_of_global.min() >= 0
        self.maps[name] = (int(n_global), _gather_map(self.ctx, pos_of_global),
                           self.ctx.zeros(n_global))

    def run(self):
        if self.nranks > 1:
            import torch.distributed as dist
            dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            capi.check(self.ctx.lib.sfem_vec_copy(self.seg, P(self.send), P(self.recv), self.ctx.stream), 'sfem_vec_copy')
        out = {}
        for name, (n, pos, dst) in self.maps.items():
            _gather(self.ctx, n, pos, self.recv, dst)
            out[name] = dst
        return out


def _flatten_contribs(contrib_ptr, contrib_code, slot):
    """Contribution lists of the CSR slots ``slot``: (lengths per slot, element-buffer codes, concatenated)."""
    cp = np.asarray(contrib_ptr, dtype=np.int64)
    slot = np.asarray(slot, dtype=np.int64)
    lens = cp[slot + 1] - cp[slot]
    tot = int(lens.sum())
    idx = np.repeat(cp[slot], lens) + (np.arange(tot, dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens))
    return lens, np.asarray(contrib_code)[idx].astype(np.int64)


def plan_local_gather(ops, nce, cell_family_end=None):
    """Gather maps of the owned rows of one or more operators that read ONE element buffer, re-addressed to a buffer
    that holds only the cells those rows touch.  ``ops``: list of ``(lens, codes)`` from :func:`_flatten_contribs`;
    ``nce`` doubles per cell; codes >= ``cell_family_end`` belong to a further family (Robin facets) that is kept whole
    behind the cells.  Returns ``(cells, [(contrib_ptr, contrib_code), ...])``.  The order of the contributions of a
    slot is untouched, so the local assembly sums exactly what the global one sums (bit-identical values)."""
    allc = []
    for _, codes in ops:
        c = codes if cell_family_end is None else codes[codes < cell_family_end]
        allc.append(np.unique(c // nce))
    cells = np.unique(np.concatenate(allc)) if allc else np.zeros(0, dtype=np.int64)
    out = []
    for lens, codes in ops:
        new = np.empty_like(codes)
        is_cell = np.ones(len(codes), dtype=bool) if cell_family_end is None else codes < cell_family_end
        cc = codes[is_cell]
        new[is_cell] = np.searchsorted(cells, cc // nce) * nce + cc % nce
        if cell_family_end is not None:
            new[~is_cell] = len(cells) * nce + (codes[~is_cell] - cell_family_end)
        ptr = np.concatenate([[0], np.cumsum(lens)])
        if len(new) and (new.max() >= 2 ** 31 or ptr[-1] >= 2 ** 31):
            raise ValueError("local element buffer exceeds int32 addressing")
        out.append((ptr.astype(np.int32), new.astype(np.int32)))
    return cells, out


class _LocalAssembly:
    """Assembly of the OWNED rows of one scalar level straight into the rank's local operator: the element kernel runs
    over the cells that touch an owned row only (1/N of the mesh + one cell layer), the gather map is the global one
    restricted to the owned rows and re-addressed -- no rank assembles the global matrix (strong scaling of the
    assembly phase of BASELINE config 5)."""

    def __init__(self, ctx, lev, lp, op: _LocalOp):
        from .device import cell_geometry
        self.ctx, self.lev, self.lp, self.op = ctx, lev, lp, op
        pat = lev.pattern
        self.nce = lev.ndof_cell ** 2
        fb = int(lev.facet_base)
        lens, codes = _flatten_contribs(pat.contrib_ptr, pat.contrib_code, op.slot_host)
        cells, maps = plan_local_gather([(lens, codes)], self.nce, fb if lev.nf else None)
        self.cells = cells
        self.nc = len(cells)
        self.contrib_ptr = ctx.up(maps[0][0], np.int32)
        self.contrib_code = ctx.up(np.concatenate([maps[0][1], [0]]), np.int32)
        self.geo = ctx.up(np.ascontiguousarray(cell_geometry(lev.mesh)[:, cells]), np.float64)
        cd = dm.p2_cell_dofs(lev.mesh) if lev.degree == 2 else dm.p1_cell_dofs(lev.mesh)
        self.celldofs = ctx.up(np.ascontiguousarray(cd[cells].T), np.int32)      # GLOBAL dof ids: velocity / mu are read from global vectors
        self.facet_len = int(pat.buffer_len - fb) if lev.nf else 0
        self.E = ctx.zeros(max(self.nc * self.nce + self.facet_len, 1))
        # Dirichlet data in the local column numbering (owned | hole | ghosts)
        l2g = np.full(lp.n_loc, -1, dtype=np.int64)
        present = np.flatnonzero(lp.g2l >= 0)
        l2g[lp.g2l[present]] = present
        self.l2g = l2g
        flag = np.where(l2g >= 0, lev.bc_flag_host[np.maximum(l2g, 0)], 0).astype(np.uint8)
        self.flag = ctx.up(flag, np.uint8)
        self.val_map = _gather_map(ctx, np.maximum(l2g, 0))
        self.val = ctx.zeros(lp.n_loc)
        self.rhs = ctx.zeros(max(lp.n_own, 1))

    def assemble(self, D, ux=None, uy=None, mu_const=0.0, mu_nodal=None, clamp=False, upwind=False, robin=True):
        ctx, lib, lev, A = self.ctx, self.ctx.lib, self.lev, self.op.csr
        if self.nc:
            if lev.degree == 2:
                capi.check(lib.sfem_elem_p2_advdiff(self.nc, P(self.geo), P(self.celldofs), float(D), P(ux), P(uy), P(self.E),
                                                    ctx.stream), 'sfem_elem_p2_advdiff (local)')
            else:
                capi.check(lib.sfem_elem_p1_advdiff(self.nc, P(self.geo), P(self.celldofs), float(D), P(ux), P(uy),
                                                    int(bool(upwind)), P(self.E), ctx.stream), 'sfem_elem_p1_advdiff (local)')
        if lev.nf:
            F = self.E[self.nc * self.nce:]
            if robin:
                fn = lib.sfem_facet_p2_robin if lev.degree == 2 else lib.sfem_facet_p1_robin
                capi.check(fn(lev.nf, P(lev.fgeo), P(lev.fdofs), float(mu_const), P(mu_nodal), int(bool(clamp)), P(F),
                              ctx.stream), 'sfem_facet_robin (local)')
            else:
                capi.check(lib.sfem_vec_set(self.facet_len, 0.0, P(F), ctx.stream), 'sfem_vec_set')
        capi.check(lib.sfem_gather_csr(A.nnz, P(self.contrib_ptr), P(self.contrib_code), P(self.E), P(A.vals), ctx.stream),
                   'sfem_gather_csr (local)')

    def apply_bc(self, rhs=None, with_values=True):
        """Symmetric elimination of the owned rows (ghost columns included); ``with_values``: the level's Dirichlet
        values (system level), else zeros (multigrid levels, Stokes velocity block)."""
        ctx, lib, A = self.ctx, self.ctx.lib, self.op.csr
        if with_values:
            _gather(ctx, self.lp.n_loc, self.val_map, self.lev.bc_val, self.val)
        else:
            capi.check(lib.sfem_vec_set(self.lp.n_loc, 0.0, P(self.val), ctx.stream), 'sfem_vec_set')
        rhs = self.rhs if rhs is None else rhs
        capi.check(lib.sfem_apply_dirichlet(self.lp.n_own, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), P(rhs), P(self.flag),
                                            P(self.val), 1, ctx.stream), 'sfem_apply_dirichlet (local)')


class DistScalarProblem:
    """Row-partitioned view of a :class:`ScalarProblem` (which every rank holds and assembles in full).

    ``nb = 2``: the velocity block of the Taylor-Hood solver (two interleaved right-hand sides per V-cycle).
    ``world``: shared communicator (planning only; call :meth:`finalize` after ``world.commit()``); without one the
    problem owns a private communicator and is ready after construction."""

    def __init__(self, prob: ScalarProblem, rank: int, nranks: int, group=None, replicate_below: int = 20000,
                 emulate=False, world: Optional[DistWorld] = None, nb: int = 1, plan=None, level0_gap: int = 0):
        self.prob, self.ctx = prob, prob.ctx
        ctx, lib = self.ctx, self.ctx.lib
        self.rank, self.nranks, self.nb, self.group = rank, nranks, int(nb), group
        nd, owners, ghosts = plan if plan is not None else plan_partitions(prob, nranks, replicate_below)
        self.nd, self.owners = nd, owners
        self.n_tail = prob.levels[nd].n
        self._own_world = world is None
        self.world = world if world is not None else DistWorld(ctx, rank, nranks, self.n_tail * self.nb, group=group,
                                                               emulate=emulate)
        if self.world.vec_cap < self.n_tail * self.nb:
            raise capi.SulcusFemError("DistWorld.vec_cap is smaller than the replicated coarse level of this problem")
        self.parts = [self.world.plan(owners[l], ghosts[l], nb_max=self.nb, gap=(level0_gap if l == 0 else 0))
                      for l in range(nd)]
        self.mg = None
        self.halos = []
        if self._own_world:
            self.world.commit()
            self.finalize()

    @property
    def dist(self):
        return self.world.dist

    def finalize(self):
        """Device side (after the communicator exists): halos, local operators, multigrid handles."""
        prob, ctx, lib, nd, nb = self.prob, self.ctx, self.ctx.lib, self.nd, self.nb
        n_tail = self.n_tail
        self.halos = [DeviceHalo(ctx, lp) for lp in self.parts]
        H = prob.hierarchy
        self.A: List[_LocalOp] = []
        self.Pm: List[_LocalOp] = []
        self.Rm: List[_LocalOp] = []
        for l in range(nd):
            lp, lev = self.parts[l], prob.levels[l]
            A = _LocalOp(ctx, lev.A, lev.pattern.rowptr, lev.pattern.cols, lp.owned, lp.g2l, lp.n_loc)
            self.halos[l].attach(A.csr)
            self.A.append(A)
            T, dT = H.transfers[l], prob.transfers[l]
            if l + 1 < nd:
                lc = self.parts[l + 1]
                Pl = _LocalOp(ctx, dT.P, T.rowptr, T.cols, lp.owned, lc.g2l, lc.n_loc)
                self.halos[l + 1].attach(Pl.csr)            # reads the coarse vector with its ghosts
                Rl = _LocalOp(ctx, dT.R, T.t_rowptr, T.t_cols, lc.owned, lp.g2l, lp.n_loc)
                self.halos[l].attach(Rl.csr)
            else:
                # transition to the replicated hierarchy: P reads the full coarse vector; R keeps only the
                # owned columns (its partial results are summed over the ranks)
                ident = np.arange(n_tail, dtype=np.int64)
                Pl = _LocalOp(ctx, dT.P, T.rowptr, T.cols, lp.owned, ident, n_tail)
                own_only = np.where((lp.g2l >= 0) & (lp.g2l < lp.n_own), lp.g2l, -1)
                Rl = _LocalOp(ctx, dT.R, T.t_rowptr, T.t_cols, ident, own_only, lp.n_own, drop_missing=True)
            self.Pm.append(Pl)
            self.Rm.append(Rl)
        # multigrid handles: replicated tail (global levels nd..) and the row-partitioned head
        self.tail = Multigrid(ctx, prob.levels[nd:], prob.transfers[nd:], prob.mg_cheb_degree, prob.mg_eig_ratio, nb)
        nlv = nd
        IntArr, PtrArr = C.c_int * nlv, C.c_void_p * nlv

        def parr(ts):
            return PtrArr(*([t.data_ptr() for t in ts] + [None] * (nlv - len(ts))))
        k = dict(n=IntArr(*[p_.n_own for p_ in self.parts]), annz=IntArr(*[a.csr.nnz for a in self.A]),
                 arp=parr([a.csr.rowptr for a in self.A]), ac=parr([a.csr.cols for a in self.A]), av=parr([a.csr.vals for a in self.A]),
                 pnnz=IntArr(*([p_.csr.nnz for p_ in self.Pm[:-1]] + [0])),
                 prp=parr([p_.csr.rowptr for p_ in self.Pm[:-1]]), pc=parr([p_.csr.cols for p_ in self.Pm[:-1]]),
                 pv=parr([p_.csr.vals for p_ in self.Pm[:-1]]),
                 rrp=parr([r_.csr.rowptr for r_ in self.Rm[:-1]]), rc=parr([r_.csr.cols for r_ in self.Rm[:-1]]),
                 rv=parr([r_.csr.vals for r_ in self.Rm[:-1]]))
        self._keep = k
        self.mg = lib.sfem_mg_create(nlv, k['n'], k['annz'], k['arp'], k['ac'], k['av'], k['pnnz'], k['prp'], k['pc'], k['pv'],
                                     k['rrp'], k['rc'], k['rv'], None, prob.mg_cheb_degree, float(prob.mg_eig_ratio), nb)
        if not self.mg:
            raise capi.SulcusFemError("sfem_mg_create (distributed) failed: " + lib.sfem_last_error().decode())
        Pl, Rl = self.Pm[-1].csr, self.Rm[-1].csr
        capi.check(lib.sfem_mg_set_tail(self.mg, self.tail.handle, n_tail, Pl.nnz, P(Pl.rowptr), P(Pl.cols), P(Pl.vals),
                                        Rl.nnz, P(Rl.rowptr), P(Rl.cols), P(Rl.vals)), 'sfem_mg_set_tail')
        self.asm = [_LocalAssembly(ctx, prob.levels[l], self.parts[l], self.A[l]) for l in range(nd)]
        for ops in (self.Pm, self.Rm):            # transfer values depend on the mesh and the Dirichlet sets only
            for o in ops:
                o.refresh()
        lp0 = self.parts[0]
        self.owned = ctx.up(lp0.owned, np.int64)
        self.owned32 = _gather_map(ctx, lp0.owned)
        self.x = ctx.zeros(lp0.n_loc * nb)
        self.rhs = ctx.zeros(lp0.n_own * nb)
        self.last_info = None
        self._ag = None

    def refresh_operators(self):
        """Pull this rank's rows out of the freshly assembled global operators and set up the smoothers."""
        for o in self.A:
            o.refresh()
        self.tail.setup()
        capi.check(self.ctx.lib.sfem_mg_setup(self.mg, self.ctx.stream), 'sfem_mg_setup')

    def refresh(self):
        self.refresh_operators()
        f = self.prob.fine
        _gather(self.ctx, self.parts[0].n_own, self.owned32, f.rhs, self.rhs)

    def assemble_coarse_local(self, D, vel, mu, upwind, robin):
        """Multigrid levels: the row-partitioned ones from their owned cells, the replicated tail in full; then the
        smoother set-up of both handles."""
        prob, nd = self.prob, self.nd
        for l, lev in enumerate(prob.levels[1:], 1):
            cu, cv = vel[l - 1]
            if l < nd:
                self.asm[l].assemble(D, cu, cv, mu, None, False, upwind=upwind, robin=robin)
                self.asm[l].apply_bc(with_values=False)
            else:
                lev.assemble(D, cu, cv, mu, None, False, upwind=upwind, robin=robin)
                lev.apply_bc(1)
        self.tail.setup()
        capi.check(self.ctx.lib.sfem_mg_setup(self.mg, self.ctx.stream), 'sfem_mg_setup')

    def assemble_local(self, D, ux=None, uy=None, mu_const=0.0, mu_nodal=None, clamp=False, bc_values=None, robin=True,
                       coarse_mu: Optional[float] = None):
        """Distributed counterpart of ``ScalarProblem.assemble`` (symmetric elimination): every rank assembles ITS rows
        of the system level and of the row-partitioned multigrid levels; ``ux`` / ``uy`` / ``mu_nodal`` are full
        (replicated) vectors.  Bit-identical to extracting the rows of the globally assembled operators."""
        prob, ctx = self.prob, self.ctx
        f = prob.fine
        f.set_bc_values(bc_values if bc_values is not None else {i: 0.0 for i in f.bc_dofs})
        a0 = self.asm[0]
        a0.assemble(D, ux, uy, mu_const, mu_nodal, clamp, robin=robin)
        capi.check(ctx.lib.sfem_vec_set(self.parts[0].n_own, 0.0, P(self.rhs), ctx.stream), 'sfem_vec_set')
        a0.apply_bc(rhs=self.rhs, with_values=True)
        vel = prob._coarse_velocity(ux, uy) if ux is not None else [(None, None)] * len(prob.transfers)
        self.assemble_coarse_local(D, vel, mu_const if coarse_mu is None else coarse_mu, ux is not None, robin)

    def solve(self, method='cg', rtol=1e-13, maxit=400, restart=80):
        ctx, f = self.ctx, self.prob.fine
        lib = ctx.lib
        n_own = self.parts[0].n_own
        if not hasattr(self, '_x0g'):
            self._x0g = ctx.zeros(f.n)
        capi.check(lib.sfem_vec_select(f.n, P(f.bc_flag), P(f.bc_val), None, P(self._x0g), ctx.stream), 'sfem_vec_select')
        capi.check(lib.sfem_vec_set(self.parts[0].n_loc, 0.0, P(self.x), ctx.stream), 'sfem_vec_set')
        _gather(ctx, n_own, self.owned32, self._x0g, self.x)
        info = (C.c_double * 4)()
        A = self.A[0].csr
        if method == 'cg':
            rc = lib.sfem_krylov_cg(n_own, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), self.mg, P(self.rhs),
                                    P(self.x), float(rtol), int(maxit), info, ctx.stream)
        elif method == 'fgmres':
            rc = lib.sfem_krylov_fgmres(n_own, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), self.mg, P(self.rhs),
                                        P(self.x), float(rtol), int(restart), int(maxit), info, ctx.stream)
        else:
            raise ValueError(f"unknown method {method!r}")
        capi.check(rc, f'sfem_krylov_{method} (distributed)')
        if self.world.error():
            raise capi.SulcusFemError("multi-GPU exchange timed out (a peer did not answer)")
        self.last_info = {'iterations': int(info[0]), 'relres': float(info[1]), 'converged': bool(info[2]),
                          'estimate': float(info[3]), 'method': method, 'ranks': self.nranks}
        return self.x[:n_own]

    def gather(self):
        """The full (replicated) solution vector on every rank."""
        if self._ag is None:
            n = self.prob.fine.n
            counts = [int((self.owners[0] == q).sum()) for q in range(self.nranks)]
            ag = AllGather(self.ctx, self.nranks, counts, group=self.group)
            pos = np.empty(n, dtype=np.int64)
            for q in range(self.nranks):
                g = np.flatnonzero(self.owners[0] == q)
                pos[g] = q * ag.seg + np.arange(len(g))
            ag.add_field('x', n, pos)
            self._ag = ag
        n_own = self.parts[0].n_own
        capi.check(self.ctx.lib.sfem_vec_copy(n_own, P(self.x), P(self._ag.send), self.ctx.stream), 'sfem_vec_copy')
        return self._ag.run()['x']

    def close(self):
        if getattr(self, 'mg', None):
            self.ctx.lib.sfem_mg_destroy(self.mg)
            self.mg = None
        for h in self.halos:
            if h.handle:
                self.ctx.lib.sfem_halo_destroy(h.handle)
                h.handle = None
        if self._own_world:
            self.world.close()


class DistStokesProblem:
    """Row-partitioned Taylor-Hood solve (reference ``solvers.py:237-306`` over several GPUs).

    Every rank holds the full :class:`StokesProblem` and assembles the global block views K, B, B^T, Mp (replicated
    assembly); this class extracts the rank's rows, numbers the local unknowns so that the owned part of a Stokes
    vector is contiguous (``csrc/sfem_stokes.cu``: [u owned, interleaved | p owned | pad | u ghosts | p ghosts]) and runs
    MINRES with the same preconditioner as on one GPU: row-partitioned 2-right-hand-side velocity multigrid,
    pressure-mass Chebyshev sweep, lubrication coarse correction (its 11 coefficients all-reduced in-kernel)."""

    def __init__(self, stokes, world: DistWorld, replicate_below: int = 200000):
        from .device import StokesProblem  # noqa: F401  (type only)
        self.stokes, self.ctx, self.world = stokes, stokes.ctx, world
        ctx, rank, nranks = self.ctx, world.rank, world.nranks
        self.rank, self.nranks = rank, nranks
        mesh = stokes.mesh
        n2, nv = stokes.n2, stokes.nv
        pb, pbt, bt_code, mp = dm.stokes_block_plans(mesh)
        self._pat = (pb, pbt, mp)
        self._bt_code = bt_code

        def extra0(owner0):
            # the divergence block reads velocity pairs j = col // 2 from the rows of the pressure owner
            return [(pb.rowptr, pb.cols.astype(np.int64) // 2, owner0[:nv])]
        plan = plan_partitions(stokes.vel, nranks, replicate_below, extra_ops0=extra0)
        owner0 = plan[1][0]
        self.owner_u, self.owner_p = owner0, owner0[:nv].copy()
        nv_own = int((self.owner_p == rank).sum())
        self.hole = pt.stokes_gaps(nv_own, 0)[0]               # pairs between owned and ghost velocity dofs
        self.vel = DistScalarProblem(stokes.vel, rank, nranks, group=world.group, world=world, nb=2, plan=plan,
                                     level0_gap=self.hole)
        lu = self.vel.parts[0]
        n2_own, n2_gh = lu.n_own, len(lu.ghost)
        # pressure partition: ghosts read by B^T (rows = velocity owner) and by the mass matrix
        row_owner_bt = np.repeat(owner0, 2)                    # interleaved velocity rows 2 j + c
        ghosts_p = pt.ghost_sets(self.owner_p, nranks, [(pbt.rowptr, pbt.cols, row_owner_bt), (mp.rowptr, mp.cols, self.owner_p)])
        self.gap_p = pt.stokes_gaps(nv_own, n2_gh)[1]          # pad + velocity ghosts sit between owned and ghost pressure
        self.lp = world.plan(self.owner_p, ghosts_p, nb_max=1, gap=self.gap_p)
        assert self.lp.n_own == nv_own
        self.n2_own, self.nv_own = n2_own, nv_own
        self.n_own = 2 * n2_own + nv_own
        self.n_alloc = 2 * lu.n_loc + len(self.lp.ghost)       # 2 (n2_own + hole + n2_gh) + nv_gh
        self.nv_alloc = self.lp.n_loc
        self.handle = None

    def finalize(self):
        ctx, lib, st = self.ctx, self.ctx.lib, self.stokes
        self.vel.finalize()
        pb, pbt, mp = self._pat
        lu, lp = self.vel.parts[0], self.lp
        n2, nv = st.n2, st.nv
        self.halo_p = DeviceHalo(ctx, lp)
        # B: owned pressure rows; columns = interleaved velocity scalars in K's local numbering (2 g2l[j] + c)
        col_b = np.full(2 * n2, -1, dtype=np.int64)
        present = lu.g2l >= 0
        col_b[0::2] = np.where(present, 2 * lu.g2l, -1)
        col_b[1::2] = np.where(present, 2 * lu.g2l + 1, -1)
        self.B = _LocalOp(ctx, st.B, pb.rowptr, pb.cols, lp.owned, col_b, 2 * lu.n_loc)
        # B^T: owned velocity rows (interleaved), columns = local pressure ids (relative to the start of the p part)
        rows_bt = np.stack([2 * lu.owned, 2 * lu.owned + 1], axis=1).ravel()
        self.BT = _LocalOp(ctx, st.BT, pbt.rowptr, pbt.cols, rows_bt, lp.g2l, lp.n_loc)
        self.halo_p.attach(self.BT.csr)
        self.Mp = _LocalOp(ctx, st.Mp, mp.rowptr, mp.cols, lp.owned, lp.g2l, lp.n_loc)
        self.halo_p.attach(self.Mp.csr)
        # lubrication correction restricted to the owned pressure dofs
        nz, zt = 0, [None] * 6
        q = st.schur
        if q is not None:
            if q.nz > 1024:
                raise capi.SulcusFemError("Schur correction too large")
            own_only = np.where((lp.g2l >= 0) & (lp.g2l < lp.n_own), lp.g2l, -1)
            rp, lc, slot = pt.localize_csr(q.zt_rowptr, q.zt_cols, np.arange(q.nz), own_only, drop_missing=True)
            self._z = (ctx.up(rp, np.int32), ctx.up(np.concatenate([lc, [0]]), np.int32),
                       ctx.up(np.concatenate([np.asarray(q.zt_vals)[slot], [0.0]]), np.float64),
                       ctx.up(np.asarray(q.zidx)[lp.owned], np.int32), ctx.up(np.asarray(q.zw)[lp.owned], np.float64),
                       ctx.up(q.C.ravel(), np.float64))
            nz, zt = q.nz, [P(t) for t in self._z]
        K = self.vel.A[0].csr
        B, BT, Mp = self.B.csr, self.BT.csr, self.Mp.csr
        self.handle = lib.sfem_stokes_create_part(self.n2_own, self.nv_own, K.nnz, P(K.rowptr), P(K.cols), P(K.vals),
                                                  B.nnz, P(B.rowptr), P(B.cols), P(B.vals),
                                                  P(BT.rowptr), P(BT.cols), P(BT.vals),
                                                  Mp.nnz, P(Mp.rowptr), P(Mp.cols), P(Mp.vals), self.vel.mg, nz, *zt,
                                                  BT.nnz, self.n_alloc, self.nv_alloc)
        if not self.handle:
            raise capi.SulcusFemError("sfem_stokes_create_part failed: " + lib.sfem_last_error().decode())
        # ---- local assembly plans (owned cells only): B / B^T share one divergence element buffer, Mp has its own
        from .device import cell_geometry
        geo_all = cell_geometry(st.mesh)
        ops_b = [_flatten_contribs(pb.contrib_ptr, pb.contrib_code, self.B.slot_host),
                 _flatten_contribs(pbt.contrib_ptr, self._bt_code, self.BT.slot_host)]
        cells_b, maps_b = plan_local_gather(ops_b, 36)
        self._div = dict(nc=len(cells_b), geo=ctx.up(np.ascontiguousarray(geo_all[:, cells_b]), np.float64),
                         E=ctx.zeros(max(36 * len(cells_b), 1)),
                         b=(ctx.up(maps_b[0][0], np.int32), ctx.up(np.concatenate([maps_b[0][1], [0]]), np.int32)),
                         bt=(ctx.up(maps_b[1][0], np.int32), ctx.up(np.concatenate([maps_b[1][1], [0]]), np.int32)))
        cells_m, maps_m = plan_local_gather([_flatten_contribs(mp.contrib_ptr, mp.contrib_code, self.Mp.slot_host)], 9)
        self._mass = dict(nc=len(cells_m), geo=ctx.up(np.ascontiguousarray(geo_all[:, cells_m]), np.float64),
                          E=ctx.zeros(max(9 * len(cells_m), 1)),
                          m=(ctx.up(maps_m[0][0], np.int32), ctx.up(np.concatenate([maps_m[0][1], [0]]), np.int32)))
        # Dirichlet data in the local layout: values (pairs with their ghosts; the pressure part is zero) and flags
        a0 = self.vel.asm[0]
        gl = np.zeros(self.n_alloc + 2)
        have = np.flatnonzero(a0.l2g >= 0)
        gil = np.asarray(st.g_il_host)
        gl[2 * have] = gil[2 * a0.l2g[have]]
        gl[2 * have + 1] = gil[2 * a0.l2g[have] + 1]
        self.g_loc = ctx.up(gl, np.float64)
        fl = np.where(a0.l2g >= 0, st.vel.fine.bc_flag_host[np.maximum(a0.l2g, 0)], 0).astype(np.uint8)
        self.flag_il_loc = ctx.up(np.repeat(fl, 2), np.uint8)
        # gather maps: solver layout of the global problem ([u interleaved | p]) -> owned part of the local vectors
        src = np.concatenate([rows_bt, 2 * n2 + lp.owned])
        self.own_map = _gather_map(ctx, src)
        self.b = ctx.zeros(self.n_alloc + 2)
        self.x = ctx.zeros(self.n_alloc + 2)
        self.last_info = None
        # all-gather of the solution into blocked global fields ux | uy | p
        counts = [2 * int((self.owner_u == r).sum()) + int((self.owner_p == r).sum()) for r in range(self.nranks)]
        ag = AllGather(ctx, self.nranks, counts, group=self.world.group)
        pos_ux, pos_uy, pos_p = np.empty(n2, np.int64), np.empty(n2, np.int64), np.empty(nv, np.int64)
        for r in range(self.nranks):
            gu = np.flatnonzero(self.owner_u == r)
            gp = np.flatnonzero(self.owner_p == r)
            pos_ux[gu] = r * ag.seg + 2 * np.arange(len(gu))
            pos_uy[gu] = r * ag.seg + 2 * np.arange(len(gu)) + 1
            pos_p[gp] = r * ag.seg + 2 * len(gu) + np.arange(len(gp))
        ag.add_field('ux', n2, pos_ux)
        ag.add_field('uy', n2, pos_uy)
        ag.add_field('p', nv, pos_p)
        self._ag = ag

    def refresh(self):
        """After ``StokesProblem.assemble(bc_mode=1)``: this rank's rows of every operator + the right-hand side."""
        self.vel.refresh_operators()
        for o in (self.B, self.BT, self.Mp):
            o.refresh()
        _gather(self.ctx, self.n_own, self.own_map, self.stokes.rhs_il, self.b)

    def assemble_local(self):
        """Distributed counterpart of ``StokesProblem.assemble(bc_mode=1)`` (direct block assembly, device.py): every
        rank assembles ITS rows of K, B, B^T, Mp and of the row-partitioned velocity multigrid levels from the cells
        that touch them, lifts the Dirichlet data into its part of the right-hand side and eliminates -- bit-identical
        to extracting the rows of the globally assembled, eliminated blocks."""
        ctx, lib, vel = self.ctx, self.ctx.lib, self.vel
        n2o, nvo = self.n2_own, self.nv_own
        a0 = vel.asm[0]
        K, B, BT, Mp = vel.A[0].csr, self.B.csr, self.BT.csr, self.Mp.csr
        a0.assemble(1.0, None, None, robin=False)                               # un-eliminated scalar stiffness rows
        d = self._div
        if d['nc']:
            capi.check(lib.sfem_elem_th_div(d['nc'], P(d['geo']), P(d['E']), ctx.stream), 'sfem_elem_th_div (local)')
        for M, (cp, cc) in ((B, d['b']), (BT, d['bt'])):
            capi.check(lib.sfem_gather_csr(M.nnz, P(cp), P(cc), P(d['E']), P(M.vals), ctx.stream), 'sfem_gather_csr (local)')
        # lifting  b = -[K g_u ; B g_u]  on the owned rows (g_u with its ghosts), then b = g on the Dirichlet rows
        r = self.b
        capi.check(lib.sfem_vec_set(self.n_alloc, 0.0, P(r), ctx.stream), 'sfem_vec_set')
        K.spmv(self.g_loc, y=r, b=r, mode=1, nb=2)
        rp = r[2 * n2o:]
        B.spmv(self.g_loc, y=rp, b=rp, mode=1)
        capi.check(lib.sfem_vec_select(2 * n2o, P(self.flag_il_loc), P(self.g_loc), P(r), P(r), ctx.stream), 'sfem_vec_select')
        a0.apply_bc(with_values=False)                                           # K: rows and columns of Dirichlet dofs
        capi.check(lib.sfem_csr_zero_flagged(nvo, P(B.rowptr), P(B.cols), P(B.vals), None, P(self.flag_il_loc), ctx.stream),
                   'sfem_csr_zero_flagged')
        capi.check(lib.sfem_csr_zero_flagged(2 * n2o, P(BT.rowptr), P(BT.cols), P(BT.vals), P(self.flag_il_loc), None,
                                             ctx.stream), 'sfem_csr_zero_flagged')
        # preconditioner data: velocity multigrid levels and the pressure mass matrix
        vel.assemble_coarse_local(1.0, [(None, None)] * len(vel.prob.transfers), 0.0, False, False)
        m = self._mass
        if m['nc']:
            capi.check(lib.sfem_elem_p1_mass(m['nc'], P(m['geo']), P(m['E']), ctx.stream), 'sfem_elem_p1_mass (local)')
        capi.check(lib.sfem_gather_csr(Mp.nnz, P(m['m'][0]), P(m['m'][1]), P(m['E']), P(Mp.vals), ctx.stream),
                   'sfem_gather_csr (local)')

    def solve(self, rtol=1e-12, maxit=2000):
        ctx, lib = self.ctx, self.ctx.lib
        capi.check(lib.sfem_vec_set(self.n_alloc, 0.0, P(self.x), ctx.stream), 'sfem_vec_set')
        info = (C.c_double * 4)()
        x0 = getattr(self.stokes, 'x0_il', None)
        if x0 is not None:       # same starting vector and anchored stopping level as the single-GPU solve
            if getattr(self, 'xref', None) is None:
                self.xref = ctx.zeros(self.n_alloc)
            capi.check(lib.sfem_vec_set(self.n_alloc, 0.0, P(self.xref), ctx.stream), 'sfem_vec_set')
            _gather(ctx, self.n_own, self.own_map, self.stokes.g_il, self.xref)
            _gather(ctx, self.n_own, self.own_map, x0, self.x)
            capi.check(lib.sfem_stokes_solve_from(self.handle, P(self.b), P(self.x), P(self.xref), float(rtol), int(maxit),
                                                  info, ctx.stream), 'sfem_stokes_solve_from (distributed)')
        else:
            _gather(ctx, self.n_own, self.own_map, self.stokes.g_il, self.x)   # initial guess: the Dirichlet values
            capi.check(lib.sfem_stokes_solve(self.handle, P(self.b), P(self.x), float(rtol), int(maxit), info, ctx.stream),
                       'sfem_stokes_solve (distributed)')
        if self.world.error():
            raise capi.SulcusFemError("multi-GPU exchange timed out (a peer did not answer)")
        self.last_info = {'iterations': int(info[0]), 'relres': float(info[1]), 'converged': bool(info[2]),
                          'estimate': float(info[3]), 'method': 'minres', 'ranks': self.nranks}
        return self.x[:self.n_own]

    def gather(self):
        """(ux, uy, p): the full (replicated, blocked) fields on every rank."""
        capi.check(self.ctx.lib.sfem_vec_copy(self.n_own, P(self.x), P(self._ag.send), self.ctx.stream), 'sfem_vec_copy')
        out = self._ag.run()
        return out['ux'], out['uy'], out['p']

    def close(self):
        if self.handle:
            self.ctx.lib.sfem_stokes_destroy(self.handle)
            self.handle = None
        if getattr(self, 'halo_p', None) is not None and self.halo_p.handle:
            self.ctx.lib.sfem_halo_destroy(self.halo_p.handle)
            self.halo_p.handle = None
        self.vel.close()
