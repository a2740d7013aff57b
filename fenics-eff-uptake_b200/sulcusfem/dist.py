"""Multi-GPU solve path for large refined meshes (BASELINE config 5): one process per GPU.

The reference is serial.  Here the P2 concentration system and its multigrid hierarchy are
row-partitioned in x-slabs over the ranks; halo exchange and the Krylov all-reduces run as kernels that
store straight into the peers' memory over NVLink (``csrc/sfem_dist.cu``) -- ``torch.distributed`` is
used once, to all-gather the 64-byte CUDA-IPC handles of the mailboxes.

Scope of this version: the symmetric (pure-diffusion / Robin) solve, i.e. multigrid-preconditioned CG.
Assembly is replicated (every rank assembles the global matrices with the single-GPU kernels -- <1 % of
the solve time -- and extracts its rows with ``sfem_csr_extract``); the solve itself is distributed.
Levels with fewer than ``replicate_below`` unknowns are kept on every rank and solved redundantly after a
vector all-reduce of the restricted residual.
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
        self.slot = ctx.up(slot, np.int32)
        self.src = global_csr
        self.ctx = ctx

    def refresh(self):
        capi.check(self.ctx.lib.sfem_csr_extract(self.csr.nnz, P(self.slot), P(self.src.vals), P(self.csr.vals), self.ctx.stream),
                   'sfem_csr_extract')


def level_coordinates(prob: ScalarProblem, l: int) -> np.ndarray:
    lev = prob.levels[l]
    return dm.p2_dof_coordinates(lev.mesh) if lev.degree == 2 else lev.mesh.coords


def plan_partitions(prob: ScalarProblem, nranks: int, replicate_below: int = 20000):
    """Owner arrays + ghost sets of the row-partitioned levels (identical on every rank)."""
    nl = len(prob.levels)
    nd = 0
    while nd < nl - 1 and prob.levels[nd].n >= replicate_below:
        nd += 1
    nd = max(nd, 1)                                         # at least the system level is distributed
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


class DistScalarProblem:
    """Row-partitioned view of a :class:`ScalarProblem` (which every rank holds and assembles in full)."""

    def __init__(self, prob: ScalarProblem, rank: int, nranks: int, group=None, replicate_below: int = 20000,
                 emulate=False):
        self.prob, self.ctx = prob, prob.ctx
        ctx, lib = self.ctx, self.ctx.lib
        self.rank, self.nranks = rank, nranks
        nd, owners, ghosts = plan_partitions(prob, nranks, replicate_below)
        self.nd = nd
        n_tail = prob.levels[nd].n
        header = int(lib.sfem_dist_header_words(nranks, n_tail))
        parts_all, words = mailbox_layout(prob, nranks, nd, owners, ghosts, header)
        self.parts_all = parts_all
        self.parts = [row[rank] for row in parts_all]
        self.dist = DistContext(ctx, rank, nranks, words, n_tail, group=group, emulate=emulate)
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
                own_only = np.where(lp.g2l < lp.n_own, lp.g2l, -1)
                Rl = _LocalOp(ctx, dT.R, T.t_rowptr, T.t_cols, ident, own_only, lp.n_own, drop_missing=True)
            self.Pm.append(Pl)
            self.Rm.append(Rl)
        # multigrid handles: replicated tail (global levels nd..) and the row-partitioned head
        self.tail = Multigrid(ctx, prob.levels[nd:], prob.transfers[nd:], prob.mg_cheb_degree, prob.mg_eig_ratio, 1)
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
                                     k['rrp'], k['rc'], k['rv'], None, prob.mg_cheb_degree, float(prob.mg_eig_ratio), 1)
        if not self.mg:
            raise capi.SulcusFemError("sfem_mg_create (distributed) failed: " + lib.sfem_last_error().decode())
        Pl, Rl = self.Pm[-1].csr, self.Rm[-1].csr
        capi.check(lib.sfem_mg_set_tail(self.mg, self.tail.handle, n_tail, Pl.nnz, P(Pl.rowptr), P(Pl.cols), P(Pl.vals),
                                        Rl.nnz, P(Rl.rowptr), P(Rl.cols), P(Rl.vals)), 'sfem_mg_set_tail')
        lp0 = self.parts[0]
        self.owned = ctx.up(lp0.owned, np.int64)
        self.x = ctx.zeros(lp0.n_loc)
        self.rhs = ctx.zeros(lp0.n_own)
        self.last_info = None

    def refresh(self):
        """Pull this rank's rows out of the freshly assembled global operators and set up the smoothers."""
        for ops in (self.A, self.Pm, self.Rm):
            for o in ops:
                o.refresh()
        self.tail.setup()
        capi.check(self.ctx.lib.sfem_mg_setup(self.mg, self.ctx.stream), 'sfem_mg_setup')
        f = self.prob.fine
        self.rhs.copy_(f.rhs[self.owned])

    def solve(self, rtol=1e-13, maxit=400):
        ctx, f = self.ctx, self.prob.fine
        torch = _torch()
        x0 = f.bc_val * f.bc_flag.to(torch.float64)
        self.x.zero_()
        self.x[:self.parts[0].n_own].copy_(x0[self.owned])
        info = (C.c_double * 4)()
        A = self.A[0].csr
        rc = ctx.lib.sfem_krylov_cg(self.parts[0].n_own, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), self.mg, P(self.rhs),
                                    P(self.x), float(rtol), int(maxit), info, ctx.stream)
        capi.check(rc, 'sfem_krylov_cg (distributed)')
        if self.dist.error():
            raise capi.SulcusFemError("multi-GPU exchange timed out (a peer did not answer)")
        self.last_info = {'iterations': int(info[0]), 'relres': float(info[1]), 'converged': bool(info[2]),
                          'estimate': float(info[3]), 'method': 'cg', 'ranks': self.nranks}
        return self.x[:self.parts[0].n_own]

    def close(self):
        if getattr(self, 'mg', None):
            self.ctx.lib.sfem_mg_destroy(self.mg)
            self.mg = None
        for h in self.halos:
            if h.handle:
                self.ctx.lib.sfem_halo_destroy(h.handle)
                h.handle = None
        self.dist.close()
