"""Device-side problem objects: upload the host maps once, then assemble / solve / reduce on the GPU.

PyTorch is used only as the device-memory allocator and stream provider (tensors are passed to the
C ABI as raw pointers); every arithmetic step of the solve path is a kernel of ``libsulcusfem.so``.
There is no CPU fallback: constructing any of these classes without a CUDA device raises.

Replaces the dolfin objects the reference builds in ``simulation.py:128-130,146`` (function spaces:
DOF maps + sparsity) and the work of ``solve(a == L, ...)`` in ``solvers.py:55,84,151,213,298``.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence

import numpy as np

from . import capi
from . import dofmap as dm
from . import hierarchy as hy
from .hostmesh import HostMesh


def _torch():
    import torch
    return torch


class Context:
    """Library handle + CUDA device + stream."""
    _inst = None

    def __init__(self, device=None):
        torch = _torch()
        self.lib = capi.load()
        if not torch.cuda.is_available():
            raise capi.SulcusFemError("no CUDA device: sulcusfem has no CPU fallback")
        self.device = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        torch.cuda.set_device(self.device)
        if dm.DEVICE_SORT is None:
            dm.DEVICE_SORT = self.device         # large CSR patterns / gather maps are sorted on the device from now on

    @classmethod
    def get(cls):
        if cls._inst is None:
            cls._inst = Context()
        return cls._inst

    @property
    def stream(self):
        # raw handle of torch's current stream of this thread (the C call: torch.cuda.current_stream() costs ~15 us of
        # Python per use, ~30 uses per sweep case)
        torch = _torch()
        try:
            return C.c_void_p(torch._C._cuda_getCurrentRawStream(self.device.index))
        except AttributeError:
            return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def up(self, arr, dtype):
        """Host array -> device tensor.  Arrays that live in pinned memory (the fields :meth:`down` returned, i.e. the
        ``Function.values`` a previous solve produced) are copied asynchronously at full PCIe rate on the current
        stream; anything else goes through the blocking pageable path."""
        torch = _torch()
        a = np.ascontiguousarray(arr, dtype=dtype)
        t = torch.from_numpy(a)
        if a is arr and t.numel() >= (1 << 15) and t.is_pinned():
            return t.to(self.device, non_blocking=True)          # stream-ordered; `arr` is kept alive by its owner
        return t.to(self.device, non_blocking=False)

    def down(self, t):
        """Device tensor -> numpy array backed by PINNED host memory (torch's caching host allocator), so the
        download runs at full PCIe rate and a later :meth:`up` of the same array does too."""
        torch = _torch()
        h = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
        h.copy_(t, non_blocking=False)
        return h.numpy()

    def zeros(self, n, dtype=None):
        torch = _torch()
        return torch.zeros(int(n), dtype=dtype or torch.float64, device=self.device)

    def empty(self, n, dtype=None):
        torch = _torch()
        return torch.empty(int(n), dtype=dtype or torch.float64, device=self.device)


P = capi.ptr


def _arr(m):
    """numpy view of a marker container (MeshMarkers, dolfin MeshFunction or plain array)."""
    return np.asarray(m.array() if hasattr(m, "array") else m)


class DeviceCsr:
    """CSR matrix in HBM (int32 indices, FP64 values), padded for the TMA-staged SpMV engine, with its
    tile plan registered in the library (keyed by the device address of rowptr)."""
    PAD = 16
    STAGED_CAP = 768             # entries per tile (one shared-memory stage)
    STAGED_ROWS = 64             # rows per tile (64: 128 consumer threads per CTA, 128: 256)
    SELL_SIGMA = 256             # sorting window of the sliced-ELL mirror (rows)
    SELL_MIN_ROWS = 2048         # smaller matrices never get a mirror (launch-latency bound anyway)
    SELL_MAX_FILL = 1.5          # mirrors that would store > 1.5x the non-zeros are not built (irregular rows)

    def __init__(self, ctx: Context, nrows, ncols, rowptr, cols, vals=None, staged=True, sell=True):
        torch = _torch()
        self.ctx = ctx
        self.nrows, self.ncols = int(nrows), int(ncols)
        self.nnz = int(len(cols))
        rp = np.zeros(self.nrows + 1 + self.PAD, dtype=np.int32)
        rp[:self.nrows + 1] = rowptr
        rp[self.nrows + 1:] = rowptr[-1]
        cc = np.zeros(self.nnz + self.PAD, dtype=np.int32)
        cc[:self.nnz] = cols
        self.rowptr = ctx.up(rp, np.int32)
        self.cols = ctx.up(cc, np.int32)
        self.vals_buf = torch.zeros(self.nnz + self.PAD, dtype=torch.float64, device=ctx.device)
        self.vals = self.vals_buf[:self.nnz]
        if vals is not None:
            self.vals.copy_(ctx.up(vals, np.float64))
        self.tile_row = None
        self.ntiles = 0
        if staged and self.nrows > 0:
            tr = np.zeros(self.nrows + 1, dtype=np.int32)
            nt = ctx.lib.sfem_staged_plan(self.nrows, rp.ctypes.data_as(C.c_void_p), self.STAGED_CAP, self.STAGED_ROWS,
                                          tr.ctypes.data_as(C.c_void_p))
            if nt > 0:
                self.ntiles = int(nt)
                self.tile_row = ctx.up(tr[:nt + 1], np.int32)
                capi.check(ctx.lib.sfem_staged_register(P(self.rowptr), self.nrows, P(self.tile_row), self.ntiles,
                                                        self.STAGED_CAP, self.STAGED_ROWS), 'sfem_staged_register')
        self.sell = None
        if sell and self.nrows >= self.SELL_MIN_ROWS and self.nnz > 0:
            from . import sell as sl
            plan = sl.build_plan_device(self.rowptr[:self.nrows + 1], self.cols[:self.nnz], self.SELL_SIGMA)
            nparts = int(ctx.lib.sfem_sell_parts())
            if plan['padded'] <= self.SELL_MAX_FILL * self.nnz:
                parts = sl.partition_slices(plan['slice_ptr'].cpu().numpy(), plan['padded'], plan['nslices'], nparts)
                self.sell = dict(
                    fill=plan['padded'] / self.nnz, nslices=plan['nslices'], padded=plan['padded'],
                    slice_ptr=plan['slice_ptr'], perm=plan['perm'], scols=plan['scols'], src=plan['src'],
                    parts=ctx.up(parts, np.int32), nparts=nparts,
                    svals=torch.zeros(max(plan['padded'], 1), dtype=torch.float64, device=ctx.device))
                k = self.sell
                capi.check(ctx.lib.sfem_sell_register(P(self.rowptr), P(self.vals_buf), self.nrows, plan['nslices'],
                                                      P(k['slice_ptr']), P(k['perm']), P(k['scols']), P(k['src']),
                                                      P(k['svals']), plan['padded'], P(k['parts']), nparts), 'sfem_sell_register')

    def mark_dirty(self):
        """Tell the library that the CSR values were written outside of it (torch ops on ``vals``)."""
        if self.sell is not None:
            self.ctx.lib.sfem_sell_mark_dirty(P(self.rowptr))

    def __del__(self):
        try:
            if getattr(self, 'tile_row', None) is not None:
                self.ctx.lib.sfem_staged_unregister(P(self.rowptr))
                self.tile_row = None
            if getattr(self, 'sell', None) is not None:
                self.ctx.lib.sfem_sell_unregister(P(self.rowptr))
                self.sell = None
        except Exception:
            pass

    def spmv(self, x, y=None, b=None, mode=0, staged=False, nb=1, sell=False):
        """y = A x (0), b - A x (1), y += A x (2).  ``staged=True`` / ``sell=True`` insist on the TMA-staged /
        sliced-ELL engine (error if the matrix is too small / has no plan); otherwise the library picks."""
        ctx = self.ctx
        if y is None:
            y = ctx.empty(self.nrows * nb)
        fn = ctx.lib.sfem_spmv_csr_f64_sell if sell else (
            ctx.lib.sfem_spmv_csr_f64_staged if staged else ctx.lib.sfem_spmv_csr_f64_nb)
        capi.check(fn(self.nrows, self.ncols, self.nnz, P(self.rowptr), P(self.cols), P(self.vals_buf),
                      P(x), P(b), P(y), mode, nb, ctx.stream), 'sfem_spmv_csr_f64')
        return y

    def to_scipy(self):
        import scipy.sparse as sp
        return sp.csr_matrix((self.vals.cpu().numpy(), self.cols[:self.nnz].cpu().numpy(),
                              self.rowptr[:self.nrows + 1].cpu().numpy()), shape=(self.nrows, self.ncols))


def cell_geometry(mesh: HostMesh) -> np.ndarray:
    """[6][nc] SoA vertex coordinates x0,y0,x1,y1,x2,y2 (memoised on the mesh: the velocity level, the scalar level, the
    Stokes problem and the functional plan of one mesh all upload it)."""
    g = getattr(mesh, '_cell_geometry', None)
    if g is None:
        p = mesh.coords[mesh.cells]
        g = np.ascontiguousarray(p.reshape(mesh.num_cells, 6).T)
        mesh._cell_geometry = g                 # (kept writable: torch.from_numpy warns on read-only arrays; nobody writes it)
    return g


def facet_geometry(mesh: HostMesh, facets: np.ndarray) -> np.ndarray:
    e = mesh.edges[facets]
    return np.ascontiguousarray(np.concatenate([mesh.coords[e[:, 0]], mesh.coords[e[:, 1]]], axis=1).T)


class ScalarLevel:
    """One scalar Lagrange space (P2 system level or P1 multigrid level) on one mesh."""

    def __init__(self, ctx: Context, mesh: HostMesh, bc_markers: np.ndarray, degree: int,
                 dirichlet_ids: Sequence[int], robin_id: Optional[int]):
        self.ctx, self.mesh, self.degree = ctx, mesh, degree
        nc = mesh.num_cells
        self.nc = nc
        self.n, cd, f, fd, self.pattern = dm.scalar_level_plan(mesh, bc_markers, degree, robin_id)
        self.ndof_cell = cd.shape[1]
        nfam = 1
        self.nf = 0
        if robin_id is not None:
            nfam = 2
            self.nf = len(f)
            self.fgeo = ctx.up(facet_geometry(mesh, f), np.float64) if self.nf else None
            self.fdofs = ctx.up(np.ascontiguousarray(fd.T), np.int32) if self.nf else None
            self.robin_facets = f
        pat = self.pattern
        self.A = DeviceCsr(ctx, self.n, self.n, pat.rowptr, pat.cols)
        self.contrib_ptr = ctx.up(pat.contrib_ptr, np.int32)
        self.contrib_code = ctx.up(pat.contrib_code, np.int32)
        self.E = ctx.zeros(max(pat.buffer_len, 1))
        self.facet_base = pat.family_base[1] if nfam > 1 else pat.buffer_len
        self.geo = ctx.up(cell_geometry(mesh), np.float64)
        self.celldofs = ctx.up(np.ascontiguousarray(cd.T), np.int32)
        # Dirichlet data (later ids overwrite earlier ones on shared dofs, like a list of dolfin bcs)
        flag = np.zeros(self.n, dtype=np.uint8)
        self.bc_dofs = {}
        for i in dirichlet_ids:
            d = dm.dirichlet_dofs_p2(mesh, bc_markers, i) if degree == 2 else dm.dirichlet_dofs_p1(mesh, bc_markers, i)
            self.bc_dofs[i] = d
            flag[d] = 1
        self.bc_flag_host = flag
        self.bc_flag = ctx.up(flag, np.uint8)
        self.bc_val = ctx.zeros(self.n)
        self.rhs = ctx.zeros(self.n)

    def set_bc_values(self, values: Dict[int, object]):
        """values[id] = scalar or array over the dofs of that id; applied in dict order."""
        key = tuple((i, float(v)) for i, v in values.items()) if all(np.isscalar(v) for v in values.values()) else None
        if key is not None and getattr(self, '_bc_key', None) == key:
            return None                                   # same constants as last time: already on the device
        self._bc_key = key
        g = np.zeros(self.n)
        for i, v in values.items():
            g[self.bc_dofs[i]] = v
        self.bc_val.copy_(self.ctx.up(g, np.float64))
        return g

    def assemble(self, D: float, ux=None, uy=None, mu_const: float = 0.0, mu_nodal=None, clamp=False,
                 upwind=False, robin=True):
        ctx, lib = self.ctx, self.ctx.lib
        if self.degree == 2:
            capi.check(lib.sfem_elem_p2_advdiff(self.nc, P(self.geo), P(self.celldofs), float(D), P(ux), P(uy),
                                                P(self.E), ctx.stream), 'sfem_elem_p2_advdiff')
        else:
            capi.check(lib.sfem_elem_p1_advdiff(self.nc, P(self.geo), P(self.celldofs), float(D), P(ux), P(uy),
                                                int(bool(upwind)), P(self.E), ctx.stream), 'sfem_elem_p1_advdiff')
        if self.nf:
            F = self.E[self.facet_base:]
            if robin:
                fn = lib.sfem_facet_p2_robin if self.degree == 2 else lib.sfem_facet_p1_robin
                capi.check(fn(self.nf, P(self.fgeo), P(self.fdofs), float(mu_const), P(mu_nodal), int(bool(clamp)),
                              P(F), ctx.stream), 'sfem_facet_robin')
            else:
                capi.check(lib.sfem_vec_set(int(F.numel()), 0.0, P(F), ctx.stream), 'sfem_vec_set')
        capi.check(lib.sfem_gather_csr(self.A.nnz, P(self.contrib_ptr), P(self.contrib_code), P(self.E), P(self.A.vals),
                                       ctx.stream), 'sfem_gather_csr')

    def apply_bc(self, mode: int, rhs=None):
        ctx = self.ctx
        rhs = self.rhs if rhs is None else rhs
        capi.check(ctx.lib.sfem_apply_dirichlet(self.n, self.A.nnz, P(self.A.rowptr), P(self.A.cols), P(self.A.vals), P(rhs),
                                                P(self.bc_flag), P(self.bc_val), mode, ctx.stream), 'sfem_apply_dirichlet')
        return rhs


class DeviceTransfer:
    def __init__(self, ctx: Context, T: hy.Transfer, fine_bc: np.ndarray, coarse_bc: np.ndarray):
        keep = (fine_bc[np.repeat(np.arange(T.n_fine), np.diff(T.rowptr))] == 0) & (coarse_bc[T.cols] == 0)
        vals = np.where(keep, T.vals, 0.0)
        self.P = DeviceCsr(ctx, T.n_fine, T.n_coarse, T.rowptr, T.cols, vals)
        self.R = DeviceCsr(ctx, T.n_coarse, T.n_fine, T.t_rowptr, T.t_cols, vals[T.t_perm])
        self.nested = T.nested
        if not T.nested:
            # unfiltered restriction + reciprocal row sums: weighted average used to carry the
            # velocity to the non-nested coarse mesh (preconditioner data only)
            self.Rfull = DeviceCsr(ctx, T.n_coarse, T.n_fine, T.t_rowptr, T.t_cols, T.t_vals)
            rs = np.add.reduceat(T.t_vals, T.t_rowptr[:-1].astype(np.int64)) if len(T.t_vals) else np.ones(T.n_coarse)
            rs[np.diff(T.t_rowptr) == 0] = 1.0
            self.rinv = ctx.up(1.0 / rs, np.float64)


class Multigrid:
    """sfem_mg handle over [system level, P1 levels...]."""

    def __init__(self, ctx: Context, levels: List[ScalarLevel], transfers: List[DeviceTransfer],
                 cheb_degree=2, eig_ratio=8.0, nb=1):
        self.ctx, self.levels, self.transfers = ctx, levels, transfers
        self.nb = int(nb)
        nl = len(levels)
        self.coarse_inv = ctx.zeros(levels[-1].n ** 2) if (nl > 1 or levels[-1].n <= 2048) else None
        IntArr, PtrArr = C.c_int * nl, C.c_void_p * nl

        def parr(ts):
            vals = [t.data_ptr() if t is not None else None for t in ts] + [None] * (nl - len(ts))
            return PtrArr(*vals)
        self._keep = dict(
            n=IntArr(*[l.n for l in levels]), annz=IntArr(*[l.A.nnz for l in levels]),
            arp=parr([l.A.rowptr for l in levels]), ac=parr([l.A.cols for l in levels]), av=parr([l.A.vals for l in levels]),
            pnnz=IntArr(*([t.P.nnz for t in transfers] + [0])),
            prp=parr([t.P.rowptr for t in transfers]), pc=parr([t.P.cols for t in transfers]), pv=parr([t.P.vals for t in transfers]),
            rrp=parr([t.R.rowptr for t in transfers]), rc=parr([t.R.cols for t in transfers]), rv=parr([t.R.vals for t in transfers]))
        k = self._keep
        self.handle = ctx.lib.sfem_mg_create(nl, k['n'], k['annz'], k['arp'], k['ac'], k['av'], k['pnnz'], k['prp'], k['pc'],
                                             k['pv'], k['rrp'], k['rc'], k['rv'], P(self.coarse_inv), int(cheb_degree),
                                             float(eig_ratio), self.nb)
        if not self.handle:
            raise capi.SulcusFemError("sfem_mg_create failed: " + ctx.lib.sfem_last_error().decode())

    def setup(self):
        ctx = self.ctx
        if self.coarse_inv is not None:
            c = self.levels[-1]
            capi.check(ctx.lib.sfem_dense_inverse_csr(c.n, P(c.A.rowptr), P(c.A.cols), P(c.A.vals), P(self.coarse_inv),
                                                      ctx.stream), 'sfem_dense_inverse_csr')
        capi.check(ctx.lib.sfem_mg_setup(self.handle, ctx.stream), 'sfem_mg_setup')

    def setup_fine(self):
        capi.check(self.ctx.lib.sfem_mg_setup_fine(self.handle, self.ctx.stream), 'sfem_mg_setup_fine')

    def vcycle(self, b, x=None):
        x = self.ctx.empty(self.levels[0].n * self.nb) if x is None else x
        capi.check(self.ctx.lib.sfem_mg_vcycle(self.handle, P(b), P(x), self.ctx.stream), 'sfem_mg_vcycle')
        return x

    def lambda_max(self):
        out = (C.c_double * len(self.levels))()
        capi.check(self.ctx.lib.sfem_mg_lambda_max(self.handle, out), 'sfem_mg_lambda_max')
        return list(out)

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.ctx.lib.sfem_mg_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class ScalarProblem:
    """P2 scalar advection-diffusion-Robin problem with its multigrid hierarchy (one mesh)."""

    def __init__(self, mesh: HostMesh, bc_markers: np.ndarray, dirichlet_ids=(1, 2), robin_id: Optional[int] = 4,
                 hierarchy: Optional[hy.Hierarchy] = None, ctx: Optional[Context] = None,
                 cheb_degree=None, eig_ratio=None, nb=1):
        import os
        cheb_degree = int(os.environ.get('SFEM_CHEB_DEGREE', 2)) if cheb_degree is None else cheb_degree
        eig_ratio = float(os.environ.get('SFEM_EIG_RATIO', 8.0)) if eig_ratio is None else eig_ratio
        self.ctx = ctx or Context.get()
        self.mesh = mesh
        self.hierarchy = hierarchy or hy.build_hierarchy(mesh)
        H = self.hierarchy
        self.fine = ScalarLevel(self.ctx, mesh, bc_markers, 2, dirichlet_ids, robin_id)
        self.levels = [self.fine]
        for m in H.meshes:
            mk = bc_markers if m is mesh else hy.level_markers(m)['bc_markers'].values
            self.levels.append(ScalarLevel(self.ctx, m, mk, 1, dirichlet_ids, robin_id))
        self.transfers = [DeviceTransfer(self.ctx, T, self.levels[l].bc_flag_host, self.levels[l + 1].bc_flag_host)
                          for l, T in enumerate(H.transfers)]
        self.mg_cheb_degree, self.mg_eig_ratio = cheb_degree, eig_ratio
        self.mg = Multigrid(self.ctx, self.levels, self.transfers, cheb_degree, eig_ratio, nb)
        self.n = self.fine.n
        self.x = self.ctx.zeros(self.n)
        self.last_info = None

    def _coarse_velocity(self, ux, uy):
        """P1 nodal velocity on every multigrid level (prefix for nested, weighted average otherwise)."""
        out = []
        cu, cv = ux, uy
        for l, T in enumerate(self.transfers):
            lev = self.levels[l + 1]
            if T.nested:
                cu, cv = cu[:lev.n], cv[:lev.n]
            else:
                lib, ctx = self.ctx.lib, self.ctx
                nu, nv_ = ctx.empty(lev.n), ctx.empty(lev.n)
                for src, dst in ((cu, nu), (cv, nv_)):
                    T.Rfull.spmv(src, dst)
                    capi.check(lib.sfem_vec_pointwise_mul(lev.n, 1.0, P(T.rinv), P(dst), P(dst), ctx.stream), 'pointwise_mul')
                cu, cv = nu, nv_
            out.append((cu.contiguous(), cv.contiguous()))
        return out

    def assemble(self, D, ux=None, uy=None, mu_const=0.0, mu_nodal=None, clamp=False, bc_values=None,
                 bc_mode=1, robin=True, coarse_mu: Optional[float] = None, fine=True, reuse_coarse=False):
        """Assemble the system level (+ BCs) and rediscretise the multigrid levels.  ``fine=False``:
        the system-level values were written by the caller (Stokes: K extracted from the Taylor-Hood
        matrix); only the coarse levels are assembled.  ``reuse_coarse=True`` (mu sweeps, ``solvers.frozen_coarse_levels``):
        keep the coarse levels of the previous full assembly (same D, no velocity) and redo the system level's smoother
        set-up only; returns False -- and assembles in full -- when there is nothing compatible to reuse."""
        f = self.fine
        key = (float(D), bool(robin), int(bc_mode))
        if reuse_coarse and fine and ux is None and getattr(self, '_coarse_key', None) == key:
            f.set_bc_values(bc_values if bc_values is not None else {i: 0.0 for i in f.bc_dofs})
            f.assemble(D, None, None, mu_const, mu_nodal, clamp, robin=robin)
            capi.check(self.ctx.lib.sfem_vec_set(f.n, 0.0, P(f.rhs), self.ctx.stream), 'sfem_vec_set')
            f.apply_bc(bc_mode)
            self.bc_mode = bc_mode
            self.mg.setup_fine()
            return True
        self._coarse_key = key if (fine and ux is None) else None
        self._coarse_mu = float(mu_const if coarse_mu is None else coarse_mu)
        if fine:
            f.set_bc_values(bc_values if bc_values is not None else {i: 0.0 for i in f.bc_dofs})
            f.assemble(D, ux, uy, mu_const, mu_nodal, clamp, robin=robin)
            capi.check(self.ctx.lib.sfem_vec_set(f.n, 0.0, P(f.rhs), self.ctx.stream), 'sfem_vec_set')
            f.apply_bc(bc_mode)
        self.bc_mode = bc_mode
        vel = self._coarse_velocity(ux, uy) if ux is not None else [(None, None)] * len(self.transfers)
        # coarse levels are preconditioner data only: a spatially varying mu is represented there by
        # the constant `coarse_mu` (mean of mu over the Robin boundary)
        for l, lev in enumerate(self.levels[1:]):
            cu, cv = vel[l]
            lev.assemble(D, cu, cv, mu_const if coarse_mu is None else coarse_mu, None, False,
                         upwind=ux is not None, robin=robin)
            lev.apply_bc(1)
        self.mg.setup()

    def solve(self, method='cg', rtol=1e-13, maxit=400, restart=80, x0=None):
        ctx, f = self.ctx, self.fine
        if x0 is None:          # initial guess: the Dirichlet values, zero elsewhere
            capi.check(ctx.lib.sfem_vec_select(f.n, P(f.bc_flag), P(f.bc_val), None, P(self.x), ctx.stream), 'sfem_vec_select')
        else:
            capi.check(ctx.lib.sfem_vec_copy(f.n, P(x0), P(self.x), ctx.stream), 'sfem_vec_copy')
        info = (C.c_double * 4)()
        A = f.A
        if method == 'cg':
            rc = ctx.lib.sfem_krylov_cg(f.n, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), self.mg.handle, P(f.rhs), P(self.x),
                                        float(rtol), int(maxit), info, ctx.stream)
            capi.check(rc, 'sfem_krylov_cg')
        elif method == 'fgmres':
            rc = ctx.lib.sfem_krylov_fgmres(f.n, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), self.mg.handle, P(f.rhs), P(self.x),
                                            float(rtol), int(restart), int(maxit), info, ctx.stream)
            capi.check(rc, 'sfem_krylov_fgmres')
        else:
            raise ValueError(f"unknown method {method!r}")
        self.last_info = {'iterations': int(info[0]), 'relres': float(info[1]), 'converged': bool(info[2]),
                          'estimate': float(info[3]), 'method': method}
        return self.x


    # ------------------------------------------------------------------ batched Robin sweep (sfem_krylov_cg_batch)
    BATCH_MAX = 16

    def batch_operators(self, D: float, bc_values: Dict[int, float]):
        """The two value arrays of A_l(mu) = A0_l + mu M_l on the pattern of EVERY level -- A0_l = D K_l with identity
        Dirichlet rows, M_l = the Robin boundary mass matrix with ZERO Dirichlet rows / columns -- and, for the system
        level, the lifted right-hand sides b0, bM (b(mu) = b0 + mu bM).  Built with the same element / facet / gather /
        Dirichlet kernels as a single assembly, once per (D, Dirichlet constants); a sweep then needs no per-mu
        assembly of the operators at all."""
        key = (float(D), tuple((int(i), float(v)) for i, v in bc_values.items()))
        ops = getattr(self, '_batch_ops', None)
        if ops is not None and ops['key'] == key:
            return ops
        ctx, lib, f = self.ctx, self.ctx.lib, self.fine
        if not f.nf:
            raise capi.SulcusFemError("batched Robin sweep: the problem has no Robin boundary")
        f.set_bc_values(bc_values)
        vals0, valsM = [], []
        b0 = bM = None
        for l, lev in enumerate(self.levels):
            # A0_l = D K_l: element kernel, facet family zeroed, symmetric elimination (level 0: lifting into b0)
            lev.assemble(float(D), None, None, 0.0, None, False, robin=False)
            rhs0 = ctx.zeros(lev.n)
            lev.apply_bc(1, rhs=rhs0)
            vals0.append(lev.A.vals_buf.clone())
            # M_l: element kernel with D = 0 (zeros), facet kernel with mu = 1; level 0: bM = -M g; then the rows and
            # columns of the Dirichlet dofs are zeroed (their rows of A0 + mu M stay identity rows, b stays g there)
            lev.assemble(0.0, None, None, 1.0, None, False, robin=True)
            if l == 0:
                zeros = ctx.zeros(lev.n)
                bM = ctx.zeros(lev.n)
                lev.A.spmv(lev.bc_val, y=bM, b=zeros, mode=1)
                capi.check(lib.sfem_vec_select(lev.n, P(lev.bc_flag), P(zeros), P(bM), P(bM), ctx.stream), 'sfem_vec_select')
                b0 = rhs0
            capi.check(lib.sfem_csr_zero_flagged(lev.n, P(lev.A.rowptr), P(lev.A.cols), P(lev.A.vals), P(lev.bc_flag),
                                                 P(lev.bc_flag), ctx.stream), 'sfem_csr_zero_flagged')
            valsM.append(lev.A.vals_buf.clone())
            lev.A.mark_dirty()
        self._coarse_key = None                  # no level holds an assembled A(mu) any more
        nl = len(self.levels)
        PtrArr = C.c_void_p * nl
        self._batch_ops = dict(key=key, vals0=vals0, valsM=valsM, b0=b0, bM=bM,
                               p0=PtrArr(*[t.data_ptr() for t in vals0]), pM=PtrArr(*[t.data_ptr() for t in valsM]))
        return self._batch_ops

    def solve_batch(self, D: float, mus: Sequence[float], bc_values: Dict[int, float], rtol=1e-13, maxit=400,
                    mu_ref: Optional[float] = None, shared_coarse=False):
        """Solve the pure-diffusion Robin problem for every mu of ``mus`` (at most ``BATCH_MAX``) in one batched PCG.
        The multigrid hierarchy is assembled for ``mu_ref`` (default: geometric mean of the positive mus): it supplies
        the structure and the dense inverse of the coarsest level; every column smooths with its own operator on
        every level (``shared_coarse=True``: only on the system level, the coarse levels of ``mu_ref`` serve all
        columns -- the first version, kept for comparison).  Returns (X, infos): X the interleaved solutions [n][nb] on
        the device (a buffer owned by the problem, overwritten by the next call of the same width), infos one
        ``last_info`` dictionary per mu."""
        ctx, f = self.ctx, self.fine
        mus = [float(m) for m in mus]
        nb = len(mus)
        if not 1 <= nb <= self.BATCH_MAX:
            raise ValueError(f"solve_batch takes 1..{self.BATCH_MAX} coefficients")
        if min(mus) < 0.0:
            raise ValueError("solve_batch: Robin coefficients must be >= 0")
        ops = self.batch_operators(D, bc_values)
        if mu_ref is None:
            pos = [m for m in mus if m > 0.0]
            mu_ref = float(np.exp(np.mean(np.log(pos)))) if pos else 0.0
        self.assemble(float(D), mu_const=float(mu_ref), bc_values=bc_values)      # hierarchy of A(mu_ref)
        capi.check(ctx.lib.sfem_vec_select(f.n, P(f.bc_flag), P(f.bc_val), None, P(self.x), ctx.stream), 'sfem_vec_select')
        # one result buffer per batch width, reused by later calls (a stable address keeps the captured iteration graph
        # valid): the returned X is overwritten by the next solve_batch of the same width -- copy out what must survive
        bufs = self.__dict__.setdefault('_batch_X', {})
        X = bufs.get(nb)
        if X is None:
            X = bufs[nb] = ctx.empty(f.n * nb)
        info = (C.c_double * (4 * nb))()
        h_mu = (C.c_double * nb)(*mus)
        A = f.A
        nl = len(self.levels)
        p0, pM = ops['p0'], ops['pM']
        if shared_coarse:
            PtrArr = C.c_void_p * nl
            p0 = PtrArr(*([ops['vals0'][0].data_ptr()] + [None] * (nl - 1)))
            pM = PtrArr(*([ops['valsM'][0].data_ptr()] + [None] * (nl - 1)))
        rc = ctx.lib.sfem_krylov_cg_batch(f.n, A.nnz, P(A.rowptr), P(A.cols), nl, p0, pM, nb, h_mu, float(mu_ref),
                                          self.mg.handle, P(ops['b0']), P(ops['bM']), P(self.x), P(X), float(rtol),
                                          int(maxit), info, ctx.stream)
        capi.check(rc, 'sfem_krylov_cg_batch')
        infos = [{'iterations': int(info[4 * c]), 'relres': float(info[4 * c + 1]), 'converged': bool(info[4 * c + 2]),
                  'estimate': float(info[4 * c + 3]), 'method': 'cg_batch', 'batch': nb, 'mu_ref': float(mu_ref)}
                 for c in range(nb)]
        self.last_info = infos[-1]
        return X, infos

    def batch_column(self, X, nb: int, c: int):
        """Column c of an interleaved batch as its own device vector."""
        out = self.ctx.empty(self.n)
        capi.check(self.ctx.lib.sfem_batch_column(self.n, int(nb), int(c), P(X), P(out), self.ctx.stream), 'sfem_batch_column')
        return out


class StokesProblem:
    """Taylor-Hood Stokes system (solvers.py:237-306): assembled with dolfin's clique pattern, solved in
    block form (K shared by both velocity components, B, B^T) by ``sfem_stokes_solve``."""

    def __init__(self, mesh: HostMesh, bc_markers: np.ndarray, hierarchy: Optional[hy.Hierarchy] = None,
                 ctx: Optional[Context] = None, velocity_ids=(1, 4, 3), schur_correction=True):
        self.ctx = ctx or Context.get()
        ctx = self.ctx
        self.mesh = mesh
        self.n2 = n2 = dm.p2_num_dofs(mesh)
        self.nv = nv = mesh.num_vertices
        self.n = 2 * n2 + nv
        self.nc = mesh.num_cells
        self._full = None                      # full Taylor-Hood CSR + gather map: built on first use (parity / export)
        self.geo = ctx.up(cell_geometry(mesh), np.float64)
        # velocity block: scalar stiffness hierarchy with the velocity Dirichlet set, 2 interleaved RHS
        self.vel = ScalarProblem(mesh, bc_markers, dirichlet_ids=tuple(velocity_ids), robin_id=None,
                                 hierarchy=hierarchy, ctx=ctx, nb=2)
        # ---- divergence blocks assembled directly (host plan, once): B (nv x 2 n2) and B^T (2 n2 x nv) with the
        # velocity columns / rows in interleaved numbering, both gathered from ONE element buffer EB [nc][3][12]
        # (sfem_elem_th_div: rows = the cell's pressure dofs, columns = [u_x x6 | u_y x6])
        pb, pbt, bt_code, mp = dm.stokes_block_plans(mesh)
        self.B = DeviceCsr(ctx, nv, 2 * n2, pb.rowptr, pb.cols)
        self.BT = DeviceCsr(ctx, 2 * n2, nv, pbt.rowptr, pbt.cols)
        self._b_map = (ctx.up(pb.contrib_ptr, np.int32), ctx.up(pb.contrib_code, np.int32))
        self._bt_map = (ctx.up(pbt.contrib_ptr, np.int32), ctx.up(bt_code, np.int32))
        self.EB = ctx.zeros(36 * self.nc)
        self.g_il = ctx.zeros(self.n)           # Dirichlet values in solver layout [u interleaved | p = 0]
        self.flag_il = None                     # uint8 [2 n2]: interleaved velocity Dirichlet flags
        # pressure mass matrix (P1)
        self.Mp = DeviceCsr(ctx, nv, nv, mp.rowptr, mp.cols)
        self._mp = (ctx.up(mp.contrib_ptr, np.int32), ctx.up(mp.contrib_code, np.int32), ctx.zeros(mp.buffer_len))
        # coarse pressure correction of the Schur-complement preconditioner (lubrication operator)
        self.schur = None
        nz = 0
        zt = [None] * 6
        if schur_correction:
            from . import schur as sc
            self.schur = sc.lubrication_correction(mesh, np.asarray(bc_markers))
            if self.schur is not None:
                q = self.schur
                nz = q.nz
                self._z = (ctx.up(q.zt_rowptr, np.int32), ctx.up(q.zt_cols, np.int32), ctx.up(q.zt_vals, np.float64),
                           ctx.up(q.zidx, np.int32), ctx.up(q.zw, np.float64), ctx.up(q.C.ravel(), np.float64))
                zt = [P(t) for t in self._z]
        # Dirichlet data on W.sub(0): both components on every velocity id
        self.bc_markers = bc_markers
        self.velocity_ids = tuple(velocity_ids)
        self.bc_flag_host = np.zeros(self.n, dtype=np.uint8)
        self.bc_flag = None
        self.bc_val = ctx.zeros(self.n)
        self.rhs = ctx.zeros(self.n)            # blocked layout [ux | uy | p] (as assembled)
        self.rhs_il = ctx.zeros(self.n)         # solver layout [u interleaved | p]
        self.x_il = ctx.zeros(self.n)
        self.x = ctx.zeros(self.n)              # blocked layout
        self.last_info = None
        K = self.vel.fine.A
        self.handle = ctx.lib.sfem_stokes_create(n2, nv, K.nnz, P(K.rowptr), P(K.cols), P(K.vals),
                                                 self.B.nnz, P(self.B.rowptr), P(self.B.cols), P(self.B.vals),
                                                 P(self.BT.rowptr), P(self.BT.cols), P(self.BT.vals),
                                                 self.Mp.nnz, P(self.Mp.rowptr), P(self.Mp.cols), P(self.Mp.vals),
                                                 self.vel.mg.handle, nz, *zt)
        if not self.handle:
            raise capi.SulcusFemError("sfem_stokes_create failed: " + ctx.lib.sfem_last_error().decode())

    def set_bcs(self, values_by_id):
        """values_by_id: ordered {id: (gx over dofs, gy over dofs)} -- later ids overwrite earlier."""
        g = np.zeros(self.n)
        flag = np.zeros(self.n, dtype=np.uint8)
        for i, (gx, gy) in values_by_id.items():
            d = dm.dirichlet_dofs_p2(self.mesh, self.bc_markers, i)
            g[d] = gx
            g[d + self.n2] = gy
            flag[d] = 1
            flag[d + self.n2] = 1
        if not np.array_equal(flag[:self.n2], self.vel.fine.bc_flag_host) or \
                not np.array_equal(flag[self.n2:2 * self.n2], self.vel.fine.bc_flag_host):
            raise capi.SulcusFemError("velocity Dirichlet set differs from the preconditioner's")
        self.bc_flag_host = flag
        self.bc_flag = self.ctx.up(flag, np.uint8)
        self.bc_val.copy_(self.ctx.up(g, np.float64))
        # solver layout (interleaved velocity): flags, values, scratch for the block elimination
        n2 = self.n2
        fil = np.empty(2 * n2, dtype=np.uint8)
        fil[0::2], fil[1::2] = flag[:n2], flag[n2:2 * n2]
        gil = np.zeros(self.n)
        gil[0:2 * n2:2], gil[1:2 * n2:2] = g[:n2], g[n2:2 * n2]
        self.flag_il = self.ctx.up(fil, np.uint8)
        self.x0_il = self.x0_il_host = None      # a starting vector built for other boundary data is stale
        self.g_il_host = gil                     # Dirichlet values in solver layout (read by the row-partitioned assembly)
        self.g_il.copy_(self.ctx.up(gil, np.float64))
        self._zero_n2 = self.ctx.zeros(n2)
        self._scratch_n2 = self.ctx.zeros(n2)
        return g

    def set_channel_flow_guess(self, L: float, H: float):
        """Starting vector of the Krylov solve for the reference's Stokes set-up (``solvers.py:252-264``: inlet profile
        ``(4 y (H - y), 0)`` on id 1, no-slip on the walls, natural outflow): the analytic channel flow of that profile --
        ``u = (4 y (H - y), 0)`` for ``y > 0``, zero below the channel floor (inside a sulcus), ``p = 8 (L - x)`` -- with
        the Dirichlet dofs set to their values (call after :meth:`set_bcs`).  Away from the sulcus mouth this IS the
        solution, so MINRES starts from a ~10x smaller residual; :meth:`solve` keeps the stopping level anchored to the
        plain guess (``sfem_stokes_solve_from``), i.e. the same absolute residual, fewer iterations."""
        n2, nv = self.n2, self.nv
        X2 = dm.p2_dof_coordinates(self.mesh)
        y = X2[:, 1]
        x0 = np.zeros(self.n)
        x0[0:2 * n2:2] = np.where(y > 0.0, 4.0 * y * (H - y), 0.0)
        x0[2 * n2:] = 8.0 * (L - self.mesh.coords[:, 0])
        fl = self.bc_flag_host
        fil = np.zeros(self.n, dtype=bool)
        fil[0:2 * n2:2], fil[1:2 * n2:2] = fl[:n2] != 0, fl[n2:2 * n2] != 0
        x0[fil] = self.g_il_host[fil]
        self.x0_il_host = x0
        self.x0_il = self.ctx.up(x0, np.float64)

    def _to_solver_layout(self, blocked, out):
        n2, lib, ctx = self.n2, self.ctx.lib, self.ctx
        capi.check(lib.sfem_vec_interleave2(n2, P(blocked[:n2]), P(blocked[n2:2 * n2]), P(out), ctx.stream), 'interleave')
        out[2 * n2:].copy_(blocked[2 * n2:])

    # ------------------------------------------------------------------ full Taylor-Hood matrix (lazy)
    def _ensure_full(self):
        """dolfin's clique pattern of the mixed space (225 entries per cell, structural zeros included), its gather
        map and the slot maps to the block views.  The solver never needs it; it is built on first use for parity
        checks / export of the assembled matrix."""
        if self._full is not None:
            return self._full
        ctx, n2, nv = self.ctx, self.n2, self.nv
        cd = dm.th_cell_dofs(self.mesh)
        pat = dm.build_pattern(self.n, self.n, [(cd, cd)])
        A = DeviceCsr(ctx, self.n, self.n, pat.rowptr, pat.cols, sell=False)
        rows = np.repeat(np.arange(self.n, dtype=np.int64), np.diff(pat.rowptr.astype(np.int64)))
        cols = pat.cols.astype(np.int64)
        kslot = np.flatnonzero((rows < n2) & (cols < n2))
        Kp = self.vel.fine.pattern
        if len(kslot) != Kp.nnz or not np.array_equal(cols[kslot], Kp.cols):
            raise capi.SulcusFemError("Taylor-Hood velocity block does not match the P2 scalar pattern")

        def sub_slots(sel, new_rows, new_cols, want):
            slot = np.flatnonzero(sel)
            r, c = new_rows[slot], new_cols[slot]
            order = np.lexsort((c, r))
            if len(slot) != want.nnz or not np.array_equal(c[order], want.cols[:want.nnz].cpu().numpy()):
                raise capi.SulcusFemError("Taylor-Hood divergence block does not match the directly built pattern")
            return ctx.up(slot[order], np.int32)
        il_rows = 2 * (rows % n2) + rows // n2          # interleaved velocity numbering (valid for rows < 2 n2)
        il_cols = 2 * (cols % n2) + cols // n2
        self._full = dict(
            pattern=pat, A=A, contrib_ptr=ctx.up(pat.contrib_ptr, np.int32), contrib_code=ctx.up(pat.contrib_code, np.int32),
            E=ctx.zeros(pat.buffer_len), k_slot=ctx.up(kslot, np.int32),
            bt_slot=sub_slots((rows < 2 * n2) & (cols >= 2 * n2), il_rows, cols - 2 * n2, self.BT),
            b_slot=sub_slots((rows >= 2 * n2) & (cols < 2 * n2), rows - 2 * n2, il_cols, self.B))
        return self._full

    @property
    def pattern(self):
        return self._ensure_full()['pattern']

    @property
    def A(self):
        return self._ensure_full()['A']

    def _assemble_preconditioner(self):
        # coarse velocity levels + pressure mass matrix
        ctx, lib = self.ctx, self.ctx.lib
        self.vel.assemble(1.0, robin=False, fine=False)
        cp, cc, E = self._mp
        capi.check(lib.sfem_elem_p1_mass(self.nc, P(self.geo), P(E), ctx.stream), 'sfem_elem_p1_mass')
        capi.check(lib.sfem_gather_csr(self.Mp.nnz, P(cp), P(cc), P(E), P(self.Mp.vals), ctx.stream), 'sfem_gather_csr')

    def assemble(self, bc_mode=1, full=None):
        """Assemble the Stokes system (+ Dirichlet).

        ``bc_mode=1`` (symmetric elimination; what :meth:`solve` needs), ``full=False`` (default): the block views
        K, B, B^T are assembled DIRECTLY -- 72 element entries per cell instead of the 225 of the mixed-space
        matrix -- and eliminated in place; the right-hand side is lifted with two SpMVs.  The numbers are the ones
        the full path produces (same element arithmetic, same gather order): the tests compare them bit for bit.
        ``full=True`` (forced for ``bc_mode=0``, dolfin's identity-row mode): assemble the Taylor-Hood matrix with
        dolfin's clique pattern into :attr:`A`, apply the conditions there and extract the blocks."""
        ctx, lib = self.ctx, self.ctx.lib
        torch = _torch()
        if full is None:
            full = bc_mode != 1
        self.bc_mode = bc_mode
        if full or bc_mode != 1:
            F = self._ensure_full()
            A = F['A']
            capi.check(lib.sfem_elem_th_stokes(self.nc, P(self.geo), P(F['E']), ctx.stream), 'sfem_elem_th_stokes')
            capi.check(lib.sfem_gather_csr(A.nnz, P(F['contrib_ptr']), P(F['contrib_code']), P(F['E']), P(A.vals),
                                           ctx.stream), 'sfem_gather_csr')
            self.rhs.zero_()
            capi.check(lib.sfem_apply_dirichlet(self.n, A.nnz, P(A.rowptr), P(A.cols), P(A.vals), P(self.rhs),
                                                P(self.bc_flag), P(self.bc_val), bc_mode, ctx.stream), 'sfem_apply_dirichlet')
            if bc_mode != 1:
                return
            K = self.vel.fine.A
            capi.check(lib.sfem_csr_extract(K.nnz, P(F['k_slot']), P(A.vals), P(K.vals), ctx.stream), 'sfem_csr_extract')
            capi.check(lib.sfem_csr_extract(self.B.nnz, P(F['b_slot']), P(A.vals), P(self.B.vals), ctx.stream), 'sfem_csr_extract')
            capi.check(lib.sfem_csr_extract(self.BT.nnz, P(F['bt_slot']), P(A.vals), P(self.BT.vals), ctx.stream), 'sfem_csr_extract')
            self._to_solver_layout(self.rhs, self.rhs_il)
            self._assemble_preconditioner()
            return
        # ---- direct block assembly
        n2, nv = self.n2, self.nv
        f = self.vel.fine
        K = f.A
        f.assemble(1.0, robin=False)                                  # un-eliminated scalar stiffness (P2 kernel, D = 1)
        capi.check(lib.sfem_elem_th_div(self.nc, P(self.geo), P(self.EB), ctx.stream), 'sfem_elem_th_div')
        for M, (cp, cc) in ((self.B, self._b_map), (self.BT, self._bt_map)):
            capi.check(lib.sfem_gather_csr(M.nnz, P(cp), P(cc), P(self.EB), P(M.vals), ctx.stream), 'sfem_gather_csr')
        # lifting  b = -[K g_u ; B g_u]  (g_u = Dirichlet values, zero elsewhere), then b = g on the Dirichlet rows
        r = self.rhs_il
        capi.check(lib.sfem_vec_set(self.n, 0.0, P(r), ctx.stream), 'sfem_vec_set')
        K.spmv(self.g_il[:2 * n2], y=r[:2 * n2], b=r[:2 * n2], mode=1, nb=2)      # r_u = 0 - K g_u (both components)
        self.B.spmv(self.g_il[:2 * n2], y=r[2 * n2:], b=r[2 * n2:], mode=1)        # r_p = 0 - B g_u
        capi.check(lib.sfem_vec_select(2 * n2, P(self.flag_il), P(self.g_il), P(r), P(r), ctx.stream), 'sfem_vec_select')
        # symmetric elimination of the blocks
        capi.check(lib.sfem_apply_dirichlet(n2, K.nnz, P(K.rowptr), P(K.cols), P(K.vals), P(self._scratch_n2), P(f.bc_flag),
                                            P(self._zero_n2), 1, ctx.stream), 'sfem_apply_dirichlet')
        capi.check(lib.sfem_csr_zero_flagged(nv, P(self.B.rowptr), P(self.B.cols), P(self.B.vals), None, P(self.flag_il),
                                             ctx.stream), 'sfem_csr_zero_flagged')
        capi.check(lib.sfem_csr_zero_flagged(2 * n2, P(self.BT.rowptr), P(self.BT.cols), P(self.BT.vals), P(self.flag_il), None,
                                             ctx.stream), 'sfem_csr_zero_flagged')
        # blocked copy of the right-hand side (public layout [ux | uy | p])
        capi.check(lib.sfem_vec_deinterleave2(n2, P(r), P(self.rhs[:n2]), P(self.rhs[n2:2 * n2]), ctx.stream), 'deinterleave')
        capi.check(lib.sfem_vec_copy(nv, P(r[2 * n2:]), P(self.rhs[2 * n2:]), ctx.stream), 'sfem_vec_copy')
        self._assemble_preconditioner()

    def solve(self, rtol=1e-12, maxit=2000):
        ctx, lib = self.ctx, self.ctx.lib
        torch = _torch()
        if getattr(self, 'bc_mode', None) != 1:
            raise capi.SulcusFemError("StokesProblem.solve needs assemble(bc_mode=1)")
        n2 = self.n2
        info = (C.c_double * 4)()
        x0 = getattr(self, 'x0_il', None)
        if x0 is not None:       # analytic channel flow as the starting vector, stopping level anchored to the plain guess
            capi.check(lib.sfem_vec_copy(self.n, P(x0), P(self.x_il), ctx.stream), 'sfem_vec_copy')
            rc = lib.sfem_stokes_solve_from(self.handle, P(self.rhs_il), P(self.x_il), P(self.g_il), float(rtol), int(maxit),
                                            info, ctx.stream)
        else:
            capi.check(lib.sfem_vec_copy(self.n, P(self.g_il), P(self.x_il), ctx.stream), 'sfem_vec_copy')   # Dirichlet values
            rc = lib.sfem_stokes_solve(self.handle, P(self.rhs_il), P(self.x_il), float(rtol), int(maxit), info, ctx.stream)
        capi.check(rc, 'sfem_stokes_solve')
        self.last_info = {'iterations': int(info[0]), 'relres': float(info[1]), 'converged': bool(info[2]),
                          'estimate': float(info[3]), 'method': 'minres'}
        capi.check(lib.sfem_vec_deinterleave2(n2, P(self.x_il), P(self.x[:n2]), P(self.x[n2:2 * n2]), ctx.stream), 'deinterleave')
        capi.check(lib.sfem_vec_copy(self.nv, P(self.x_il[2 * n2:]), P(self.x[2 * n2:]), ctx.stream), 'sfem_vec_copy')
        return self.x[:n2], self.x[n2:2 * n2], self.x[2 * n2:]

    def __del__(self):
        try:
            if getattr(self, 'handle', None):
                self.ctx.lib.sfem_stokes_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


class FunctionalPlan:
    """Facet / cell groups of the reference's integration measures (mesh.py:721-737), built once."""
    GROUPS = ['left', 'right', 'top', 'bottom', 'bottom_left', 'sulcus', 'bottom_right', 'y0_ext', 'mouth']

    def __init__(self, mesh: HostMesh, markers: dict, domain_type: str, ctx: Optional[Context] = None):
        self.ctx = ctx or Context.get()
        ctx = self.ctx
        self.mesh, self.domain_type = mesh, domain_type
        bc = _arr(markers['bc_markers'])
        ent_cell, ent_loc, ptr = [], [], [0]

        def add(facets, cells, locs):
            ent_cell.append(np.asarray(cells, dtype=np.int64))
            ent_loc.append(np.asarray(locs, dtype=np.int64))
            ptr.append(ptr[-1] + len(facets))
        for mid in (1, 2, 3, 4):
            add(*dm.boundary_facets(mesh, bc, mid))
        if domain_type == 'sulcus':
            bs = _arr(markers['bottom_segment_markers'])
            y0 = _arr(markers['y0_markers'])
            dmk = _arr(markers['domain_markers'])
            for mid in (5, 6, 7):
                add(*dm.boundary_facets(mesh, bs, mid))
            add(*dm.boundary_facets(mesh, y0, 10))
            # interior y0 facets: channel-side (cell marker 2) trace, analysis.py:217-237
            f = np.flatnonzero((y0 == 10) & ~mesh.edge_on_boundary)
            c0, c1 = mesh.edge_cells[f, 0], mesh.edge_cells[f, 1]
            chan0, chan1 = dmk[c0] == 2, dmk[c1] == 2
            if np.any(chan0 & chan1):
                raise ValueError("interior y0 facet with channel cells on both sides")
            sel = chan0 | chan1
            f = f[sel]
            cells = np.where(chan0[sel], c0[sel], c1[sel])
            locs = np.where(chan0[sel], mesh.edge_local[f, 0], mesh.edge_local[f, 1])
            add(f, cells, locs)
        self.ngroups = len(ptr) - 1
        self.grp_ptr = ctx.up(np.asarray(ptr), np.int32)
        ec = np.concatenate(ent_cell) if ent_cell else np.zeros(0)
        el = np.concatenate(ent_loc) if ent_loc else np.zeros(0)
        self.ent_cell = ctx.up(np.concatenate([ec, [0]]), np.int32)
        self.ent_local = ctx.up(np.concatenate([el, [0]]), np.int32)
        self.geo = ctx.up(cell_geometry(mesh), np.float64)
        self.celldofs = ctx.up(np.ascontiguousarray(dm.p2_cell_dofs(mesh).T), np.int32)
        if domain_type == 'sulcus':
            self.cell_marker = ctx.up(_arr(markers['domain_markers']), np.int32)
            self.nmarkers = 3
        else:
            self.cell_marker = None
            self.nmarkers = 1
        self.out_f = ctx.zeros(self.ngroups * 8)
        self.out_c = ctx.zeros(self.nmarkers * 2)

    def evaluate(self, c, ux=None, uy=None, D=1.0, mu_const=0.0, mu_nodal=None):
        """Returns (facet[ngroups,8], cell[nmarkers,2]) as numpy arrays (one D2H copy each)."""
        ctx, lib, m = self.ctx, self.ctx.lib, self.mesh
        capi.check(lib.sfem_facet_functionals(self.ngroups, P(self.grp_ptr), P(self.ent_cell), P(self.ent_local), P(self.geo),
                                              P(self.celldofs), m.num_cells, P(c), P(ux), P(uy), float(D), float(mu_const),
                                              P(mu_nodal), P(self.out_f), ctx.stream), 'sfem_facet_functionals')
        if self.cell_marker is None:
            # marker pointer NULL -> every cell counts for marker 0
            capi.check(lib.sfem_cell_functionals(m.num_cells, P(self.geo), P(self.celldofs), None, 1, P(c), P(self.out_c),
                                                 ctx.stream), 'sfem_cell_functionals')
        else:
            capi.check(lib.sfem_cell_functionals(m.num_cells, P(self.geo), P(self.celldofs), P(self.cell_marker), self.nmarkers,
                                                 P(c), P(self.out_c), ctx.stream), 'sfem_cell_functionals')
        return (self.out_f.cpu().numpy().reshape(self.ngroups, 8).copy(),
                self.out_c.cpu().numpy().reshape(self.nmarkers, 2).copy())
