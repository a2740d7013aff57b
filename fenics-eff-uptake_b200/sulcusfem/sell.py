"""Host-side plan of the sliced-ELL (SELL-32-sigma) mirror of a CSR matrix (layout: include/sulcusfem.h).

Built once per sparsity pattern; the device kernels (csrc/sfem_spmv_sell.cu) only ever see the five arrays
returned here plus the mirror's value array, which the library refreshes from the CSR values.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

SLICE = 32


@dataclass
class SellPlan:
    nrows: int
    nslices: int
    padded: int                 # total stored positions (multiple of 32), including padding
    slice_ptr: np.ndarray       # int32 [nslices + 1]
    perm: np.ndarray            # int32 [nslices * 32], CSR row of every lane, -1 = none
    scols: np.ndarray           # int32 [padded], -1 on padding
    src: np.ndarray             # int32 [padded], CSR slot of every position, -1 on padding

    @property
    def fill(self) -> float:
        """stored positions / non-zeros (1.0 = no padding)."""
        nnz = int((self.src >= 0).sum())
        return self.padded / nnz if nnz else 1.0


def build_plan(rowptr: np.ndarray, cols: np.ndarray, sigma: int = 256) -> SellPlan:
    """Rows are sorted by decreasing length inside windows of ``sigma`` rows (stable: equal-length neighbours
    stay neighbours, which keeps the epilogue accesses and the gathers of x local), cut into slices of 32 and
    every slice is padded to its longest row.  ``sigma`` is rounded up to a multiple of 32; sigma <= 32 keeps
    the row order."""
    rowptr = np.asarray(rowptr, dtype=np.int64)
    cols = np.asarray(cols)
    n = len(rowptr) - 1
    nslices = (n + SLICE - 1) // SLICE
    npad = nslices * SLICE
    lens = np.zeros(npad, dtype=np.int64)
    lens[:n] = np.diff(rowptr)
    sigma = max(SLICE, ((int(sigma) + SLICE - 1) // SLICE) * SLICE)
    pos = np.arange(npad, dtype=np.int64)
    if sigma > SLICE and npad:
        maxlen = int(lens.max()) if npad else 0
        key = (pos // sigma) * (maxlen + 1) + (maxlen - lens)
        order = np.argsort(key, kind='stable')
    else:
        order = pos
    perm = np.where(order < n, order, -1).astype(np.int32)
    slens = lens[order].reshape(nslices, SLICE).max(axis=1) if nslices else np.zeros(0, dtype=np.int64)
    slice_ptr = np.zeros(nslices + 1, dtype=np.int64)
    np.cumsum(slens * SLICE, out=slice_ptr[1:])
    padded = int(slice_ptr[-1])
    if padded >= 2 ** 31:
        raise ValueError("sliced-ELL mirror exceeds int32 indexing")
    scols = np.full(padded, -1, dtype=np.int32)
    src = np.full(padded, -1, dtype=np.int32)
    nnz = int(rowptr[n]) if n else 0
    if nnz:
        where = np.empty(npad, dtype=np.int64)           # position (slice * 32 + lane) of every row
        where[order] = pos
        row_of = np.repeat(np.arange(n, dtype=np.int64), lens[:n])
        k = np.arange(nnz, dtype=np.int64) - rowptr[row_of]
        w = where[row_of]
        dest = slice_ptr[w // SLICE] + k * SLICE + (w % SLICE)
        scols[dest] = cols[:nnz]
        src[dest] = np.arange(nnz, dtype=np.int32)
    return SellPlan(n, nslices, padded, slice_ptr.astype(np.int32), perm, scols, src)


def build_plan_device(rowptr, cols, sigma: int = 256):
    """:func:`build_plan` on the device (torch tensors in, torch tensors out; same integer arithmetic and the same
    stable sort, hence bit-identical arrays -- ``tests/test_gpu_core.py``).  The plan of a refined mesh's operators is
    ~2 s of numpy scatter work per matrix on the host and 170 MB of uploads; built where the matrix already lives it
    costs milliseconds.  Returns a dict with ``nrows, nslices, padded`` (ints) and ``slice_ptr, perm, scols, src``
    (int32 device tensors)."""
    import torch
    dev = rowptr.device
    rp = rowptr.to(torch.int64)
    n = int(rp.numel()) - 1
    nslices = (n + SLICE - 1) // SLICE
    npad = nslices * SLICE
    lens = torch.zeros(npad, dtype=torch.int64, device=dev)
    lens[:n] = rp[1:] - rp[:-1]
    sigma = max(SLICE, ((int(sigma) + SLICE - 1) // SLICE) * SLICE)
    pos = torch.arange(npad, dtype=torch.int64, device=dev)
    if sigma > SLICE and npad:
        maxlen = int(lens.max().item())
        key = torch.div(pos, sigma, rounding_mode='floor') * (maxlen + 1) + (maxlen - lens)
        order = torch.sort(key, stable=True).indices
    else:
        order = pos
    perm = torch.where(order < n, order, torch.full_like(order, -1)).to(torch.int32)
    slens = lens[order].reshape(nslices, SLICE).max(dim=1).values if nslices else torch.zeros(0, dtype=torch.int64, device=dev)
    slice_ptr = torch.zeros(nslices + 1, dtype=torch.int64, device=dev)
    slice_ptr[1:] = torch.cumsum(slens * SLICE, 0)
    padded = int(slice_ptr[-1].item()) if nslices else 0
    if padded >= 2 ** 31:
        raise ValueError("sliced-ELL mirror exceeds int32 indexing")
    scols = torch.full((max(padded, 1),), -1, dtype=torch.int32, device=dev)
    src = torch.full((max(padded, 1),), -1, dtype=torch.int32, device=dev)
    nnz = int(rp[n].item()) if n else 0
    if nnz:
        where = torch.empty(npad, dtype=torch.int64, device=dev)
        where[order] = pos
        row_of = torch.repeat_interleave(torch.arange(n, dtype=torch.int64, device=dev), lens[:n])
        k = torch.arange(nnz, dtype=torch.int64, device=dev) - rp[row_of]
        w = where[row_of]
        dest = slice_ptr[torch.div(w, SLICE, rounding_mode='floor')] + k * SLICE + (w % SLICE)
        scols[dest] = cols[:nnz].to(torch.int32)
        src[dest] = torch.arange(nnz, dtype=torch.int32, device=dev)
    return dict(nrows=n, nslices=nslices, padded=padded, slice_ptr=slice_ptr.to(torch.int32), perm=perm,
                scols=scols[:max(padded, 1)], src=src[:max(padded, 1)])


def partition_slices(slice_ptr: np.ndarray, padded: int, nslices: int, nparts: int) -> np.ndarray:
    """:func:`partition` from the bare arrays (host copy of ``slice_ptr``)."""
    starts = np.asarray(slice_ptr[:-1], dtype=np.int64) // SLICE
    total = int(padded) // SLICE
    targets = (total * np.arange(nparts + 1, dtype=np.int64)) // max(nparts, 1)
    parts = np.searchsorted(starts, targets, side='left').astype(np.int32)
    parts[0] = 0
    parts[-1] = nslices
    return parts


def partition(plan: SellPlan, nparts: int) -> np.ndarray:
    """First slice of each of ``nparts`` contiguous parts holding (nearly) equal numbers of column-steps;
    int32 [nparts + 1], parts[0] = 0, parts[-1] = nslices.  The device kernel gives every warp a run of
    consecutive parts, i.e. one contiguous span of the mirror."""
    starts = plan.slice_ptr[:-1].astype(np.int64) // SLICE
    total = plan.padded // SLICE
    targets = (total * np.arange(nparts + 1, dtype=np.int64)) // max(nparts, 1)
    parts = np.searchsorted(starts, targets, side='left').astype(np.int32)
    parts[0] = 0
    parts[-1] = plan.nslices
    return parts


def to_dense_rows(plan: SellPlan, vals: np.ndarray):
    """(rows, cols, vals) triplets stored by the plan -- used by the tests to check the layout."""
    s = np.repeat(np.arange(plan.nslices), np.diff(plan.slice_ptr.astype(np.int64)))
    off = np.arange(plan.padded) - plan.slice_ptr.astype(np.int64)[s]
    lane = off % SLICE
    rows = plan.perm[s * SLICE + lane]
    ok = plan.src >= 0
    return rows[ok], plan.scols[ok], np.asarray(vals)[plan.src[ok]]
