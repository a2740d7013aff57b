"""Domain decomposition of one level for the multi-GPU solve path (host, numpy).

The reference is serial; BASELINE config 5 (mesh-convergence meshes of several million DOFs) is served
by row-partitioning every operator over the ranks (one process per GPU):

* ``slab_owner``      recursive-coordinate-bisection-style owner map: DOFs sorted by x (the channel is
                      10:1, so x-slabs give <= 2 neighbours and a one-element-column halo) cut into
                      equal counts.  Any other owner array (e.g. a METIS partition) plugs in unchanged.
* ``LevelPartition``  for one rank and one level: owned DOFs (local ids 0..n_own-1, ascending global id),
                      ghost DOFs (n_own.., grouped by owner rank), the symmetric neighbour list, the
                      send lists in the order of each receiver's ghost list, and the mailbox channel
                      layout of *every* rank (computed redundantly, so set-up needs no communication
                      beyond the IPC handle all-gather).
* ``localize_csr``    rows owned by the rank, columns renumbered to local ids.

Every rank runs the same deterministic code on the same global arrays, hence all ranks agree on all
offsets without talking to each other.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Sequence, Tuple

import numpy as np


def slab_owner(x: np.ndarray, nranks: int, y: np.ndarray = None) -> np.ndarray:
    """Owner rank of every DOF: equal-count slabs along x (ties broken by y, then index)."""
    n = len(x)
    order = np.lexsort((np.arange(n), y if y is not None else np.zeros(n), x))
    owner = np.empty(n, dtype=np.int32)
    bounds = (np.arange(nranks + 1, dtype=np.int64) * n) // nranks
    for r in range(nranks):
        owner[order[bounds[r]:bounds[r + 1]]] = r
    return owner


def ghost_sets(owner: np.ndarray, nranks: int,
               operators: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]]) -> List[np.ndarray]:
    """ghost[r] = sorted global DOFs of this level that rank r reads but does not own.

    ``operators``: (rowptr, cols, row_owner) of every operator whose *columns* live on this level;
    ``row_owner[i]`` is the rank that owns row i (rows may live on another level).
    """
    n = len(owner)
    keys = []
    for rowptr, cols, row_owner in operators:
        rp = np.asarray(rowptr, dtype=np.int64)
        c = np.asarray(cols, dtype=np.int64)
        ro = np.repeat(np.asarray(row_owner, dtype=np.int64), np.diff(rp))
        sel = ro != owner[c]
        keys.append(ro[sel] * n + c[sel])
    allk = np.unique(np.concatenate(keys)) if keys else np.zeros(0, dtype=np.int64)
    kr = allk // n
    out = []
    for r in range(nranks):
        out.append((allk[kr == r] % n).astype(np.int64))
    return out


@dataclass
class LevelPartition:
    rank: int
    nranks: int
    n_global: int
    owned: np.ndarray                 # global ids, ascending
    ghost: np.ndarray                 # global ids grouped by owner rank (ascending rank, ascending id)
    g2l: np.ndarray                   # global -> local id (-1: not present on this rank)
    neighbors: List[int] = field(default_factory=list)
    send_idx: List[np.ndarray] = field(default_factory=list)   # per neighbour: local owned ids to send
    recv_off: List[int] = field(default_factory=list)          # first ghost local id filled by the neighbour
    recv_cnt: List[int] = field(default_factory=list)
    # mailbox channel layout (8-byte words from the mailbox base), per neighbour
    peer_data_off: List[int] = field(default_factory=list)
    peer_flag_off: List[int] = field(default_factory=list)
    my_data_off: List[int] = field(default_factory=list)
    my_flag_off: List[int] = field(default_factory=list)
    cap: List[int] = field(default_factory=list)
    mailbox_end: int = 0              # first free word of this rank's mailbox after this level's channels
    all_mailbox_ends: List[int] = field(default_factory=list)
    ghost_gap: int = 0                # unused local ids between the owned and the ghost DOFs

    @property
    def n_own(self) -> int:
        return int(len(self.owned))

    @property
    def n_loc(self) -> int:
        return int(len(self.owned) + self.ghost_gap + len(self.ghost))


def _even(v: int) -> int:
    return (int(v) + 1) & ~1


def partition_level(owner: np.ndarray, nranks: int, rank: int, ghosts: List[np.ndarray],
                    mailbox_base: Sequence[int], nb_max: int = 2, ghost_gap: int = 0) -> LevelPartition:
    """Partition data of ``rank`` for one level.

    ``ghosts``: output of :func:`ghost_sets`.  ``mailbox_base[q]``: first free word of rank q's mailbox
    before this level; the returned ``mailbox_end`` values of all ranks are obtained by calling this
    function for every q (or :func:`mailbox_ends`).  Channels hold ``nb_max`` values per DOF.
    ``ghost_gap``: unused local ids between the owned and the ghost DOFs of THIS rank (the Taylor-Hood solver keeps
    the owned pressure entries of a Stokes vector in that hole, so that the owned part of the vector is contiguous).
    """
    owner = np.asarray(owner)
    n = len(owner)
    owned = [np.flatnonzero(owner == q).astype(np.int64) for q in range(nranks)]
    # symmetric neighbour relation: q is a neighbour of r if either needs data from the other
    needs = np.zeros((nranks, nranks), dtype=np.int64)       # needs[r, q] = #dofs r reads from q
    for r in range(nranks):
        if len(ghosts[r]):
            needs[r] = np.bincount(owner[ghosts[r]], minlength=nranks)
    nbr = (needs + needs.T) > 0
    np.fill_diagonal(nbr, False)

    # mailbox layout of every rank: for rank q, one receive channel per neighbour p (ascending p)
    # channel = [data parity 0 | data parity 1 | 2 flags], cap words per parity
    chan = {}                                # (receiver q, sender p) -> (data_off, flag_off, cap)
    ends = []
    for q in range(nranks):
        off = _even(mailbox_base[q])
        for p in range(nranks):
            if nbr[q, p]:
                # same slot size both ways; every double travels as a flagged pair of words (csrc/sfem_dist.h)
                cap = _even(max(2 * int(max(needs[q, p], needs[p, q])) * nb_max, 2))
                chan[(q, p)] = (off, off + 2 * cap, cap)
                off += 2 * cap + 2
        ends.append(off)

    mine, gh = owned[rank], ghosts[rank]
    gh_owner = owner[gh] if len(gh) else np.zeros(0, dtype=np.int64)
    order = np.lexsort((gh, gh_owner))
    gh = gh[order]
    gh_owner = gh_owner[order]
    g2l = np.full(n, -1, dtype=np.int64)
    g2l[mine] = np.arange(len(mine))
    gap = int(ghost_gap)
    g2l[gh] = len(mine) + gap + np.arange(len(gh))
    lp = LevelPartition(rank, nranks, n, mine, gh, g2l, ghost_gap=gap)
    for q in range(nranks):
        if not nbr[rank, q]:
            continue
        lp.neighbors.append(q)
        # what q needs from me, in the order of q's ghost list (ascending global id within my rank's group)
        gq = ghosts[q]
        want = np.sort(gq[owner[gq] == rank]) if len(gq) else np.zeros(0, dtype=np.int64)
        lp.send_idx.append(g2l[want].astype(np.int32))
        sel = np.flatnonzero(gh_owner == q)
        lp.recv_off.append(int(len(mine) + gap + (sel[0] if len(sel) else 0)))
        lp.recv_cnt.append(int(len(sel)))
        d_off, f_off, cap = chan[(q, rank)]              # I write into q's channel for sender = me
        lp.peer_data_off.append(d_off); lp.peer_flag_off.append(f_off)
        d_off, f_off, cap2 = chan[(rank, q)]             # q writes into my channel for sender = q
        lp.my_data_off.append(d_off); lp.my_flag_off.append(f_off)
        assert cap == cap2
        lp.cap.append(cap)
    lp.mailbox_end = ends[rank]
    lp.all_mailbox_ends = ends          # of every rank: the next level's ``mailbox_base``
    return lp


def localize_csr(rowptr: np.ndarray, cols: np.ndarray, rows: np.ndarray, col_g2l: np.ndarray, vals: np.ndarray = None,
                 drop_missing: bool = False):
    """Rows ``rows`` (global ids) of a CSR matrix with columns renumbered by ``col_g2l``.

    Returns (rowptr_loc, cols_loc, slot) where ``slot`` are the positions of the kept entries in the
    global arrays (``vals_loc = vals[slot]``; also usable as the device extraction map).
    ``drop_missing``: silently drop entries whose column is not on the rank (restriction onto the
    replicated coarse hierarchy uses only owned columns); otherwise a missing column is an error.
    """
    rp = np.asarray(rowptr, dtype=np.int64)
    rows = np.asarray(rows, dtype=np.int64)
    lens = rp[rows + 1] - rp[rows]
    start = np.repeat(rp[rows], lens)
    within = np.arange(int(lens.sum()), dtype=np.int64) - np.repeat(np.cumsum(lens) - lens, lens)
    slot = start + within
    lc = col_g2l[np.asarray(cols, dtype=np.int64)[slot]]
    if drop_missing:
        keep = lc >= 0
        rid = np.repeat(np.arange(len(rows), dtype=np.int64), lens)[keep]
        slot, lc = slot[keep], lc[keep]
        lens = np.bincount(rid, minlength=len(rows))
    elif len(lc) and lc.min() < 0:
        raise ValueError("localize_csr: a column is neither owned nor ghost on this rank")
    rp_loc = np.concatenate([[0], np.cumsum(lens)]).astype(np.int32)
    return rp_loc, lc.astype(np.int32), slot.astype(np.int64)


def stokes_gaps(nv_own: int, n2_ghost: int) -> Tuple[int, int]:
    """Local numbering of a row-partitioned Taylor-Hood vector
    ``[u owned, interleaved | p owned | pad | u ghosts, interleaved | p ghosts]`` (csrc/sfem_stokes.cu):
    returns ``(hole, gap_p)`` -- ``hole`` = velocity PAIRS between the owned and the ghost velocity dofs in K's column
    numbering (the owned pressure entries live there), ``gap_p`` = entries between the owned and the ghost pressure
    dofs in B^T's column numbering (pad + the velocity ghosts)."""
    hole = (int(nv_own) + 1) // 2
    return hole, (2 * hole - int(nv_own)) + 2 * int(n2_ghost)


def emulate_exchange(parts: List[LevelPartition], xs: List[np.ndarray], nb: int = 1) -> None:
    """numpy model of the device halo exchange (tests): fills the ghost entries of every rank's vector."""
    for lp, x in zip(parts, xs):
        for k, q in enumerate(lp.neighbors):
            src = parts[q]
            kk = src.neighbors.index(lp.rank)
            sent = xs[q].reshape(-1, nb)[src.send_idx[kk]]
            assert len(sent) == lp.recv_cnt[k]
            x.reshape(-1, nb)[lp.recv_off[k]:lp.recv_off[k] + lp.recv_cnt[k]] = sent
