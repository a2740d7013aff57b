"""Sweep sharding (BASELINE config 4): the reference's studies are serial Python loops over
independent cases (``no_advection_analysis_A.py:1306-1347``, ``no_advection_analysis_B.py:110-141``,
``adv_diff_analysis.py:201-260``).  Cases share nothing, so they are dealt round-robin to the ranks (one
process per GPU) and only the result dictionaries are gathered -- no data-path collective."""
from __future__ import annotations

import threading
from typing import Callable, List, Sequence

# Within one GPU a case at the reference's mesh size (~120 k dofs) is launch-latency bound -- every kernel of its Krylov
# graphs occupies a few SMs for a few microseconds -- so several cases can run at once: ``streams`` worker threads, each
# with its own CUDA stream and its own device problems (cache slot, see ``current_slot``); the host mesh, patterns and
# multigrid hierarchy of a geometry are shared.  The C ABI is thread-safe per handle and releases the GIL.
_tls = threading.local()
_warm_slots = set()          # slots whose worker has run at least one case (its device problems exist)


def current_slot() -> int:
    """Device-problem slot of the calling worker thread (0 outside ``run_concurrent``): part of the per-mesh cache keys
    of ``solvers`` / ``analysis`` so that concurrent cases never share device buffers."""
    return getattr(_tls, 'slot', 0)


def shard_cases(cases: Sequence, rank: int, world: int) -> List:
    """Cases of ``rank``: static round-robin (case i -> rank i mod world)."""
    return [c for i, c in enumerate(cases) if i % world == rank]


def run_concurrent(indexed_cases: Sequence, run_case: Callable, streams: int) -> List:
    """Run ``(index, case)`` pairs on ``streams`` worker threads, one CUDA stream and one device-problem slot each;
    returns ``(index, result)`` in index order.  The first case a slot ever runs is taken under a lock (it builds that
    slot's device problems: pattern sorts and uploads are not worth overlapping and the per-mesh caches fill in a fixed
    order); later sweeps start all workers at once."""
    import torch
    streams = max(1, min(int(streams), len(indexed_cases)))
    if streams == 1:
        return [(i, run_case(c)) for i, c in indexed_cases]
    out, errors = {}, []
    first = threading.Lock()
    device = torch.cuda.current_device()

    def worker(slot):
        try:
            torch.cuda.set_device(device)
            _tls.slot = slot
            st = torch.cuda.Stream()
            with torch.cuda.stream(st):
                mine = indexed_cases[slot::streams]
                for k, (i, c) in enumerate(mine):
                    if k == 0 and slot not in _warm_slots:
                        with first:                     # first case ever of this slot: builds its device problems
                            out[i] = run_case(c)
                            st.synchronize()
                        _warm_slots.add(slot)
                    else:
                        out[i] = run_case(c)
                st.synchronize()
        except BaseException as e:              # re-raised in the caller
            errors.append(e)
        finally:
            _tls.slot = 0
    torch.cuda.current_stream().synchronize()
    threads = [threading.Thread(target=worker, args=(s,), name=f"sfem-sweep-{s}") for s in range(streams)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return sorted(out.items(), key=lambda t: t[0])


def run_sharded(cases: Sequence, run_case: Callable, rank: int = 0, world: int = 1, gather: bool = True, streams: int = 1):
    """Run this rank's share; with ``gather`` and an initialised ``torch.distributed`` group every rank
    receives the full list of (case index, result) pairs in case order.  ``streams`` > 1: the rank's cases run
    concurrently on that many CUDA streams (``run_concurrent``)."""
    mine = run_concurrent([(i, c) for i, c in enumerate(cases) if i % world == rank], run_case, streams)
    if world == 1 or not gather:
        return mine
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    return sorted((p for part in parts for p in part), key=lambda t: t[0])
