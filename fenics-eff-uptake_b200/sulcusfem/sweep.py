"""Sweep sharding (BASELINE config 4): the reference's studies are serial Python loops over
independent cases (``no_advection_analysis_A.py:1306-1347``, ``no_advection_analysis_B.py:110-141``,
``adv_diff_analysis.py:201-260``).  Cases share nothing, so they are dealt round-robin to the ranks (one
process per GPU) and only the result dictionaries are gathered -- no data-path collective."""
from __future__ import annotations

from typing import Callable, List, Sequence


def shard_cases(cases: Sequence, rank: int, world: int) -> List:
    """Cases of ``rank``: static round-robin (case i -> rank i mod world)."""
    return [c for i, c in enumerate(cases) if i % world == rank]


def run_sharded(cases: Sequence, run_case: Callable, rank: int = 0, world: int = 1, gather: bool = True):
    """Run this rank's share; with ``gather`` and an initialised ``torch.distributed`` group every rank
    receives the full list of (case index, result) pairs in case order."""
    mine = [(i, run_case(c)) for i, c in enumerate(cases) if i % world == rank]
    if world == 1 or not gather:
        return mine
    import torch.distributed as dist
    parts = [None] * world
    dist.all_gather_object(parts, mine)
    return sorted((p for part in parts for p in part), key=lambda t: t[0])
