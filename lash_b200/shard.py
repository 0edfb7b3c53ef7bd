"""Multi-GPU partitioning (one process per GPU).

Sketching: genomes are independent units (reference: one rayon task per file, utils.rs:450-452)
-> longest-processing-time-first assignment by base count, no collective.
Distance: output rows are independent -> each rank takes a contiguous row range; for the
triangular (same_files) case the ranges are cut so every rank gets the same number of pairs.
The only exchange step is one broadcast of the sketch set, done by the caller with NCCL.
"""
from __future__ import annotations

import math
from typing import Sequence


def genome_shard(sizes: Sequence[int], rank: int, world: int) -> list[int]:
    """Indices of the genomes rank `rank` sketches (LPT greedy, deterministic on every rank)."""
    loads = [0] * world
    mine: list[int] = []
    for g in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        r = min(range(world), key=lambda j: (loads[j], j))
        loads[r] += sizes[g]
        if r == rank:
            mine.append(g)
    return sorted(mine)


def row_shard(n_rows: int, rank: int, world: int, triangular: bool) -> tuple[int, int]:
    """[begin, end) reference rows of rank `rank`."""
    def cut(r: int) -> int:
        if r <= 0:
            return 0
        if r >= world:
            return n_rows
        if not triangular:
            return (n_rows * r) // world
        # rows [0, x) hold x(x+1)/2 pairs: solve x(x+1)/2 = total * r / world
        total = n_rows * (n_rows + 1) / 2
        x = (math.sqrt(1 + 8 * total * r / world) - 1) / 2
        return min(n_rows, max(0, int(round(x))))
    return cut(rank), cut(rank + 1)


def pair_count(rows: tuple[int, int], n_qry: int, triangular: bool) -> int:
    b, e = rows
    if triangular:
        return e * (e + 1) // 2 - b * (b + 1) // 2
    return (e - b) * n_qry


def read_shard(n_records: int, rank: int, world: int) -> tuple[int, int]:
    """[begin, end) records of ONE sample that rank `rank` sketches (config "100 Gbp of reads -> one
    sketch per sample").  Every rank folds the world's partial sketches afterwards (all-gather of
    world x reg_bytes, then lash_sketch_merge_dev): registers are order-free, so any partition of the
    records gives the same sketch."""
    return (n_records * rank) // world, (n_records * (rank + 1)) // world
