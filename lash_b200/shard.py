"""Multi-GPU partitioning (one process per GPU).

Sketching: genomes are independent units (reference: one rayon task per file, utils.rs:450-452)
-> longest-processing-time-first assignment by base count, no collective.
Distance: output rows are independent -> each rank takes a contiguous row range; for the
triangular (same_files) case the ranges are cut so every rank gets the same number of pairs.
The only exchange step is one broadcast of the sketch set, done by the caller with NCCL.
"""
from __future__ import annotations

import heapq
import math
from typing import Sequence


def genome_shards(sizes: Sequence[int], world: int) -> list[list[int]]:
    """Every rank's genome indices at once (LPT greedy: largest genome first onto the least loaded rank, ties to the
    lower rank; deterministic, so every rank computes the same partition).  O(n log n)."""
    heap = [(0, r) for r in range(world)]
    out: list[list[int]] = [[] for _ in range(world)]
    for g in sorted(range(len(sizes)), key=lambda i: (-sizes[i], i)):
        load, r = heapq.heappop(heap)
        out[r].append(g)
        heapq.heappush(heap, (load + sizes[g], r))
    return [sorted(s) for s in out]


def genome_shard(sizes: Sequence[int], rank: int, world: int) -> list[int]:
    """Indices of the genomes rank `rank` sketches."""
    return genome_shards(sizes, world)[rank]


def weighted_counts(n: int, weights: Sequence[float]) -> list[int]:
    """n units split over the ranks in proportion to `weights` (largest-remainder rounding, every rank with a positive
    weight gets at least one unit when n allows): link-aware genome shards -- a rank's weight is the host-to-device rate
    it measured while all ranks copy at once, so every rank finishes pushing its share at the same time."""
    w = [max(float(x), 0.0) for x in weights]
    tot = sum(w)
    if tot <= 0:
        w, tot = [1.0] * len(w), float(len(w))
    exact = [n * x / tot for x in w]
    counts = [int(e) for e in exact]
    order = sorted(range(len(w)), key=lambda i: (-(exact[i] - counts[i]), i))
    for i in order[: n - sum(counts)]:
        counts[i] += 1
    return counts


def gather_permutation(shards: Sequence[Sequence[int]], pad_to: int | None = None) -> list[int]:
    """Sketches come back from an all-gather rank after rank, each rank's block padded to `pad_to` rows (default: the
    largest shard).  perm[g] = row of global genome g in that rank-concatenated array, so
    `gathered.index_select(0, perm)` is the register array in list order (what the reference's Vec<S> is, utils.rs:507)."""
    pad = max((len(s) for s in shards), default=0) if pad_to is None else pad_to
    n = sum(len(s) for s in shards)
    perm = [-1] * n
    for r, s in enumerate(shards):
        if len(s) > pad:
            raise ValueError("shard larger than the padded block")
        for t, g in enumerate(s):
            if not 0 <= g < n or perm[g] != -1:
                raise ValueError("shards must partition range(n)")
            perm[g] = r * pad + t
    return perm


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
