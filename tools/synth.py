"""Deterministic synthetic genomes for tests and bench.py.

A shared random "ancestor" of i.i.d. uniform ACGT; genome g = ancestor with substitutions at rate
mu_g (log-uniform in [1e-3, 0.3]) so Mash distances span (0, 1].  numpy only (CPU); bench.py has a
torch twin that builds the same kind of data directly in HBM.
"""
from __future__ import annotations

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def ancestor_codes(length: int, seed: int = 42) -> np.ndarray:
    return np.random.default_rng(seed).integers(0, 4, size=length, dtype=np.uint8)


def mutation_rate(g: int, seed: int = 42) -> float:
    r = np.random.default_rng(seed + 1000003 * (g + 1))
    return float(10 ** r.uniform(-3.0, np.log10(0.3)))


def mutate(codes: np.ndarray, mu: float, seed: int) -> np.ndarray:
    r = np.random.default_rng(seed)
    hit = r.random(len(codes)) < mu
    out = codes.copy()
    out[hit] = (out[hit] + r.integers(1, 4, size=int(hit.sum()), dtype=np.uint8)) & 3
    return out


def to_ascii(codes: np.ndarray) -> bytes:
    return ACGT[codes].tobytes()


def genomes(n: int, length: int, seed: int = 42) -> list[list[bytes]]:
    """n single-record genomes (uppercase ACGT)."""
    anc = ancestor_codes(length, seed)
    return [[to_ascii(mutate(anc, mutation_rate(g, seed), seed + g))] for g in range(n)]


def dirty_genome(length: int, k: int, seed: int = 7) -> list[bytes]:
    """Multi-record genome with lowercase runs, N runs, IUPAC codes, records shorter than k, an
    empty record and a record that becomes shorter than k only after filtering."""
    r = np.random.default_rng(seed)
    recs: list[bytes] = []
    remaining = length
    while remaining > 0:
        n = int(min(remaining, r.integers(1, max(2, length // 6))))
        remaining -= n
        s = bytearray(to_ascii(r.integers(0, 4, size=n, dtype=np.uint8)))
        for _ in range(int(r.integers(0, 4))):  # dirt
            a = int(r.integers(0, max(1, n)))
            b = int(min(n, a + r.integers(1, 40)))
            kind = int(r.integers(0, 3))
            if kind == 0:
                s[a:b] = bytes(s[a:b]).lower()
            elif kind == 1:
                s[a:b] = b"N" * (b - a)
            else:
                s[a:b] = bytes(r.choice(np.frombuffer(b"RYKMSWnacgt-*", dtype=np.uint8), size=b - a))
        recs.append(bytes(s))
    recs.insert(1, b"ACGT"[: max(0, min(4, k - 1))])   # shorter than k
    recs.insert(2, b"")                                  # empty
    recs.append(b"A" * (k - 1) + b"n" * 5 + b"")        # k-1 valid bases after filtering
    recs.append(b"acgtn" * 10)                           # nothing survives the filter
    recs.append(to_ascii(r.integers(0, 4, size=k, dtype=np.uint8)))  # exactly one k-mer
    return recs
