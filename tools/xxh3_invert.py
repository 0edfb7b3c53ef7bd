"""Inverse of XXH3-64 on 8-byte inputs (the 4..8-byte short path is a bijection of u64 for a fixed
seed).  Used by tests to build adversarial k-mers whose hash has a chosen bit pattern, e.g. >= 32
leading zeros after the index bits -- a 2^-32 event that random genomes never exercise."""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1
MX2 = 0x9FB21C651E98DF25
MX2_INV = pow(MX2, -1, 1 << 64)
SECRET_X_8_16 = 0xC73AB174C5ECD5A2


def _rotl(x, r):
    return ((x << r) | (x >> (64 - r))) & M64


def _bswap32(x):
    return int.from_bytes(x.to_bytes(4, "little"), "big")


def _linear_inverse():
    # L(h) = h ^ rotl(h,49) ^ rotl(h,24) as a 64x64 matrix over GF(2); invert by Gauss-Jordan
    n = 64
    A = np.zeros((n, 2 * n), dtype=np.uint8)
    for j in range(n):
        col = (1 << j) ^ _rotl(1 << j, 49) ^ _rotl(1 << j, 24)
        for i in range(n):
            A[i, j] = (col >> i) & 1
        A[j, n + j] = 1
    for c in range(n):
        piv = next(r for r in range(c, n) if A[r, c])
        A[[c, piv]] = A[[piv, c]]
        for r in range(n):
            if r != c and A[r, c]:
                A[r] ^= A[c]
    inv = A[:, n:]
    return [int(sum(int(inv[i, j]) << i for i in range(n))) for j in range(n)]  # columns as ints


_LINV = _linear_inverse()


def xxh3_64_le64(v: int, seed: int) -> int:
    s = seed ^ (_bswap32(seed & 0xFFFFFFFF) << 32)
    bitflip = (SECRET_X_8_16 - s) & M64
    in64 = ((v >> 32) | ((v & 0xFFFFFFFF) << 32)) & M64
    h = in64 ^ bitflip
    h ^= _rotl(h, 49) ^ _rotl(h, 24)
    h = (h * MX2) & M64
    h ^= (h >> 35) + 8
    h = (h * MX2) & M64
    return h ^ (h >> 28)


def invert_xxh3_64_le64(h: int, seed: int) -> int:
    """The u64 v with xxh3_64_with_seed(v.to_le_bytes(), seed) == h."""
    d = h ^ (h >> 28) ^ (h >> 56)
    c = (d * MX2_INV) & M64
    b = c ^ ((c >> 35) + 8)
    a = (b * MX2_INV) & M64
    h0 = 0
    for j in range(64):
        if (a >> j) & 1:
            h0 ^= _LINV[j]
    s = seed ^ (_bswap32(seed & 0xFFFFFFFF) << 32)
    in64 = h0 ^ ((SECRET_X_8_16 - s) & M64)
    return ((in64 >> 32) | ((in64 & 0xFFFFFFFF) << 32)) & M64


def kmer_to_seq(v: int, k: int) -> bytes:
    return bytes(b"ACGT"[(v >> (2 * (k - 1 - i))) & 3] for i in range(k))


def revcomp_value(v: int, k: int) -> int:
    r = 0
    for i in range(k):
        r = (r << 2) | (3 - ((v >> (2 * i)) & 3))
    return r


def adversarial_32mers(targets, seed: int) -> list[bytes]:
    """For each target hash that admits one, the 32-mer whose CANONICAL value hashes to it."""
    out = []
    for h in targets:
        v = invert_xxh3_64_le64(h, seed)
        if v <= revcomp_value(v, 32):   # v is its own canonical form
            out.append(kmer_to_seq(v, 32))
    return out
