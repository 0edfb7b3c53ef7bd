"""Host-side filter + 2-bit packer (numpy): what the host does before lash_sketch_push.

Follows the reference's front end: filter_out_n (src/utils.rs:33-41: keep uppercase A/C/G/T only,
delete everything else) and kmerutils' 2-bit alphabet A=0 C=1 G=2 T=3 (utils.rs:464).  Packed
format of the C ABI: four bases per byte, first base in the two most significant bits.
"""
from __future__ import annotations

import numpy as np

_LUT = np.full(256, 255, dtype=np.uint8)
for _i, _c in enumerate(b"ACGT"):
    _LUT[_c] = _i


def encode_record(seq: bytes | np.ndarray) -> np.ndarray:
    """filter_out_n + 2-bit codes (one uint8 per kept base)."""
    a = np.frombuffer(seq, dtype=np.uint8) if isinstance(seq, (bytes, bytearray, memoryview)) else seq
    codes = _LUT[a]
    return codes[codes != 255]


def pack_codes(codes: np.ndarray) -> np.ndarray:
    """Dense 2-bit packing, first base in the high bits of each byte."""
    n = len(codes)
    pad = (-n) % 4
    if pad:
        codes = np.concatenate([codes, np.zeros(pad, dtype=np.uint8)])
    q = codes.reshape(-1, 4)
    return ((q[:, 0] << 6) | (q[:, 1] << 4) | (q[:, 2] << 2) | q[:, 3]).astype(np.uint8)


def padded_bytes(n_bases: int) -> int:
    """== lash_sketch_padded_bytes"""
    b = (n_bases + 3) // 4
    return ((b + 15) // 16) * 16 + 16


class PackedBatch:
    """A push buffer under construction: spans of genomes packed back to back at 16-byte offsets."""

    def __init__(self):
        self.chunks: list[np.ndarray] = []
        self.spans: list[tuple[int, int, int, int, int]] = []  # genome, byte_off, n_bases, rec_first, n_rec
        self.rec_start: list[int] = []
        self.n_bytes = 0

    def add_genome(self, genome: int, records: list[bytes], k: int | None = None) -> None:
        """Append one span holding `records` of genome slot `genome`.  Records shorter than k may be
        dropped by the host (the kernel ignores them anyway, utils.rs:460-462); k=None keeps all."""
        codes = [encode_record(r) for r in records]
        if k is not None:
            codes = [c for c in codes if len(c) >= k]
        lens = [len(c) for c in codes]
        n_bases = int(sum(lens))
        packed = pack_codes(np.concatenate(codes) if codes else np.zeros(0, dtype=np.uint8))
        room = padded_bytes(n_bases)
        buf = np.zeros(room, dtype=np.uint8)
        buf[: len(packed)] = packed
        rec_first = len(self.rec_start)
        n_rec = len(codes)
        if n_rec > 1:
            self.rec_start.extend(np.concatenate([[0], np.cumsum(lens)]).astype(np.uint64).tolist())
        self.spans.append((genome, self.n_bytes, n_bases, rec_first, n_rec))
        self.chunks.append(buf)
        self.n_bytes += room

    def buffer(self) -> np.ndarray:
        return np.concatenate(self.chunks) if self.chunks else np.zeros(16, dtype=np.uint8)
