"""CPU oracle for lash's hot paths (ctypes binding of oracle/lash_oracle.c).

TEST INFRASTRUCTURE ONLY.  Import from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; never from lash_b200/.  Parity status: see the header of
lash_oracle.c ("parity unpinned" except XXH3).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liblash_oracle.so")

HMH, HLL, ULL = 0, 1, 2
FGRA, ML = 0, 1
BINOMIAL, POISSON = 0, 1


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "lash_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        u64, i32, dbl, vp = C.c_uint64, C.c_int, C.c_double, C.c_void_p
        L.lo_xxh3_64_le64.restype = u64
        L.lo_xxh3_64_le64.argtypes = [u64, u64]
        L.lo_xxh3_128_le32.restype = None
        L.lo_xxh3_128_le32.argtypes = [C.c_uint32, u64, C.POINTER(u64), C.POINTER(u64)]
        L.lo_filter_out_n.restype = C.c_size_t
        L.lo_filter_out_n.argtypes = [vp, C.c_size_t, vp]
        L.lo_mask_bits.restype = u64
        L.lo_mask_bits.argtypes = [u64, i32]
        L.lo_canonical_kmers.restype = C.c_size_t
        L.lo_canonical_kmers.argtypes = [vp, C.c_size_t, i32, vp]
        L.lo_reg_bytes.restype = C.c_size_t
        L.lo_reg_bytes.argtypes = [i32, i32]
        L.lo_add_kmer.restype = None
        L.lo_add_kmer.argtypes = [i32, i32, vp, u64, u64]
        L.lo_sketch_add_record.restype = None
        L.lo_sketch_add_record.argtypes = [i32, i32, i32, u64, vp, C.c_size_t, vp]
        L.lo_sketch_genomes.restype = i32
        L.lo_sketch_genomes.argtypes = [i32, i32, i32, u64, vp, vp, vp, u64, vp, i32]
        L.lo_hll_len.restype = dbl
        L.lo_hll_len.argtypes = [vp, i32, C.POINTER(i32)]
        L.lo_ull_fgra.restype = dbl
        L.lo_ull_fgra.argtypes = [vp, i32]
        L.lo_ull_ml.restype = dbl
        L.lo_ull_ml.argtypes = [vp, i32]
        L.lo_ull_ml_stats.restype = None
        L.lo_ull_ml_stats.argtypes = [vp, i32, C.POINTER(u64), vp]
        L.lo_ull_merge.restype = None
        L.lo_ull_merge.argtypes = [vp, vp, vp, i32]
        L.lo_ull_register_contributions.restype = C.POINTER(dbl)
        L.lo_ull_estimation_factor.restype = dbl
        L.lo_ull_estimation_factor.argtypes = [i32]
        L.lo_hmh_cardinality.restype = dbl
        L.lo_hmh_cardinality.argtypes = [vp]
        L.lo_hmh_similarity.restype = dbl
        L.lo_hmh_similarity.argtypes = [vp, vp]
        L.lo_hmh_counts.restype = None
        L.lo_hmh_counts.argtypes = [vp, vp, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        L.lo_hmh_expected_collisions.restype = dbl
        L.lo_hmh_expected_collisions.argtypes = [dbl, dbl]
        L.lo_compute_distance_f64.restype = dbl
        L.lo_compute_distance_f64.argtypes = [dbl, i32, i32]
        L.lo_compute_distance_f32.restype = C.c_float
        L.lo_compute_distance_f32.argtypes = [C.c_float, i32, i32]
        L.lo_cardinality.restype = dbl
        L.lo_cardinality.argtypes = [i32, i32, i32, vp, C.POINTER(i32)]
        L.lo_dist.restype = i32
        L.lo_dist.argtypes = [i32, i32, i32, i32, i32, i32, vp, u64, vp, u64, i32, vp, vp, i32]
        L.lo_selfcheck.restype = i32
        assert L.lo_selfcheck() == 1, 'oracle secret constants corrupted'
        _lib = L
    return _lib


def _ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def reg_dtype(algo: int):
    return np.uint16 if algo == HMH else np.uint8


def reg_count(algo: int, p: int) -> int:
    return 16384 if algo == HMH else (1 << p)


def xxh3_64_le64(v: int, seed: int) -> int:
    return int(lib().lo_xxh3_64_le64(v & (2**64 - 1), seed & (2**64 - 1)))


def xxh3_128_le32(w: int, seed: int) -> int:
    lo, hi = C.c_uint64(), C.c_uint64()
    lib().lo_xxh3_128_le32(w & 0xFFFFFFFF, seed & (2**64 - 1), C.byref(lo), C.byref(hi))
    return (hi.value << 64) | lo.value


def filter_out_n(seq: bytes) -> bytes:
    a = np.frombuffer(seq, dtype=np.uint8)
    out = np.empty(max(len(a), 1), dtype=np.uint8)
    n = lib().lo_filter_out_n(_ptr(a), len(a), _ptr(out))
    return out[:n].tobytes()


def canonical_kmers(filtered: bytes, k: int) -> np.ndarray:
    a = np.frombuffer(filtered, dtype=np.uint8)
    out = np.empty(max(len(a), 1), dtype=np.uint64)
    n = lib().lo_canonical_kmers(_ptr(a), len(a), k, _ptr(out))
    return out[:n].copy()


def sketch_genomes(algo: int, p: int, k: int, seed: int, genomes: list[list[bytes]], threads: int = 1) -> np.ndarray:
    """genomes: list of genomes, each a list of raw record sequences (bytes).  Returns
    [n_genomes, reg_count] registers (uint8, or uint16 for HMH)."""
    recs = [r for g in genomes for r in g]
    rec_off = np.zeros(len(recs) + 1, dtype=np.uint64)
    if recs:
        rec_off[1:] = np.cumsum([len(r) for r in recs], dtype=np.uint64)
    seqs = np.frombuffer(b"".join(recs) or b"\0", dtype=np.uint8)
    gen_rec = np.zeros(len(genomes) + 1, dtype=np.uint64)
    gen_rec[1:] = np.cumsum([len(g) for g in genomes], dtype=np.uint64)
    regs = np.zeros((len(genomes), reg_count(algo, p)), dtype=reg_dtype(algo))
    rc = lib().lo_sketch_genomes(algo, p, k, seed, _ptr(seqs), _ptr(rec_off), _ptr(gen_rec), len(genomes), _ptr(regs), threads)
    if rc != 0:
        raise ValueError(f"lo_sketch_genomes failed: {rc}")
    return regs


def cardinality(algo: int, p: int, estimator: int, regs: np.ndarray) -> float:
    regs = np.ascontiguousarray(regs)
    st = C.c_int(0)
    return float(lib().lo_cardinality(algo, p, estimator, _ptr(regs), C.byref(st)))


def ull_ml_stats(regs: np.ndarray, p: int):
    regs = np.ascontiguousarray(regs, dtype=np.uint8)
    S = C.c_uint64()
    b = np.zeros(66, dtype=np.int32)
    lib().lo_ull_ml_stats(_ptr(regs), p, C.byref(S), _ptr(b))
    return S.value, b


def ull_merge(a: np.ndarray, b: np.ndarray, p: int) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint8)
    b = np.ascontiguousarray(b, dtype=np.uint8)
    out = np.empty_like(a)
    lib().lo_ull_merge(_ptr(a), _ptr(b), _ptr(out), p)
    return out


def hmh_counts(a: np.ndarray, b: np.ndarray):
    a = np.ascontiguousarray(a, dtype=np.uint16)
    b = np.ascontiguousarray(b, dtype=np.uint16)
    c, n = C.c_uint32(), C.c_uint32()
    lib().lo_hmh_counts(_ptr(a), _ptr(b), C.byref(c), C.byref(n))
    return c.value, n.value


def dist(algo: int, p: int, k: int, estimator: int, model: int, fp32: bool, ref: np.ndarray, qry: np.ndarray,
         triangular: bool = False, threads: int = 1, return_flags: bool = False):
    """Dense [n_ref, n_qry] Mash distances (NaN-prefilled; triangular leaves j>i untouched)."""
    ref = np.ascontiguousarray(ref, dtype=reg_dtype(algo))
    qry = np.ascontiguousarray(qry, dtype=reg_dtype(algo))
    out = np.full((ref.shape[0], qry.shape[0]), np.nan, dtype=np.float32 if fp32 else np.float64)
    flags = np.zeros((ref.shape[0], qry.shape[0]), dtype=np.uint8)
    rc = lib().lo_dist(algo, p, k, estimator, model, int(fp32), _ptr(ref), ref.shape[0], _ptr(qry), qry.shape[0],
                       int(triangular), _ptr(out), _ptr(flags), threads)
    if rc != 0:
        raise ValueError(f"lo_dist failed: {rc}")
    return (out, flags) if return_flags else out
