"""Python host mirror of the reference's operator interface for the two hot paths.

  sketch side: ``KmerSketch`` (src/utils.rs:377-386) / ``sketch_files`` (utils.rs:439-510)
  dist side:   ``hmh_distance`` / ``ull_distance`` / ``hll_distance`` (utils.rs:84-373) + the
               ``print_dist`` name rule (main.rs:452-456)

Everything computes through the C ABI (liblash_gpu.so); nothing here does sketch or distance
arithmetic on the CPU.
"""
from __future__ import annotations

import ctypes as C
from typing import Callable, Sequence

import numpy as np

from . import capi
from .capi import ALGO_HLL, ALGO_HMH, ALGO_ULL, TEXT_RECORD_SEP, Span, TextSpan, check, lib
from .pack import PackedBatch

ALGO_BY_NAME = {"hmh": ALGO_HMH, "hll": ALGO_HLL, "ull": ALGO_ULL}  # main.rs:210-246
EST_BY_NAME = {"fgra": capi.EST_FGRA, "ml": capi.EST_ML}            # utils.rs:214-218


def reg_dtype(algo: int):
    return np.uint16 if algo == ALGO_HMH else np.uint8


def reg_count(algo: int, p: int) -> int:
    return 16384 if algo == ALGO_HMH else (1 << p)


class Context:
    """lash_ctx: one per GPU."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        check(lib().lash_ctx_create(device, C.byref(self._h)))

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            lib().lash_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Sketcher:
    """lash_sketcher: accumulators for n_genomes sketches of one (algo, p, k, seed)."""

    def __init__(self, ctx: Context, algo: int, p: int, k: int, seed: int, n_genomes: int):
        self.ctx, self.algo, self.p, self.k, self.n_genomes = ctx, algo, p, k, n_genomes
        self._h = C.c_void_p()
        self._keep = []  # host buffers that must outlive the async copies
        check(lib().lash_sketch_open(ctx.handle, algo, p, k, seed & (2**64 - 1), n_genomes, C.byref(self._h)))

    def push_batch(self, batch: PackedBatch) -> int:
        buf = batch.buffer()
        spans = (Span * max(len(batch.spans), 1))()
        for i, (g, off, nb, rf, nr) in enumerate(batch.spans):
            spans[i] = Span(g, off, nb, rf, nr, 0)
        recs = np.asarray(batch.rec_start, dtype=np.uint64)
        ticket = C.c_uint64()
        check(lib().lash_sketch_push(self._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes, spans, len(batch.spans),
                                     recs.ctypes.data_as(C.c_void_p) if len(recs) else None, len(recs), C.byref(ticket)))
        self._keep.append((buf, spans, recs))
        return ticket.value

    def push_raw(self, packed_ptr: int, n_bytes: int, spans, n_spans: int, rec_ptr=None, n_rec: int = 0, dev: bool = False) -> int:
        ticket = C.c_uint64()
        fn = lib().lash_sketch_push_dev if dev else lib().lash_sketch_push
        check(fn(self._h, C.c_void_p(packed_ptr), n_bytes, spans, n_spans, rec_ptr, n_rec, C.byref(ticket)))
        return ticket.value

    def push_text(self, genomes: Sequence[tuple[int, Sequence[bytes]]]) -> int:
        """lash_sketch_push_ascii: genomes = [(slot, [raw record bytes, ...])].  Records of a genome are joined with the
        in-band separator; the device does filter_out_n + 2-bit packing (utils.rs:33-41,464)."""
        sep = bytes([TEXT_RECORD_SEP])
        parts, spans, off = [], [], 0
        for slot, recs in genomes:
            recs = list(recs)
            if any(sep in r for r in recs):
                raise ValueError("a sequence byte equals LASH_TEXT_RECORD_SEP")
            body = sep.join(recs)
            spans.append((slot, off, len(body), len(recs)))
            room = (len(body) + 15) // 16 * 16 + 16
            parts.append(body + b"\xff" * (room - len(body)))   # padding the device must ignore
            off += room
        buf = np.frombuffer(b"".join(parts) or b"\0" * 16, dtype=np.uint8)
        arr = (TextSpan * max(len(spans), 1))()
        for i, (slot, o, n, nr) in enumerate(spans):
            arr[i] = TextSpan(slot, o, n, nr, 0)
        ticket = C.c_uint64()
        check(lib().lash_sketch_push_ascii(self._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes, arr, len(spans), C.byref(ticket)))
        self._keep.append((buf, arr))
        return ticket.value

    def wait_copied(self, ticket: int):
        check(lib().lash_sketch_wait_copied(self._h, ticket))

    def sync(self):
        check(lib().lash_sketch_sync(self._h))
        self._keep.clear()

    def fetch(self, first: int = 0, n: int | None = None) -> np.ndarray:
        n = self.n_genomes - first if n is None else n
        out = np.empty((n, reg_count(self.algo, self.p)), dtype=reg_dtype(self.algo))
        check(lib().lash_sketch_fetch(self._h, first, n, out.ctypes.data_as(C.c_void_p)))
        self._keep.clear()
        return out

    def regs_dev(self) -> int:
        p = C.c_void_p()
        check(lib().lash_sketch_regs_dev(self._h, C.byref(p)))
        return p.value

    def reset(self):
        check(lib().lash_sketch_reset(self._h))

    def set_stream(self, cuda_stream: int | None):
        check(lib().lash_sketch_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    def stats(self) -> tuple[float, int]:
        ms, n = C.c_double(), C.c_uint64()
        check(lib().lash_sketch_stats(self._h, C.byref(ms), C.byref(n)))
        return ms.value, n.value

    def close(self):
        if self._h:
            lib().lash_sketch_close(self._h)
            self._h = C.c_void_p()
        self._keep.clear()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def sketch_genomes(ctx: Context, algo: int, p: int, k: int, seed: int, genomes: Sequence[Sequence[bytes]],
                   genomes_per_push: int = 64) -> np.ndarray:
    """sketch_files (utils.rs:439-510) for in-memory inputs: genomes[g] is the list of record
    sequences of file g (raw bytes, unfiltered).  Returns registers [n_genomes, reg_count] in list order."""
    with Sketcher(ctx, algo, p, k, seed, max(len(genomes), 1)) as sk:
        for g0 in range(0, len(genomes), genomes_per_push):
            b = PackedBatch()
            for g in range(g0, min(len(genomes), g0 + genomes_per_push)):
                b.add_genome(g, list(genomes[g]))
            sk.push_batch(b)
        regs = sk.fetch()
    return regs[: len(genomes)]


def sketch_genomes_text(ctx: Context, algo: int, p: int, k: int, seed: int, genomes: Sequence[Sequence[bytes]],
                        genomes_per_push: int = 64) -> np.ndarray:
    """sketch_genomes through lash_sketch_push_ascii: raw record bytes go to the GPU, which filters and packs them."""
    with Sketcher(ctx, algo, p, k, seed, max(len(genomes), 1)) as sk:
        for g0 in range(0, len(genomes), genomes_per_push):
            sk.push_text([(g, genomes[g]) for g in range(g0, min(len(genomes), g0 + genomes_per_push))])
        regs = sk.fetch()
    return regs[: len(genomes)]


def merge(ctx: Context, algo: int, p: int, dst: np.ndarray, src: np.ndarray) -> np.ndarray:
    """UltraLogLog::merge / HyperLogLog::union / hyperminhash merge of register arrays (returns a new array)."""
    out = np.ascontiguousarray(dst, dtype=reg_dtype(algo)).copy()
    src = np.ascontiguousarray(src, dtype=reg_dtype(algo))
    assert out.shape == src.shape
    n = out.shape[0] if out.ndim == 2 else 1
    check(lib().lash_sketch_merge(ctx.handle, algo, p, out.ctypes.data_as(C.c_void_p), src.ctypes.data_as(C.c_void_p), n))
    return out


def cardinality(ctx: Context, algo: int, p: int, estimator: int, regs: np.ndarray) -> np.ndarray:
    regs = np.ascontiguousarray(regs, dtype=reg_dtype(algo))
    out = np.empty(regs.shape[0], dtype=np.float64)
    check(lib().lash_cardinality(ctx.handle, algo, p, estimator, regs.ctypes.data_as(C.c_void_p), regs.shape[0],
                                 out.ctypes.data_as(C.c_void_p)))
    return out


def dist(ctx: Context, algo: int, p: int, k: int, estimator: int, model: int, fp32: bool, ref: np.ndarray, qry: np.ndarray,
         triangular: bool = False) -> tuple[np.ndarray, int]:
    """lash_dist.  Returns (distances, warning).  Dense [n_ref, n_qry]; triangular: packed lower
    triangle (row i at i*(i+1)/2)."""
    ref = np.ascontiguousarray(ref, dtype=reg_dtype(algo))
    same = qry is ref
    qry = ref if same else np.ascontiguousarray(qry, dtype=reg_dtype(algo))
    n_ref, n_qry = ref.shape[0], qry.shape[0]
    dt = np.float32 if fp32 else np.float64
    out = np.full(n_ref * (n_ref + 1) // 2 if triangular else n_ref * n_qry, np.nan, dtype=dt)
    w = check(lib().lash_dist(ctx.handle, algo, p, k, estimator, model, int(fp32), ref.ctypes.data_as(C.c_void_p), n_ref,
                              qry.ctypes.data_as(C.c_void_p), n_qry, int(triangular), out.ctypes.data_as(C.c_void_p)))
    return (out if triangular else out.reshape(n_ref, n_qry)), w


def dist_stream(ctx: Context, algo: int, p: int, k: int, estimator: int, model: int, fp32: bool, ref: np.ndarray,
                qry: np.ndarray, triangular: bool, rows_per_block: int, on_block: Callable[[int, np.ndarray], None]) -> int:
    ref = np.ascontiguousarray(ref, dtype=reg_dtype(algo))
    qry = ref if qry is ref else np.ascontiguousarray(qry, dtype=reg_dtype(algo))
    n_qry = qry.shape[0]
    dt = np.float32 if fp32 else np.float64

    def _cb(user, row0, nrows, ptr):
        arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_float if fp32 else C.c_double)), shape=(nrows * n_qry,))
        on_block(int(row0), arr.view(dt).reshape(nrows, n_qry))
        return 0

    cb = capi.DIST_BLOCK_CB(_cb)
    return check(lib().lash_dist_stream(ctx.handle, algo, p, k, estimator, model, int(fp32), ref.ctypes.data_as(C.c_void_p),
                                        ref.shape[0], qry.ctypes.data_as(C.c_void_p), n_qry, int(triangular), rows_per_block,
                                        cb, None))


def dist_stats(ctx: Context) -> tuple[float, int]:
    ms, n = C.c_double(), C.c_uint64()
    check(lib().lash_dist_stats(ctx.handle, C.byref(ms), C.byref(n)))
    return ms.value, n.value


# --------------------------------------------------------------------------------------------------
# the reference's three *_distance entry points (utils.rs:84-94, 186-197, 290-299)
# --------------------------------------------------------------------------------------------------
def _distance(ctx, algo, p, k, estimator, model, fp32, reference_names, ref_regs, query_names, qry_regs, create_matrix,
              same_files, emit):
    """Shared body.  emit(rows) receives, per reference, a list of (ref_name, query_name, distance)
    -- the reference emits `frac` and applies compute_distance in print_dist (main.rs:455); here
    compute_distance is fused into the kernel, and the name-equality => 0 rule (main.rs:452-453) is
    applied on the host, which owns the names.  Row order is deterministic (list order); the
    reference's is HashMap order, only the *set* of pairs is contractual."""
    d, w = dist(ctx, algo, p, k, estimator, model, fp32, ref_regs, ref_regs if same_files else qry_regs, triangular=same_files)
    if create_matrix:
        emit([("", q, 1.0) for q in query_names])  # header row signal, utils.rs:133-146
    for i, rn in enumerate(reference_names):
        if same_files:
            row = d[i * (i + 1) // 2: i * (i + 1) // 2 + i + 1]
            cols = query_names[: i + 1]
        else:
            row, cols = d[i], query_names
        emit([(rn, qn, (0.0 if qn == rn else float(x))) for qn, x in zip(cols, row)])
    return w


def hmh_distance(ctx, k, model, fp32, reference_names, ref_regs, query_names, qry_regs, create_matrix, same_files, emit):
    return _distance(ctx, ALGO_HMH, 14, k, 0, model, fp32, reference_names, ref_regs, query_names, qry_regs, create_matrix,
                     same_files, emit)


def ull_distance(ctx, p, k, model, fp32, reference_names, ref_regs, query_names, qry_regs, estimator, create_matrix, same_files,
                 emit):
    if estimator not in EST_BY_NAME:
        raise ValueError("estimator needs to be either fgra or ml")  # utils.rs:217
    return _distance(ctx, ALGO_ULL, p, k, EST_BY_NAME[estimator], model, fp32, reference_names, ref_regs, query_names,
                     qry_regs, create_matrix, same_files, emit)


def hll_distance(ctx, p, k, model, fp32, reference_names, ref_regs, query_names, qry_regs, create_matrix, same_files, emit):
    return _distance(ctx, ALGO_HLL, p, k, 0, model, fp32, reference_names, ref_regs, query_names, qry_regs, create_matrix,
                     same_files, emit)
