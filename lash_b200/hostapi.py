"""ctypes binding of include/lash_host.h (liblash_host.so, the C++ host layer).

The product's host side is C++ (lash_b200/host/); this module only exposes its C ABI to the Python
tests and to bench.py, one prototype per exported symbol.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Iterator, Sequence

import numpy as np

from . import capi

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_lib", "liblash_host.so")

E_IO, E_FORMAT, E_PARAMS = -10, -11, -12


class SketchFilesStats(C.Structure):
    """struct lash_sketch_files_stats"""
    _fields_ = [("n_records", C.c_uint64), ("n_bases_in", C.c_uint64), ("n_bases_kept", C.c_uint64), ("n_pushes", C.c_uint64),
                ("seconds_total", C.c_double), ("gpu_kernel_ms", C.c_double), ("seconds_open", C.c_double),
                ("seconds_workers", C.c_double), ("seconds_drain", C.c_double)]


u64, i32, vp, sz, cp = C.c_uint64, C.c_int, C.c_void_p, C.c_size_t, C.c_char_p
PROTOTYPES = {
    "lash_host_last_error": (cp, []),
    "lash_fastx_open": (i32, [cp, C.POINTER(vp)]),
    "lash_fastx_next": (i32, [vp, C.POINTER(vp), C.POINTER(sz), C.POINTER(vp), C.POINTER(sz)]),
    "lash_fastx_close": (i32, [vp]),
    "lash_host_filter_pack": (i32, [vp, sz, vp, C.POINTER(u64), i32]),
    "lash_host_pack_has_simd": (i32, []),
    "lash_host_pack_isa": (i32, []),
    "lash_host_sketch_files_regs": (i32, [vp, i32, i32, i32, u64, C.POINTER(cp), u64, i32, u64, vp, C.POINTER(SketchFilesStats)]),
    "lash_host_sketch_files": (i32, [vp, i32, i32, i32, u64, C.POINTER(cp), u64, cp, i32, C.POINTER(SketchFilesStats)]),
    "lash_host_pack_files_dry": (i32, [C.POINTER(cp), u64, i32, i32, u64, C.POINTER(SketchFilesStats)]),
    "lash_host_release_pinned": (i32, []),
    "lash_host_set_ingest_mode": (i32, [i32]),
    "lash_host_write_parameters": (i32, [cp, i32, i32, i32, u64]),
    "lash_host_write_sketches": (i32, [cp, i32, i32, vp, u64, i32]),
    "lash_host_read_sketches": (i32, [cp, i32, C.POINTER(i32), u64, vp]),
    "lash_host_dist": (i32, [vp, cp, cp, cp, cp, i32, i32, i32, i32, i32]),
    "lash_host_dist_rows": (i32, [vp, cp, cp, cp, cp, i32, i32, i32, i32, i32, i32]),
    "lash_host_format_fixed6_f64": (i32, [C.c_double, C.c_char_p]),
    "lash_host_format_fixed6_f32": (i32, [C.c_float, C.c_char_p]),
    "lash_host_format_fixed6_bulk": (sz, [vp, sz, vp]),
}

_lib = None


def lib_path() -> str:
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C lash_b200/host)")
        capi.lib()  # liblash_gpu.so first (liblash_host.so links against it)
        L = C.CDLL(_SO)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class HostError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lash_host error {code}: {msg}")
        self.code = code


def check(rc: int) -> int:
    if rc < 0:
        raise HostError(rc, lib().lash_host_last_error().decode("utf-8", "replace"))
    return rc


def read_fastx(path: str) -> Iterator[tuple[bytes, bytes]]:
    """(id, seq) per record, like needletail's reader: seq has line breaks removed."""
    h = vp()
    check(lib().lash_fastx_open(os.fsencode(path), C.byref(h)))
    try:
        idp, seqp, idn, seqn = vp(), vp(), sz(), sz()
        while True:
            rc = check(lib().lash_fastx_next(h, C.byref(idp), C.byref(idn), C.byref(seqp), C.byref(seqn)))
            if rc == 0:
                return
            yield C.string_at(idp, idn.value), C.string_at(seqp, seqn.value)
    finally:
        lib().lash_fastx_close(h)


def filter_pack(seq: bytes, packed: np.ndarray | None = None, n_bases: int = 0, simd: bool | int = True) -> tuple[np.ndarray, int]:
    """filter_out_n + 2-bit pack; appends to an existing packed stream when given.
    simd: False/0 scalar table, True/1 the default SIMD path, 2 AVX2+BMI2, 3 AVX-512 VBMI2 (where the CPU has it)."""
    need = (n_bases + len(seq) + 3) // 4 + 16
    buf = np.zeros(need, dtype=np.uint8)
    if packed is not None:
        buf[: (n_bases + 3) // 4] = packed[: (n_bases + 3) // 4]
    nb = u64(n_bases)
    a = np.frombuffer(seq, dtype=np.uint8) if len(seq) else np.zeros(1, dtype=np.uint8)
    check(lib().lash_host_filter_pack(a.ctypes.data_as(vp), len(seq), buf.ctypes.data_as(vp), C.byref(nb), int(simd)))
    return buf[: (nb.value + 3) // 4].copy(), nb.value


def _files_arg(files: Sequence[str]):
    arr = (cp * max(len(files), 1))()
    for i, f in enumerate(files):
        arr[i] = os.fsencode(f)
    return arr


def sketch_files_regs(ctx, algo: int, p: int, k: int, seed: int, files: Sequence[str], threads: int = 0,
                      chunk_bytes: int = 0) -> tuple[np.ndarray, SketchFilesStats]:
    rb = capi.lib().lash_sketch_reg_bytes(algo, p)
    out = np.zeros((len(files), rb // (2 if algo == capi.ALGO_HMH else 1)), dtype=np.uint16 if algo == capi.ALGO_HMH else np.uint8)
    st = SketchFilesStats()
    check(lib().lash_host_sketch_files_regs(ctx.handle, algo, p, k, seed & (2**64 - 1), _files_arg(files), len(files), threads,
                                            chunk_bytes, out.ctypes.data_as(vp), C.byref(st)))
    return out, st


def pack_files_dry(files: Sequence[str], k: int, threads: int = 0, chunk_bytes: int = 0) -> SketchFilesStats:
    """Parse + filter + pack only (no GPU): the host ingest ceiling."""
    st = SketchFilesStats()
    check(lib().lash_host_pack_files_dry(_files_arg(files), len(files), k, threads, chunk_bytes, C.byref(st)))
    return st


def sketch_files(ctx, algo: int, p: int, k: int, seed: int, files: Sequence[str], output_name: str, threads: int = 0) -> SketchFilesStats:
    """sketch_files::<S> + the parameters JSON the `sketch` sub-command writes next to it."""
    st = SketchFilesStats()
    check(lib().lash_host_sketch_files(ctx.handle, algo, p, k, seed & (2**64 - 1), _files_arg(files), len(files),
                                       os.fsencode(output_name), threads, C.byref(st)))
    check(lib().lash_host_write_parameters(os.fsencode(output_name), algo, p, k, seed & (2**64 - 1)))
    return st


def write_sketches(path: str, algo: int, p: int, regs: np.ndarray, threads: int = 1) -> None:
    regs = np.ascontiguousarray(regs)
    check(lib().lash_host_write_sketches(os.fsencode(path), algo, p, regs.ctypes.data_as(vp), regs.shape[0], threads))


def read_sketches(path: str, algo: int, n: int, p: int = 0) -> tuple[np.ndarray, int]:
    pp = i32(p)
    if algo != capi.ALGO_HMH and p == 0:
        # two passes: first record tells the precision
        probe = np.zeros(1 << 26, dtype=np.uint8) if n else np.zeros(1, dtype=np.uint8)
        check(lib().lash_host_read_sketches(os.fsencode(path), algo, C.byref(pp), min(n, 1), probe.ctypes.data_as(vp)))
    rb = capi.lib().lash_sketch_reg_bytes(algo, pp.value if algo != capi.ALGO_HMH else 14)
    out = np.zeros((n, rb // (2 if algo == capi.ALGO_HMH else 1)), dtype=np.uint16 if algo == capi.ALGO_HMH else np.uint8)
    check(lib().lash_host_read_sketches(os.fsencode(path), algo, C.byref(pp), n, out.ctypes.data_as(vp)))
    return out, pp.value


def dist(ctx, ref_prefix: str, query_prefix: str, output_file: str, estimator: str = "fgra", model: int = 1, dm: bool = False,
         fp32: bool = False, threads: int = 1, fused: bool = True) -> int:
    return check(lib().lash_host_dist(ctx.handle, os.fsencode(ref_prefix), os.fsencode(query_prefix), os.fsencode(output_file),
                                      estimator.encode(), model, int(dm), int(fp32), threads, int(fused)))


def dist_rows(ctx, ref_prefix: str, query_prefix: str, output_file: str, rank: int, world: int, estimator: str = "fgra",
              model: int = 1, dm: bool = False, fp32: bool = False, threads: int = 1) -> int:
    return check(lib().lash_host_dist_rows(ctx.handle, os.fsencode(ref_prefix), os.fsencode(query_prefix), os.fsencode(output_file),
                                           estimator.encode(), model, int(dm), int(fp32), threads, rank, world))


def format_fixed6_bulk(values: np.ndarray) -> list[str]:
    """The fused writer's formatter over an array (fast path + exact fallback)."""
    v = np.ascontiguousarray(values, dtype=np.float64)
    buf = np.zeros(341 * max(len(v), 1), dtype=np.uint8)
    n = lib().lash_host_format_fixed6_bulk(v.ctypes.data_as(vp), len(v), buf.ctypes.data_as(vp))
    return buf[:n].tobytes().decode().split("\n")[:-1]


def format_fixed6(v: float, fp32: bool = False) -> str:
    buf = C.create_string_buffer(400)
    n = lib().lash_host_format_fixed6_f32(v, buf) if fp32 else lib().lash_host_format_fixed6_f64(v, buf)
    return buf.raw[:n].decode()
