"""ctypes binding of include/lash_gpu.h (one prototype per exported symbol)."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LASH_GPU_LIB: load a tuning build of the same library (tools/, never the tests)
_SO = os.environ.get("LASH_GPU_LIB") or os.path.join(_HERE, "_lib", "liblash_gpu.so")

ALGO_HMH, ALGO_HLL, ALGO_ULL = 0, 1, 2
EST_FGRA, EST_ML = 0, 1
MODEL_BINOMIAL, MODEL_POISSON, MODEL_FRAC = 0, 1, 2
W_HLL_BIAS_REGIME = 1


class LashError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"lash_gpu error {code}: {msg}")
        self.code = code


class Span(C.Structure):
    """struct lash_span"""
    _fields_ = [("genome", C.c_uint64), ("byte_off", C.c_uint64), ("n_bases", C.c_uint64), ("rec_first", C.c_uint64),
                ("n_rec", C.c_uint32), ("rec_len", C.c_uint32)]


class TextSpan(C.Structure):
    """struct lash_text_span"""
    _fields_ = [("genome", C.c_uint64), ("byte_off", C.c_uint64), ("n_bytes", C.c_uint64), ("n_rec", C.c_uint32), ("reserved", C.c_uint32)]


TEXT_RECORD_SEP = 0x01
DIST_BLOCK_CB = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p)

u64, i32, vp, sz = C.c_uint64, C.c_int, C.c_void_p, C.c_size_t
PROTOTYPES = {
    # name: (restype, argtypes)
    "lash_gpu_last_error": (C.c_char_p, []),
    "lash_gpu_abi_version": (i32, []),
    "lash_gpu_device_count": (i32, []),
    "lash_ctx_create": (i32, [i32, C.POINTER(vp)]),
    "lash_ctx_destroy": (i32, [vp]),
    "lash_ctx_device": (i32, [vp]),
    "lash_bind_thread_to_device": (i32, [i32]),
    "lash_host_alloc": (i32, [sz, C.POINTER(vp)]),
    "lash_host_free": (i32, [vp]),
    "lash_sketch_reg_bytes": (sz, [i32, i32]),
    "lash_sketch_padded_bytes": (u64, [u64]),
    "lash_sketch_open": (i32, [vp, i32, i32, i32, u64, u64, C.POINTER(vp)]),
    "lash_sketch_push": (i32, [vp, vp, u64, C.POINTER(Span), C.c_uint32, vp, u64, C.POINTER(u64)]),
    "lash_sketch_push_dev": (i32, [vp, vp, u64, C.POINTER(Span), C.c_uint32, vp, u64, C.POINTER(u64)]),
    "lash_sketch_push_ascii": (i32, [vp, vp, u64, C.POINTER(TextSpan), C.c_uint32, C.POINTER(u64)]),
    "lash_sketch_push_ascii_dev": (i32, [vp, vp, u64, C.POINTER(TextSpan), C.c_uint32, C.POINTER(u64)]),
    "lash_sketch_wait_copied": (i32, [vp, u64]),
    "lash_sketch_sync": (i32, [vp]),
    "lash_sketch_fetch": (i32, [vp, u64, u64, vp]),
    "lash_sketch_regs_dev": (i32, [vp, C.POINTER(vp)]),
    "lash_sketch_reset": (i32, [vp]),
    "lash_sketch_set_stream": (i32, [vp, vp]),
    "lash_sketch_stats": (i32, [vp, C.POINTER(C.c_double), C.POINTER(u64)]),
    "lash_sketch_close": (i32, [vp]),
    "lash_sketch_merge_dev": (i32, [vp, i32, i32, vp, vp, u64, vp]),
    "lash_sketch_merge": (i32, [vp, i32, i32, vp, vp, u64]),
    "lash_dist": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, u64, vp, u64, i32, vp]),
    "lash_dist_dev": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, u64, vp, u64, vp, vp, i32, u64, u64, vp, vp, vp]),
    "lash_cardinality_dev": (i32, [vp, i32, i32, i32, vp, u64, vp, vp]),
    "lash_cardinality": (i32, [vp, i32, i32, i32, vp, u64, vp]),
    "lash_dist_stream": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, u64, vp, u64, i32, u64, DIST_BLOCK_CB, vp]),
    "lash_dist_stream_rows": (i32, [vp, i32, i32, i32, i32, i32, i32, vp, u64, vp, u64, i32, u64, u64, u64, DIST_BLOCK_CB, vp]),
    "lash_dist_stats": (i32, [vp, C.POINTER(C.c_double), C.POINTER(u64)]),
    "lash_dist_set_checksum": (i32, [vp, i32]),
    "lash_dist_checksum": (i32, [vp, C.POINTER(u64), C.POINTER(u64)]),
    "lash_dist_checksum_dev": (i32, [vp, i32, vp, u64, i32, u64, u64, vp, vp]),
}

_lib = None


def lib_path() -> str:
    return _SO


def lib() -> C.CDLL:
    """Load liblash_gpu.so.  Fails loudly if the CUDA extension has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            raise ImportError(f"{_SO} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C lash_b200/csrc).  There is no CPU fallback.")
        L = C.CDLL(_SO)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> int:
    """Raise on negative return codes, pass warnings (positive) through."""
    if rc < 0:
        raise LashError(rc, lib().lash_gpu_last_error().decode("utf-8", "replace"))
    return rc
