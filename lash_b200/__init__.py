"""lash_b200 -- B200 (sm_100a) implementation of lash's two hot paths behind a C ABI.

The product is ``lash_b200/_lib/liblash_gpu.so`` (CUDA, built from ``lash_b200/csrc``; contract in
``include/lash_gpu.h``).  This package is the thin Python host used by tests and bench.py:
ctypes bindings (:mod:`lash_b200.capi`), the host-side 2-bit packer (:mod:`lash_b200.pack`) and a
mirror of the reference's sketch/dist operator interface (:mod:`lash_b200.ops`).
There is no CPU fallback: importing works without a GPU, computing does not.
"""
from .capi import (ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, MODEL_BINOMIAL, MODEL_FRAC, MODEL_POISSON, LashError, lib,
                   lib_path)

__all__ = ["ALGO_HLL", "ALGO_HMH", "ALGO_ULL", "EST_FGRA", "EST_ML", "MODEL_BINOMIAL", "MODEL_FRAC", "MODEL_POISSON", "LashError",
           "lib", "lib_path"]
