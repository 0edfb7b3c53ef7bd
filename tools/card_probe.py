"""Kernel time of the per-sketch cardinalities (lash_cardinality_dev) on device-resident sketches:
    python tools/card_probe.py            (HLL p=14 and HMH, 10 000 sketches; LASH_HLL_KERNEL=float times the sequential HLL kernel)"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, ops  # noqa: E402
from lash_b200.capi import check, lib  # noqa: E402

rng = np.random.default_rng(0)
dev = torch.device("cuda", 0)
with ops.Context(0) as ctx:
    for name, algo, p, est, regs in (
            ("hll p=14", ALGO_HLL, 14, 0, np.clip(np.floor(8.0 - np.log2(-np.log(rng.random((10000, 16384))))) + 1, 1, 51).astype(np.uint8)),
            ("hmh", ALGO_HMH, 14, 0, (((np.clip(np.floor(8.0 - np.log2(-np.log(rng.random((10000, 16384))))), 0, 40).astype(np.uint16) + 1) << 10) | 5).astype(np.uint16)),
            ("ull p=10 fgra", ALGO_ULL, 10, 0, (4 * (rng.integers(3, 12, size=(10000, 1024)) + 9) + rng.integers(0, 4, size=(10000, 1024))).astype(np.uint8))):
        d = torch.from_numpy(regs.view(np.uint8).reshape(-1)).to(dev)
        card = torch.empty(regs.shape[0], dtype=torch.float64, device=dev)
        st = torch.cuda.Stream(dev)
        ts = []
        with torch.cuda.stream(st):
            for r in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(st)
                check(lib().lash_cardinality_dev(ctx.handle, algo, p, est, C.c_void_p(d.data_ptr()), regs.shape[0], C.c_void_p(card.data_ptr()), C.c_void_p(st.cuda_stream)))
                e1.record(st)
                e1.synchronize()
                ts.append(e0.elapsed_time(e1))
        print(f"{name}: {regs.shape[0]} sketches, cardinality kernel {min(ts):.3f} ms", flush=True)
