"""One all-vs-all launch on simulated sketches, for ncu captures and A/B timings of the dist kernels.
    python tools/dist_probe.py ull-ml|ull-fgra|hll|hmh|hmh-small [n] [p]
hmh-small: sketches of ~10^5 k-mers, so every pair takes expectedCollision's double-loop branch (LASH_HMH_EC=loop: per pair).
"""
import json
import sys

import numpy as np

sys.path.insert(0, ".")
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, ops  # noqa: E402
from tools import bench_configs as B  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
kind = args[0] if args else "ull-ml"
n = int(args[1]) if len(args) > 1 else 2000
p = int(args[2]) if len(args) > 2 else (14 if kind == "hll" else 10)
rng = np.random.default_rng(7)
m = 1 << p
lvl = np.clip(np.floor(12.0 - np.log2(-np.log(rng.random((n, m))))), 0, 30).astype(np.int64)   # max nlz of ~4000 hashes (clipped: no register outside the pair tables)
if kind.startswith("ull"):
    regs = (4 * (lvl + p - 1) + rng.integers(0, 4, size=(n, m))).astype(np.uint8)
    algo, est = ALGO_ULL, (EST_ML if kind == "ull-ml" else EST_FGRA)
elif kind == "hll":
    regs = (lvl + 1).astype(np.uint8)
    algo, est = ALGO_HLL, 0
else:
    m = 16384
    per_reg = 2.6 if kind == "hmh-small" else 7.0
    lvl = np.clip(np.floor(per_reg - np.log2(-np.log(rng.random((n, m))))), 0, 40).astype(np.int64)
    regs = ((lvl << 10) | rng.integers(0, 1024, size=(n, m))).astype(np.uint16)
    regs[1::2] = np.where(rng.random((n // 2 + n % 2 if False else regs[1::2].shape[0], m)) < 0.3, regs[0::2][: regs[1::2].shape[0]], regs[1::2])
    algo, est, p = ALGO_HMH, 0, 14
B.PROFILE = "--profile" in sys.argv
with ops.Context(0) as ctx:
    print(json.dumps(B.dist_case(ctx, f"{kind} p={p} n={n}", algo, p, 16, est, regs)), flush=True)
