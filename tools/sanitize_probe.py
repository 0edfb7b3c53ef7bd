"""Small end-to-end pass over every kernel, for compute-sanitizer (memcheck / racecheck) runs:
    compute-sanitizer --tool memcheck python tools/sanitize_probe.py
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, ops  # noqa: E402
from tools import synth  # noqa: E402

rng = np.random.default_rng(0)
with ops.Context(0) as ctx:
    genomes = synth.genomes(5, 40_000, seed=1) + [synth.dirty_genome(30_000, 21, seed=2)]
    for algo, p, k in ((ALGO_ULL, 10, 16), (ALGO_ULL, 14, 21), (ALGO_HLL, 12, 21), (ALGO_HMH, 14, 16), (ALGO_ULL, 16, 31)):
        regs = ops.sketch_genomes(ctx, algo, p, k, 42, genomes)
        for est in ((EST_FGRA, EST_ML) if algo == ALGO_ULL else (0,)):
            d, _ = ops.dist(ctx, algo, p, k, est, 1, False, regs, regs)
            t, _ = ops.dist(ctx, algo, p, k, est, 0, True, regs[:3], regs, triangular=False)
            tri, _ = ops.dist(ctx, algo, p, k, est, 1, False, regs, regs, triangular=True)
            ops.dist_stream(ctx, algo, p, k, est, 1, False, regs, regs, True, 2, lambda r0, b: None)
        ops.cardinality(ctx, algo, p, 0, regs)
        ops.merge(ctx, algo, p, regs, regs[::-1].copy())
    # wider problems: several tiles per kernel, ragged edges
    m = 1 << 10
    ull = (4 * (rng.integers(3, 12, size=(150, m)) + 9) + rng.integers(0, 4, size=(150, m))).astype(np.uint8)
    for est in (EST_FGRA, EST_ML):
        ops.dist(ctx, ALGO_ULL, 10, 16, est, 1, False, ull[:70], ull)
    hll = rng.integers(0, 30, size=(100, 1 << 12)).astype(np.uint8)
    ops.dist(ctx, ALGO_HLL, 12, 21, 0, 1, False, hll[:45], hll)
print("sanitize probe done")
