"""Small end-to-end pass over every kernel, for compute-sanitizer runs:
    compute-sanitizer --tool memcheck  python tools/sanitize_probe.py            (every kernel, every output layout)
    compute-sanitizer --tool racecheck python tools/sanitize_probe.py --small    (reduced: racecheck is ~1000x slower)
racecheck looks at the shared-memory hazards the design rests on: the sketch kernel's shared `red.*` cells + deferred
atomics + flush, the staged tiles / pair tables / tile counters of the dist kernels, the bit-stream staging of the text
pack kernel."""
import ctypes as C
import sys

import numpy as np

sys.path.insert(0, ".")
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, capi, ops  # noqa: E402
from tools import synth  # noqa: E402

small = "--small" in sys.argv
rng = np.random.default_rng(0)
with ops.Context(0) as ctx:
    if small:
        genomes = synth.genomes(2, 200_000, seed=1) + [synth.dirty_genome(20_000, 16, seed=2)]
        cases = ((ALGO_ULL, 10, 16), (ALGO_HLL, 12, 21), (ALGO_HMH, 14, 16))
    else:
        genomes = synth.genomes(5, 40_000, seed=1) + [synth.dirty_genome(30_000, 21, seed=2)]
        cases = ((ALGO_ULL, 10, 16), (ALGO_ULL, 14, 21), (ALGO_HLL, 12, 21), (ALGO_HMH, 14, 16), (ALGO_ULL, 16, 31))
    for algo, p, k in cases:
        regs = ops.sketch_genomes(ctx, algo, p, k, 42, genomes)
        regs_t = ops.sketch_genomes_text(ctx, algo, p, k, 42, genomes)          # device-side filter + pack
        assert np.array_equal(regs, regs_t)
        for est in ((EST_FGRA, EST_ML) if algo == ALGO_ULL else (0,)):
            d, _ = ops.dist(ctx, algo, p, k, est, 1, False, regs, regs)
            if not small:
                t, _ = ops.dist(ctx, algo, p, k, est, 0, True, regs[:3], regs, triangular=False)
                tri, _ = ops.dist(ctx, algo, p, k, est, 1, False, regs, regs, triangular=True)
                ops.dist_stream(ctx, algo, p, k, est, 1, False, regs, regs, True, 2, lambda r0, b: None)
        ops.cardinality(ctx, algo, p, 0, regs)
        ops.merge(ctx, algo, p, regs, regs[::-1].copy())
    # wider problems: several tiles per kernel, ragged edges, the checksum kernel
    capi.check(capi.lib().lash_dist_set_checksum(ctx.handle, 1))
    m = 1 << 10
    n = 64 if small else 150
    ull = (4 * (rng.integers(3, 12, size=(n, m)) + 9) + rng.integers(0, 4, size=(n, m))).astype(np.uint8)
    for est in (EST_FGRA, EST_ML):
        ops.dist(ctx, ALGO_ULL, 10, 16, est, 1, False, ull[: n // 2 + 3], ull)
    ull[3, 7] = 0                                             # an empty register: the tiles of sketch 3 use the ML table form,
    ull[5, 9] = 4 * 10 + 4 + 4 * 29                           # ... sketch 5 is flagged on G-sum tiles (28 levels above the rest)
    ops.dist(ctx, ALGO_ULL, 10, 16, EST_ML, 1, False, ull[: n // 2 + 3], ull)
    hll = rng.integers(0, 30, size=(n, 1 << 12)).astype(np.uint8)
    ops.dist(ctx, ALGO_HLL, 12, 21, 0, 1, False, hll[: n // 2 - 3], hll)
    hll_w = rng.integers(5, 20, size=(n, 1 << 12)).astype(np.uint8)   # K4i: windows above 0, a flagged sketch (hll_pair_exact)
    hll_w[2, 11] = 50
    ops.dist(ctx, ALGO_HLL, 12, 21, 0, 1, False, hll_w[: n // 2 - 3], hll_w)
    if not small:                                             # K4i's 64 x 64 tile shape needs a grid of >= 8 tiles per SM
        hll_7 = rng.integers(3, 25, size=(2300, 1 << 7)).astype(np.uint8)
        hll_7[100, 5] = 58
        ops.dist(ctx, ALGO_HLL, 7, 21, 0, 1, False, hll_7, hll_7)
    # HMH sketches of few k-mers: slots, term vectors and the expected-collision tile product (two tiles when not --small)
    ns = 20 if small else 140
    dense3 = ((rng.integers(1, 12, size=(3, 16384)) << 10) | rng.integers(0, 1024, size=(3, 16384))).astype(np.uint16)
    sparse = np.where(rng.random((ns, 16384)) < 0.05, ((rng.integers(1, 6, size=(ns, 16384)) << 10) | rng.integers(0, 1024, size=(ns, 16384))), 0).astype(np.uint16)
    sparse[1::2] = np.where(rng.random((ns // 2, 16384)) < 0.5, sparse[0::2], sparse[1::2])
    ops.dist(ctx, ALGO_HMH, 14, 16, 0, 1, False, np.concatenate([sparse, dense3]), sparse, triangular=False)
    ops.dist(ctx, ALGO_HMH, 14, 16, 0, 1, False, sparse, sparse, triangular=True)
    hmh = ((rng.integers(0, 12, size=(n, 16384)) << 10) | rng.integers(0, 1024, size=(n, 16384))).astype(np.uint16)
    hmh[0, ::3] = 0
    ops.dist(ctx, ALGO_HMH, 14, 16, 0, 1, False, hmh[: n // 2 + 1], hmh)
    s, c = C.c_uint64(), C.c_uint64()
    capi.check(capi.lib().lash_dist_checksum(ctx.handle, C.byref(s), C.byref(c)))
    assert c.value == (n // 2 + 1) * n
print("sanitize probe done")
