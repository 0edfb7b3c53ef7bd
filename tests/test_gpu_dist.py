"""GPU all-vs-all distance vs the CPU oracle, through the C ABI.

Mirrors what a test of hmh_distance / ull_distance / hll_distance (src/utils.rs:84-373) +
compute_distance (src/main.rs:415-423) would check.

Tolerance (north_star: 1e-12 relative in f64, 1e-6 with --fp32), written out:
  * per-sketch cardinalities: <= 8 ulp (only pow/log differ between CUDA and glibc; every sum runs
    in register order on both sides and is bit-identical);
  * distances (the per-pair epilogues use correctly rounded pow / log / log1p, csrc/ddmath.cuh, so the union estimate and the
    cardinalities equal glibc's except where glibc itself is an ulp off: measured max relative error 9e-15 (FGRA),
    1.7e-13 (ML), 3e-16 (HLL, HMH) on the BASELINE shapes, 99.7-99.95 % of the cells bit-identical);  the asserted bound is
    |d_gpu - d_cpu| <= 1e-12*|d_cpu| + C(s), where C(s) = 64*eps/(s*k) is the
    first-order effect on d of a 4-ulp change of the union estimate U through
    s = (a+b-U)/U  (d = -ln(2s/(1+s))/k  =>  |dd/dU * U| ~ 1/(s*k)): the Jaccard subtraction
    cancels catastrophically for unrelated genomes, so a pure relative 1e-12 on d is not a
    property of the formula itself (any two libm's differ there).  C(s) is < 1e-12 whenever
    s > 1e-3/k; the test also asserts that >= 99.9 % of the cells meet the plain 1e-12 bound and prints the achieved figures.
"""
import numpy as np
import pytest

from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, MODEL_BINOMIAL, MODEL_POISSON, LashError
from lash_b200 import ops
from lash_b200.capi import W_HLL_BIAS_REGIME
from tools import synth

pytestmark = pytest.mark.gpu
EPS = np.finfo(np.float64).eps


def _sketches(oracle, algo, p, k, n, length, seed=42):
    """CPU-oracle sketches as inputs, so this file tests the dist path in isolation."""
    return oracle.sketch_genomes(algo, p, k, seed, synth.genomes(n, length, seed=seed), threads=8)


def _assert_close_f64(got, exp, frac, k, what, strict_frac=0.999):
    """The stated bound, and the achieved figures on stdout (pytest -s / -rP shows them): worst relative error, fraction of
    cells within plain 1e-12 relative, fraction bit-identical."""
    assert got.shape == exp.shape
    both_nan = np.isnan(got) & np.isnan(exp)
    s = frac / (2.0 - frac)
    with np.errstate(divide="ignore", invalid="ignore"):
        cond = 64 * EPS / (np.maximum(s, 1e-300) * k)
    tol = 1e-12 * np.abs(exp) + np.minimum(cond, 1.0)
    err = np.abs(got - exp)
    bad = ~((err <= tol) | both_nan)
    assert not bad.any(), f"{what}: {bad.sum()} cells off, worst {np.nanmax(err[bad])} at {np.argwhere(bad)[0]}"
    strict = (err <= 1e-12 * np.abs(exp)) | both_nan
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.where(err == 0, 0.0, err / np.abs(exp))
    print(f"[f64 parity] {what}: cells {got.size}, max_rel_err {np.nanmax(np.where(both_nan, 0.0, rel)):.3e}, "
          f"within 1e-12: {strict.mean():.6f}, bit-identical: {((got == exp) | both_nan).mean():.6f}")
    assert strict.mean() > strict_frac, f"{what}: only {strict.mean():.4f} of cells within plain 1e-12 relative"


CASES = [
    # algo, p, k, estimator, n, genome length
    (ALGO_ULL, 10, 16, EST_FGRA, 40, 200_000),
    (ALGO_ULL, 10, 16, EST_ML, 40, 200_000),
    (ALGO_ULL, 14, 21, EST_FGRA, 12, 300_000),
    (ALGO_ULL, 14, 21, EST_ML, 12, 300_000),
    (ALGO_HLL, 14, 21, 0, 12, 2_000_000),
    (ALGO_HLL, 10, 16, 0, 40, 200_000),
    (ALGO_HMH, 14, 16, 0, 10, 1_500_000),
]


@pytest.mark.parametrize("algo,p,k,est,n,length", CASES)
@pytest.mark.parametrize("model", [MODEL_POISSON, MODEL_BINOMIAL])
def test_dist_matches_oracle_f64(oracle, gpu_ctx, algo, p, k, est, n, length, model):
    regs = _sketches(oracle, algo, p, k, n, length)
    ref, qry = regs[: n // 2 + 3], regs[n // 3:]
    exp = oracle.dist(algo, p, k, est, model, False, ref, qry, threads=8)
    frac = oracle.dist(algo, p, k, est, 2, False, ref, qry, threads=8)
    got, w = ops.dist(gpu_ctx, algo, p, k, est, model, False, ref, qry)
    assert w == 0
    _assert_close_f64(got, exp, frac, k, f"algo={algo} p={p} est={est} model={model}")
    assert np.isfinite(got).all() and (got >= 0).all() and (got <= 1).all()
    assert (got < 0.5).any(), "related genomes should be close"


@pytest.mark.parametrize("algo,p,k,est,n,length", CASES[:2] + CASES[4:5] + CASES[6:])
def test_dist_fp32(oracle, gpu_ctx, algo, p, k, est, n, length):
    """--fp32: frac is cast to f32 and ln/powf run in f32 (utils.rs:176,277,364; main.rs:415-423)."""
    regs = _sketches(oracle, algo, p, k, n, length)
    for model in (MODEL_POISSON, MODEL_BINOMIAL):
        exp = oracle.dist(algo, p, k, est, model, True, regs, regs, threads=8)
        got, _ = ops.dist(gpu_ctx, algo, p, k, est, model, True, regs, regs)
        assert got.dtype == np.float32
        np.testing.assert_allclose(got, exp, rtol=1e-6, atol=2e-7)


@pytest.mark.parametrize("algo,p,k,est", [(ALGO_ULL, 10, 16, EST_FGRA), (ALGO_ULL, 10, 16, EST_ML), (ALGO_HLL, 12, 21, 0),
                                          (ALGO_HMH, 14, 16, 0)])
def test_cardinalities(oracle, gpu_ctx, algo, p, k, est):
    regs = _sketches(oracle, algo, p, k, 9, 700_000)
    got = ops.cardinality(gpu_ctx, algo, p, est, regs)
    exp = np.array([oracle.cardinality(algo, p, est, r) for r in regs])
    assert np.all(np.abs(got - exp) <= 8 * EPS * np.abs(exp)), (got, exp)
    if algo == ALGO_HLL:  # raw HLL estimate is +,*,/ only: bit-exact
        assert np.array_equal(got, exp)


@pytest.mark.parametrize("p", [4, 7, 12, 14, 18])
def test_hll_cardinalities_warp_parallel_equal_the_sequential_sum(oracle, gpu_ctx, p):
    """card_hll_int_kernel: 32 lanes sum 2^-r as integers when the sketch's registers span at most 29 levels (exact, so equal
    to the reference's sequential f64 loop bit for bit), one lane runs that loop otherwise.  Raw estimates have no libm call:
    bit-exact; linear counting adds one log (<= 4 ulp); the bias regime is flagged as NaN on both sides."""
    rng = np.random.default_rng(p)
    m = 1 << p
    rows = []
    for lg in (-4.0, -1.0, 1.0, 5.0, 12.0, 25.0):                       # nearly empty ... huge
        if lg < 0:
            hit = rng.random(m) < 2.0 ** lg
            rows.append((hit * np.clip(rng.geometric(0.5, size=m), 1, 64 - p + 1)).astype(np.uint8))
        else:
            rows.append(np.clip(np.floor(lg - np.log2(-np.log(rng.random(m)))) + 1, 1, 64 - p + 1).astype(np.uint8))
    wide = rows[3].copy(); wide[1] = 64 - p + 1                          # spans more than 29 levels: sequential path
    edge = np.full(m, 7, dtype=np.uint8); edge[2] = 7 + 28               # exactly the last level inside the window
    over = np.full(m, 7, dtype=np.uint8); over[2] = 7 + 29               # first level outside
    rows += [wide, edge, over, np.zeros(m, dtype=np.uint8), np.full(m, 64 - p + 1, dtype=np.uint8)]
    regs = np.stack(rows)
    got = ops.cardinality(gpu_ctx, ALGO_HLL, p, 0, regs)
    exp = np.array([oracle.cardinality(ALGO_HLL, p, 0, r) for r in regs])
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    assert np.all(np.abs(got[ok] - exp[ok]) <= 4 * EPS * np.abs(exp[ok])), (got, exp)
    raw = ok & (regs.min(axis=1) > 0)                                    # no empty register -> no linear counting, no log
    assert raw.sum() >= 4
    np.testing.assert_array_equal(got[raw], exp[raw])


def test_hmh_cardinalities_warp_parallel_and_sequential_paths(oracle, gpu_ctx):
    """card_hmh_kernel sums 2^-lz warp-parallel as integers when a sketch's leading-zero counts span at most 29 levels (exact:
    equal to the sequential loop), sequentially otherwise: dense, sparse (empty registers), wide-span, edge-of-window, empty
    and saturated sketches against the oracle (the estimate adds beta(ez): a few ulp of libm)."""
    rng = np.random.default_rng(21)
    m = 16384
    def sk(per_reg, sparse=1.0):
        lvl = np.clip(np.floor(per_reg - np.log2(-np.log(rng.random(m)))), 0, 45).astype(np.int64)
        hit = rng.random(m) < sparse
        return np.where(hit, ((lvl + 1) << 10) | rng.integers(0, 1024, size=m), 0).astype(np.uint16)
    rows = [sk(7.0), sk(2.0), sk(0.5, 0.3), sk(0.0, 0.01), sk(20.0)]
    wide = sk(7.0); wide[5] = (51 << 10) | 3; wide[6] = (1 << 10) | 9          # spans 50 levels: sequential path
    edge = np.full(m, (9 << 10) | 1, dtype=np.uint16); edge[3] = ((9 + 28) << 10) | 5   # last level inside the window
    over = np.full(m, (9 << 10) | 1, dtype=np.uint16); over[3] = ((9 + 29) << 10) | 5   # first level outside
    rows += [wide, edge, over, np.zeros(m, dtype=np.uint16), np.full(m, (50 << 10) | 1023, dtype=np.uint16)]
    regs = np.stack(rows)
    got = ops.cardinality(gpu_ctx, ALGO_HMH, 14, 0, regs)
    exp = np.array([oracle.cardinality(ALGO_HMH, 14, 0, r) for r in regs])
    fin = np.isfinite(exp)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.all(np.abs(got[fin] - exp[fin]) <= 8 * EPS * np.abs(exp[fin])), (got, exp)


def test_triangular_packed_equals_dense_lower(oracle, gpu_ctx):
    """same_files rule (utils.rs:158-160,256-258,350-352): only j <= i, diagonal included."""
    regs = _sketches(oracle, ALGO_ULL, 10, 16, 45, 100_000)
    dense, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_FGRA, MODEL_POISSON, False, regs, regs)
    tri, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_FGRA, MODEL_POISSON, False, regs, regs, triangular=True)
    n = len(regs)
    assert tri.shape == (n * (n + 1) // 2,)
    for i in range(n):
        assert np.array_equal(tri[i * (i + 1) // 2: i * (i + 1) // 2 + i + 1], dense[i, : i + 1])
    # symmetry of the union: d(i,j) == d(j,i) bit for bit for ULL/HLL (same merged registers, same order)
    assert np.array_equal(dense, dense.T)
    # a sketch against itself: union == itself, s = 1, frac = 1, d = 0 exactly (no name rule needed)
    assert np.all(np.diag(dense) == 0.0)


def test_stream_blocks_equal_dense(oracle, gpu_ctx):
    regs = _sketches(oracle, ALGO_HLL, 10, 21, 70, 60_000)
    dense, _ = ops.dist(gpu_ctx, ALGO_HLL, 10, 21, 0, MODEL_POISSON, False, regs[:50], regs)
    got = np.full_like(dense, np.nan)
    seen = []

    def on_block(row0, block):
        got[row0: row0 + block.shape[0]] = block
        seen.append((row0, block.shape[0]))

    ops.dist_stream(gpu_ctx, ALGO_HLL, 10, 21, 0, MODEL_POISSON, False, regs[:50], regs, False, 16, on_block)
    assert seen == [(0, 16), (16, 16), (32, 16), (48, 2)]
    assert np.array_equal(got, dense)


def _checksum_of(cells: np.ndarray) -> tuple[int, int]:
    bits = cells.view(np.uint64 if cells.dtype == np.float64 else np.uint32).astype(np.uint64)
    return int(bits.sum(dtype=np.uint64)), int(cells.size)


@pytest.mark.parametrize("fp32", [False, True])
def test_checksum_of_checksums_is_tiling_independent(oracle, gpu_ctx, fp32):
    """lash_dist_set_checksum: the device-side wrapping sum of the output bit patterns equals the host sum of the same
    cells, for one call, for row blocks, for row ranges (the multi-GPU tiling) and for the packed triangle."""
    import ctypes as C

    from lash_b200 import capi
    L = capi.lib()
    regs = _sketches(oracle, ALGO_ULL, 10, 16, 75, 60_000)
    n = len(regs)

    def read():
        s, c = C.c_uint64(), C.c_uint64()
        capi.check(L.lash_dist_checksum(gpu_ctx.handle, C.byref(s), C.byref(c)))
        return s.value, c.value

    capi.check(L.lash_dist_set_checksum(gpu_ctx.handle, 1))
    try:
        dense, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, fp32, regs, regs)
        assert read() == _checksum_of(dense)
        tri, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, fp32, regs, regs, triangular=True)
        want_tri = _checksum_of(tri)
        assert read() == want_tri and want_tri[1] == n * (n + 1) // 2
        ops.dist_stream(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, fp32, regs, regs, True, 16, lambda r0, b: None)
        assert read() == want_tri
        # row ranges, as two ranks would compute them: the sums add up (mod 2^64)
        rp = regs.ctypes.data_as(C.c_void_p)
        cb = capi.DIST_BLOCK_CB(lambda u, r0, nr, ptr: 0)
        total = [0, 0]
        for b, e in ((0, 31), (31, n)):
            capi.check(L.lash_dist_stream_rows(gpu_ctx.handle, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, int(fp32), rp, n, rp, n, 1, b, e, 7, cb, None))
            s, c = read()
            total = [(total[0] + s) % 2**64, total[1] + c]
        assert tuple(total) == want_tri
    finally:
        capi.check(L.lash_dist_set_checksum(gpu_ctx.handle, 0))
    ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, fp32, regs[:5], regs[:5])
    assert read() == (0, 0)   # switched off again


def test_small_sketches_hit_small_range_paths(oracle, gpu_ctx):
    """Tiny inputs: most registers empty -> ULL FGRA small-range correction (c0..c10, sigma),
    ULL ML with b[0], b[1] contributions, HLL linear counting, HMH with few collisions."""
    rng = np.random.default_rng(1)
    gs = [[synth.to_ascii(rng.integers(0, 4, size=int(n), dtype=np.uint8))] for n in (16, 40, 100, 300, 1000, 3000, 10_000)]
    gs.append([b""])  # completely empty sketch
    for algo, p, est in ((ALGO_ULL, 10, EST_FGRA), (ALGO_ULL, 10, EST_ML), (ALGO_ULL, 4, EST_FGRA), (ALGO_ULL, 4, EST_ML),
                         (ALGO_HLL, 10, 0)):
        regs = oracle.sketch_genomes(algo, p, 16, 42, gs)
        exp, flags = oracle.dist(algo, p, 16, est, MODEL_POISSON, False, regs, regs, return_flags=True)
        frac = oracle.dist(algo, p, 16, est, 2, False, regs, regs)
        got, w = ops.dist(gpu_ctx, algo, p, 16, est, MODEL_POISSON, False, regs, regs)
        assert (w == W_HLL_BIAS_REGIME) == bool(flags.any())
        ok = flags == 0
        both_nan = np.isnan(got) & np.isnan(exp)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
        _assert_close_f64(np.where(both_nan, 0, got)[ok], np.where(both_nan, 0, exp)[ok], np.nan_to_num(frac)[ok], 16,
                          f"small algo={algo} p={p} est={est}")
        card_g = ops.cardinality(gpu_ctx, algo, p, est, regs)
        card_c = np.array([oracle.cardinality(algo, p, est, r) for r in regs])
        fin = np.isfinite(card_c)
        assert np.array_equal(np.isfinite(card_g), fin)
        assert np.all(np.abs(card_g[fin] - card_c[fin]) <= 8 * EPS * np.abs(card_c[fin]))


def test_hmh_small_sketches_precomputed_expected_collisions(oracle, gpu_ctx):
    """hyperminhash's expectedCollision runs a 64 x 1024 loop of pows when both cardinalities are <= 2^19 (utils.rs:150-180 ->
    similarity()).  K4m's small-sketch path stores a term vector per small sketch and sums the products in the loop's order
    (setup_hmh_ec, hmh_ec_fill_kernel, hmh_ec_gemm_kernel): related small genomes (C > 0, so the value matters), large ones
    mixed in (closed form), an empty sketch, rectangular / triangular / row-ranged calls."""
    import ctypes as C

    from lash_b200 import capi
    lengths = [3000, 20_000, 20_000, 90_000, 90_000, 250_000, 250_000, 400_000, 1_500_000, 1_500_000, 60_000, 60_000, 8000]
    gs = []
    for g, n in enumerate(lengths):
        base = synth.ancestor_codes(n, seed=100 + n)          # equal lengths share an ancestor -> collisions
        gs.append([synth.to_ascii(synth.mutate(base, 0.01 * (1 + g % 3), seed=g))])
    gs.append([b""])
    regs = oracle.sketch_genomes(ALGO_HMH, 14, 16, 42, gs, threads=8)
    n = len(regs)
    cards = np.array([oracle.cardinality(ALGO_HMH, 14, 0, r) for r in regs])
    assert (cards[:8] <= 524288).all() and (cards[8:10] > 524288).all()
    exp = oracle.dist(ALGO_HMH, 14, 16, 0, MODEL_POISSON, False, regs, regs, threads=8)
    frac = oracle.dist(ALGO_HMH, 14, 16, 0, 2, False, regs, regs, threads=8)
    assert ((frac > 0) & (frac < 1)).sum() >= 8, "the fixture must hold small pairs with collisions beyond the expected ones"
    got, w = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, MODEL_POISSON, False, regs, regs)
    assert w == 0
    _assert_close_f64(got, exp, frac, 16, "HMH small sketches, dense")
    # rectangular, reference and query sets with different small subsets
    ref, qry = regs[2:11], regs[[0, 9, 4, 13, 7, 12]]
    got_r, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, MODEL_POISSON, False, ref, qry)
    np.testing.assert_array_equal(got_r, got[2:11][:, [0, 9, 4, 13, 7, 12]])
    # packed triangle and row ranges (slots relative to the row range)
    tri, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, MODEL_BINOMIAL, False, regs, regs, triangular=True)
    full_b, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, MODEL_BINOMIAL, False, regs, regs)
    np.testing.assert_array_equal(tri, full_b[np.tril_indices(n)])
    L = capi.lib()
    rows = {}

    def _cb(user, row0, nrows, ptr):
        block = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), (nrows * n,)).reshape(nrows, n)
        for r in range(nrows):
            rows[row0 + r] = block[r, : row0 + r + 1].copy()    # triangular: cells j > i are unspecified
        return 0

    cb = capi.DIST_BLOCK_CB(_cb)
    rp = regs.ctypes.data_as(C.c_void_p)
    for b, e in ((0, 5), (5, n)):
        capi.check(L.lash_dist_stream_rows(gpu_ctx.handle, ALGO_HMH, 14, 16, 0, MODEL_BINOMIAL, 0, rp, n, rp, n, 1, b, e, 3, cb, None))
    for i in range(n):
        np.testing.assert_array_equal(rows[i], full_b[i, : i + 1])


def test_hmh_small_sketch_tile_product_equals_the_per_pair_loop(gpu_ctx, tmp_path):
    """hmh_ec_gemm_kernel (128 x 128 tiles, 8 x 8 pairs per thread, exact early end of the row loop) against the per-pair
    41 x 1024 loop of the same library (LASH_HMH_EC=loop, read once per process -> a child): 300 simulated sketches of 150 to
    4 x 10^5 k-mers -- several tiles, ragged edges, cardinalities orders of magnitude apart inside one tile -- plus a few large
    ones; same arithmetic in the same order, so the distances must be equal bit for bit."""
    import subprocess
    import sys
    rng = np.random.default_rng(9)
    n, m = 300, 16384
    per_reg = np.exp(rng.uniform(np.log(0.01), np.log(25.0), size=n))          # k-mers per register
    per_reg[::37] = 400.0                                                       # large sketches in between (closed form)
    regs = np.zeros((n, m), dtype=np.uint16)
    for i in range(n):
        hit = rng.random(m) < 1.0 - np.exp(-per_reg[i])
        lvl = np.clip(np.floor(np.log2(max(per_reg[i], 1.0)) - np.log2(-np.log(rng.random(m)))), 0, 40).astype(np.int64)
        regs[i] = np.where(hit, ((lvl + 1) << 10) | rng.integers(0, 1024, size=m), 0).astype(np.uint16)
    regs[1::2] = np.where(rng.random((n // 2, m)) < 0.4, regs[0::2], regs[1::2])                # collisions between neighbours
    np.save(tmp_path / "regs.npy", regs)
    child = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from lash_b200 import ALGO_HMH, ops\n"
        "regs = np.load(%r)\n"
        "with ops.Context(0) as ctx:\n"
        "    d, _ = ops.dist(ctx, ALGO_HMH, 14, 16, 0, 2, False, regs, regs, triangular=True)\n"
        "    r, _ = ops.dist(ctx, ALGO_HMH, 14, 16, 0, 2, False, regs[100:290], regs[20:170])\n"
        "np.save(%r, d); np.save(%r, r)\n" % (str(__import__('os').path.dirname(__import__('os').path.dirname(__file__))), str(tmp_path / "regs.npy"),
                                               str(tmp_path / "loop.npy"), str(tmp_path / "loop_rect.npy")))
    subprocess.run([sys.executable, "-c", child], check=True, env=dict(__import__('os').environ, LASH_HMH_EC="loop"))
    cards = ops.cardinality(gpu_ctx, ALGO_HMH, 14, 0, regs)
    assert (cards <= 524288).sum() > 250 and (cards > 524288).sum() >= 5
    tri, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, 2, False, regs, regs, triangular=True)
    loop = np.load(tmp_path / "loop.npy")
    assert ((loop > 0) & (loop < 1)).sum() > 100
    np.testing.assert_array_equal(tri, loop)
    rect, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, 2, False, regs[100:290], regs[20:170])
    np.testing.assert_array_equal(rect, np.load(tmp_path / "loop_rect.npy"))


def test_hll_bias_regime_is_flagged_not_silently_different(oracle, gpu_ctx):
    """HLL++ estimates in (threshold, 5m] need Google's empirical bias tables (not reproducible
    offline): both sides must flag those cells instead of inventing a number."""
    regs = oracle.sketch_genomes(ALGO_HLL, 10, 16, 42, synth.genomes(3, 3000, seed=3))
    exp, flags = oracle.dist(ALGO_HLL, 10, 16, 0, MODEL_POISSON, False, regs, regs, return_flags=True)
    assert flags.any()
    got, w = ops.dist(gpu_ctx, ALGO_HLL, 10, 16, 0, MODEL_POISSON, False, regs, regs)
    assert w == W_HLL_BIAS_REGIME
    np.testing.assert_array_equal(got, exp)


def test_saturated_ull_registers_large_range_path(oracle, gpu_ctx):
    """Registers >= 252 (unreachable for genomes, part of the estimator): FGRA large-range term."""
    rng = np.random.default_rng(4)
    regs = rng.integers(200, 256, size=(6, 256), dtype=np.uint8)
    regs[0, :] = 255
    for est in (EST_FGRA, EST_ML):
        exp = oracle.dist(ALGO_ULL, 8, 16, est, MODEL_BINOMIAL, False, regs, regs)
        got, _ = ops.dist(gpu_ctx, ALGO_ULL, 8, 16, est, MODEL_BINOMIAL, False, regs, regs)
        frac = oracle.dist(ALGO_ULL, 8, 16, est, 2, False, regs, regs)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
        m = ~np.isnan(exp)
        _assert_close_f64(np.where(m, got, 0), np.where(m, exp, 0), np.nan_to_num(frac), 16, f"saturated registers est={est}", strict_frac=0.9)


@pytest.mark.parametrize("p", [10, 14])
def test_huge_cardinality_ull_sketches_stay_on_the_table_path(oracle, gpu_ctx, p):
    """Sketches of very large inputs (config 4: 10^11 k-mers -> ~2^22 hashes per register at p=14) have no
    register near 4p-4; the FGRA pair table is anchored at the smallest register present, so they are
    neither wrong nor pushed to the per-pair exact path.  A mixed set (one small sketch among them) and
    an out-of-window register exercise the sentinel -> exact fallback."""
    rng = np.random.default_rng(p)
    m = 1 << p
    n = 12
    def simulated(log2_per_reg, rows):
        # max nlz of 2^log2 hashes: Gumbel-like; u = p - 1 + nlz, two random sub-bits
        nlz = np.clip(np.floor(log2_per_reg - np.log2(-np.log(rng.random((rows, m))))), 0, 60 - p).astype(np.int64)
        return (4 * (nlz + p - 1) + rng.integers(0, 4, size=(rows, m))).astype(np.uint8)
    big = simulated(22.0, n)
    big[1] = big[0]
    big[1, ::7] = np.maximum(big[1, ::7], big[2, ::7])          # a near-duplicate pair (large Jaccard)
    for regs in (big, np.concatenate([big[:6], simulated(6.0, 3), simulated(14.0, 3)])):
        regs = regs.copy()
        regs[5, 3] = 251                                         # far above any window: exact per-pair path
        exp = oracle.dist(ALGO_ULL, p, 21, EST_FGRA, MODEL_POISSON, False, regs, regs)
        frac = oracle.dist(ALGO_ULL, p, 21, EST_FGRA, 2, False, regs, regs)
        got, _ = ops.dist(gpu_ctx, ALGO_ULL, p, 21, EST_FGRA, MODEL_POISSON, False, regs, regs)
        _assert_close_f64(got, exp, frac, 21, f"huge-cardinality p={p}")
        tri, _ = ops.dist(gpu_ctx, ALGO_ULL, p, 21, EST_FGRA, MODEL_POISSON, False, regs, regs, triangular=True)
        il = np.tril_indices(len(regs))
        np.testing.assert_array_equal(tri, got[il])


def test_ull_merge_in_packed_domain_is_exhaustively_correct(oracle, gpu_ctx):
    """ULL union is pack(unpack(a)|unpack(b)), not max(a,b).  Every ordered pair of valid register
    bytes for p=8 goes through the kernel's SIMD merge; the union's ML statistics are integers, so
    the resulting estimates must match the oracle to the last few ulps on every row."""
    p = 8
    valid = np.array([0, 4 * p - 4, 4 * p, 4 * p + 2] + list(range(4 * p + 4, 256)), dtype=np.uint8)
    nv = len(valid)
    m = 1 << p
    assert nv <= m
    a = np.zeros((nv, m), dtype=np.uint8)
    b = np.zeros((nv, m), dtype=np.uint8)
    for i, v in enumerate(valid):
        a[i, :nv] = v           # row i: constant register v ...
        b[i, :nv] = np.roll(valid, i)  # ... against every valid value, rotated
    for est in (EST_ML, EST_FGRA):
        exp = oracle.dist(ALGO_ULL, p, 16, est, MODEL_BINOMIAL, False, a, b)
        got, _ = ops.dist(gpu_ctx, ALGO_ULL, p, 16, est, MODEL_BINOMIAL, False, a, b)
        frac = oracle.dist(ALGO_ULL, p, 16, est, 2, False, a, b)
        np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
        mm = ~np.isnan(exp)
        _assert_close_f64(np.where(mm, got, 0), np.where(mm, exp, 0), np.nan_to_num(frac), 16, f"exhaustive merge est={est}", strict_frac=0.9)


def test_dist_error_behaviour(oracle, gpu_ctx):
    regs = _sketches(oracle, ALGO_ULL, 10, 16, 4, 20_000)
    with pytest.raises(LashError):   # main.rs:421 "model needs to be 0 or 1"
        ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_FGRA, 3, False, regs, regs)
    with pytest.raises(LashError):   # utils.rs:217 "estimator needs to be either fgra or ml"
        ops.dist(gpu_ctx, ALGO_ULL, 10, 16, 5, MODEL_POISSON, False, regs, regs)
    with pytest.raises(LashError):
        ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_FGRA, MODEL_POISSON, False, regs, regs[:2], triangular=True)
    with pytest.raises(ValueError):
        ops.ull_distance(gpu_ctx, 10, 16, 1, False, ["a"], regs[:1], ["a"], regs[:1], "mle", False, True, lambda r: None)


def test_reference_style_emit_interface(oracle, gpu_ctx):
    """ull_distance(..., emit) as main.rs:557-566 drives it: rows per reference, lower triangle when
    same_files, d = 0 whenever the two names are equal (main.rs:452-453), header row with --dm."""
    names = [f"g{i}.fa" for i in range(6)]
    regs = _sketches(oracle, ALGO_ULL, 10, 16, 6, 80_000)
    rows = []
    ops.ull_distance(gpu_ctx, 10, 16, MODEL_POISSON, False, names, regs, names, regs, "fgra", True, True, rows.append)
    assert rows[0] == [("", q, 1.0) for q in names]
    body = rows[1:]
    assert [len(r) for r in body] == [1, 2, 3, 4, 5, 6]
    exp = oracle.dist(ALGO_ULL, 10, 16, EST_FGRA, MODEL_POISSON, False, regs, regs)
    for i, r in enumerate(body):
        for j, (rn, qn, d) in enumerate(r):
            assert (rn, qn) == (names[i], names[j])
            assert d == (0.0 if i == j else pytest.approx(exp[i, j], rel=1e-11))
    # different file sets, duplicate name across them -> forced zero
    rows = []
    ops.hll_distance(gpu_ctx, 10, 16, MODEL_POISSON, False, ["x", "y"], oracle.sketch_genomes(ALGO_HLL, 10, 16, 42, synth.genomes(2, 50_000)),
                     ["y", "z"], oracle.sketch_genomes(ALGO_HLL, 10, 16, 42, synth.genomes(2, 50_000, seed=1)), False, False, rows.append)
    assert len(rows) == 2 and rows[1][0][:2] == ("y", "y") and rows[1][0][2] == 0.0 and rows[0][0][2] > 0.0


@pytest.mark.parametrize("p,n_ref,n_qry", [(4, 5, 9), (7, 37, 101), (12, 70, 33), (14, 40, 67)])
def test_hll_recoded_kernel_ragged_tiles_and_mixed_empties(oracle, gpu_ctx, p, n_ref, n_qry):
    """K4h (registers recoded to the high word of 2^-r, zero count only in chunks where both sides hold an empty
    register): simulated sketches from nearly empty (linear counting) to huge, tiles that are not multiples of
    32 x 64, chunks with empties on one side only.  The HLL sum is +,*,/ of exact powers of two in register
    order on both sides, so everything up to frac is bit-identical; the poisson distance adds one log."""
    rng = np.random.default_rng(100 + p)
    m = 1 << p

    def simulated(log2_per_reg, rows):
        if log2_per_reg < 0:                                     # sparse: a fraction 2^log2 of the registers is hit once
            hit = rng.random((rows, m)) < 2.0 ** log2_per_reg
            return (hit * np.clip(rng.geometric(0.5, size=(rows, m)), 1, 64 - p + 1)).astype(np.uint8)
        rho = np.clip(np.floor(log2_per_reg - np.log2(-np.log(rng.random((rows, m))))) + 1, 1, 64 - p + 1)
        return rho.astype(np.uint8)

    def mixed(n):
        kinds = [-3.0, -0.5, 2.0, 9.0, 22.0]
        rows = [simulated(kinds[i % len(kinds)], 1)[0] for i in range(n)]
        rows[0][:] = 0                                           # a completely empty sketch
        if n > 3:
            rows[3][: m // 2] = 0                                # empties in the first half only
        return np.stack(rows)

    ref, qry = mixed(n_ref), mixed(n_qry)
    exp, flags = oracle.dist(ALGO_HLL, p, 21, 0, MODEL_POISSON, False, ref, qry, return_flags=True)
    frac_exp = oracle.dist(ALGO_HLL, p, 21, 0, 2, False, ref, qry)
    got, w = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, MODEL_POISSON, False, ref, qry)
    frac_got, _ = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, 2, False, ref, qry)
    assert (w == W_HLL_BIAS_REGIME) == bool(flags.any())
    np.testing.assert_array_equal(np.isnan(got), np.isnan(exp))
    ok = ~np.isnan(exp)
    # frac = 2s/(1+s) with s from exact sums and a bit-exact raw estimate; linear counting adds one log (glibc vs CUDA)
    np.testing.assert_allclose(frac_got[ok], frac_exp[ok], rtol=1e-13, atol=1e-300)
    _assert_close_f64(np.where(ok, got, 0), np.where(ok, exp, 0), np.nan_to_num(frac_exp), 21, f"hll recoded p={p}")
    # triangular / row-streamed forms of the same kernel agree with the dense one bit for bit
    sq = np.concatenate([ref, qry])[: min(n_ref + n_qry, 90)]
    dense, _ = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, MODEL_POISSON, False, sq, sq)
    tri, _ = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, MODEL_POISSON, False, sq, sq, triangular=True)
    np.testing.assert_array_equal(tri, dense[np.tril_indices(len(sq))])


@pytest.mark.parametrize("p,n", [(5, 150), (7, 150), (12, 150), (14, 150), (7, 2300)])
def test_hll_fixed_point_kernel_is_bit_identical_in_and_out_of_its_window(oracle, gpu_ctx, p, n):
    """K4i sums min(v(ra), v(rb)) in integers, exact while every register lies within 29 levels of the tile's smallest one;
    sketches that reach above the window are flagged per tile and their pairs redone by the sequential f64 loop.  Without
    empty registers the estimate is alpha*m^2/sum -- no libm call -- so `frac` must equal the oracle's BIT FOR BIT on every
    pair: in-window sketches (several tiles, ragged edges), sketches with a planted register just inside (lo + 28), just
    outside (lo + 29) and far outside the window, a sketch that lowers the window of its tiles, and all of them again as
    query columns.  n = 2300 is large enough for the launcher to pick the 64 x 64 tile shape (8 rows per thread)."""
    rng = np.random.default_rng(p)
    m = 1 << p
    regs = np.clip(np.floor(9.0 - np.log2(-np.log(rng.random((n, m))))) + 1, 6, 30).astype(np.uint8)
    lo = int(regs.min())
    assert lo == 6
    regs[:, 0] = lo                          # every sketch touches the minimum: the window of every tile is [6, 34]
    regs[3, 5] = lo + 28                     # last level inside
    regs[40, 9] = lo + 29                    # first level outside -> flagged
    regs[41, m - 1] = 64 - p + 1             # the largest register HLL can hold
    regs[100, 17] = lo + 33
    regs[120, :] = np.clip(regs[120].astype(np.int64) - 4, 2, 255).astype(np.uint8)   # lowers lo to 2 in its tiles -> neighbours flagged there
    regs[n - 1, 3] = lo + 31
    for ref, qry in ((regs, regs), (regs[30:75], regs[90:])):
        exp = oracle.dist(ALGO_HLL, p, 21, 0, 2, False, ref, qry, threads=8)
        got, w = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, 2, False, ref, qry)
        assert w == 0
        np.testing.assert_array_equal(got, exp)
    tri, _ = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, 2, False, regs, regs, triangular=True)
    full, _ = ops.dist(gpu_ctx, ALGO_HLL, p, 21, 0, 2, False, regs, regs)
    np.testing.assert_array_equal(tri, full[np.tril_indices(n)])


@pytest.mark.parametrize("n_ref,n_qry", [(5, 9), (37, 101), (70, 33)])
def test_hmh_word_kernel_ragged_tiles_and_mixed_empties(oracle, gpu_ctx, n_ref, n_qry, monkeypatch):
    """K4m (two u16 registers per 32-bit word, C = N - NZ, N counted only in chunks where both sides hold an empty
    register): simulated sketches from empty to full, with planted equal registers, tiles that are not multiples of
    32 x 64, empties on one side only.  C and N are integers, so the result must equal the generic kernel's bit for bit."""
    rng = np.random.default_rng(7 * n_ref + n_qry)
    M = 16384

    def simulated(fill, rows):
        lz = np.clip(rng.geometric(0.5, size=(rows, M)), 1, 50).astype(np.uint16)
        regs = (lz << 10) | rng.integers(0, 1024, size=(rows, M), dtype=np.uint16)
        return np.where(rng.random((rows, M)) < fill, regs, 0).astype(np.uint16)

    def mixed(n, base):
        fills = [0.002, 0.3, 0.97, 1.0, 1.0]
        rows = [simulated(fills[i % len(fills)], 1)[0] for i in range(n)]
        for i in range(n):                                        # related sketches: share registers with a common base
            share = rng.random(M) < (0.05 + 0.9 * rng.random())
            rows[i] = np.where(share & (rows[i] != 0), base, rows[i]).astype(np.uint16)
        rows[0][:] = 0                                            # a completely empty sketch
        if n > 3:
            rows[3][: M // 2] = 0                                 # empties in the first half only
            rows[2][1::2] = 0                                     # every high halfword of the register words empty
        return np.stack(rows)

    base = simulated(1.0, 1)[0]
    ref, qry = mixed(n_ref, base), mixed(n_qry, base)
    for i in range(min(n_ref, n_qry)):                            # C and N of a few pairs against plain numpy
        c, n = oracle.hmh_counts(ref[i], qry[i])
        assert c == int(((ref[i] == qry[i]) & (ref[i] != 0)).sum()) and n == int(((ref[i] != 0) | (qry[i] != 0)).sum())
    got, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, 2, False, ref, qry)            # frac
    exp = oracle.dist(ALGO_HMH, 14, 16, 0, 2, False, ref, qry, threads=8)
    assert ((got > 0) & (got < 1)).any() and (got == 0).any()
    np.testing.assert_allclose(got, exp, rtol=1e-12, atol=1e-300)
    sq = np.concatenate([ref, qry])[:90]
    dense, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, MODEL_POISSON, False, sq, sq)
    tri, _ = ops.dist(gpu_ctx, ALGO_HMH, 14, 16, 0, MODEL_POISSON, False, sq, sq, triangular=True)
    np.testing.assert_array_equal(tri, dense[np.tril_indices(len(sq))])
    # the generic kernel (LASH_HMH_KERNEL=generic is read once per process: ask a child)
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        np.save(f"{tmp}/sq.npy", sq)
        code = ("import numpy as np, sys; sys.path.insert(0, '.'); from lash_b200 import ops, ALGO_HMH\n"
                f"sq = np.load('{tmp}/sq.npy')\n"
                "with ops.Context(0) as c: d, _ = ops.dist(c, ALGO_HMH, 14, 16, 0, 1, False, sq, sq)\n"
                f"np.save('{tmp}/d.npy', d)\n")
        import os
        subprocess.check_call([sys.executable, "-c", code], env={**os.environ, "LASH_HMH_KERNEL": "generic"},
                              cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        np.testing.assert_array_equal(np.load(f"{tmp}/d.npy"), dense)


@pytest.mark.parametrize("p", [5, 6, 10, 12])
def test_ml_g_sum_tiles_equal_table_tiles_bit_for_bit(oracle, gpu_ctx, tmp_path, p):
    """K4c on tiles whose sketches hold no empty / small-range register takes S from min(G_a, G_b) sums instead of the 64-bit
    contribution table (dist_tables.cuh).  S and b[] are integers, so the distances must equal those of the
    table form (LASH_ML_S=table, read once per process -> a child) bit for bit: plain G-sum tiles, a register at the last
    level inside (k = 27), one just above (k = 28: the sketch is flagged, exact path), one outside the pair table, and
    tiles that fall back to the table because one of their sketches has an empty or a small-range register."""
    import subprocess
    import sys
    rng = np.random.default_rng(p)
    m = 1 << p
    n = 200                                                       # 13 row tiles x 4 column tiles
    lvl = np.clip(np.floor(7.0 - np.log2(-np.log(rng.random((n, m))))), 0, 27).astype(np.int64)
    regs = (4 * p + 4 + 4 * lvl + rng.integers(0, 4, size=(n, m))).astype(np.uint8)
    regs[1::2] = np.where(rng.random((n // 2, m)) < 0.5, regs[0::2], regs[1::2])   # related neighbours
    lo = 4 * p + 4
    regs[5, 3] = lo + 4 * 27 + 2                                  # k = 27: last level of the G-sum window
    regs[70, 9] = lo + 4 * 28 + 1                                 # k = 28: flagged
    regs[71, 0] = 251                                             # outside the pair table
    regs[130, 17] = 0                                             # an empty register: its tiles use the table
    regs[131, 4] = lo - 2                                         # a small-range register
    regs[199, :] = lo                                             # the largest G everywhere (32-bit batches at their maximum)
    np.save(tmp_path / "regs.npy", regs)
    child = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from lash_b200 import ALGO_ULL, EST_ML, ops\n"
        "regs = np.load(%r)\n"
        "with ops.Context(0) as ctx:\n"
        "    d, _ = ops.dist(ctx, ALGO_ULL, %d, 16, EST_ML, 2, False, regs, regs)\n"
        "    r, _ = ops.dist(ctx, ALGO_ULL, %d, 16, EST_ML, 2, False, regs[60:140], regs[100:])\n"
        "np.save(%r, d); np.save(%r, r)\n" % (str(__import__('os').path.dirname(__import__('os').path.dirname(__file__))), str(tmp_path / "regs.npy"),
                                               p, p, str(tmp_path / "table.npy"), str(tmp_path / "table_rect.npy")))
    subprocess.run([sys.executable, "-c", child], check=True, env=dict(__import__('os').environ, LASH_ML_S="table"))
    dense, w = ops.dist(gpu_ctx, ALGO_ULL, p, 16, EST_ML, 2, False, regs, regs)
    assert w == 0
    np.testing.assert_array_equal(dense, np.load(tmp_path / "table.npy"))
    rect, _ = ops.dist(gpu_ctx, ALGO_ULL, p, 16, EST_ML, 2, False, regs[60:140], regs[100:])
    np.testing.assert_array_equal(rect, np.load(tmp_path / "table_rect.npy"))
    tri, _ = ops.dist(gpu_ctx, ALGO_ULL, p, 16, EST_ML, 2, False, regs, regs, triangular=True)
    np.testing.assert_array_equal(tri, dense[np.tril_indices(n)])
    exp = oracle.dist(ALGO_ULL, p, 16, EST_ML, 2, False, regs[:64], regs[:64], threads=8)
    np.testing.assert_allclose(dense[:64, :64], exp, rtol=1e-11, atol=1e-15)


def test_ml_two_kernel_form_matches_fused_in_every_output_layout(oracle, gpu_ctx, tmp_path):
    """ULL ML runs as a tile kernel that stores the pair statistics + ml_finish_kernel.  The scratch is
    addressed by output cell, so every output layout is its own case: dense, packed triangle, streamed dense blocks
    (out_row0 != 0, triangular or not).  All must equal the fused single-kernel form (LASH_ML_KERNEL=fused, read once
    per process -> a child process) bit for bit, and the oracle within tolerance."""
    import subprocess
    import sys
    regs = _sketches(oracle, ALGO_ULL, 10, 16, 75, 60_000)
    regs[7] = 0                                                   # an empty sketch (S == 0 branch)
    regs[9, 5] = 251                                              # outside the pair table: exact per-pair path
    np.save(tmp_path / "regs.npy", regs)
    child = (
        "import sys, numpy as np; sys.path.insert(0, %r)\n"
        "from lash_b200 import ALGO_ULL, EST_ML, MODEL_POISSON, ops\n"
        "regs = np.load(%r)\n"
        "with ops.Context(0) as ctx:\n"
        "    d, _ = ops.dist(ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, False, regs, regs)\n"
        "np.save(%r, d)\n" % (str(__import__('os').path.dirname(__import__('os').path.dirname(__file__))), str(tmp_path / "regs.npy"),
                             str(tmp_path / "fused.npy")))
    env = dict(__import__('os').environ, LASH_ML_KERNEL="fused")
    subprocess.run([sys.executable, "-c", child], check=True, env=env)
    fused = np.load(tmp_path / "fused.npy")
    dense, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, False, regs, regs)
    np.testing.assert_array_equal(dense, fused)
    exp = oracle.dist(ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, False, regs, regs)
    frac = oracle.dist(ALGO_ULL, 10, 16, EST_ML, 2, False, regs, regs)
    both_nan = np.isnan(dense) & np.isnan(exp)
    _assert_close_f64(np.where(both_nan, 0, dense), np.where(both_nan, 0, exp), np.nan_to_num(frac), 16, "ml two-kernel")
    tri, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, False, regs, regs, triangular=True)
    np.testing.assert_array_equal(tri, dense[np.tril_indices(len(regs))])
    for triangular in (False, True):
        got = np.full_like(dense, np.nan)

        def on_block(row0, block):
            got[row0: row0 + block.shape[0]] = block

        ops.dist_stream(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_POISSON, False, regs, regs, triangular, 7, on_block)
        if triangular:
            il = np.tril_indices(len(regs))
            np.testing.assert_array_equal(got[il], dense[il])
        else:
            np.testing.assert_array_equal(got, dense)
    # a rectangular problem with different sets on the two sides, fp32 output
    d32, _ = ops.dist(gpu_ctx, ALGO_ULL, 10, 16, EST_ML, MODEL_BINOMIAL, True, regs[:20], regs[30:])
    e32 = oracle.dist(ALGO_ULL, 10, 16, EST_ML, MODEL_BINOMIAL, True, regs[:20], regs[30:])
    np.testing.assert_allclose(d32, e32, rtol=1e-6, atol=2e-7)
