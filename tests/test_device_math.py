"""The DEVICE arithmetic headers (lash_b200/csrc/hash.cuh, registers.cuh) compiled with g++ through a small intrinsic
shim (tests/host_shim/device_math.cpp) and checked on the CPU: the hash forms the kernels' fast paths use, the register
algebra, and the rules that decide which hashes the fast paths may handle.  The GPU tests check the same things end to
end on registers; these run where there is no GPU, against python-xxhash (an independent implementation of the frozen
XXH3 spec) and the oracle.
"""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "host_shim", "device_math.cpp")
HDRS = [os.path.join(ROOT, "lash_b200", "csrc", h) for h in ("hash.cuh", "registers.cuh")]
SEED = 42
U64 = np.uint64


@pytest.fixture(scope="module")
def dm(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("dm") / "libdevice_math.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, SRC])
    L = C.CDLL(out)
    vp, u64, i32, u32 = C.c_void_p, C.c_uint64, C.c_int, C.c_uint32
    L.dm_xxh3_64.argtypes = [vp, u64, u64, vp]
    L.dm_pre.argtypes = [vp, u64, u64, i32, vp, vp]
    L.dm_xxh3_128.argtypes = [vp, u64, u64, vp, vp]
    L.dm_cell.argtypes = [i32, vp, u64, u64, i32, vp, vp]
    L.dm_ull_fast.argtypes = [vp, u64, i32, i32, vp, vp, vp]
    for f in (L.dm_ull_update, L.dm_ull_merge1, L.dm_ull_merge4):
        f.argtypes, f.restype = [u32, u32], u32
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _values(n, bits, seed=0):
    rng = np.random.default_rng(seed)
    v = rng.integers(0, 1 << 63, size=n, dtype=np.uint64) * U64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if bits < 64:
        v &= U64((1 << bits) - 1)
    edge = np.array([0, 1, (1 << bits) - 1, 1 << (bits - 1), 0x0123456789ABCDEF & ((1 << bits) - 1)], dtype=np.uint64)
    return np.concatenate([edge, v])


def test_device_xxh3_64_is_xxh3(dm):
    xxhash = pytest.importorskip("xxhash")
    for seed in (SEED, 0, 93, (1 << 64) - 1):
        v = _values(3000, 64, seed=seed & 0xFFFF)
        out = np.empty_like(v)
        dm.dm_xxh3_64(_p(v), len(v), seed, _p(out))
        exp = np.array([xxhash.xxh3_64_intdigest(int(x).to_bytes(8, "little"), seed=seed) for x in v], dtype=np.uint64)
        assert np.array_equal(out, exp)


@pytest.mark.parametrize("narrow", [1, 0])
def test_pre_xorshift_forms_and_their_high_word_variants(dm, narrow):
    """xxh3_64_narrow_pre / wide_pre return g with h = g ^ (g >> 28); the *_hi variants (second multiply without its
    low word, every multiply as IMAD.WIDE + two chained mads) must be exactly g >> 32."""
    v = _values(50_000, 32 if narrow else 64, seed=7)
    g = np.empty_like(v)
    ghi = np.empty(len(v), dtype=np.uint32)
    full = np.empty_like(v)
    dm.dm_pre(_p(v), len(v), SEED, narrow, _p(g), _p(ghi))
    dm.dm_xxh3_64(_p(v), len(v), SEED, _p(full))
    assert np.array_equal(g ^ (g >> U64(28)), full)
    assert np.array_equal(ghi, (g >> U64(32)).astype(np.uint32))


def test_device_xxh3_128_of_4_bytes(dm):
    xxhash = pytest.importorskip("xxhash")
    w = _values(3000, 32, seed=3).astype(np.uint32)
    lo, hi = np.empty(len(w), dtype=np.uint64), np.empty(len(w), dtype=np.uint64)
    dm.dm_xxh3_128(_p(w), len(w), SEED, _p(lo), _p(hi))
    for x, a, b in zip(w[:1500], lo, hi):
        d = xxhash.xxh3_128_intdigest(int(x).to_bytes(4, "little"), seed=SEED)
        assert (int(b) << 64) | int(a) == d


@pytest.mark.parametrize("p,drop4", [(10, 1), (3, 1), (11, 1), (12, 0), (14, 0), (10, 0)])
def test_ull_fast_path_bits_or_rare(dm, p, drop4):
    """SmemAcc<ULL>::prep: from the pre-xorshift high word alone the fast path must either produce the exact
    (index, nlz) of the finished hash or declare the hash rare (-> exact path); it may never produce a wrong bit.  Random
    hashes plus hashes built to sit on every boundary (all-zero fields of every length after the index)."""
    rng = np.random.default_rng(p * 2 + drop4)
    g = rng.integers(0, 1 << 63, size=200_000, dtype=np.uint64) * U64(2) + rng.integers(0, 2, size=200_000, dtype=np.uint64)
    # adversarial: index | z zeros | one | random tail, for every z that fits the high word and a few beyond it
    adv = []
    for z in range(0, 40):
        for _ in range(50):
            idx = int(rng.integers(0, 1 << p))
            below = 63 - p - z
            adv.append((idx << (64 - p)) | (1 << below) | int(rng.integers(0, 1 << below)))
    h_adv = np.array(adv, dtype=np.uint64)
    # invert h = g ^ (g >> 28) to get the g whose finished hash is h_adv
    g_adv = h_adv ^ (h_adv >> U64(28)) ^ (h_adv >> U64(56))
    g = np.concatenate([g, g_adv])
    h = g ^ (g >> U64(28))
    ghi = (g >> U64(32)).astype(np.uint32)
    idx, nlz, rare = (np.empty(len(g), dtype=np.uint32) for _ in range(3))
    dm.dm_ull_fast(_p(ghi), len(g), p, drop4, _p(idx), _p(nlz), _p(rare))
    exp_idx = (h >> U64(64 - p)).astype(np.uint32)
    body = (h << U64(p)) | U64((1 << p) - 1)                      # ultraloglog: nlz = clz(~(~h << p))
    exp_nlz = np.array([64 - int(x).bit_length() for x in body], dtype=np.uint32)
    assert np.array_equal(idx, exp_idx)
    fast = rare == 0
    assert np.array_equal(nlz[fast], exp_nlz[fast])
    # rare exactly when the bits the fast path looks at are all zero
    looked_at = 32 - p - (4 if drop4 else 0)
    assert np.array_equal(rare == 1, exp_nlz >= looked_at)
    assert rare[: 200_000].mean() < 4 * 2.0 ** -looked_at + 1e-4


@pytest.mark.parametrize("algo,p", [(1, 8), (1, 14), (2, 10), (2, 14), (2, 20), (0, 14)])
def test_cell_from_kmer_matches_the_oracle_sketch_of_one_kmer(dm, oracle, algo, p):
    """Cell<ALGO>::from_kmer (exact path, global accumulators, flush) against the oracle sketching a single 32-mer."""
    from tools import synth
    from tools.xxh3_invert import revcomp_value
    rng = np.random.default_rng(algo * 100 + p)
    vals = []
    while len(vals) < 40:
        v = int(rng.integers(0, 1 << 63)) * 2 + int(rng.integers(0, 2))
        if v <= revcomp_value(v, 32):                              # its own canonical form
            vals.append(v)
    v = np.array(vals, dtype=np.uint64)
    idx, val = np.empty(len(v), dtype=np.uint32), np.empty(len(v), dtype=np.uint32)
    dm.dm_cell(algo, _p(v), len(v), SEED, p, _p(idx), _p(val))
    for x, i, r in zip(vals, idx, val):
        seq = bytes(b"ACGT"[(x >> (2 * (31 - j))) & 3] for j in range(32))
        regs = oracle.sketch_genomes(algo, p, 32, SEED, [[seq]])[0]
        nz = np.flatnonzero(regs)
        assert list(nz) == [int(i)]
        # ULL: one update of an empty register gives pack(1 << u) = 4u; HLL/HMH: the value itself
        assert int(regs[i]) == (4 * int(r) if algo == 2 else int(r))
    assert synth is not None


def test_ull_register_algebra_exhaustive(dm, oracle):
    p = 8
    valid = np.array([0, 4 * p - 4, 4 * p, 4 * p + 2] + list(range(4 * p + 4, 256)), dtype=np.uint8)
    a = np.repeat(valid, len(valid))
    b = np.tile(valid, len(valid))
    m = 1 << p
    assert len(a) % m == 0
    merge = lambda x, y: np.concatenate([oracle.ull_merge(x[o:o + m], y[o:o + m], p) for o in range(0, len(x), m)])  # noqa: E731
    exp = merge(a, b)
    got1 = np.array([dm.dm_ull_merge1(int(x), int(y)) for x, y in zip(a, b)], dtype=np.uint8)
    assert np.array_equal(got1, exp)
    n4 = len(a) // 4 * 4
    wa, wb = a[:n4].view(np.uint32), b[:n4].view(np.uint32)
    got4 = np.array([dm.dm_ull_merge4(int(x), int(y)) for x, y in zip(wa, wb)], dtype=np.uint32).view(np.uint8)
    assert np.array_equal(got4, exp[:n4])
    # update(r, u) == merge(r, pack(1 << u)) for every valid register and every u the hash can produce
    for u in range(p - 1, 64):
        regs = np.zeros(m, dtype=np.uint8)
        regs[: len(valid)] = valid
        single = np.zeros(m, dtype=np.uint8)
        single[: len(valid)] = 4 * u
        exp_u = oracle.ull_merge(regs, single, p)[: len(valid)]
        got_u = np.array([dm.dm_ull_update(int(r), u) for r in valid], dtype=np.uint8)
        assert np.array_equal(got_u, exp_u), u


def test_shim_sources_are_the_product_headers():
    """The shim includes the product headers by relative path -- no copies that could drift."""
    text = open(SRC).read()
    assert '#include "../../lash_b200/csrc/registers.cuh"' in text
    for h in HDRS:
        assert os.path.exists(h)
    assert "LASH_HOST_SHIM" not in open(os.path.join(ROOT, "lash_b200", "csrc", "Makefile")).read()


# ------------------------------------------------------------------------------------------------------------------
# estimator epilogues (lash_b200/csrc/estimators.cuh) on the CPU
# ------------------------------------------------------------------------------------------------------------------
EST_SRC = os.path.join(ROOT, "tests", "host_shim", "estimators.cpp")
EPS = np.finfo(np.float64).eps


@pytest.fixture(scope="module")
def est(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("est") / "libestimators.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, EST_SRC])
    L = C.CDLL(out)
    d, i32, u32, u64, vp = C.c_double, C.c_int, C.c_uint32, C.c_uint64, C.c_void_p
    L.dm_ull_reg.argtypes, L.dm_ull_reg.restype = [i32], d
    L.dm_hll_len.argtypes, L.dm_hll_len.restype = [d, u32, i32, vp], d
    L.dm_fgra.argtypes, L.dm_fgra.restype = [d, vp, i32], d
    L.dm_ml.argtypes, L.dm_ml.restype = [u64, vp, i32, u32], d
    L.dm_hmh_card.argtypes, L.dm_hmh_card.restype = [d, d], d
    L.dm_hmh_similarity.argtypes, L.dm_hmh_similarity.restype = [u32, u32, d, d], d
    L.dm_hmh_ec_term.argtypes, L.dm_hmh_ec_term.restype = [i32, i32, d], d
    L.dm_hmh_ec.argtypes, L.dm_hmh_ec.restype = [d, d], d
    L.dm_hmh_ec_rows.argtypes, L.dm_hmh_ec_rows.restype = [], i32
    L.dm_mash64.argtypes, L.dm_mash64.restype = [d, i32, i32], d
    L.dm_mash32.argtypes, L.dm_mash32.restype = [C.c_float, i32, i32], C.c_float
    return L


def _simulated(algo, p, log2_per_reg, rng):
    """Registers of a sketch that saw about 2^log2_per_reg hashes per register (negative: mostly empty)."""
    m = 1 << p
    if log2_per_reg < 0:
        hit = rng.random(m) < 2.0 ** log2_per_reg
        lvl = np.clip(rng.geometric(0.5, size=m) - 1, 0, 60 - p)
    else:
        hit = np.ones(m, dtype=bool)
        lvl = np.clip(np.floor(log2_per_reg - np.log2(-np.log(rng.random(m)))), 0, 60 - p).astype(np.int64)
    if algo == 1:
        return (hit * (lvl + 1)).astype(np.uint8)
    # ULL: register of the OR of a few bits around the top one
    return (hit * (4 * (lvl + p - 1) + rng.integers(0, 4, size=m))).astype(np.uint8)


def _close(a, b, ulps=8):
    return (np.isnan(a) and np.isnan(b)) or a == b or abs(a - b) <= ulps * EPS * abs(b)


@pytest.mark.parametrize("p", [4, 10, 14])
def test_hll_len_epilogue(est, oracle, p):
    rng = np.random.default_rng(p)
    for lg in (-6.0, -2.0, 0.5, 3.0, 9.0, 20.0):
        regs = _simulated(1, p, lg, rng)
        s = 0.0
        for r in regs:                              # register order, exact powers of two (what the kernels add)
            s += 2.0 ** -int(r)
        bias = C.c_int(0)
        got = est.dm_hll_len(s, int((regs == 0).sum()), p, C.byref(bias))
        exp = oracle.cardinality(1, p, 0, regs)
        assert _close(got, exp), (p, lg, got, exp)
        assert bool(bias.value) == bool(np.isnan(exp))


@pytest.mark.parametrize("p", [3, 4, 10, 14])
def test_ull_fgra_and_ml_epilogues(est, oracle, p):
    """ull_fgra_finalize from (sum, counts) and ull_ml_finalize from (S, b[]) -- the statistics the tile kernels hand
    to the epilogue -- against the oracle's estimate from the registers, across small-range, normal and saturated
    sketches.  The ML solver's exact power-of-two / exponent constructions must leave every iteration unchanged."""
    rng = np.random.default_rng(100 + p)
    off = 4 * p + 4
    reg_tab = [est.dm_ull_reg(i) for i in range(256)]
    cases = [_simulated(2, p, lg, rng) for lg in (-5.0, -1.5, 0.0, 2.0, 7.0, 12.0, 25.0)]
    sat = _simulated(2, p, 12.0, rng)
    sat[:: 5] = rng.integers(252, 256, size=len(sat[:: 5]))          # saturated registers: FGRA large-range term
    cases.append(sat)
    small = np.zeros(1 << p, dtype=np.uint8)
    small[:: 3] = rng.choice([4 * p - 4, 4 * p, 4 * p + 2], size=len(small[:: 3]))   # only small-range registers
    cases.append(small)
    cases.append(np.zeros(1 << p, dtype=np.uint8))                   # empty sketch
    for regs in cases:
        s, cnt = 0.0, np.zeros(8, dtype=np.uint32)
        for r in regs:
            r = int(r)
            r2 = r - off
            if r2 < 0:
                cnt[0] += r2 < -8
                cnt[1] += r2 == -8
                cnt[2] += r2 == -4
                cnt[3] += r2 == -2
            elif r < 252:
                s += reg_tab[r2]
            else:
                cnt[4 + r - 252] += 1
        got = est.dm_fgra(s, _p(cnt), p)
        exp = oracle.cardinality(2, p, 0, regs)
        assert _close(got, exp), ("fgra", p, got, exp)
        S, b = oracle.ull_ml_stats(regs, p)
        got = est.dm_ml(S, _p(np.ascontiguousarray(b, dtype=np.int32)), p, int(regs[0]))
        exp = oracle.cardinality(2, p, 1, regs)
        assert _close(got, exp) or (np.isinf(got) and np.isinf(exp)), ("ml", p, got, exp)


def test_hmh_and_mash_epilogues(est, oracle):
    rng = np.random.default_rng(5)
    m = 16384
    lvl = np.clip(np.floor(7.0 - np.log2(-np.log(rng.random((2, m))))), 0, 40).astype(np.int64)
    regs = ((lvl + 1) << 10 | rng.integers(0, 1024, size=(2, m))).astype(np.uint16)
    regs[1, ::3] = regs[0, ::3]                                      # shared registers -> collisions
    cards = []
    for r in regs:
        s, ez = 0.0, 0.0
        for v in r:
            lz = int(v) >> 10
            ez += lz == 0
            s += 2.0 ** -lz
        got = est.dm_hmh_card(s, ez)
        exp = oracle.cardinality(0, 14, 0, r)
        assert _close(got, exp), (got, exp)
        cards.append(exp)
    c, n = oracle.hmh_counts(regs[0], regs[1])
    sim = est.dm_hmh_similarity(c, n, cards[0], cards[1])
    frac_exp = oracle.dist(0, 14, 16, 0, 2, False, regs[:1], regs[1:])[0, 0]
    s = max(sim, 0.0)
    assert _close(2.0 * s / (1.0 + s), frac_exp), (sim, frac_exp)
    for frac in (1.0, 0.5, 1e-3, 1e-9, 0.0):
        for k in (16, 21, 31):
            assert _close(est.dm_mash64(frac, k, 1), min(-np.log(frac) / k, 1.0) if frac > 0 else 1.0, ulps=4)
            assert _close(est.dm_mash64(frac, k, 0), 1.0 - frac ** (1.0 / k), ulps=4)
            assert est.dm_mash64(frac, k, 2) == frac
            f32 = np.float32(frac)
            assert abs(float(est.dm_mash32(f32, k, 0)) - float(np.float32(1) - np.power(f32, np.float32(1) / np.float32(k)))) <= 2e-7


def test_hmh_expected_collision_terms_and_sum(est, oracle):
    d_ = C.c_double
    """K4m's small-sketch path (estimators.cuh: hmh_ec_term, kHmhEcRows): the term of rows beyond kHmhEcRows is exactly +0 for
    every cardinality that takes the loop (so cutting the 64 x 1024 loop at row 41 changes nothing), row 41 still holds a
    non-zero term, and the loop sum equals the oracle's expectedCollision (glibc pow there, correctly rounded pow here)."""
    rows = est.dm_hmh_ec_rows()
    assert rows == 41
    for n in (1.0, 17.0, 1234.5, 65536.0, 524288.0):
        for i in range(rows + 1, 64):
            for j in (1, 2, 511, 1023, 1024):
                t = est.dm_hmh_ec_term(i, j, n)
                assert t == 0.0 and not np.signbit(t), (i, j, n, t)
    assert est.dm_hmh_ec_term(41, 1024, 524288.0) != 0.0
    # b = (1024 + j) / 2^(24 + i) as the reference builds it (a division by a power of two) is the product used here
    for i in (1, 7, 40):
        for j in (1, 333, 1024):
            b1 = (1024.0 + j) / 2.0 ** (24 + i)
            b2 = (1024.0 + j + 1.0) / 2.0 ** (24 + i)
            n = 70000.0
            t = est.dm_hmh_ec_term(i, j, n)
            ref = (1.0 - b2) ** n - (1.0 - b1) ** n
            assert abs(t - ref) <= 4 * EPS, (i, j, t, ref)
    L = oracle.lib()
    for n, m in ((3.0, 2.0), (1000.0, 10.0), (52345.25, 480000.0), (524288.0, 524288.0), (100.0, 100.0)):
        got = est.dm_hmh_ec(n, m)
        exp = L.lo_hmh_expected_collisions(n, m)
        assert abs(got - exp) <= 1e-13 * max(abs(exp), 1.0), (n, m, got, exp)
    # the closed form above 2^19 is untouched
    assert _close(est.dm_hmh_ec(3e6, 2e6), L.lo_hmh_expected_collisions(3e6, 2e6), ulps=4)
    # the tile product's early end: wherever the rule fires, the partial sum IS the full 41-row sum
    est.dm_hmh_ec_early.argtypes, est.dm_hmh_ec_early.restype = [d_, d_, C.c_void_p, C.c_void_p], C.c_int
    full, early = C.c_double(0), C.c_double(0)
    stops = []
    rng = np.random.default_rng(3)
    cases = [(1.0, 1.0), (2.0, 524288.0), (17.0, 3.0), (524288.0, 524288.0), (1000.0, 1000.0), (65536.0, 9.0), (300000.0, 120000.5)]
    cases += [tuple(np.exp(rng.uniform(0, np.log(524288.0), size=2))) for _ in range(12)]
    for n, m in cases:
        stop = est.dm_hmh_ec_early(float(n), float(m), C.byref(full), C.byref(early))
        assert early.value == full.value, (n, m, stop, early.value, full.value)
        stops.append(stop)
    assert min(stops) < 30, stops     # the rule does fire (typically after 23-27 rows)


# ------------------------------------------------------------------------------------------------------------------
# pair tables of the distance kernels (lash_b200/csrc/dist_tables.cuh) on the CPU
# ------------------------------------------------------------------------------------------------------------------
def _valid_registers(p):
    return np.array([0, 4 * p - 4, 4 * p, 4 * p + 2] + list(range(4 * p + 4, 256)), dtype=np.uint8)


def _merge_pairs(oracle, a, b, p):
    m = 1 << p
    pad = (-len(a)) % m
    a2, b2 = np.concatenate([a, np.zeros(pad, np.uint8)]), np.concatenate([b, np.zeros(pad, np.uint8)])
    out = np.concatenate([oracle.ull_merge(a2[o:o + m], b2[o:o + m], p) for o in range(0, len(a2), m)])
    return out[: len(a)]


@pytest.mark.parametrize("p,base_shift", [(10, 0), (10, 24), (14, 0), (14, 60), (4, 0)])
def test_fgra_pair_table_reproduces_the_merge_and_its_contribution(dm, est, oracle, p, base_shift):
    """Every entry of the 128 x 128 table dist_fgra_tab_kernel builds (for a window anchored at `base`) equals
    REGISTER_CONTRIBUTIONS[merge(ra, rb) - (4p+4)] of the two registers the codes stand for, or the sentinel exactly
    when a register is outside the window or the merged register needs the small- / large-range treatment."""
    dm.dm_fgra_code.argtypes, dm.dm_fgra_code.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    dm.dm_fgra_table.argtypes, dm.dm_fgra_table.restype = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p], C.c_double
    dm.dm_ull_merge_fast.argtypes, dm.dm_ull_merge_fast.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    reg = np.array([est.dm_ull_reg(i) for i in range(256)], dtype=np.float64)
    off, base = 4 * p + 4, 4 * p - 4 + base_shift
    table = np.empty(128 * 128, dtype=np.float64)
    sentinel = dm.dm_fgra_table(base, p, _p(reg), _p(table))
    table = table.reshape(128, 128)
    # the kernel's contract: base <= every non-empty register present (base = 4p-4, or the sets' smallest register
    # rounded down to a multiple of 4), so the registers a table can meet are 0 and the valid ones >= base
    valid = _valid_registers(p)
    valid = valid[(valid == 0) | (valid >= base)]
    codes = np.array([dm.dm_fgra_code(int(r), base) for r in valid])
    inside = codes != 127
    # recoding: 0 <-> empty, 1..126 <-> base .. base+125, everything above 127; injective inside the window
    assert codes[0] == 0
    for r, c in zip(valid[1:], codes[1:]):
        assert c == (int(r) - base + 1 if int(r) <= base + 125 else 127)
    assert dm.dm_fgra_code(max(base - 2, 1), base) == 127                 # far below the window (cannot happen): flagged
    a = np.repeat(valid[inside], inside.sum())
    b = np.tile(valid[inside], inside.sum())
    merged = _merge_pairs(oracle, a, b, p).astype(np.int64)
    fast = np.array([dm.dm_ull_merge_fast(int(x), int(y)) for x, y in zip(a, b)])
    assert np.array_equal(fast, merged)                                   # the ALU merge of K4 agrees too
    ca, cb = np.repeat(codes[inside], inside.sum()), np.tile(codes[inside], inside.sum())
    got = table[ca, cb]
    in_range = (merged >= off) & (merged < 252)
    assert np.array_equal(got[in_range], reg[merged[in_range] - off])
    assert np.all(got[~in_range] == sentinel)
    assert np.all(table[127, :] == sentinel) and np.all(table[:, 127] == sentinel)


@pytest.mark.parametrize("p", [4, 10, 14])
def test_ml_pair_tables_sum_to_the_ml_statistics(dm, oracle, p):
    """R / W tables of dist_ml_tab_kernel: summing R (mod 2^64) and counting the bits of W over the registers of two
    sketches gives exactly the (S, b[]) hash4j's contribute() produces on the merged sketch."""
    dm.dm_fgra_code.argtypes, dm.dm_fgra_code.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    dm.dm_ml_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    R, W = np.empty(128 * 128, dtype=np.uint64), np.empty(128 * 128, dtype=np.uint32)
    dm.dm_ml_tables(p, _p(R), _p(W))
    rng = np.random.default_rng(p)
    m, base = 1 << p, 4 * p - 4
    for lg in (-3.0, 0.0, 3.0, 10.0, 20.0):
        a, b = _simulated(2, p, lg, rng), _simulated(2, p, lg + 1.0, rng)
        a[a > base + 125] = base + 125
        b[b > base + 125] = base + 125                                    # keep every register inside the table window
        ca = np.array([dm.dm_fgra_code(int(r), base) for r in a])
        cb = np.array([dm.dm_fgra_code(int(r), base) for r in b])
        assert ca.max() < 127 and cb.max() < 127
        e = ca * 128 + cb
        S = int(np.sum(R[e].astype(object)) % (1 << 64))
        bits = np.zeros(66, dtype=np.int64)
        for j in range(32):
            bits[j] = int(((W[e] >> np.uint32(j)) & np.uint32(1)).sum())
        S_exp, b_exp = oracle.ull_ml_stats(oracle.ull_merge(a, b, p), p)
        assert S == S_exp
        assert np.array_equal(bits, b_exp.astype(np.int64))


@pytest.mark.parametrize("p", [4, 10, 14, 20, 26])
def test_ml_g_sum_form_of_S_equals_the_contribution_sum(dm, oracle, p):
    """K4c's G-sum tiles (dist_tables.cuh): for sketches without empty or small-range registers whose top levels lie within
    27 of the smallest one (k0),  S = (sum of min(G_a, G_b)) << (36 - p - k0)  -  sum_j b[j] << (63 - j - p)  (mod 2^64) is
    hash4j's sum of contribute() over the merged sketch -- through the staged formats the tile kernel uses (packed query
    words, 32-bit batches of eight)."""
    dm.dm_ml_gs_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_int, C.c_uint32, C.c_void_p, C.c_void_p]
    dm.dm_ml_gs_pair.restype = C.c_uint64
    dm.dm_ml_gs_max_reg.argtypes, dm.dm_ml_gs_max_reg.restype = [C.c_int, C.c_uint32], C.c_uint32
    dm.dm_ml_gs_k.argtypes, dm.dm_ml_gs_k.restype = [C.c_uint32, C.c_int], C.c_uint32
    dm.dm_ml_ret_of.argtypes, dm.dm_ml_ret_of.restype = [C.c_uint32, C.c_int], C.c_uint64
    dm.dm_ull_merge1.argtypes, dm.dm_ull_merge1.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    rng = np.random.default_rng(p)
    m = min(1 << p, 1 << 14)                       # the identity is per register: 2^14 registers are plenty
    base = 4 * p + 4
    gsum, bits = C.c_uint64(0), np.zeros(32, dtype=np.int32)
    for shape, kmin in (("genome", 0), ("genome", 2), ("flat", 0), ("flat", 2), ("all-low", 1), ("all-high", 0)):
        if kmin + p > 36:
            continue
        kmax = min(kmin + 27, 28)                                   # the pair table ends inside k = 29 (codes <= 126)
        if shape == "genome":
            lvl = np.clip(np.floor(kmin + 6.0 - np.log2(-np.log(rng.random((2, m))))), kmin, kmax).astype(np.int64)
        elif shape == "flat":
            lvl = rng.integers(kmin, kmax + 1, size=(2, m))
        elif shape == "all-low":
            lvl = np.full((2, m), kmin)                             # eight copies of the largest term in every batch
        else:
            lvl = np.full((2, m), kmax)
            lvl[0, 0] = kmin
        regs = (base + 4 * lvl + rng.integers(0, 4, size=(2, m))).astype(np.uint8)
        k0 = dm.dm_ml_gs_k(int(regs.min()), p)
        assert k0 >= kmin and int(regs.max()) <= dm.dm_ml_gs_max_reg(p, k0) and k0 + p <= 36
        S = dm.dm_ml_gs_pair(_p(regs[0]), _p(regs[1]), m, p, k0, C.byref(gsum), _p(bits))
        if m == 1 << p:
            S_exp, b_exp = oracle.ull_ml_stats(oracle.ull_merge(regs[0], regs[1], p), p)
            assert np.array_equal(bits.astype(np.int64), b_exp.astype(np.int64)[:32]) and not b_exp[32:].any()
            assert S == S_exp, (p, shape, hex(S), hex(S_exp))
        else:
            # larger precisions: the same identity register by register against the header's own contribute()
            S_exp = 0
            for ra, rb in zip(regs[0][:2048], regs[1][:2048]):
                S_exp += dm.dm_ml_ret_of(dm.dm_ull_merge1(int(ra), int(rb)), p)
            k0s = dm.dm_ml_gs_k(int(regs[:, :2048].min()), p)
            S2 = dm.dm_ml_gs_pair(_p(regs[0]), _p(regs[1]), 2048, p, k0s, C.byref(gsum), _p(bits))
            assert S2 == S_exp % (1 << 64), (p, shape)


@pytest.mark.parametrize("k", [1, 2, 5, 12, 14, 15, 16, 17, 21, 24, 31, 32])
def test_funnel_shift_kmer_windows_equal_string_level_canonical_kmers(dm, oracle, k):
    """kmer_windows.cuh (the sketch kernel's k-mer extraction: funnel-shift windows of the packed stream and of its per-word
    reverse complement, min, mask) against the oracle's k-mers for every start position, on random sequence plus
    homopolymers and reverse-complement palindromes."""
    from lash_b200 import hostapi
    dm.dm_kmers.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    rng = np.random.default_rng(k)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)) for n in (k, k + 1, 63, 64, 65, 500)]
    seqs += [b"A" * 70, b"T" * 70, b"ACGT" * 20, b"AATT" * 18 + b"GC"]
    for seq in seqs:
        if len(seq) < k:
            continue
        packed, nb = hostapi.filter_pack(seq, simd=0)
        assert nb == len(seq)
        buf = np.concatenate([packed, np.zeros(16, dtype=np.uint8)])
        out = np.zeros(len(seq) - k + 1, dtype=np.uint64)
        dm.dm_kmers(_p(buf), len(packed), len(seq), k, _p(out))
        exp = oracle.canonical_kmers(seq, k)
        assert np.array_equal(out, np.asarray(exp, dtype=np.uint64)), (k, len(seq))


@pytest.mark.parametrize("p", [3, 10, 14, 20, 26])
def test_ull_smem_cell_flush_equals_sequential_updates(dm, oracle, p):
    """The sketch kernel keeps, per register, two words of "seen nlz" bits and converts them once at flush
    (ull_cell_to_reg).  For sets of hashes that fall into one register -- a few, many, with very long zero runs -- the
    converted cell must equal the register ultraloglog reaches by sequential add()s (ull_update from the empty register)."""
    dm.dm_ull_cell.argtypes, dm.dm_ull_cell.restype = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p], C.c_uint32
    dm.dm_ull_update.argtypes, dm.dm_ull_update.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    rng = np.random.default_rng(p)
    for n in (1, 2, 3, 7, 50, 400):
        for zmax in (4, 20, 64 - p):
            idx = int(rng.integers(0, 1 << p))
            hs = []
            for _ in range(n):
                z = int(rng.integers(0, zmax + 1))                        # leading zeros after the index
                below = 63 - p - z
                tail = ((1 << below) | int(rng.integers(0, 1 << below))) if below >= 0 else 0
                hs.append((idx << (64 - p)) | tail)
            h = np.array(hs, dtype=np.uint64)
            out_idx, w = C.c_uint32(), np.zeros(2, dtype=np.uint32)
            reg = dm.dm_ull_cell(_p(h), n, p, C.byref(out_idx), _p(w))
            assert out_idx.value == idx
            exp = 0
            for x in hs:
                body = ((x << p) & ((1 << 64) - 1)) | ((1 << p) - 1)          # nlz = clz(~(~h << p))
                nlz = 64 - body.bit_length()
                exp = dm.dm_ull_update(exp, nlz + p - 1)
            assert reg == exp, (p, n, zmax)


@pytest.mark.parametrize("npl,p", [(12, 10), (12, 11), (16, 14), (27, 16), (12, 4)])
def test_ml_bit_sliced_counters_count_exactly(dm, npl, p):
    """MlAccT: b[j] += bit j of W over all 2^p registers, kept as vertical (bit-sliced) counters -- Harley-Seal steps
    of 8 patterns, eight carries folded per 64 registers, one ripple through the upper planes.  The extracted counts
    must equal plain popcounts for sparse, dense and all-ones patterns (counts up to 2^p need p+1 planes)."""
    dm.dm_ml_counters.argtypes = [C.c_int, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
    dm.dm_ml_counters_generic.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p]
    rng = np.random.default_rng(npl * 100 + p)
    n = 1 << p
    chunk = min(n, 128)
    for kind in ("sparse", "dense", "ones", "levels"):
        if kind == "sparse":
            w = (rng.random((n, 32)) < 0.03)
        elif kind == "dense":
            w = (rng.random((n, 32)) < 0.6)
        elif kind == "ones":
            w = np.ones((n, 32), dtype=bool)
        else:                                                             # what ULL registers produce: (4 | y) << k
            k = np.clip(rng.geometric(0.3, size=n), 0, 29)
            pat = ((4 | rng.integers(0, 4, size=n)).astype(np.uint64) << k.astype(np.uint64)) & np.uint64(0xFFFFFFFF)
            w = ((pat[:, None] >> np.arange(32, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
        words = (w.astype(np.uint64) << np.arange(32, dtype=np.uint64)[None, :]).sum(axis=1).astype(np.uint32)
        exp = w.sum(axis=0).astype(np.int32)
        bb = np.zeros(66, dtype=np.int32)
        dm.dm_ml_counters(npl, _p(words), n, chunk, _p(bb))
        assert np.array_equal(bb[:32], exp), (kind, "tab kernel path")
        assert not bb[32:].any()
        bb[:] = 0
        dm.dm_ml_counters_generic(_p(words), n, 16 if n >= 16 else 8, p + 1, _p(bb))
        assert np.array_equal(bb[:32], exp), (kind, "generic path")


def test_hll_recoded_registers_are_high_words_of_powers_of_two(dm):
    """K4h: v = 0x3FF00000 - (r << 20) is the high word of the double 2^-r (low word 0), so for every pair of register
    bytes hiloint2double(min(va, vb), 0) == 2^-max(ra, rb); has_zero_byte is exact on every kind of word."""
    dm.dm_hll_recode.argtypes = [C.c_uint32, C.c_void_p, C.c_void_p]
    out, z = np.zeros(4, dtype=np.uint32), C.c_int(0)
    v = np.zeros(256, dtype=np.uint32)
    for r in range(0, 256, 4):
        w = r | ((r + 1) << 8) | ((r + 2) << 16) | ((r + 3) << 24)
        dm.dm_hll_recode(w, _p(out), C.byref(z))
        v[r:r + 4] = out
        assert bool(z.value) == (r == 0)
    as_double = (v.astype(np.uint64) << np.uint64(32)).view(np.float64)
    assert np.array_equal(as_double, 2.0 ** -np.arange(256, dtype=np.float64))
    a, b = np.meshgrid(np.arange(256), np.arange(256))
    m = np.minimum(v[a], v[b])
    assert np.array_equal((m.astype(np.uint64) << np.uint64(32)).view(np.float64), 2.0 ** -np.maximum(a, b).astype(np.float64))
    rng = np.random.default_rng(0)
    for w in list(rng.integers(0, 1 << 32, size=2000)) + [0, 0x01010101, 0x80808080, 0x00ffffff, 0xff00ffff, 0x7f7f7f00, 0x01000101]:
        dm.dm_hll_recode(int(w), _p(out), C.byref(z))
        assert bool(z.value) == any(((int(w) >> (8 * i)) & 0xff) == 0 for i in range(4)), hex(int(w))


def test_hll_fixed_point_terms_and_exact_pair_sums(dm):
    """K4i (dist_tables.cuh: hll_int_recode): v(r) = 2^(28 - (r - lo)) inside the window, 0 above it; min of two terms is the
    term of the larger register; and for registers inside the window the integer sum, scaled, equals the reference's
    sequential f64 loop bit for bit (no partial sum of that loop rounds) -- including batches that hold eight copies of the
    largest term, and the both-empty count when lo = 0."""
    dm.dm_hll_int_recode.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p]
    dm.dm_hll_int_pair_sum.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_void_p]
    dm.dm_hll_int_pair_sum.restype = C.c_double
    out = np.zeros(4, dtype=np.uint32)
    for lo in (0, 1, 5, 23, 200):
        v = np.zeros(256, dtype=np.uint64)
        for r in range(lo - lo % 4, 256, 4):
            rr = [max(r + i, lo) for i in range(4)]             # bytes below lo never occur (lo is the minimum)
            w = rr[0] | (rr[1] << 8) | (rr[2] << 16) | (rr[3] << 24)
            dm.dm_hll_int_recode(w, lo, _p(out))
            for i in range(4):
                v[rr[i]] = out[i]
        for r in range(lo, 256):
            assert int(v[r]) == ((1 << (28 - (r - lo))) if r - lo <= 28 else 0), (lo, r, int(v[r]))
        a, b = np.meshgrid(np.arange(lo, 256), np.arange(lo, 256))
        assert np.array_equal(np.minimum(v[a], v[b]), v[np.maximum(a, b)])
    rng = np.random.default_rng(11)
    z = C.c_uint32(0)
    for p, lo, shape in ((14, 4, "genome"), (14, 0, "small"), (18, 2, "genome"), (10, 7, "flat"), (14, 3, "all-lo")):
        m = 1 << p
        for _ in range(3):
            if shape == "genome":
                regs = np.clip(np.floor(lo + 4.5 - np.log2(-np.log(rng.random((2, m))))), lo, lo + 28).astype(np.uint8)
            elif shape == "small":
                regs = (rng.random((2, m)) < 0.3) * np.clip(rng.geometric(0.5, size=(2, m)), 1, 28).astype(np.uint8)
            elif shape == "flat":
                regs = rng.integers(lo, lo + 29, size=(2, m)).astype(np.uint8)
            else:
                regs = np.full((2, m), lo, dtype=np.uint8)
            regs = np.ascontiguousarray(regs, dtype=np.uint8)
            assert regs.min() >= lo
            regs[0, 0] = lo                                        # the window is anchored at the smallest register present
            got = dm.dm_hll_int_pair_sum(_p(regs[0]), _p(regs[1]), m, lo, C.byref(z))
            mx = np.maximum(regs[0], regs[1])
            seq = 0.0
            for t in (2.0 ** -mx.astype(np.float64)):             # the reference's loop: sequential f64 adds in register order
                seq += t
            assert got == seq, (p, lo, shape, got, seq)
            if lo == 0:
                assert z.value == int((mx == 0).sum())


@pytest.mark.parametrize("p,k,drop4", [(10, 16, 1), (10, 12, 1), (14, 21, 0), (8, 31, 1), (12, 16, 0)])
def test_cpu_emulation_of_the_ull_sketch_path_end_to_end(dm, oracle, p, k, drop4):
    """The arithmetic of the sketch kernel's ULL path chained on the CPU, piece by piece as the kernel does it:
    funnel-shift canonical k-mers -> pre-xorshift hash, high word only -> fast (index, bit) or "rare" -> exact rule ->
    OR into the two-word cell -> conversion at flush.  The registers must equal the oracle's sketch of the same genome
    (the GPU tests check the same end to end on the device; this one runs without one)."""
    from lash_b200 import hostapi
    from tools import synth
    dm.dm_kmers.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    dm.dm_ull_cell_to_reg.argtypes, dm.dm_ull_cell_to_reg.restype = [C.c_uint32, C.c_uint32, C.c_int], C.c_uint32
    seq = synth.genomes(1, 120_000, seed=p * 100 + k)[0][0]
    packed, nb = hostapi.filter_pack(seq, simd=0)
    buf = np.concatenate([packed, np.zeros(16, dtype=np.uint8)])
    kmers = np.zeros(len(seq) - k + 1, dtype=np.uint64)
    dm.dm_kmers(_p(buf), len(packed), len(seq), k, _p(kmers))
    narrow = 1 if k <= 16 else 0
    g = np.empty_like(kmers)
    ghi = np.empty(len(kmers), dtype=np.uint32)
    dm.dm_pre(_p(kmers), len(kmers), SEED, narrow, _p(g), _p(ghi))
    idx, nlz, rare = (np.empty(len(kmers), dtype=np.uint32) for _ in range(3))
    dm.dm_ull_fast(_p(ghi), len(kmers), p, drop4, _p(idx), _p(nlz), _p(rare))
    # rare hashes go through the exact rule on the finished hash
    h = np.empty_like(kmers)
    dm.dm_xxh3_64(_p(kmers), len(kmers), SEED, _p(h))
    for i in np.flatnonzero(rare):
        body = ((int(h[i]) << p) & ((1 << 64) - 1)) | ((1 << p) - 1)
        nlz[i] = 64 - body.bit_length()
        idx[i] = int(h[i]) >> (64 - p)
    # the cell: word 0 bit j <=> nlz = 31 - j, word 1 bit j <=> nlz = 63 - j
    w = np.zeros((1 << p, 2), dtype=np.uint32)
    word = (nlz >= 32).astype(np.int64)
    bit = np.where(nlz >= 32, 63 - nlz.astype(np.int64), 31 - nlz.astype(np.int64))
    np.bitwise_or.at(w, (idx.astype(np.int64), word), (np.uint32(1) << bit.astype(np.uint32)))
    regs = np.array([dm.dm_ull_cell_to_reg(int(a), int(b), p) for a, b in w], dtype=np.uint8)
    exp = oracle.sketch_genomes(2, p, k, SEED, [[seq]])[0]
    assert np.array_equal(regs, exp)
    assert (regs != 0).mean() > 0.9


@pytest.mark.parametrize("p,k", [(14, 21), (10, 16), (4, 12), (16, 31)])
def test_cpu_emulation_of_the_hll_sketch_path_end_to_end(dm, oracle, p, k):
    """Same for HLL: index from the low p bits of the finished hash's low word, rho from the PRE-xorshift high word
    (the xorshift never moves its highest set bit), max into the register."""
    from lash_b200 import hostapi
    from tools import synth
    dm.dm_kmers.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p]
    dm.dm_hll_fast.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    seq = synth.genomes(1, 100_000, seed=p * 100 + k)[0][0]
    packed, nb = hostapi.filter_pack(seq, simd=0)
    buf = np.concatenate([packed, np.zeros(16, dtype=np.uint8)])
    kmers = np.zeros(len(seq) - k + 1, dtype=np.uint64)
    dm.dm_kmers(_p(buf), len(packed), len(seq), k, _p(kmers))
    g = np.empty_like(kmers)
    ghi = np.empty(len(kmers), dtype=np.uint32)
    dm.dm_pre(_p(kmers), len(kmers), SEED, 1 if k <= 16 else 0, _p(g), _p(ghi))
    # adversarial additions: pre-xorshift words with a zero / tiny high word
    g = np.concatenate([g, np.array([0x00000000_12345678, 0x00000007_9abcdef0, 0x0000000f_00000000, 0xf0000000_00000001], dtype=np.uint64)])
    idx, rho, rare = (np.empty(len(g), dtype=np.uint32) for _ in range(3))
    dm.dm_hll_fast(_p(g), len(g), p, _p(idx), _p(rho), _p(rare))
    h = g ^ (g >> U64(28))
    exp_idx = (h & U64((1 << p) - 1)).astype(np.uint32)
    wv = h >> U64(p)                                                      # streaming_algorithms: rho = clz64(w) - p + 1
    exp_rho = np.array([(64 - int(x).bit_length()) - p + 1 for x in wv], dtype=np.uint32)
    assert np.array_equal(idx, exp_idx)
    assert np.array_equal(rare == 1, (h >> U64(32)) == 0)
    assert np.array_equal(rho[rare == 0], exp_rho[rare == 0])
    n = len(kmers)
    regs = np.zeros(1 << p, dtype=np.uint32)
    np.maximum.at(regs, exp_idx[:n].astype(np.int64), np.where(rare[:n] == 1, exp_rho[:n], rho[:n]))
    assert np.array_equal(regs.astype(np.uint8), oracle.sketch_genomes(1, p, k, SEED, [[seq]])[0])


def test_cpu_emulation_of_the_pair_table_distance_paths(dm, est, oracle):
    """K4b / K4c chained on the CPU: registers -> codes relative to the smallest register present -> table lookups in
    register order (FGRA: running f64 sum; ML: wrapping S and bit counts) -> estimator epilogue -> Jaccard -> frac.
    With the host libm on both sides the result is the oracle's to the last bit."""
    from tools import synth
    dm.dm_fgra_code.argtypes, dm.dm_fgra_code.restype = [C.c_uint32, C.c_uint32], C.c_uint32
    dm.dm_fgra_table.argtypes, dm.dm_fgra_table.restype = [C.c_uint32, C.c_int, C.c_void_p, C.c_void_p], C.c_double
    dm.dm_ml_tables.argtypes = [C.c_int, C.c_void_p, C.c_void_p]
    p, k = 10, 16
    regs = oracle.sketch_genomes(2, p, k, SEED, synth.genomes(5, 150_000, seed=3))
    reg_tab = np.array([est.dm_ull_reg(i) for i in range(256)], dtype=np.float64)
    off = 4 * p + 4
    base = max(4 * p - 4, int(regs[regs != 0].min()) & ~3)               # regmin_kernel + the kernel's anchoring
    T = np.empty(128 * 128, dtype=np.float64)
    sentinel = dm.dm_fgra_table(base, p, _p(reg_tab), _p(T))
    R, W = np.empty(128 * 128, dtype=np.uint64), np.empty(128 * 128, dtype=np.uint32)
    dm.dm_ml_tables(p, _p(R), _p(W))
    code = np.vectorize(lambda r: dm.dm_fgra_code(int(r), base))(regs)
    code_ml = np.vectorize(lambda r: dm.dm_fgra_code(int(r), 4 * p - 4))(regs)
    assert code.max() < 127 and code_ml.max() < 127
    zeros8 = np.zeros(8, dtype=np.uint32)
    card_f = [oracle.cardinality(2, p, 0, r) for r in regs]
    card_m = [oracle.cardinality(2, p, 1, r) for r in regs]
    frac_f = oracle.dist(2, p, k, 0, 2, False, regs, regs)
    frac_m = oracle.dist(2, p, k, 1, 2, False, regs, regs)
    for i in range(len(regs)):
        for j in range(len(regs)):
            vals = T[code[i].astype(np.int64) * 128 + code[j]]
            assert vals.max() < sentinel
            total = float(np.cumsum(vals)[-1])                              # register order, one running sum
            U = est.dm_fgra(total, _p(zeros8), p)
            s = max((card_f[i] + card_f[j] - U) / U, 0.0)
            assert 2.0 * s / (1.0 + s) == frac_f[i, j], ("fgra", i, j)
            e = code_ml[i].astype(np.int64) * 128 + code_ml[j]
            S = int(np.sum(R[e].astype(object)) % (1 << 64))
            b = np.zeros(66, dtype=np.int32)
            for bit in range(32):
                b[bit] = int(((W[e] >> np.uint32(bit)) & np.uint32(1)).sum())
            merged0 = int(oracle.ull_merge(regs[i], regs[j], p)[0])
            U = est.dm_ml(S, _p(b), p, merged0)
            s = max((card_m[i] + card_m[j] - U) / U, 0.0)
            assert 2.0 * s / (1.0 + s) == frac_m[i, j], ("ml", i, j)


# ---- ddmath.cuh: the correctly rounded pow / log / log1p of the per-pair epilogues ----------------------------------
DD_SRC = os.path.join(ROOT, "tests", "host_shim", "ddmath.cpp")


@pytest.fixture(scope="module")
def ddlib(tmp_path_factory):
    out = str(tmp_path_factory.mktemp("ddmath") / "libddmath.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-shared", "-fPIC", "-o", out, DD_SRC])
    L = C.CDLL(out)
    for name, args in (("dm_pow_cr", [C.c_double, C.c_double]), ("dm_log_cr", [C.c_double]), ("dm_log1p_cr", [C.c_double])):
        getattr(L, name).restype = C.c_double
        getattr(L, name).argtypes = args
    for name in ("dm_pow_cr_many", "dm_pow_libm_many"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_long]
    for name in ("dm_log_cr_many", "dm_log_libm_many"):
        getattr(L, name).argtypes = [C.c_void_p, C.c_void_p, C.c_long]
    return L


def test_ddmath_pow_log_log1p_are_correctly_rounded(ddlib):
    """pow_cr / log_cr / log1p_cr (double-double, rounded once) against mpmath at 400 bits rounded to nearest: equal on every
    argument, over the ranges the epilogues see (FGRA sums 1e-3..1e4 with exponent -1/tau, frac in (0, 1] with 1/k, arguments
    next to 1, tiny log1p arguments where the second-order term decides the rounding)."""
    import mpmath as mp
    mp.mp.prec = 400
    rng = np.random.default_rng(1)
    y0 = -1.0 / 0.8194911375910897
    xs = np.concatenate([10 ** rng.uniform(-3, 4, 1500), rng.uniform(0.5, 2.0, 500), 2.0 ** rng.integers(-20, 20, 30).astype(float),
                         1 + rng.uniform(-1e-6, 1e-6, 200)])
    for x in xs:
        for y in (y0, 1 / 16.0, 1 / 21.0, float(rng.uniform(-3, 3))):
            assert ddlib.dm_pow_cr(float(x), y) == float(mp.power(mp.mpf(float(x)), mp.mpf(y))), (x, y)
    for x in np.concatenate([10 ** rng.uniform(-12, 0, 2000), rng.uniform(0.9, 1.0, 800), 1 - 10 ** rng.uniform(-16, -1, 600), 10 ** rng.uniform(0, 300, 100)]):
        assert ddlib.dm_log_cr(float(x)) == float(mp.log(mp.mpf(float(x)))), x
    for x in np.concatenate([10 ** rng.uniform(-20, 6, 2000), -10 ** rng.uniform(-20, -0.01, 1000), rng.uniform(0.5e-16, 5e-16, 800),
                             -rng.uniform(0.5e-16, 5e-16, 800)]):
        assert ddlib.dm_log1p_cr(float(x)) == float(mp.log1p(mp.mpf(float(x)))), x
    # the cases the library handles: signs of zero, infinities, NaN, non-positive arguments
    assert ddlib.dm_log_cr(1.0) == 0.0 and not np.signbit(ddlib.dm_log_cr(1.0)) and ddlib.dm_log_cr(0.0) == -np.inf
    assert np.isnan(ddlib.dm_log_cr(-1.0)) and np.isnan(ddlib.dm_log_cr(np.nan)) and ddlib.dm_pow_cr(0.0, 1 / 16.0) == 0.0
    assert ddlib.dm_pow_cr(4.0, 0.5) == 2.0 and ddlib.dm_pow_cr(1.0, 123.456) == 1.0 and np.isnan(ddlib.dm_pow_cr(np.nan, 2.0))


def test_ddmath_agrees_with_glibc_except_where_glibc_is_not_correctly_rounded(ddlib):
    """Why the epilogues use it: glibc's pow / log (what the oracle calls) are within 0.52 ulp, so they equal the correctly
    rounded value in all but ~0.1 % / ~0.01 % of the arguments, and never differ by more than one ulp."""
    rng = np.random.default_rng(2)
    x = np.ascontiguousarray(10 ** rng.uniform(-2, 3.5, 400_000))
    a, b = np.empty_like(x), np.empty_like(x)
    ddlib.dm_pow_cr_many(x.ctypes.data, -1.0 / 0.8194911375910897, a.ctypes.data, len(x))
    ddlib.dm_pow_libm_many(x.ctypes.data, -1.0 / 0.8194911375910897, b.ctypes.data, len(x))
    assert (a != b).mean() < 0.003 and np.max(np.abs(a - b) / np.spacing(b)) <= 1.0
    x = np.ascontiguousarray(10 ** rng.uniform(-8, 0, 400_000))
    ddlib.dm_log_cr_many(x.ctypes.data, a.ctypes.data, len(x))
    ddlib.dm_log_libm_many(x.ctypes.data, b.ctypes.data, len(x))
    assert (a != b).mean() < 0.001 and np.max(np.abs(a - b) / np.spacing(np.abs(b))) <= 1.0
