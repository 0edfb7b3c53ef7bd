"""The CPU oracle against every fixed point available offline (-m "not gpu").

The reference holds no tests, fixtures or golden vectors, so the oracle is
"parity unpinned" against lash itself; what CAN be pinned is pinned here:
  * XXH3 against python-xxhash vectors (tests/golden/xxh3_kat.json) -- exact;
  * canonical k-mers against brute-force string code (tests/golden/kmers.json) -- exact;
  * register updates against an independent pure-Python restatement (tests/golden/sketch_py.json);
  * ULL constants against hash4j's literal table entries (tests/golden/ull_constants.json);
  * every estimator against a second pure-Python restatement (tests/golden/estimators.json, tools/estimators_py.py);
  * estimator unbiasedness by simulation, merge/union algebra, quirks of the reference front end.
"""
import ctypes as C
import json
import os

import numpy as np
import pytest

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


def test_xxh3_known_answers(oracle):
    kat = _gold("xxh3_kat.json")
    assert len(kat["xxh3_64_le64"]) >= 500
    for v, s, h in kat["xxh3_64_le64"]:
        assert oracle.xxh3_64_le64(int(v), int(s)) == int(h)
    for w, s, h in kat["xxh3_128_le32"]:
        assert oracle.xxh3_128_le32(int(w), int(s)) == int(h)
    c = _gold("ull_constants.json")["xxh3_kat_from_survey"]
    assert oracle.xxh3_64_le64(0, 42) == int(c["xxh3_64(le64(0),42)"], 16)
    assert oracle.xxh3_128_le32(0, 42) == int(c["xxh3_128(le32(0),42)"], 16)


def test_canonical_kmers_golden(oracle):
    for case in _gold("kmers.json"):
        got = oracle.canonical_kmers(case["seq"].encode(), case["k"])
        assert [str(int(x)) for x in got] == case["kmers"], case["k"]


def test_filter_and_mask_quirks(oracle):
    # utils.rs:33-41: lowercase and everything else is DELETED, flanks joined
    assert oracle.filter_out_n(b"ACgtNNRYGT\nAC-*T") == b"ACGTACT"
    assert oracle.filter_out_n(b"") == b""
    L = oracle.lib()
    assert L.lo_mask_bits(2**64 - 1, 32) == 2**64 - 1       # utils.rs:59-60
    assert L.lo_mask_bits(2**64 - 1, 16) == 2**32 - 1
    assert L.lo_mask_bits(0xE0000123, 14) == 0x0000123       # strips Kmer32bit's length nibble
    # records shorter than k yield nothing (utils.rs:460-462)
    assert len(oracle.canonical_kmers(b"ACGTACG", 8)) == 0
    assert len(oracle.canonical_kmers(b"ACGTACGT", 8)) == 1


def test_registers_match_independent_python_restatement(oracle):
    algo_id = {"hmh": oracle.HMH, "hll": oracle.HLL, "ull": oracle.ULL}
    for case in _gold("sketch_py.json"):
        regs = oracle.sketch_genomes(algo_id[case["algo"]], case["p"], case["k"], case["seed"],
                                     [[r.encode() for r in case["records"]]])[0]
        assert len(regs) == case["n_regs"]
        exp = np.zeros(case["n_regs"], dtype=regs.dtype)
        for i, v in case["nonzero"].items():
            exp[int(i)] = v
        assert np.array_equal(regs, exp), case["algo"]


def test_ull_constants_match_hash4j_literals(oracle):
    c = _gold("ull_constants.json")
    tab = oracle.lib().lo_ull_register_contributions()
    for i, v in enumerate(c["register_contributions_0_5"]):
        assert tab[i] == pytest.approx(v, rel=2e-16), i
    for p, v in c["estimation_factors"].items():
        assert oracle.lib().lo_ull_estimation_factor(int(p)) == pytest.approx(v, rel=4e-16)


def test_ull_sequential_add_equals_pack_of_or_and_is_not_max(oracle):
    """ultraloglog add() is order free and merge is pack(unpack|unpack), not a byte max."""
    L = oracle.lib()
    L.lo_ull_add.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    L.lo_ull_add.restype = None
    rng = np.random.default_rng(0)
    p = 6
    hs = rng.integers(0, 2**64, size=4000, dtype=np.uint64)
    a = np.zeros(1 << p, dtype=np.uint8)
    b = np.zeros(1 << p, dtype=np.uint8)
    ab = np.zeros(1 << p, dtype=np.uint8)
    ba = np.zeros(1 << p, dtype=np.uint8)
    for h in hs[:2000]:
        L.lo_ull_add(a.ctypes.data, p, int(h))
        L.lo_ull_add(ab.ctypes.data, p, int(h))
    for h in hs[2000:]:
        L.lo_ull_add(b.ctypes.data, p, int(h))
        L.lo_ull_add(ab.ctypes.data, p, int(h))
    for h in hs[::-1]:
        L.lo_ull_add(ba.ctypes.data, p, int(h))
    assert np.array_equal(ab, ba)                                   # order free
    assert np.array_equal(oracle.ull_merge(a, b, p), ab)            # merge == sketch of the union
    assert np.array_equal(oracle.ull_merge(a, a, p), a)             # idempotent
    assert np.array_equal(oracle.ull_merge(a, np.zeros_like(a), p), a)
    # there are inputs where a byte max is wrong
    x = np.zeros(8, dtype=np.uint8)
    y = np.zeros(8, dtype=np.uint8)
    x[0] = 4 * 10          # top bit 10, nothing below
    y[0] = 4 * 9           # top bit 9
    assert oracle.ull_merge(x, y, 3)[0] == 4 * 10 + 2 != max(x[0], y[0])


def _random_sketch(oracle, algo, p, n, rng):
    L = oracle.lib()
    regs = np.zeros(1 << p, dtype=np.uint8)
    fn = L.lo_ull_add if algo == oracle.ULL else L.lo_hll_push_hash64
    fn.argtypes = [C.c_void_p, C.c_int, C.c_uint64]
    fn.restype = None
    for h in rng.integers(0, 2**64, size=n, dtype=np.uint64):
        fn(regs.ctypes.data, p, int(h))
    return regs


@pytest.mark.parametrize("algo_name,p,est,rse", [("ULL", 8, 0, 0.0489), ("ULL", 8, 1, 0.0476), ("HLL", 8, 0, 0.065)])
def test_estimators_unbiased_on_random_hashes(oracle, algo_name, p, est, rse):
    """FGRA rse = sqrt(0.6119/m); ML slightly better; HLL ~ 1.04/sqrt(m)."""
    algo = getattr(oracle, algo_name)
    rng = np.random.default_rng(11)
    n = 20000
    errs = np.array([oracle.cardinality(algo, p, est, _random_sketch(oracle, algo, p, n, rng)) / n - 1 for _ in range(24)])
    assert abs(errs.mean()) < 3 * rse / np.sqrt(len(errs)) + 0.005
    assert 0.5 * rse < errs.std() < 1.6 * rse


def test_small_range_estimates(oracle):
    rng = np.random.default_rng(5)
    for n in (1, 5, 50, 500):
        for est in (0, 1):
            e = np.mean([oracle.cardinality(oracle.ULL, 10, est, _random_sketch(oracle, oracle.ULL, 10, n, rng)) for _ in range(8)])
            assert abs(e / n - 1) < 0.12, (n, est, e)
        e = np.mean([oracle.cardinality(oracle.HLL, 10, 0, _random_sketch(oracle, oracle.HLL, 10, n, rng)) for _ in range(8)])
        assert abs(e / n - 1) < 0.12, (n, e)
    empty = np.zeros(1024, dtype=np.uint8)
    assert oracle.cardinality(oracle.ULL, 10, 0, empty) == 0.0
    assert oracle.cardinality(oracle.ULL, 10, 1, empty) == 0.0
    assert oracle.cardinality(oracle.HLL, 10, 0, empty) == 0.0


def test_ml_stats_are_bit_counts_of_the_unpacked_prefix(oracle):
    """b[j] = number of registers whose unpacked prefix has bit j+p-1 set -- the identity the GPU
    kernel's bit-sliced counters rely on."""
    rng = np.random.default_rng(2)
    p = 7
    regs = _random_sketch(oracle, oracle.ULL, p, 3000, rng)
    regs[:5] = [0, 4 * p - 4, 4 * p, 4 * p + 2, 4 * p + 4]
    S, b = oracle.ull_ml_stats(regs, p)
    exp = np.zeros(66, dtype=np.int64)
    for r in regs:
        r = int(r)
        hp = ((4 | (r & 3)) << (((r >> 2) - 2) & 63)) & (2**64 - 1) if r else 0
        w = hp >> (p - 1)
        for j in range(64):
            exp[j] += (w >> j) & 1
    assert np.array_equal(b[:64], exp[:64])


def test_distance_formula_and_dist_driver(oracle):
    L = oracle.lib()
    assert L.lo_compute_distance_f64(1.0, 16, 1) == 0.0
    assert L.lo_compute_distance_f64(0.0, 16, 1) == 1.0       # min(1, +inf)
    assert L.lo_compute_distance_f64(0.0, 16, 0) == 1.0
    assert L.lo_compute_distance_f64(0.5, 16, 1) == pytest.approx(np.log(2) / 16)
    assert L.lo_compute_distance_f64(0.5, 21, 0) == pytest.approx(1 - 0.5 ** (1 / 21))
    from tools import synth
    gs = synth.genomes(6, 60_000)
    for algo, p, est in ((oracle.ULL, 10, 0), (oracle.ULL, 10, 1), (oracle.HLL, 10, 0), (oracle.HMH, 14, 0)):
        regs = oracle.sketch_genomes(algo, p, 16, 42, gs, threads=4)
        d = oracle.dist(algo, p, 16, est, 1, False, regs, regs, threads=4)
        assert np.allclose(d, d.T, rtol=1e-12) and np.all(np.diag(d) < 1e-3)
        t = oracle.dist(algo, p, 16, est, 1, False, regs, regs, triangular=True, threads=2)
        assert np.isnan(t[np.triu_indices(6, 1)]).all()
        assert np.array_equal(t[np.tril_indices(6)], d[np.tril_indices(6)])
        d32 = oracle.dist(algo, p, 16, est, 1, True, regs, regs, threads=4)
        assert d32.dtype == np.float32 and np.allclose(d32, d, rtol=1e-5, atol=1e-6)
    # threads do not change results
    regs = oracle.sketch_genomes(oracle.ULL, 10, 16, 42, gs, threads=1)
    assert np.array_equal(regs, oracle.sketch_genomes(oracle.ULL, 10, 16, 42, gs, threads=5))


def test_estimators_match_the_independent_python_restatement(oracle):
    """tests/golden/estimators.json: registers -> cardinalities, union estimate, frac and distances computed by
    tools/estimators_py.py, a pure-Python restatement written from the published algorithms, not
    from the oracle: ULL FGRA incl. small-range and saturated registers, ULL ML (Ertl's solver), ULL merge, HLL++ len()
    incl. linear counting and the flagged bias regime, HyperMinHash cardinality / similarity (the n <= 2^19 double loop),
    Jaccard -> frac -> Mash distance.  Same libm on both sides: agreement to a few ulp."""
    eps = np.finfo(np.float64).eps

    def close(got, exp_repr, ulps=16):
        exp = float(exp_repr)
        if np.isnan(exp):
            return bool(np.isnan(got))
        if np.isinf(exp) or exp == 0.0:
            return got == exp
        return abs(got - exp) <= ulps * eps * abs(exp)

    cases = _gold("estimators.json")
    assert len(cases) >= 30
    seen = set()
    for c in cases:
        p = c["p"]
        if c["algo"] == "ull":
            est = 0 if c["estimator"] == "fgra" else 1
            a, b = np.array(c["a"], dtype=np.uint8), np.array(c["b"], dtype=np.uint8)
            assert list(oracle.ull_merge(a, b, p)) == c["merged"]
            assert close(oracle.cardinality(oracle.ULL, p, est, a), c["card_a"]), (c["estimator"], p, "card_a")
            assert close(oracle.cardinality(oracle.ULL, p, est, b), c["card_b"])
            assert close(oracle.cardinality(oracle.ULL, p, est, np.array(c["merged"], dtype=np.uint8)), c["union"])
            fr = oracle.dist(oracle.ULL, p, 16, est, 2, False, a[None, :], b[None, :])[0, 0]
            # frac = 2s/(1+s) with s = (a+b-U)/U: the subtraction amplifies the few-ulp differences of the estimates
            exp_fr = float(c["frac"])
            assert (np.isnan(fr) and np.isnan(exp_fr)) or abs(fr - exp_fr) <= 1e-12 * max(abs(exp_fr), 1e-3), (p, fr, exp_fr)
            d1 = oracle.dist(oracle.ULL, p, 16, est, 1, False, a[None, :], b[None, :])[0, 0]
            d0 = oracle.dist(oracle.ULL, p, 21, est, 0, False, a[None, :], b[None, :])[0, 0]
            for got, key in ((d1, "d_poisson_k16"), (d0, "d_binomial_k21")):
                exp = float(c[key])
                assert (np.isnan(got) and np.isnan(exp)) or abs(got - exp) <= 1e-11, (key, got, exp)
            seen.add(("ull", c["estimator"], any(r >= 252 for r in c["a"]), any(0 < r < 4 * p + 4 for r in c["a"])))
        elif c["algo"] == "hll":
            a, b = np.array(c["a"], dtype=np.uint8), np.array(c["b"], dtype=np.uint8)
            assert close(oracle.cardinality(oracle.HLL, p, 0, a), c["card_a"])
            assert close(oracle.cardinality(oracle.HLL, p, 0, b), c["card_b"])
            fr, flags = oracle.dist(oracle.HLL, p, 16, 0, 2, False, a[None, :], b[None, :], return_flags=True)
            assert bool(flags[0, 0]) == bool(c["bias_regime"][2])
            if not any(c["bias_regime"]):
                exp_fr = float(c["frac"])
                assert abs(fr[0, 0] - exp_fr) <= 1e-12 * max(abs(exp_fr), 1e-3)
            seen.add(("hll", any(c["bias_regime"]), 0 in c["a"]))
        else:
            a, b = np.zeros(16384, dtype=np.uint16), np.zeros(16384, dtype=np.uint16)
            for i, v in c["a"].items():
                a[int(i)] = v
            for i, v in c["b"].items():
                b[int(i)] = v
            assert close(oracle.cardinality(oracle.HMH, 14, 0, a), c["card_a"])
            assert close(oracle.cardinality(oracle.HMH, 14, 0, b), c["card_b"])
            fr = oracle.dist(oracle.HMH, 14, 16, 0, 2, False, a[None, :], b[None, :])[0, 0]
            assert abs(fr - float(c["frac"])) <= 1e-12 and fr > 0.1
            seen.add(("hmh",))
    # the fixture exercises every branch it is meant to
    assert ("ull", "fgra", True, False) in seen or ("ull", "fgra", True, True) in seen       # saturated registers
    assert any(s[0] == "ull" and s[3] for s in seen)                                         # small-range registers
    assert ("hll", True, True) in seen or ("hll", True, False) in seen                       # bias regime flagged
    assert any(s[0] == "hll" and not s[1] and s[2] for s in seen)                            # linear counting
