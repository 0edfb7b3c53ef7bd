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
