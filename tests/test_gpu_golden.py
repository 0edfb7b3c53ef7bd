"""GPU path against the COMMITTED golden fixtures (tests/golden/, made by tools/make_golden.py from an
independent pure-Python restatement + python-xxhash) -- no oracle in the loop -- and, at BASELINE.json's
full size, through size-independent properties of the domain (the oracle cannot run 5 Gbp in seconds):
split-invariance, idempotence, merge-of-shares, symmetry and the exact zero diagonal."""
import json
import os

import numpy as np
import pytest

from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, MODEL_POISSON, ops
from lash_b200.capi import Span, check, lib
from lash_b200.ops import Sketcher, sketch_genomes

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ALGO = {"hmh": ALGO_HMH, "hll": ALGO_HLL, "ull": ALGO_ULL}


def test_registers_equal_the_golden_fixture(gpu_ctx):
    for case in json.load(open(os.path.join(GOLD, "sketch_py.json"))):
        recs = [r.encode() for r in case["records"]]
        got = sketch_genomes(gpu_ctx, ALGO[case["algo"]], case["p"], case["k"], case["seed"], [recs])[0]
        exp = np.zeros(case["n_regs"], dtype=got.dtype)
        for i, v in case["nonzero"].items():
            exp[int(i)] = v
        assert np.array_equal(got, exp), (case["algo"], case["p"], case["k"])


def test_fgra_constants_on_device_match_hash4j_literals(gpu_ctx):
    """A sketch with ONE distinct register value r has FGRA estimate factor(p) * (m * contribution[r])^(-1/tau):
    the device tables must reproduce hash4j's literal table entries / factors (tests/golden/ull_constants.json)."""
    c = json.load(open(os.path.join(GOLD, "ull_constants.json")))
    tau = 0.8194911375910897
    for p_str, factor in c["estimation_factors"].items():
        p = int(p_str)
        m = 1 << p
        for idx, contrib in enumerate(c["register_contributions_0_5"]):
            r = idx + 4 * p + 4
            regs = np.full((1, m), r, dtype=np.uint8)
            got = ops.cardinality(gpu_ctx, ALGO_ULL, p, EST_FGRA, regs)[0]
            exp = factor * (m * contrib) ** (-1.0 / tau)
            assert abs(got - exp) <= 1e-12 * exp, (p, r, got, exp)


def test_full_size_config2_properties(gpu_ctx):
    """BASELINE configs[1] at full size: 1000 x 5 Mbp, ULL p=10 k=16 seed 42, FGRA all-vs-all.
    (a) one push == 25 pushes of 40 genomes == pushing everything twice (split-invariance, idempotence);
    (b) sketching each genome in two halves that overlap by k-1 bases and folding with lash_sketch_merge
        equals the sketch of the whole (every k-mer start exactly once);
    (c) the distance matrix is symmetric, its packed triangle equals the dense lower triangle bit for bit,
        and identical sketches are at distance exactly 0 (-ln(1)/k)."""
    import torch

    import bench
    n_g, length, p, k = 1000, 5_000_000, 10, 16
    dev = torch.device("cuda", 0)
    buf, stride = bench.make_packed_genomes(torch, dev, n_g, length, 42, 0)
    spans = (Span * n_g)()
    for i in range(n_g):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    with Sketcher(gpu_ctx, ALGO_ULL, p, k, 42, n_g) as sk:
        sk.push_raw(buf.data_ptr(), n_g * stride, spans, n_g, None, 0, dev=True)
        one = sk.fetch()
        sk.reset()
        for g0 in range(0, n_g, 40):                       # 25 pushes, then everything once more
            sub = (Span * 40)(*[Span(g0 + i, i * stride, length, 0, 1, 0) for i in range(40)])
            sk.push_raw(buf.data_ptr() + g0 * stride, 40 * stride, sub, 40, None, 0, dev=True)
        sk.push_raw(buf.data_ptr(), n_g * stride, spans, n_g, None, 0, dev=True)
        many = sk.fetch()
    assert np.array_equal(one, many)
    assert (one != 0).all()                                # 5 M k-mers over 1024 registers: none stays empty
    # (b) halves: bases [0, h + k - 1) and [h, length), h a multiple of 64 so that the second half starts on a
    # 16-byte boundary of the packed buffer
    h = (length // 2) // 64 * 64
    with Sketcher(gpu_ctx, ALGO_ULL, p, k, 42, n_g) as a, Sketcher(gpu_ctx, ALGO_ULL, p, k, 42, n_g) as b:
        first = (Span * n_g)(*[Span(i, i * stride, h + k - 1, 0, 1, 0) for i in range(n_g)])
        second = (Span * n_g)(*[Span(i, i * stride + h // 4, length - h, 0, 1, 0) for i in range(n_g)])
        a.push_raw(buf.data_ptr(), n_g * stride, first, n_g, None, 0, dev=True)
        b.push_raw(buf.data_ptr(), n_g * stride, second, n_g, None, 0, dev=True)
        ra, rb = a.fetch(), b.fetch()
    assert not np.array_equal(ra, one)
    assert np.array_equal(ops.merge(gpu_ctx, ALGO_ULL, p, ra, rb), one)
    # (c)
    for est in (EST_FGRA, EST_ML):
        dense, _ = ops.dist(gpu_ctx, ALGO_ULL, p, k, est, MODEL_POISSON, False, one[:600], one[:600])
        tri, _ = ops.dist(gpu_ctx, ALGO_ULL, p, k, est, MODEL_POISSON, False, one[:600], one[:600], triangular=True)
        assert np.array_equal(tri, dense[np.tril_indices(600)])
        assert np.array_equal(dense, dense.T)
        assert np.all(np.diag(dense) == 0.0) and np.all((dense >= 0.0) & (dense <= 1.0))
        assert dense[np.triu_indices(600, 1)].min() > 0.0  # distinct genomes are never at distance 0
