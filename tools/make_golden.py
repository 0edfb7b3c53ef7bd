"""Generates tests/golden/*.json.  Run in the build container (needs python-xxhash 3.7 = libxxhash 0.8.2):
    python -m tools.make_golden
Sources of truth, none of them the oracle:
  xxh3_kat.json      python-xxhash (independent C implementation of the frozen XXH3 spec)
  kmers.json         brute-force string slicing / reverse complement in Python
  ull_constants.json hash4j/ultraloglog literal table entries as recorded in SURVEY.md A.4
  sketch_py.json     registers from an independent pure-Python restatement of the three update
                     rules written from the published algorithms with Python big ints
                     (ULL as pack(OR of 1<<u), HLL/HMH as max) on python-xxhash hashes
"""
import json
import os
import random
import struct

import xxhash

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
M64 = (1 << 64) - 1


def rc(s: bytes) -> bytes:
    return s[::-1].translate(bytes.maketrans(b"ACGT", b"TGCA"))


def val(s: bytes) -> int:
    v = 0
    for c in s:
        v = (v << 2) | b"ACGT".index(c)
    return v


def canonical_kmers(seq: bytes, k: int):
    return [min(val(seq[i:i + k]), val(rc(seq[i:i + k]))) for i in range(len(seq) - k + 1)]


def clz64(x: int) -> int:
    return 64 - x.bit_length()


def ull_pack(hp: int) -> int:
    u = hp.bit_length() - 1
    return (u << 2) | (((hp >> (u - 1)) & 1) << 1 if u >= 1 else 0) | ((hp >> (u - 2)) & 1 if u >= 2 else 0)


def sketch_py(algo: str, p: int, k: int, seed: int, records):
    if algo == "ull":
        masks = [0] * (1 << p)
    elif algo == "hll":
        regs = [0] * (1 << p)
    else:
        regs = [0] * 16384
    for rec in records:
        f = bytes(c for c in rec if c in b"ACGT")
        if len(f) < k:
            continue
        for km in canonical_kmers(f, k):
            if algo == "hmh":
                h = xxhash.xxh3_128_intdigest(struct.pack("<I", km & 0xFFFFFFFF), seed)
                x, y = h >> 64, h & M64   # convention switch LO_HMH_X_IS_HIGH64 = 1
                idx = x >> 50
                lz = clz64(((x << 14) & M64) | 0x3FFF) + 1
                regs[idx] = max(regs[idx], (lz << 10) | (y & 1023))
            else:
                h = xxhash.xxh3_64_intdigest(struct.pack("<Q", km), seed)
                if algo == "hll":
                    j = h & ((1 << p) - 1)
                    w = h >> p
                    rho = clz64(w) - p + 1
                    regs[j] = max(regs[j], rho)
                else:
                    idx = h >> (64 - p)
                    nlz = clz64(((h << p) & M64) | ((1 << p) - 1))
                    masks[idx] |= 1 << (nlz + p - 1)
    if algo == "ull":
        regs = [ull_pack(m) if m else 0 for m in masks]
    return regs


def main():
    os.makedirs(OUT, exist_ok=True)
    rnd = random.Random(20261017)
    kat = {"xxh3_64_le64": [], "xxh3_128_le32": []}
    fixed = [(0, 42), (0, 0), (M64, M64), (1, 93), (0x0123456789ABCDEF, 42), (0xFFFFFFFF, 42), (1 << 32, 42)]
    for v, s in fixed + [(rnd.getrandbits(rnd.choice([8, 28, 32, 42, 64])), rnd.choice([42, 0, rnd.getrandbits(64)])) for _ in range(500)]:
        kat["xxh3_64_le64"].append([str(v), str(s), str(xxhash.xxh3_64_intdigest(struct.pack("<Q", v), s))])
        w = v & 0xFFFFFFFF
        kat["xxh3_128_le32"].append([str(w), str(s), str(xxhash.xxh3_128_intdigest(struct.pack("<I", w), s))])
    json.dump(kat, open(os.path.join(OUT, "xxh3_kat.json"), "w"))

    km = []
    for k in (1, 2, 5, 13, 14, 15, 16, 17, 21, 31, 32):
        s = bytes(rnd.choice(b"ACGT") for _ in range(80 + k))
        km.append({"k": k, "seq": s.decode(), "kmers": [str(x) for x in canonical_kmers(s, k)]})
    # palindromes and homopolymers: canonical ties
    for s in (b"ACGT" * 10, b"A" * 40, b"T" * 40, b"AT" * 20, b"GC" * 20):
        km.append({"k": 4, "seq": s.decode(), "kmers": [str(x) for x in canonical_kmers(s, 4)]})
    json.dump(km, open(os.path.join(OUT, "kmers.json"), "w"))

    consts = {
        "source": "hash4j UltraLogLog.OptimalFGRAEstimator literals as recorded in SURVEY.md Appendix A.4",
        "register_contributions_0_5": [0.8484061093359406, 0.38895829052007685, 0.5059986252327467, 0.17873835725405993,
                                       0.48074234060273024, 0.22040001471443574],
        "estimation_factors": {"3": 94.59941722950778, "10": 4824374.384717942, "11": 2.2486750611989766e7},
        "xxh3_kat_from_survey": {"xxh3_64(le64(0),42)": "0x4596708167f8eb2e",
                                 "xxh3_128(le32(0),42)": "0xe5703e4f92e590a19871214b43bdc0ac"},
    }
    json.dump(consts, open(os.path.join(OUT, "ull_constants.json"), "w"), indent=1)

    sk = []
    recs = ["".join(rnd.choice("ACGT") for _ in range(n)) for n in (700, 31, 5, 1200)]
    recs[0] = recs[0][:300] + "NNNNacgtnnRY" + recs[0][300:]
    for algo, p, k, seed in (("ull", 6, 16, 42), ("ull", 10, 21, 7), ("ull", 3, 5, 42), ("hll", 6, 16, 42), ("hll", 10, 31, 99),
                             ("hmh", 14, 16, 42), ("hmh", 14, 21, 42)):
        regs = sketch_py(algo, p, k, seed, [r.encode() for r in recs])
        nz = {str(i): v for i, v in enumerate(regs) if v}
        sk.append({"algo": algo, "p": p, "k": k, "seed": seed, "records": recs, "n_regs": len(regs), "nonzero": nz})
    json.dump(sk, open(os.path.join(OUT, "sketch_py.json"), "w"))

    # estimators: registers -> cardinalities, union estimate, frac, distances, by the pure-Python restatement of
    # tools/estimators_py.py (written from SURVEY.md Appendix A, not from the oracle)
    from tools import estimators_py as E
    est = []

    def valid_ull(p, n_hashes_log2):
        regs = []
        for _ in range(1 << p):
            if n_hashes_log2 < 0 and rnd.random() > 2.0 ** n_hashes_log2:
                regs.append(0)
                continue
            lvl = 0
            for _ in range(max(1, int(2 ** max(n_hashes_log2, 0)))):
                z = 0
                while rnd.random() < 0.5 and z < 60 - p:
                    z += 1
                lvl = max(lvl, z)
            regs.append(4 * (lvl + p - 1) + rnd.randrange(4) if lvl + p - 1 >= p + 1 else rnd.choice([4 * p - 4, 4 * p, 4 * p + 2]))
        return regs

    def hll_regs(p, n_hashes_log2):
        regs = []
        for _ in range(1 << p):
            if n_hashes_log2 < 0 and rnd.random() > 2.0 ** n_hashes_log2:
                regs.append(0)
                continue
            rho = 1
            for _ in range(max(1, int(2 ** max(n_hashes_log2, 0)))):
                z = 1
                while rnd.random() < 0.5 and z < 64 - p + 1:
                    z += 1
                rho = max(rho, z)
            regs.append(rho)
        return regs

    for p in (3, 6, 8):
        for lg in (-2.0, 0.0, 3.0, 7.0):
            a, b = valid_ull(p, lg), valid_ull(p, lg + 0.7)
            if lg == 7.0:
                a[1], b[2] = 253, 255                                   # saturated registers: FGRA large-range branch
            u = E.ull_merge(a, b)
            for name, fn in (("fgra", E.ull_fgra), ("ml", E.ull_ml)):
                ca, cb, cu = fn(a, p), fn(b, p), fn(u, p)
                fr = E.frac_from(ca, cb, cu, hll=False)
                est.append({"algo": "ull", "estimator": name, "p": p, "a": a, "b": b, "merged": u, "card_a": repr(ca), "card_b": repr(cb),
                            "union": repr(cu), "frac": repr(fr), "d_poisson_k16": repr(E.mash(fr, 16, 1)), "d_binomial_k21": repr(E.mash(fr, 21, 0))})
    for p in (4, 6, 8):
        for lg in (-2.0, 1.0, 4.0, 8.0):
            a, b = hll_regs(p, lg), hll_regs(p, lg + 0.5)
            u = [max(x, y) for x, y in zip(a, b)]
            (ca, fa), (cb, fb), (cu, fu) = E.hll_len(a, p), E.hll_len(b, p), E.hll_len(u, p)
            fr = E.frac_from(ca, cb, cu, hll=True)
            est.append({"algo": "hll", "p": p, "a": a, "b": b, "card_a": repr(ca), "card_b": repr(cb), "union": repr(cu),
                        "bias_regime": [fa, fb, fu], "frac": repr(fr), "d_poisson_k16": repr(E.mash(fr, 16, 1))})
    # HyperMinHash: two real sketches of overlapping sequences (sparse: stored as index -> value)
    base = "".join(rnd.choice("ACGT") for _ in range(6000))
    other = base[:3500] + "".join(rnd.choice("ACGT") for _ in range(2500))
    ha = sketch_py("hmh", 14, 16, 42, [base.encode()])
    hb = sketch_py("hmh", 14, 16, 42, [other.encode()])
    sim = E.hmh_similarity(ha, hb)
    est.append({"algo": "hmh", "p": 14, "a": {str(i): v for i, v in enumerate(ha) if v}, "b": {str(i): v for i, v in enumerate(hb) if v},
                "card_a": repr(E.hmh_cardinality(ha)), "card_b": repr(E.hmh_cardinality(hb)), "similarity": repr(sim),
                "frac": repr(2.0 * max(sim, 0.0) / (1.0 + max(sim, 0.0)))})
    json.dump(est, open(os.path.join(OUT, "estimators.json"), "w"))
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
