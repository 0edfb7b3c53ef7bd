#!/usr/bin/env python
"""Pin parity against the REAL lash binary -- for the first machine that has both a Rust toolchain and a B200.

This repo's parity is "unpinned" (DESIGN.md section 2): the reference is Rust + five crates that are not vendored, and the
build image has no cargo.  Everything that can be pinned offline is pinned in tests/; this script is the remaining step.
It needs no network beyond what `cargo build` of the reference needs:

    git clone https://github.com/jianshu93/lash && (cd lash && cargo build --release)
    python tools/pin_parity.py --lash lash/target/release/lash [--work /tmp/pin] [--algos ull,hll,hmh]

What it does, per algorithm (ULL p=10 / HLL p=14 / HMH; k=16 and k=21; seed 42):
  1. writes deterministic inputs: 6 synthetic genomes (tools/synth.py, the bench's generator) as FASTA, one "dirty"
     multi-record file (lower case, N runs, records shorter than k) and one gzip-compressed FASTQ of 150 bp reads;
  2. runs `lash sketch` and `lash-b200 sketch` on the same list file and compares the sketch files REGISTER BY REGISTER
     (both are read with this repo's reader: the on-disk formats are the reference's) -- must be equal;
  3. runs `lash dist` and `lash-b200 dist` (every estimator, both models, list and --dm) on the REFERENCE's sketch files
     and compares the two outputs as sets of (reference, query) -> value: equal to 6 decimals, the printed precision.
On a mismatch it says which of the recalled conventions to flip (they are single switches, identical in oracle and kernels):
  HMH registers differ everywhere   -> rebuild with -DLASH_HMH_X_IS_HIGH64=0 (oracle: LO_HMH_X_IS_HIGH64), SURVEY.md A.5
  HMH distances of TINY sketches    -> LO_HMH_CARD_TRUNC (cardinality truncated to u64 as in the Go original; for sketches of
                                       >= 10^4 k-mers the two conventions agree to ~4e-16: --self-test measures it)
  HLL registers differ              -> index / rho convention of streaming_algorithms, SURVEY.md A.3
Exit status 0 = every comparison passed.  Nothing here is imported by the product or bench.

    python tools/pin_parity.py --self-test        (no GPU, no lash binary)
fabricates "reference" registers / distances with each recalled convention flipped in turn (pure-Python restatements on
python-xxhash) and checks that the diagnosis below names exactly that switch -- so the first run against the real binary
either passes or says in one line what to flip.
"""
from __future__ import annotations

import argparse
import gzip
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALGOS = {"ull": (2, 10), "hll": (1, 14), "hmh": (0, 14)}


def write_inputs(work: str) -> str:
    from tools import synth
    os.makedirs(work, exist_ok=True)
    files = []
    for g, recs in enumerate(synth.genomes(6, 300_000, seed=42)):
        path = os.path.join(work, f"genome_{g}.fa")
        with open(path, "wb") as f:
            f.write(b">g%d\n" % g + b"\n".join(recs[0][o:o + 80] for o in range(0, len(recs[0]), 80)) + b"\n")
        files.append(path)
    dirty = os.path.join(work, "dirty.fa")
    with open(dirty, "wb") as f:
        for i, rec in enumerate(synth.dirty_genome(120_000, 21, seed=7)):
            f.write(b">r%d\n" % i + rec + b"\n")
    files.append(dirty)
    reads = os.path.join(work, "reads.fq.gz")
    rng = np.random.default_rng(3)
    src = synth.genomes(1, 200_000, seed=9)[0][0]
    with gzip.open(reads, "wb") as f:
        for i in range(4000):
            o = int(rng.integers(0, len(src) - 150))
            f.write(b"@read%d\n" % i + src[o:o + 150] + b"\n+\n" + b"I" * 150 + b"\n")
    files.append(reads)
    lst = os.path.join(work, "files.txt")
    with open(lst, "w") as f:
        f.write("\n".join(files) + "\n")
    return lst


def run(cmd, cwd):
    r = subprocess.run(cmd, cwd=cwd, capture_output=True, text=True)
    if r.returncode != 0:
        raise SystemExit(f"FAILED: {' '.join(cmd)}\n{r.stdout}\n{r.stderr}")


def read_table(path: str) -> dict:
    """TSV list or --dm matrix -> {(reference, query): text value}."""
    out = {}
    with open(path) as f:
        lines = [ln.rstrip("\n") for ln in f if ln.strip()]
    if lines and lines[0].startswith("Reference\tQuery"):
        for ln in lines[1:]:
            r, q, d = ln.split("\t")
            out[(r, q)] = d
        return out
    cols = lines[0].split("\t")[1:]
    for ln in lines[1:]:
        cells = ln.split("\t")
        for q, d in zip(cols, cells[1:]):
            out[(cells[0], q)] = d
    return out


def same_pairs(a: dict, b: dict) -> tuple[int, list]:
    """Triangular outputs hold one orientation of every unordered pair, and which one depends on the reference's
    HashMap order: compare by unordered pair."""
    ka = {tuple(sorted(k)): v for k, v in a.items()}
    kb = {tuple(sorted(k)): v for k, v in b.items()}
    bad = [(k, ka.get(k), kb.get(k)) for k in sorted(set(ka) | set(kb)) if ka.get(k) != kb.get(k)]
    return len(ka), bad


# ---- diagnosis: which recalled convention explains a mismatch ----------------------------------------------------------
M64 = (1 << 64) - 1
SWITCHES = {
    "hmh_x_low64": "HMH: x (index + leading zeros) is the LOW 64 bits of xxh3_128 -> rebuild with -DLASH_HMH_X_IS_HIGH64=0 "
                   "(csrc/registers.cuh) and LO_HMH_X_IS_HIGH64 0 (oracle/lash_oracle.c)",
    "hll_index_high": "HLL: the register index is the TOP p bits of the hash and rho counts the bits below -> swap the index / rho "
                      "convention in Cell<HLL>::from_kmer + SmemAcc<HLL>::prep (csrc) and lo_hll_push_hash64 (oracle), SURVEY.md A.3",
    "hmh_card_trunc": "HMH: cardinality() is truncated to an integer before similarity() -> set LO_HMH_CARD_TRUNC 1 (oracle) and "
                      "truncate in hmh_cardinality_from (csrc/estimators.cuh), SURVEY.md A.5",
}


def _clz64(x: int) -> int:
    return 64 - x.bit_length()


def _canonical_kmers(seq: bytes, k: int):
    f = bytes(c for c in seq if c in b"ACGT")                       # filter_out_n, utils.rs:33-41
    code = {65: 0, 67: 1, 71: 2, 84: 3}
    mask = (1 << (2 * k)) - 1
    fwd = rc = 0
    out = []
    for i, c in enumerate(f):
        b = code[c]
        fwd = ((fwd << 2) | b) & mask
        rc = (rc >> 2) | ((3 - b) << (2 * (k - 1)))
        if i >= k - 1:
            out.append(min(fwd, rc))
    return out


def registers_py(algo: str, p: int, k: int, seed: int, records, hmh_x_high=True, hll_index_low=True):
    """Pure-Python sketch of one file under the current conventions (defaults) or a flipped one."""
    import struct

    import xxhash
    regs = [0] * (16384 if algo == "hmh" else (1 << p))
    masks = [0] * (1 << p)
    for rec in records:
        kms = _canonical_kmers(rec, k)
        for km in kms:
            if algo == "hmh":
                h = xxhash.xxh3_128_intdigest(struct.pack("<I", km & 0xFFFFFFFF), seed)      # utils.rs:397
                x, y = (h >> 64, h & M64) if hmh_x_high else (h & M64, h >> 64)
                idx = x >> 50
                lz = _clz64(((x << 14) & M64) | 0x3FFF) + 1
                regs[idx] = max(regs[idx], (lz << 10) | (y & 1023))
                continue
            h = xxhash.xxh3_64_intdigest(struct.pack("<Q", km), seed)                        # utils.rs:412,428
            if algo == "hll":
                if hll_index_low:
                    j, rho = h & ((1 << p) - 1), _clz64(h >> p) - p + 1
                else:
                    j, rho = h >> (64 - p), _clz64(((h << p) & M64) | (1 << (p - 1))) + 1
                regs[j] = max(regs[j], rho)
            else:
                nlz = _clz64(((h << p) & M64) | ((1 << p) - 1))
                masks[h >> (64 - p)] |= 1 << (nlz + p - 1)
    if algo == "ull":
        regs = [(((m.bit_length() - 1) << 2) | ((m >> (m.bit_length() - 3)) & 3)) if m else 0 for m in masks]
    return regs


def diagnose_registers(algo: str, p: int, k: int, seed: int, records, reference_regs) -> str:
    """Name the convention under which a pure-Python sketch of `records` reproduces the reference's registers."""
    ref = [int(x) for x in reference_regs]
    if registers_py(algo, p, k, seed, records) == ref:
        return "current"
    if algo == "hmh" and registers_py(algo, p, k, seed, records, hmh_x_high=False) == ref:
        return "hmh_x_low64"
    if algo == "hll" and registers_py(algo, p, k, seed, records, hll_index_low=False) == ref:
        return "hll_index_high"
    return "unknown"


def hmh_frac_py(a, b, trunc: bool) -> float:
    from tools import estimators_py as E
    c = sum(1 for x, y in zip(a, b) if x != 0 and x == y)
    n = sum(1 for x, y in zip(a, b) if x != 0 or y != 0)
    ca, cb = E.hmh_cardinality(a), E.hmh_cardinality(b)
    if trunc:
        ca, cb = float(int(ca)), float(int(cb))
    s = 0.0
    if c:
        ec = E.hmh_expected_collisions(ca, cb)
        s = 0.0 if c < ec else (c - ec) / n
    s = max(s, 0.0)
    return 2.0 * s / (1.0 + s)


def diagnose_hmh_distance(a, b, reference_frac: float) -> str:
    cur, alt = hmh_frac_py(a, b, False), hmh_frac_py(a, b, True)
    if abs(cur - alt) <= 1e-13 * abs(cur) and abs(cur - reference_frac) <= 1e-13 * abs(cur):
        return "indistinguishable"   # the two conventions agree far below the path's 1e-12: the switch cannot matter here
    if abs(cur - reference_frac) <= 1e-13 * abs(cur):
        return "current"
    if abs(alt - reference_frac) <= 1e-13 * abs(alt):
        return "hmh_card_trunc"
    return "unknown"


def self_test() -> int:
    """Each flipped convention must be named; the current conventions must be recognised as such."""
    import random
    rnd = random.Random(5)
    recs = [bytes(rnd.choice(b"ACGT") for _ in range(n)) for n in (6000, 40, 2500)] + [b"ACGTNNNNacgtACGTACGTACGTACGTTTGACCA" * 30]
    ok = True

    def expect(what, got, want):
        nonlocal ok
        good = got == want
        ok = ok and good
        print(json.dumps({"self_test": what, "diagnosis": got, "expected": want, "ok": good,
                          "advice": SWITCHES.get(got, "nothing to change" if got in ("current", "indistinguishable") else "no single switch explains this")}))

    for k in (16, 21):
        for algo, p in (("ull", 10), ("hll", 12), ("hmh", 14)):
            expect(f"{algo} k={k}: reference == current conventions", diagnose_registers(algo, p, k, 42, recs, registers_py(algo, p, k, 42, recs)), "current")
        expect(f"hmh k={k}: reference built with x = low 64 bits", diagnose_registers("hmh", 14, k, 42, recs, registers_py("hmh", 14, k, 42, recs, hmh_x_high=False)),
               "hmh_x_low64")
        expect(f"hll k={k}: reference built with index = top p bits", diagnose_registers("hll", 12, k, 42, recs, registers_py("hll", 12, k, 42, recs, hll_index_low=False)),
               "hll_index_high")
    garbage = registers_py("hll", 12, 16, 42, recs)
    garbage[7] ^= 1
    expect("hll: a corrupted reference register is not explained by any switch", diagnose_registers("hll", 12, 16, 42, recs, garbage), "unknown")
    # distances: related HMH sketches (shared registers), cardinality truncated or not
    a = registers_py("hmh", 14, 16, 42, [bytes(rnd.choice(b"ACGT") for _ in range(60_000))])
    b = list(a)
    for i in range(0, 16384, 3):
        b[i] = (b[i] + 1025) & 0xFFFF or 1
    # measured: truncating the two cardinalities moves frac by ~4e-16 for sketches of >= 10^4 k-mers (the expected-collision term
    # it feeds is ~1e-6 of C) -- far below 1e-12 and invisible at {:.6}; only sketches of a handful of k-mers can tell
    expect("hmh distance, 60k k-mers: f64 vs truncated cardinalities", diagnose_hmh_distance(a, b, hmh_frac_py(a, b, True)), "indistinguishable")
    tiny_a = registers_py("hmh", 14, 16, 42, [bytes(rnd.choice(b"ACGT") for _ in range(40))])
    tiny_b = list(tiny_a)
    tiny_b[next(i for i, v in enumerate(tiny_b) if v)] = 0
    if abs(hmh_frac_py(tiny_a, tiny_b, False) - hmh_frac_py(tiny_a, tiny_b, True)) > 1e-12:
        expect("hmh distance, 25 k-mers: reference with truncated cardinalities", diagnose_hmh_distance(tiny_a, tiny_b, hmh_frac_py(tiny_a, tiny_b, True)),
               "hmh_card_trunc")
        expect("hmh distance, 25 k-mers: reference with f64 cardinalities", diagnose_hmh_distance(tiny_a, tiny_b, hmh_frac_py(tiny_a, tiny_b, False)), "current")
    print("SELF-TEST PASSED: every flipped convention is named" if ok else "SELF-TEST FAILED")
    return 0 if ok else 1


def read_records(path: str):
    """Record sequences of a FASTA / FASTQ (.gz) file, for the diagnosis (small inputs only)."""
    op = gzip.open if path.endswith(".gz") else open
    with op(path, "rb") as f:
        data = f.read()
    if data[:1] == b"@":
        lines = data.split(b"\n")
        return [lines[i + 1].strip() for i in range(0, len(lines) - 3, 4)]
    return [b"".join(part.split(b"\n")[1:]).replace(b"\r", b"") for part in data.split(b">")[1:]]


def main():
    if "--self-test" in sys.argv:
        sys.exit(self_test())
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--lash", required=True, help="path to the reference binary (cargo build --release)")
    ap.add_argument("--ours", default=os.path.join(ROOT, "lash_b200", "_lib", "lash-b200"))
    ap.add_argument("--work", default="/tmp/lash_pin_parity")
    ap.add_argument("--algos", default="ull,hll,hmh")
    a = ap.parse_args()
    from lash_b200 import hostapi
    lash, ours = os.path.abspath(a.lash), os.path.abspath(a.ours)
    lst = write_inputs(a.work)
    n_files = sum(1 for _ in open(lst))
    failures = 0
    for name in a.algos.split(","):
        algo, p = ALGOS[name]
        for k in (16, 21):
            ref_dir, our_dir = os.path.join(a.work, f"ref_{name}_k{k}"), os.path.join(a.work, f"ours_{name}_k{k}")
            for d in (ref_dir, our_dir):
                os.makedirs(d, exist_ok=True)
            common = ["sketch", "-f", lst, "-o", "db", "-k", str(k), "-a", name, "-p", str(p), "-s", "42", "-t", "4"]
            run([lash] + common, ref_dir)
            run([ours] + common, our_dir)
            r_ref, _ = hostapi.read_sketches(os.path.join(ref_dir, "db_sketches.bin"), algo, n_files, p)
            r_our, _ = hostapi.read_sketches(os.path.join(our_dir, "db_sketches.bin"), algo, n_files, p)
            ok = np.array_equal(r_ref, r_our)
            diff = int((r_ref != r_our).sum())
            line = {"check": "registers", "algo": name, "k": k, "equal": bool(ok), "registers_differing": diff}
            if not ok:   # which recalled convention reproduces the reference's registers (first file that differs)?
                files = [ln.strip() for ln in open(lst) if ln.strip()]
                g = int(np.argmax((r_ref != r_our).any(axis=1)))
                d = diagnose_registers(name, p, k, 42, read_records(files[g]), r_ref[g])
                line.update({"diagnosis": d, "advice": SWITCHES.get(d, "no single switch explains this")})
            print(json.dumps(line))
            failures += not ok
            ests = ("fgra", "ml") if name == "ull" else ("fgra",)
            for est in ests:
                for model in ("1", "0"):
                    for dm in (False, True):
                        flags = ["dist", "-q", "db", "-r", "db", "-e", est, "-m", model, "-t", "4"] + (["--dm"] if dm else [])
                        run([lash] + flags + ["-o", "ref.out"], ref_dir)
                        run([ours] + flags + ["-o", "ours.out"], ref_dir)      # ours on the REFERENCE's sketch files
                        n, bad = same_pairs(read_table(os.path.join(ref_dir, "ref.out")), read_table(os.path.join(ref_dir, "ours.out")))
                        print(json.dumps({"check": "dist", "algo": name, "k": k, "estimator": est, "model": model, "dm": dm,
                                          "pairs": n, "differing": len(bad), "first": bad[:3]}))
                        failures += bool(bad)
    print("PARITY PINNED: every comparison passed" if failures == 0 else f"{failures} comparison(s) FAILED -- see the docstring for the switches")
    sys.exit(1 if failures else 0)


if __name__ == "__main__":
    main()
