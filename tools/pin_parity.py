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
  HMH distances differ at ~1e-10    -> LO_HMH_CARD_TRUNC (cardinality truncated to u64 as in the Go original)
  HLL registers differ              -> index / rho convention of streaming_algorithms, SURVEY.md A.3
Exit status 0 = every comparison passed.  Nothing here is imported by the product, tests or bench.
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


def main():
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
            print(json.dumps({"check": "registers", "algo": name, "k": k, "equal": bool(ok), "registers_differing": diff}))
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
