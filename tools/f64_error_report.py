"""f64 distance error of the GPU path against the oracle, per estimator, on the BASELINE config shapes.
    python -m tools.f64_error_report [--quick] > gpurun_out/f64_error.jsonl
For every case: max relative error, the fraction of cells within plain 1e-12 relative, the smallest s = (a+b-U)/U above which
every cell meets 1e-12, and how often the per-sketch cardinalities (pure pow / log differences) are bit-identical."""
import argparse
import json
import sys

import numpy as np

sys.path.insert(0, ".")
import oracle as O  # noqa: E402
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, ops  # noqa: E402
from tools import synth  # noqa: E402


def report(ctx, name, algo, oalgo, p, k, est, n, length, model=1, threads=16):
    gen = synth.genomes(n, length, seed=42)
    regs = ops.sketch_genomes(ctx, algo, p, k, 42, gen)
    exp = O.dist(oalgo, p, k, est, model, False, regs, regs, triangular=True, threads=threads)
    frac = O.dist(oalgo, p, k, est, 2, False, regs, regs, triangular=True, threads=threads)
    got, _ = ops.dist(ctx, algo, p, k, est, model, False, regs, regs, triangular=True)
    il = np.tril_indices(n)
    e, f, g = exp[il], frac[il], got
    err = np.abs(g - e)
    rel = np.where(err == 0, 0.0, err / np.maximum(np.abs(e), 1e-300))
    s = f / (2.0 - f)
    strict = rel <= 1e-12
    bad_s = s[~strict]
    cg = ops.cardinality(ctx, algo, p, est, regs)
    ce = np.array([O.cardinality(oalgo, p, est, r) for r in regs])
    ulp = np.abs(cg - ce) / np.spacing(np.abs(ce))
    out = {"case": name, "cells": int(len(e)), "max_rel_err": float(rel.max()), "frac_within_1e-12": float(strict.mean()),
           "frac_bit_identical": float((g == e).mean()), "cells_off": int((~strict).sum()),
           "smallest_s_where_1e-12_always_holds": float(bad_s.max()) if len(bad_s) else 0.0,
           "d_at_that_s": float(e[~strict][np.argmax(bad_s)]) if len(bad_s) else None,
           "cells_with_0<s<1e-4": int(((s > 0) & (s < 1e-4)).sum()), "cells_d_eq_1": int((e == 1.0).sum()),
           "card_bit_identical": float((cg == ce).mean()), "card_max_ulp": float(ulp.max())}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    a = ap.parse_args()
    q = a.quick
    with ops.Context(0) as ctx:
        report(ctx, "C2 shape: ULL p=10 k=16 FGRA poisson, 5 Mbp", ALGO_ULL, O.ULL, 10, 16, EST_FGRA, 100 if q else 400, 5_000_000)
        report(ctx, "C2 shape, binomial model", ALGO_ULL, O.ULL, 10, 16, EST_FGRA, 100 if q else 400, 5_000_000, model=0)
        report(ctx, "C5 shape: ULL p=10 k=16 ML poisson, 100 kbp", ALGO_ULL, O.ULL, 10, 16, EST_ML, 200 if q else 1000, 100_000)
        report(ctx, "C3 shape: HLL p=14 k=21 poisson, 5 Mbp", ALGO_HLL, O.HLL, 14, 21, 0, 50 if q else 200, 5_000_000)
        report(ctx, "C1 shape: HMH k=16 poisson, 2 Mbp", ALGO_HMH, O.HMH, 14, 16, 0, 50 if q else 200, 2_000_000)
        report(ctx, "ULL p=14 k=21 FGRA, 1 Mbp", ALGO_ULL, O.ULL, 14, 21, EST_FGRA, 60 if q else 200, 1_000_000)


if __name__ == "__main__":
    main()
