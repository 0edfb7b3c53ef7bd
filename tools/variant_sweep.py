"""Kernel-only sketch rates for tuning builds of liblash_gpu (LASH_GPU_LIB=... python -m tools.variant_sweep).
Build variants with e.g. `make -C lash_b200/csrc VARIANT=-DLASH_MINB_256=5 OUT=../_lib/liblash_gpu_b5.so OBJDIR=_obj_b5`."""
import json
import os
import sys

sys.path.insert(0, ".")
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, ops  # noqa: E402
from tools import bench_configs as bc  # noqa: E402


def main():
    out = {"lib": os.environ.get("LASH_GPU_LIB", "default")}
    with ops.Context(0) as ctx:
        for name, algo, p, k, n, length in (("ull10k16", ALGO_ULL, 10, 16, 400, 5_000_000), ("ull10k12", ALGO_ULL, 10, 12, 400, 5_000_000),
                                            ("ull10k31", ALGO_ULL, 10, 31, 400, 5_000_000), ("hll14k21", ALGO_HLL, 14, 21, 400, 5_000_000),
                                            ("hll12k21", ALGO_HLL, 12, 21, 400, 5_000_000), ("hmh16", ALGO_HMH, 14, 16, 400, 2_000_000),
                                            ("ull14k21", ALGO_ULL, 14, 21, 400, 5_000_000)):
            r, _ = bc.sketch_case(ctx, name, algo, p, k, n, length, reps=5)
            out[name] = round(r["gbp_per_s"], 1)
        out["reads150_ull14k21"] = round(bc.reads_case(ctx, "reads", 14, 21, 8_000_000, 150, uniform=True)["gbp_per_s"], 1)
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
