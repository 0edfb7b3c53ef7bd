"""Host ingest rates on the GPU box: FASTA text (tmpfs) -> sketch_files, per ingest mode and worker count.
    python -m tools.ingest_probe [--gpu]      (dry = parse + pack/copy only; --gpu adds the real runs)"""
import argparse
import os
import sys

import numpy as np

sys.path.insert(0, ".")
from lash_b200 import ALGO_ULL, hostapi  # noqa: E402
from tools import synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpu", action="store_true")
    ap.add_argument("--threads", default="1,4,16")
    a = ap.parse_args()
    d = "/dev/shm/lash_ingest_probe"
    os.makedirs(d, exist_ok=True)
    files = []
    for i in range(32):
        p = f"{d}/g{i}.fa"
        if not os.path.exists(p):
            s = np.frombuffer(synth.genomes(1, 5_000_000, seed=i)[0][0], dtype=np.uint8)
            body = np.concatenate([s.reshape(-1, 80), np.full((len(s) // 80, 1), 10, dtype=np.uint8)], axis=1).tobytes()
            open(p, "wb").write(b">g\n" + body)
        files.append(p)
    files = files * 8
    gbp = len(files) * 5e6 / 1e9
    ctx = None
    if a.gpu:
        from lash_b200 import ops
        ctx = ops.Context(0)
    for mode, name in ((1, "packed"), (2, "ascii")):
        hostapi.lib().lash_host_set_ingest_mode(mode)
        for th in [int(x) for x in a.threads.split(",")]:
            dry = min(hostapi.pack_files_dry(files, 16, threads=th).seconds_total for _ in range(3))
            line = f"{name:6s} threads {th:2d}  dry {gbp / dry:7.2f} Gbp/s ({gbp / dry / th:5.2f}/thread)"
            if ctx:
                best, st = 1e9, None
                for _ in range(3):
                    _, s1 = hostapi.sketch_files_regs(ctx, ALGO_ULL, 10, 16, 42, files, threads=th)
                    if s1.seconds_total < best:
                        best, st = s1.seconds_total, s1
                line += (f"   gpu {gbp / best:7.2f} Gbp/s  (open {st.seconds_open * 1e3:.1f} workers {st.seconds_workers * 1e3:.1f} "
                         f"drain {st.seconds_drain * 1e3:.1f} ms, {st.n_pushes} pushes, kernels {st.gpu_kernel_ms:.1f} ms)")
            print(line, flush=True)


if __name__ == "__main__":
    main()
