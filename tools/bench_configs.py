"""Kernel-level timings of the other BASELINE configs (not the bench.py headline): run on a GPU box.
    python -m tools.bench_configs [--quick] > gpurun_out/configs.json
Everything goes through the C ABI; torch builds the synthetic inputs in HBM and provides events."""
import argparse
import ctypes as C
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402  (reuses the HBM genome generator)
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, ops  # noqa: E402
from lash_b200.capi import Span, check, lib  # noqa: E402
from lash_b200.pack import padded_bytes  # noqa: E402


PROFILE = False  # --profile: exactly one launch per case (the run is under ncu, which replays it anyway)


def sketch_case(ctx, name, algo, p, k, n_g, length, reps=5):
    dev = torch.device("cuda", 0)
    buf, stride = bench.make_packed_genomes(torch, dev, n_g, length, 42, 0)
    spans = (Span * n_g)()
    for i in range(n_g):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    sk = ops.Sketcher(ctx, algo, p, k, 42, n_g)
    stream = torch.cuda.Stream(dev)
    sk.set_stream(stream.cuda_stream)
    ts = []
    warm = 0 if PROFILE else 2
    reps = 1 if PROFILE else reps
    for r in range(reps + warm):
        sk.reset()
        ms0, _ = sk.stats()
        sk.push_raw(buf.data_ptr(), n_g * stride, spans, n_g, None, 0, dev=True)
        ms1, _ = sk.stats()
        if r >= warm:
            ts.append(ms1 - ms0)
    regs = sk.fetch()
    sk.close()
    ms = float(np.median(ts))
    return {"case": name, "algo": algo, "p": p, "k": k, "genomes": n_g, "genome_len": length, "kernel_ms": ms,
            "gbp_per_s": n_g * length / ms / 1e6}, regs


def reads_case(ctx, name, p, k, n_reads, read_len, reps=3, uniform=False):
    """config 4 shape: every 150 bp read is its own record of ONE sample -> boundary mask path."""
    dev = torch.device("cuda", 0)
    n_bases = n_reads * read_len
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    packed = torch.randint(0, 256, (padded_bytes(n_bases) + 64,), dtype=torch.uint8, device=dev, generator=g)
    rec = (np.arange(n_reads + 1, dtype=np.uint64) * read_len)
    spans = (Span * 1)(Span(0, 0, n_bases, 0, n_reads, read_len if uniform else 0))
    sk = ops.Sketcher(ctx, ALGO_ULL, p, k, 42, 1)
    stream = torch.cuda.Stream(dev)
    sk.set_stream(stream.cuda_stream)
    ts = []
    warm = 0 if PROFILE else 1
    reps = 1 if PROFILE else reps
    for r in range(reps + warm):
        sk.reset()
        ms0, _ = sk.stats()
        t0 = time.perf_counter()
        if uniform:
            sk.push_raw(packed.data_ptr(), padded_bytes(n_bases), spans, 1, None, 0, dev=True)
        else:
            sk.push_raw(packed.data_ptr(), padded_bytes(n_bases), spans, 1, rec.ctypes.data_as(C.c_void_p), len(rec), dev=True)
        ms1, _ = sk.stats()
        wall = time.perf_counter() - t0
        if r >= warm:
            ts.append((ms1 - ms0, wall * 1e3))
    sk.close()
    ms = float(np.median([a for a, _ in ts]))
    return {"case": name, "algo": ALGO_ULL, "p": p, "k": k, "reads": n_reads, "read_len": read_len, "kernel_ms(mask+sketch)": ms,
            "gbp_per_s": n_bases / ms / 1e6, "push_wall_ms_incl_8B_per_read_table_h2d": float(np.median([b for _, b in ts]))}


def dist_case(ctx, name, algo, p, k, est, regs, reps=3):
    n = regs.shape[0]
    ts = []
    warm = 0 if PROFILE else 1
    reps = 1 if PROFILE else reps
    for r in range(reps + warm):
        d, w = ops.dist(ctx, algo, p, k, est, 1, False, regs, regs, triangular=True)
        ms, _ = ops.dist_stats(ctx)
        if r >= warm:
            ts.append(ms)
    ms = float(np.median(ts))
    pairs = n * (n + 1) // 2
    cells = regs.shape[1]
    return {"case": name, "algo": algo, "p": p, "est": est, "n": n, "pairs": pairs, "kernel_ms": ms, "pairs_per_s": pairs / ms * 1e3,
            "register_merges_per_s": pairs * cells / ms * 1e3, "warn": w}


def c5_full(ctx, n=100_000, length=100_000, est=EST_ML):
    """BASELINE configs[4] at full size on ONE GPU: n ULL p=10 sketches (made by the sketch kernel from n random
    `length`-bp genomes), all-vs-all ML, lower triangle, f64, streamed to the host in row blocks (--dm shape)."""
    dev = torch.device("cuda", 0)
    stride = padded_bytes(length)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    buf = torch.randint(0, 256, (n * stride + 64,), dtype=torch.uint8, device=dev, generator=g)
    spans = (Span * n)()
    for i in range(n):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    sk = ops.Sketcher(ctx, ALGO_ULL, 10, 16, 42, n)
    t0 = time.perf_counter()
    sk.push_raw(buf.data_ptr(), n * stride, spans, n, None, 0, dev=True)
    regs = sk.fetch()
    t_sketch = time.perf_counter() - t0
    k_ms, _ = sk.stats()
    sk.close()
    del buf
    torch.cuda.empty_cache()
    seen = {"rows": 0, "blocks": 0, "checksum": 0.0}

    def on_block(row0, block):
        seen["rows"] += block.shape[0]
        seen["blocks"] += 1
        seen["checksum"] += float(block[0, 0])

    t0 = time.perf_counter()
    w = ops.dist_stream(ctx, ALGO_ULL, 10, 16, est, 1, False, regs, regs, True, 0, on_block)
    wall = time.perf_counter() - t0
    ms, launches = ops.dist_stats(ctx)
    pairs = n * (n + 1) // 2
    return {"case": f"C5 FULL: {n} x {n} ULL p=10 {'ML' if est == EST_ML else 'FGRA'} triangle, f64, streamed in row blocks to pinned host memory",
            "n": n, "pairs": pairs, "sketch_wall_s": t_sketch, "sketch_kernel_ms": k_ms, "sketch_gbp_per_s_kernel": n * length / k_ms / 1e6,
            "dist_wall_s": wall, "dist_kernel_ms": ms, "pairs_per_s_wall": pairs / wall, "pairs_per_s_kernel": pairs / ms * 1e3,
            "d2h_bytes": n * n * 8, "blocks": seen["blocks"], "rows_delivered": seen["rows"], "launches": launches, "warn": w}


def c4_full(ctx, total_gbp=100, chunk_reads=26_666_667, read_len=150):
    """BASELINE configs[3] shape: `total_gbp` Gbp of 150 bp reads of ONE sample -> one ULL p=14 k=21 sketch; a 4 Gbp
    chunk of packed reads resident in HBM is pushed repeatedly (fixed-length records: no boundary table)."""
    dev = torch.device("cuda", 0)
    n_bases = chunk_reads * read_len
    g = torch.Generator(device=dev)
    g.manual_seed(9)
    packed = torch.randint(0, 256, (padded_bytes(n_bases) + 64,), dtype=torch.uint8, device=dev, generator=g)
    spans = (Span * 1)(Span(0, 0, n_bases, 0, chunk_reads, read_len))
    sk = ops.Sketcher(ctx, ALGO_ULL, 14, 21, 42, 1)
    stream = torch.cuda.Stream(dev)
    sk.set_stream(stream.cuda_stream)
    n_push = int(round(total_gbp * 1e9 / n_bases))
    sk.push_raw(packed.data_ptr(), padded_bytes(n_bases), spans, 1, None, 0, dev=True)  # warm-up
    sk.sync()
    ms0, _ = sk.stats()
    t0 = time.perf_counter()
    for _ in range(n_push):
        sk.push_raw(packed.data_ptr(), padded_bytes(n_bases), spans, 1, None, 0, dev=True)
    sk.sync()
    wall = time.perf_counter() - t0
    ms1, _ = sk.stats()
    regs = sk.fetch()
    sk.close()
    return {"case": f"C4 FULL: {n_push} x {n_bases / 1e9:.2f} Gbp of {read_len} bp reads -> one ULL p=14 k=21 sketch (HBM-resident chunk re-pushed)",
            "bases": n_push * n_bases, "kernel_ms(mask+sketch)": ms1 - ms0, "wall_s": wall, "gbp_per_s_kernel": n_push * n_bases / (ms1 - ms0) / 1e6,
            "gbp_per_s_wall": n_push * n_bases / wall / 1e9, "nonzero_registers": int((regs != 0).sum())}


def writer_case(ctx, n=20_000, threads=0):
    """`lash dist` end to end on files: n ULL p=10 sketches on disk -> lash::dist_command (fused kernel
    epilogue + parallel `{:.6}` formatter) -> text on tmpfs; --dm matrix (lower triangle) and the TSV list."""
    import json as js
    import os
    import shutil
    import tempfile

    from lash_b200 import hostapi
    threads = threads or (os.cpu_count() or 1)
    dev = torch.device("cuda", 0)
    length = 50_000
    stride = padded_bytes(length)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    buf = torch.randint(0, 256, (n * stride + 64,), dtype=torch.uint8, device=dev, generator=g)
    spans = (Span * n)()
    for i in range(n):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    sk = ops.Sketcher(ctx, ALGO_ULL, 10, 16, 42, n)
    sk.push_raw(buf.data_ptr(), n * stride, spans, n, None, 0, dev=True)
    regs = sk.fetch()
    sk.close()
    d = tempfile.mkdtemp(prefix="lash_writer_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    out = []
    try:
        prefix = os.path.join(d, "db")
        hostapi.write_sketches(prefix + "_sketches.bin", ALGO_ULL, 10, regs, threads=threads)
        open(prefix + "_files.json", "w").write(js.dumps([f"genome_{i:06d}.fna" for i in range(n)], indent=2))
        hostapi.check(hostapi.lib().lash_host_write_parameters(prefix.encode(), ALGO_ULL, 10, 16, 42))
        for dm in (True, False):
            path = os.path.join(d, "dist.out")
            t0 = time.perf_counter()
            hostapi.dist(ctx, prefix, prefix, path, "fgra", 1, dm=dm, threads=threads, fused=True)
            wall = time.perf_counter() - t0
            cells = n * (n + 1) // 2
            out.append({"case": f"lash dist on files: {n} x {n} ULL p=10 FGRA, {'--dm matrix' if dm else 'TSV list'}, fused + {threads} formatter threads, tmpfs",
                        "cells": cells, "wall_s": wall, "cells_per_s": cells / wall, "bytes_written": os.path.getsize(path),
                        "text_gb_per_s": os.path.getsize(path) / wall / 1e9})
            os.remove(path)
    finally:
        shutil.rmtree(d, ignore_errors=True)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--profile", action="store_true", help="one launch per case, quick sizes (for ncu)")
    ap.add_argument("--writer", action="store_true", help="only the file -> dist -> text writer case")
    ap.add_argument("--full", action="store_true", help="only the full-size C4 / C5 runs (tens of seconds of GPU time)")
    a = ap.parse_args()
    if a.writer:
        with ops.Context(0) as ctx:
            for r in writer_case(ctx):
                print(json.dumps(r), flush=True)
        return
    if a.full:
        with ops.Context(0) as ctx:
            print(json.dumps(c4_full(ctx)), flush=True)
            print(json.dumps(c5_full(ctx, est=EST_FGRA)), flush=True)
            print(json.dumps(c5_full(ctx, est=EST_ML)), flush=True)
        return
    global PROFILE
    PROFILE = a.profile
    q = a.quick or a.profile
    out = []
    with ops.Context(0) as ctx:
        r, regs_ull10 = sketch_case(ctx, "C2 ULL p=10 k=16", ALGO_ULL, 10, 16, 200 if q else 1000, 5_000_000)
        out.append(r)
        r, regs_hll14 = sketch_case(ctx, "C3 HLL p=14 k=21", ALGO_HLL, 14, 21, 200 if q else 1000, 5_000_000)
        out.append(r)
        r, regs_hmh = sketch_case(ctx, "C1 HMH k=16", ALGO_HMH, 14, 16, 100 if q else 400, 2_000_000)
        out.append(r)
        r, _ = sketch_case(ctx, "ULL p=14 k=21 (genomes)", ALGO_ULL, 14, 21, 200 if q else 1000, 5_000_000)
        out.append(r)
        r, _ = sketch_case(ctx, "ULL p=10 k=31", ALGO_ULL, 10, 31, 200 if q else 1000, 5_000_000)
        out.append(r)
        r, _ = sketch_case(ctx, "ULL p=10 k=12", ALGO_ULL, 10, 12, 200 if q else 1000, 5_000_000)
        out.append(r)
        r, _ = sketch_case(ctx, "ULL p=18 k=21 (global accumulators)", ALGO_ULL, 18, 21, 50 if q else 200, 5_000_000)
        out.append(r)
        r, regs_small = sketch_case(ctx, "C5 inputs: ULL p=10 k=16 100 kbp", ALGO_ULL, 10, 16, 2000 if q else 8000, 100_000)
        out.append(r)
        out.append(reads_case(ctx, "C4 ULL p=14 k=21, 150 bp reads, one sample (rec_start table)", 14, 21, 2_000_000 if q else 20_000_000, 150))
        out.append(reads_case(ctx, "C4 ULL p=14 k=21, 150 bp reads, one sample (rec_len=150, no table)", 14, 21,
                              2_000_000 if q else 20_000_000, 150, uniform=True))
        out.append(dist_case(ctx, "C2 dist FGRA 1000x1000", ALGO_ULL, 10, 16, EST_FGRA, regs_ull10))
        out.append(dist_case(ctx, "C5 dist ML p=10 (n x n triangle)", ALGO_ULL, 10, 16, EST_ML, regs_small))
        out.append(dist_case(ctx, "dist FGRA p=10 (n x n triangle)", ALGO_ULL, 10, 16, EST_FGRA, regs_small))
        out.append(dist_case(ctx, "C3 dist HLL p=14", ALGO_HLL, 14, 21, 0, regs_hll14))
        if not a.quick:  # the same sketches several times over: enough tiles to fill the GPU (timing only)
            n_big = 2000 if q else 4000
            out.append(dist_case(ctx, f"C3 dist HLL p=14 ({n_big} sketches)", ALGO_HLL, 14, 21, 0,
                                 np.tile(regs_hll14, (n_big // regs_hll14.shape[0], 1))))
        out.append(dist_case(ctx, "C1 dist HMH", ALGO_HMH, 14, 16, 0, regs_hmh))
    for r in out:
        print(json.dumps(r))


if __name__ == "__main__":
    main()
