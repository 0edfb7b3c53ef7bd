"""One launch of every hot kernel at the BASELINE config shapes, for ncu (tools/profile_round.sh).
    ncu ... python -m tools.profile_kernels gpurun_out/<tag>_manifest.jsonl
Writes one manifest line per case: which kernels (regex on the demangled name) belong to it and how many algorithmic units
(k-mers, register pairs, text bytes) the launch processed.  tools/kernel_costs.py joins the manifest with the ncu CSV of
the same run into profiles/kernel_costs.json (instructions per unit, DRAM bytes per unit, pipe utilisation), tagged with
the hash of the CUDA sources, which bench.py checks before it prints a roofline fraction."""
import ctypes as C
import json
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, EST_FGRA, EST_ML, capi, ops  # noqa: E402
from lash_b200.capi import Span, TextSpan, check  # noqa: E402
from lash_b200.pack import padded_bytes  # noqa: E402

OUT = open(sys.argv[1], "w") if len(sys.argv) > 1 else sys.stdout
DEV = torch.device("cuda", 0)


def emit(**kw):
    OUT.write(json.dumps(kw) + "\n")
    OUT.flush()


def sketch_case(ctx, key, regex, algo, p, k, n_g, length):
    buf, stride = bench.make_packed_genomes(torch, DEV, n_g, length, 42, 0)
    spans = (Span * n_g)()
    for i in range(n_g):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    with ops.Sketcher(ctx, algo, p, k, 42, n_g) as sk:
        sk.push_raw(buf.data_ptr(), n_g * stride, spans, n_g, None, 0, dev=True)
        regs = sk.fetch()
    emit(key=key, kernels=[regex], units=n_g * (length - k + 1), unit="kmer", launches=1,
         shape=f"{n_g} genomes x {length} bp, p={p} k={k}")
    return regs


def reads_case(ctx, key, regex, p, k, n_reads, read_len):
    n_bases = n_reads * read_len
    g = torch.Generator(device=DEV)
    g.manual_seed(7)
    packed = torch.randint(0, 256, (padded_bytes(n_bases) + 64,), dtype=torch.uint8, device=DEV, generator=g)
    spans = (Span * 1)(Span(0, 0, n_bases, 0, n_reads, read_len))
    with ops.Sketcher(ctx, ALGO_ULL, p, k, 42, 1) as sk:
        sk.push_raw(packed.data_ptr(), padded_bytes(n_bases), spans, 1, None, 0, dev=True)
        sk.sync()
    emit(key="build_invalid_mask_kernel", kernels=["build_invalid_mask_kernel"], units=n_reads, unit="record", launches=1,
         shape="the boundary-mask launch of the reads push below")
    emit(key=key, kernels=[regex], units=n_reads * (read_len - k + 1), unit="kmer", launches=1,
         shape=f"{n_reads} reads x {read_len} bp of one sample, ULL p={p} k={k}")


def text_case(ctx, n_g, length):
    """FASTA bodies (80 columns + line feed) of n_g genomes as device-resident text -> the three pack kernels + the sketch."""
    buf, stride = bench.make_packed_genomes(torch, DEV, n_g, length, 42, 0)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=DEV)
    per = length + length // 80 + 1
    t_stride = (per + 15) // 16 * 16 + 16
    text = torch.full((n_g * t_stride + 64,), 10, dtype=torch.uint8, device=DEV)
    for i in range(n_g):
        pk = buf[i * stride: i * stride + (length + 3) // 4]
        codes = torch.stack([(pk >> 6) & 3, (pk >> 4) & 3, (pk >> 2) & 3, pk & 3], dim=1).reshape(-1)[:length]
        asc = lut[codes.long()]
        rows = torch.cat([asc[: length // 80 * 80].view(-1, 80), torch.full((length // 80, 1), 10, dtype=torch.uint8, device=DEV)], dim=1).reshape(-1)
        text[i * t_stride: i * t_stride + rows.numel()] = rows
        text[i * t_stride + rows.numel(): i * t_stride + rows.numel() + (length % 80)] = asc[length // 80 * 80:]
    spans = (TextSpan * n_g)()
    for i in range(n_g):
        spans[i] = TextSpan(i, i * t_stride, per, 1, 0)
    with ops.Sketcher(ctx, ALGO_ULL, 10, 16, 42, n_g) as sk:
        check(capi.lib().lash_sketch_push_ascii_dev(sk._h, C.c_void_p(text.data_ptr()), n_g * t_stride, spans, n_g, None))
        sk.sync()
    emit(key="text_pack_kernels", kernels=["text_count_kernel", "text_scan_kernel", "text_compact_kernel"], units=n_g * per, unit="text byte",
         launches=3, shape=f"{n_g} FASTA bodies x {length} bp (80 columns)")
    emit(key="sketch_kernel<ULL,k16,smem,clipped>", kernels=[r"sketch_kernel<2, 0, 0, 256>"], units=n_g * (length - 15), unit="kmer", launches=1,
         shape="the sketch launch of the same text push")


def dist_case(ctx, key, kernels, algo, p, k, est, regs, launches):
    n = regs.shape[0]
    ops.dist(ctx, algo, p, k, est, 1, False, regs, regs, triangular=True)
    emit(key=key, kernels=kernels, units=n * (n + 1) // 2 * regs.shape[1], unit="register pair", launches=launches, pairs=n * (n + 1) // 2,
         shape=f"{n} x {n} lower triangle, p={p}")


def main():
    with ops.Context(0) as ctx:
        regs_ull = sketch_case(ctx, "sketch_kernel<ULL,k16,smem>", r"sketch_kernel<2, 0, 0, 256>", ALGO_ULL, 10, 16, 400, 5_000_000)
        regs_hll = sketch_case(ctx, "sketch_kernel<HLL,wide,smem>", r"sketch_kernel<1, 2, 0, 768>", ALGO_HLL, 14, 21, 400, 5_000_000)
        regs_hmh = sketch_case(ctx, "sketch_kernel<HMH,k16,smem>", r"sketch_kernel<0, 0, 0, 768>", ALGO_HMH, 14, 16, 400, 2_000_000)
        reads_case(ctx, "sketch_kernel<ULL,wide,smem,reads>", r"sketch_kernel<2, 2, 0, 768>", 14, 21, 20_000_000, 150)
        text_case(ctx, 200, 5_000_000)
        regs_small = sketch_case(ctx, "sketch_kernel<ULL,k16,smem>@100kbp", r"sketch_kernel<2, 0, 0, 256>", ALGO_ULL, 10, 16, 8000, 100_000)
        dist_case(ctx, "dist_fgra_tab_kernel@n=1000", ["dist_fgra_tab_kernel"], ALGO_ULL, 10, 16, EST_FGRA, np.tile(regs_ull, (3, 1))[:1000], 1)
        dist_case(ctx, "dist_fgra_tab_kernel", ["dist_fgra_tab_kernel"], ALGO_ULL, 10, 16, EST_FGRA, regs_small, 1)
        dist_case(ctx, "dist_ml_tab_kernel+ml_finish_kernel", ["dist_ml_tab_kernel", "ml_finish_kernel"], ALGO_ULL, 10, 16, EST_ML, regs_small, 2)
        dist_case(ctx, "dist_hll_int_kernel", ["dist_hll_int_kernel"], ALGO_HLL, 14, 21, 0, np.tile(regs_hll, (10, 1)), 1)
        dist_case(ctx, "dist_hmh_fast_kernel", ["dist_hmh_fast_kernel"], ALGO_HMH, 14, 16, 0, np.tile(regs_hmh, (5, 1)), 1)


if __name__ == "__main__":
    main()
