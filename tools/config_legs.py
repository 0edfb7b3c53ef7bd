"""BASELINE.json configs[2..4] as legs of bench.py (any N; the driver's SCALE run shows them at N = 1, 2, 4, 8).

  C3  HLL p=14 k=21 sketch of 10,000 synthetic 5 Mbp genomes sharded across the ranks (shard.genome_shards), then the
      row-sharded poisson all-vs-all (dist_hll_int_kernel)                        -- src/utils.rs:449-510, 342-370
  C4  ONE sample of 100 Gbp of 150 bp reads split across the ranks (read groups), every rank sketches its share into a
      ULL p=14 accumulator, NCCL all-gather of world x 2^14 B, lash_sketch_merge_dev (UltraLogLog::merge, utils.rs:260)
  C5  100k x 100k ULL p=10, ML estimator, --dm shape: row ranges per rank, lash_dist_stream_rows -> pinned host blocks
                                                                                   -- src/utils.rs:248-285

These legs are STRONG-scaled: the job (inputs, outputs) is the same for every N, so the hash of the global-order
registers and the order-free checksum of all distances must be identical for N = 1, 2, 4, 8 -- bench.py
prints them and `equal_across_n_keys` names the fields to compare.  Every leg also checks itself against the CPU oracle
on a bounded sample (rank 0).  All compute goes through liblash_gpu.so; torch is plumbing (device memory, RNG for the
synthetic inputs, events, torch.distributed).
"""
from __future__ import annotations

import ctypes as C
import math
import time

import numpy as np

from lash_b200 import ALGO_HLL, ALGO_ULL, EST_FGRA, EST_ML, capi, ops, shard
from lash_b200.capi import Span, check
from lash_b200.pack import padded_bytes

SEED = 42
EPS = float(np.finfo(np.float64).eps)


class Env:
    """What a leg needs from bench.py: torch, the process group, this rank's context and stream."""

    def __init__(self, torch, dist, rank, world, local, device, ctx, stream, sm_mhz, costs):
        self.torch, self.dist, self.rank, self.world, self.local = torch, dist, rank, world, local
        self.device, self.ctx, self.stream, self.sptr = device, ctx, stream, stream.cuda_stream
        self.sm_mhz, self.costs = sm_mhz, costs
        self.L = capi.lib()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.device)

    def max_f64(self, vals):
        t = self.torch.tensor(list(vals), dtype=self.torch.float64, device=self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(x) for x in t.tolist()]

    def sum_i64(self, vals):
        """Wrapping (two's complement) sum over ranks of unsigned 64-bit values."""
        a = np.array([int(v) & (2**64 - 1) for v in vals], dtype=np.uint64).view(np.int64)
        t = self.torch.from_numpy(a.copy()).to(self.device)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return [int(x) & (2**64 - 1) for x in t.cpu().numpy().view(np.uint64).tolist()]

    def issue_frac(self, kernel_key, units_per_s):
        """Fraction of the SM issue slots (148 SMs x 4 schedulers x 32 lanes x f) the kernel keeps busy, from the
        executed instructions per unit of the committed ncu capture of THIS build (None when the capture is stale)."""
        c = (self.costs or {}).get(kernel_key)
        if not c or not c.get("warp_inst_x32_per_unit"):
            return None
        peak = 148 * 4 * 32 * self.sm_mhz * 1e6
        return {"kernel": kernel_key, "bound": c.get("bound", "sm_issue"), "inst_per_unit": c["warp_inst_x32_per_unit"],
                "unit": c.get("unit"), "units_per_s": units_per_s, "peak_lane_inst_per_s": peak,
                "frac": units_per_s * c["warp_inst_x32_per_unit"] / peak, "capture": c.get("capture")}


def as_tensor(torch, ptr: int, nbytes: int, device):
    """A raw device pointer owned by liblash_gpu.so as a torch uint8 tensor (no copy)."""
    class _Iface:
        pass
    o = _Iface()
    o.__cuda_array_interface__ = {"shape": (nbytes,), "typestr": "|u1", "data": (ptr, False), "version": 3, "strides": None}
    return torch.as_tensor(o, device=device)


def unpack_to_ascii(packed: np.ndarray, n_bases: int) -> bytes:
    c = np.stack([(packed >> 6) & 3, (packed >> 4) & 3, (packed >> 2) & 3, packed & 3], axis=1).reshape(-1)[:n_bases]
    return np.frombuffer(b"ACGT", dtype=np.uint8)[c].tobytes()


def reg_hash(torch, regs_u8) -> int:
    """Position-dependent 64-bit hash of a register array in GLOBAL genome order (wrapping multiply-add on the device)."""
    n = regs_u8.numel() // 8
    w = regs_u8.reshape(-1)[: n * 8].view(torch.int64)
    idx = torch.arange(n, dtype=torch.int64, device=regs_u8.device)
    h = (w * (idx * 2 + 1) + (idx ^ 0x5DEECE66D)).sum()
    return int(h.item()) & (2**64 - 1)


def pack_codes(torch, codes, out_rows):
    """codes [n, L] (uint8 in 0..3, L a multiple of 4) -> packed bytes written into out_rows[:, :L/4]."""
    q = codes.view(codes.shape[0], -1, 4)
    out_rows[:, : q.shape[1]] = (q[:, :, 0] << 6) | (q[:, :, 1] << 4) | (q[:, :, 2] << 2) | q[:, :, 3]


def mutated_batch(torch, device, anc, batch_id: int, n: int, seed: int, out_rows):
    """n genomes = the shared ancestor with substitutions at a per-genome rate (log-uniform in [1e-3, 0.3]); everything
    is a function of (seed, batch_id) only, so any rank generates the same genomes for the same batch."""
    g = torch.Generator(device=device)
    g.manual_seed(seed * 1000003 + batch_id + 1)
    length = anc.numel()
    mu = 10.0 ** (torch.rand(n, device=device, generator=g) * (math.log10(0.3) + 3.0) - 3.0)
    hit = torch.rand((n, length), device=device, generator=g) < mu[:, None]
    sub = torch.randint(1, 4, (n, length), dtype=torch.uint8, device=device, generator=g)
    codes = torch.where(hit, (anc[None, :] + sub) & 3, anc[None, :].expand(n, length))
    pad = (-length) % 4
    if pad:
        codes = torch.cat([codes, torch.zeros((n, pad), dtype=torch.uint8, device=device)], dim=1)
    pack_codes(torch, codes.contiguous(), out_rows)


def dist_tolerance(d, k):
    """|gpu - oracle| bound of the f64 parity tests (tests/test_gpu_dist.py): 1e-12 relative plus the amplification of a
    1-ulp difference in the union estimate through s = (a + b - U) / U -> d = -ln(2s/(1+s))/k."""
    s = np.maximum(np.exp(-np.asarray(d) * k) / 2.0, 1e-300)
    return 1e-12 * np.abs(d) + 64.0 * EPS / (s * k)


def error_stats(got, exp, k):
    """f64 distance error of GPU cells against the oracle's: the tolerance verdict plus the raw figures."""
    got = np.asarray(got, dtype=np.float64).reshape(-1)
    exp = np.asarray(exp, dtype=np.float64).reshape(-1)
    err = np.abs(got - exp)
    rel = np.where(err == 0, 0.0, err / np.maximum(np.abs(exp), 1e-300))
    ok = bool(np.all((err <= dist_tolerance(exp, k)) | (got == exp)))
    strict = rel <= 1e-12
    s = np.exp(-exp * k) / (2.0 - np.exp(-exp * k))               # poisson model: frac = exp(-d k), s = frac / (2 - frac)
    return {"ok": ok, "cells": int(len(exp)), "max_rel_err": float(rel.max()) if len(rel) else 0.0,
            "frac_within_1e-12": float(strict.mean()) if len(rel) else 1.0, "frac_bit_identical": float((got == exp).mean()) if len(rel) else 1.0,
            "smallest_s_where_1e-12_always_holds": float(s[~strict].max()) if (~strict).any() else 0.0}


def spot_check_dist(O, algo, p, k, est, regs_of, cells, got):
    """cells: [(i, j)], got: GPU distances; regs_of(i) -> register row.  Returns error_stats()."""
    exp = np.array([O.dist(algo, p, k, est, O.POISSON, False, regs_of(i)[None, :], regs_of(j)[None, :])[0, 0] for i, j in cells])
    return error_stats(got, exp, k)


# ----------------------------------------------------------------------------------------------------------------------
# C3: HLL p=14 k=21, 10,000 genomes sharded across the ranks, poisson all-vs-all
# ----------------------------------------------------------------------------------------------------------------------
def leg_c3(env: Env, n_total=10_000, length=5_000_000, steps=3):
    torch, dist = env.torch, env.dist
    P, K = 14, 21
    rb = 1 << P
    shards = shard.genome_shards([length] * n_total, env.world)
    mine = shards[env.rank]
    n_max = max(len(s) for s in shards)
    perm = shard.gather_permutation(shards)          # global genome g sits at row perm[g] of the rank-concatenated array
    import bench
    buf, stride = bench.make_packed_genomes_ids(torch, env.device, mine, length, SEED)
    spans = (Span * max(len(mine), 1))()
    for i in range(len(mine)):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    sk = ops.Sketcher(env.ctx, ALGO_HLL, P, K, SEED, n_max)
    sk.set_stream(env.sptr)
    regs_view = as_tensor(torch, sk.regs_dev(), n_max * rb, env.device)
    gath = torch.empty(env.world * n_max * rb, dtype=torch.uint8, device=env.device) if env.world > 1 else None
    perm_t = torch.from_numpy(np.asarray(perm, dtype=np.int64)).to(env.device)
    card = torch.empty(n_total, dtype=torch.float64, device=env.device)
    rows = shard.row_shard(n_total, env.rank, env.world, triangular=True)
    out = torch.empty(n_total * (n_total + 1) // 2, dtype=torch.float64, device=env.device)
    flags = torch.zeros(1, dtype=torch.int32, device=env.device)
    sums = torch.zeros(2, dtype=torch.int64, device=env.device)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    state = {}

    def step():
        ev[0].record(env.stream)
        sk.reset()
        sk.push_raw(buf.data_ptr(), len(mine) * stride, spans, len(mine), None, 0, dev=True)
        ev[1].record(env.stream)
        if env.world > 1:
            dist.all_gather_into_tensor(gath, regs_view)
            regs_all = gath.view(env.world * n_max, rb).index_select(0, perm_t)
        else:
            regs_all = regs_view.view(n_max, rb)
        ev[2].record(env.stream)
        src = C.c_void_p(regs_all.data_ptr())
        check(env.L.lash_cardinality_dev(env.ctx.handle, ALGO_HLL, P, 0, src, n_total, C.c_void_p(card.data_ptr()), C.c_void_p(env.sptr)))
        check(env.L.lash_dist_dev(env.ctx.handle, ALGO_HLL, P, K, 0, 1, 0, src, n_total, src, n_total, C.c_void_p(card.data_ptr()),
                                  C.c_void_p(card.data_ptr()), 1, rows[0], rows[1], C.c_void_p(out.data_ptr()),
                                  C.c_void_p(flags.data_ptr()), C.c_void_p(env.sptr)))
        ev[3].record(env.stream)
        state["regs_all"] = regs_all

    step()
    env.barrier()
    k_ms0, _ = sk.stats()
    # one more untimed step AFTER the barrier and without a sync: the host then runs ahead of the GPU, and the events below
    # bracket K steps of steady-state device work instead of the host's first enqueue (tile plan of thousands of spans)
    step()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record(env.stream)
    for it in range(steps):
        step()
        marks[it + 1].record(env.stream)
    env.barrier()
    k_ms1, _ = sk.stats()
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    total, sk_ms, ga_ms, di_ms, sk_kernel = env.max_f64([marks[0].elapsed_time(marks[-1]) / steps, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
                                                         ev[2].elapsed_time(ev[3]), (k_ms1 - k_ms0) / (steps + 1)])
    regs_all = state["regs_all"]
    # order-free checksum of this rank's rows, summed over the ranks; hash of the global-order registers
    sums.zero_()
    check(env.L.lash_dist_checksum_dev(env.ctx.handle, 0, C.c_void_p(out.data_ptr()), n_total, 1, rows[0], rows[1],
                                       C.c_void_p(sums.data_ptr()), C.c_void_p(env.sptr)))
    torch.cuda.synchronize(env.device)
    s = sums.cpu().numpy().view(np.uint64)
    dsum, dcells = env.sum_i64([int(s[0]), int(s[1])])
    rhash = reg_hash(torch, regs_all)
    n_pairs = n_total * (n_total + 1) // 2
    parity = None
    if env.rank == 0:
        import oracle as O
        gsel = mine[:2]
        gen = [[unpack_to_ascii(buf[i * stride: i * stride + (length + 3) // 4].cpu().numpy(), length)] for i in range(len(gsel))]
        exp = O.sketch_genomes(O.HLL, P, K, SEED, gen, threads=2)
        regs_ok = bool(np.array_equal(exp, regs_all[torch.tensor(gsel, device=env.device)].cpu().numpy()))
        rng = np.random.default_rng(3)
        ii = rng.integers(rows[0], rows[1], size=64)
        cells = [(int(i), int(rng.integers(0, i + 1))) for i in ii]
        o_idx = torch.tensor([i * (i + 1) // 2 + j for i, j in cells], device=env.device)
        got = out[o_idx].cpu().numpy()
        need = sorted({x for c in cells for x in c})
        host_regs = dict(zip(need, regs_all[torch.tensor(need, device=env.device)].cpu().numpy()))
        st = spot_check_dist(O, O.HLL, P, K, 0, lambda i: host_regs[i], cells, got)
        # the leading 96 x 96 triangle (rank 0's rows start at 0), every cell
        nb = min(96, rows[1])
        r_idx = torch.arange(nb, device=env.device)
        blk_idx = (r_idx * (r_idx + 1) // 2)[:, None] + r_idx[None, :]
        tri_ok = np.tril(np.ones((nb, nb), dtype=bool))
        got_blk = out[blk_idx.clamp(max=out.numel() - 1)].cpu().numpy()
        lead_regs = regs_all[:nb].cpu().numpy()
        exp_full = O.dist(O.HLL, P, K, 0, O.POISSON, False, lead_regs, lead_regs, threads=8)
        blk = error_stats(got_blk[tri_ok], exp_full[tri_ok], K)
        parity = {"ok": bool(regs_ok and st["ok"] and blk["ok"] and dcells == n_pairs), "registers_bit_exact_2_genomes": regs_ok,
                  "dist_64_random_cells": st, "dist_block": blk, "cells_covered_once": dcells == n_pairs}
    sk_gbps = len(mine) * length / (sk_kernel * 1e-3) / 1e9   # this rank's kernel rate
    res = {"config": "configs[2]: HLL p=14 k=21 sketch of 10,000 synthetic 5 Mbp genomes sharded across the GPUs, poisson-model all-vs-all (lower triangle, f64)",
           "scaling": "strong", "genomes": n_total, "genome_len": length, "genomes_per_gpu": len(mine),
           "gbp_per_s": n_total * length / (total * 1e-3) / 1e9, "pairs_per_s": n_pairs / (di_ms * 1e-3), "pairs": n_pairs,
           "ms_per_step": total, "phases_ms": {"sketch": sk_ms, "gather+permute": ga_ms, "cardinality+dist": di_ms},
           "step_ms_rank0": step_ms, "sketch_kernel_ms": sk_kernel, "sketch_kernel_gbp_per_s_per_gpu": sk_gbps,
           "register_merges_per_s": n_pairs * rb / (di_ms * 1e-3),
           "roofline_sketch_kernel": env.issue_frac("sketch_kernel<HLL,wide,smem>", len(mine) * (length - K + 1) / (sk_kernel * 1e-3)),
           "roofline_dist_hll_int_kernel": env.issue_frac("dist_hll_int_kernel", shard.pair_count(rows, n_total, True) * rb / (di_ms * 1e-3)),
           "registers_hash": f"{rhash:016x}", "dist_checksum": f"{dsum:016x}", "dist_cells": dcells,
           "hll_bias_flags": int(flags.item()), "parity": parity}
    sk.close()
    del buf, out, gath, regs_all, state
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------------------------------------------------
# C4: one sample of 100 Gbp of 150 bp reads -> ONE ULL p=14 sketch; shares per rank, all-gather, merge
# ----------------------------------------------------------------------------------------------------------------------
READ_LEN = 150
GROUP_READS = 32                       # 32 reads = 4800 bases = 1200 bytes: share / chunk cuts stay 16-byte aligned
GROUP_BYTES = GROUP_READS * READ_LEN // 4
CHUNK_GROUPS = 1 << 19                 # 16,777,216 reads = 2.52 Gbp = 629 MB packed per chunk


def c4_chunk(torch, device, c: int, groups: int):
    """Chunk c of the sample: `groups` read groups of uniformly random bases, a function of c only."""
    g = torch.Generator(device=device)
    g.manual_seed(977 * SEED + c)
    return torch.randint(0, 256, (groups * GROUP_BYTES + 64,), dtype=torch.uint8, device=device, generator=g)


def leg_c4(env: Env, total_bases=100_000_000_000, steps=3, e2e=True):
    torch, dist = env.torch, env.dist
    P, K = 14, 21
    rb = 1 << P
    n_groups = -(-total_bases // (GROUP_READS * READ_LEN))
    g0, g1 = shard.read_shard(n_groups, env.rank, env.world)      # this rank's read groups
    pieces = []   # (chunk tensor, byte offset, reads)
    for c in range(g0 // CHUNK_GROUPS, -(-g1 // CHUNK_GROUPS)):
        cg0, cg1 = c * CHUNK_GROUPS, min((c + 1) * CHUNK_GROUPS, n_groups)
        a, b = max(g0, cg0), min(g1, cg1)
        if b > a:
            pieces.append((c4_chunk(torch, env.device, c, cg1 - cg0), (a - cg0) * GROUP_BYTES, (b - a) * GROUP_READS))
    my_bases = (g1 - g0) * GROUP_READS * READ_LEN
    sk = ops.Sketcher(env.ctx, ALGO_ULL, P, K, SEED, 1)
    sk.set_stream(env.sptr)
    acc = as_tensor(torch, sk.regs_dev(), rb, env.device)
    gath = torch.empty(env.world * rb, dtype=torch.uint8, device=env.device)
    merged = torch.empty(rb, dtype=torch.uint8, device=env.device)
    span_of = [(Span * 1)(Span(0, off, reads * READ_LEN, 0, reads, READ_LEN)) for _, off, reads in pieces]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]

    def step():
        ev[0].record(env.stream)
        sk.reset()
        for (t, off, reads), sp in zip(pieces, span_of):
            sk.push_raw(t.data_ptr(), off + padded_bytes(reads * READ_LEN), sp, 1, None, 0, dev=True)
        ev[1].record(env.stream)
        if env.world > 1:
            dist.all_gather_into_tensor(gath, acc)                 # world x 16 KiB: the path's one exchange step
            merged.copy_(gath[:rb])
            for r in range(1, env.world):                          # UltraLogLog::merge, NOT a byte max
                check(env.L.lash_sketch_merge_dev(env.ctx.handle, ALGO_ULL, P, C.c_void_p(merged.data_ptr()),
                                                  C.c_void_p(gath[r * rb:].data_ptr()), 1, C.c_void_p(env.sptr)))
        else:
            merged.copy_(acc)
        ev[2].record(env.stream)

    step()
    env.barrier()
    k_ms0, _ = sk.stats()
    # one more untimed step AFTER the barrier and without a sync: the host then runs ahead of the GPU, and the events below
    # bracket K steps of steady-state device work instead of the host's first enqueue (tile plan of thousands of spans)
    step()
    marks = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
    marks[0].record(env.stream)
    for it in range(steps):
        step()
        marks[it + 1].record(env.stream)
    env.barrier()
    k_ms1, _ = sk.stats()
    step_ms = [marks[i].elapsed_time(marks[i + 1]) for i in range(steps)]
    total, sk_ms, mg_ms, sk_kernel = env.max_f64([marks[0].elapsed_time(marks[-1]) / steps, ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]),
                                                  (k_ms1 - k_ms0) / (steps + 1)])
    rhash = reg_hash(torch, merged)
    nonzero = int((merged != 0).sum().item())
    parity = None
    if env.rank == 0:
        import oracle as O
        # (1) a bounded sub-sample of the same reads through the same path vs the oracle: 32,768 reads = 4.9 Mbp
        n_chk = 1 << 15
        chunk0, off0, _ = pieces[0]
        sk2 = ops.Sketcher(env.ctx, ALGO_ULL, P, K, SEED, 1)
        sp = (Span * 1)(Span(0, off0, n_chk * READ_LEN, 0, n_chk, READ_LEN))
        sk2.push_raw(chunk0.data_ptr(), off0 + padded_bytes(n_chk * READ_LEN), sp, 1, None, 0, dev=True)
        got = sk2.fetch()
        sk2.close()
        host = chunk0[off0: off0 + n_chk * READ_LEN // 4].cpu().numpy()
        asc = unpack_to_ascii(host, n_chk * READ_LEN)
        exp = O.sketch_genomes(O.ULL, P, K, SEED, [[asc[i * READ_LEN:(i + 1) * READ_LEN] for i in range(n_chk)]], threads=1)
        regs_ok = bool(np.array_equal(exp, got))
        # (2) the exchange step: the GPU fold of the gathered shares equals the oracle's UltraLogLog::merge of them
        if env.world > 1:
            parts = gath.view(env.world, rb).cpu().numpy()
        else:
            # one GPU has nothing to exchange: sketch the first piece in two shares and fold them, which must give the
            # registers of the piece sketched in one go (merge(sketch(A), sketch(B)) == sketch(A u B))
            t, off, reads = pieces[0]
            half = (reads // 2) // GROUP_READS * GROUP_READS
            sk3 = ops.Sketcher(env.ctx, ALGO_ULL, P, K, SEED, 3)
            sps = (Span * 3)(Span(0, off, half * READ_LEN, 0, half, READ_LEN),
                             Span(1, off + half * READ_LEN // 4, (reads - half) * READ_LEN, 0, reads - half, READ_LEN),
                             Span(2, off, reads * READ_LEN, 0, reads, READ_LEN))
            sk3.push_raw(t.data_ptr(), off + padded_bytes(reads * READ_LEN), sps, 3, None, 0, dev=True)
            three = sk3.fetch()
            sk3.close()
            parts = three[:2]
        fold = parts[0].copy()
        for r in range(1, parts.shape[0]):
            fold = O.ull_merge(fold, parts[r], P)
        if env.world > 1:
            merge_ok = bool(np.array_equal(fold, merged.cpu().numpy()))
        else:
            gpu_fold = ops.merge(env.ctx, ALGO_ULL, P, parts[0][None, :], parts[1][None, :])[0]
            merge_ok = bool(np.array_equal(fold, gpu_fold) and np.array_equal(fold, three[2]))
        parity = {"ok": bool(regs_ok and merge_ok), "registers_bit_exact_32768_reads": regs_ok,
                  "merge_of_shares_equals_oracle_merge": merge_ok}
    res = {"config": "configs[3]: metagenome, 100 Gbp of synthetic 150 bp reads streamed into one ULL p=14 k=21 sketch per sample "
                     "(packed reads resident in HBM; reads of the sample split across the GPUs, all-gather + UltraLogLog merge)",
           "scaling": "strong", "bases": n_groups * GROUP_READS * READ_LEN, "reads": n_groups * GROUP_READS, "bases_per_gpu": my_bases,
           "pushes_per_gpu": len(pieces), "gbp_per_s": n_groups * GROUP_READS * READ_LEN / (total * 1e-3) / 1e9, "ms_per_step": total,
           "phases_ms": {"mask+sketch": sk_ms, "gather+merge": mg_ms}, "step_ms_rank0": step_ms, "kernel_ms(mask+sketch)": sk_kernel,
           "kernel_gbp_per_s_per_gpu": my_bases / (sk_kernel * 1e-3) / 1e9,
           "roofline_sketch_kernel": env.issue_frac("sketch_kernel<ULL,wide,smem,reads>", (g1 - g0) * GROUP_READS * (READ_LEN - K + 1) / (sk_kernel * 1e-3)),
           "registers_hash": f"{rhash:016x}", "nonzero_registers": nonzero, "parity": parity}
    # the same share from PINNED HOST memory (H2D inside the timed region): one pinned chunk re-pushed
    if e2e and pieces:
        sk.set_stream(None)
        t, off, reads = pieces[0]
        nb = padded_bytes(reads * READ_LEN)
        pin = C.c_void_p()
        check(env.L.lash_host_alloc(nb + 64, C.byref(pin)))
        host = np.ctypeslib.as_array(C.cast(pin, C.POINTER(C.c_uint8)), shape=(nb + 64,))
        host[:nb] = t[off: off + nb].cpu().numpy()
        sp = (Span * 1)(Span(0, 0, reads * READ_LEN, 0, reads, READ_LEN))
        n_push = len(pieces)

        def e2e_step():
            check(env.L.lash_sketch_reset(sk._h))
            for _ in range(n_push):
                check(env.L.lash_sketch_push(sk._h, pin, nb, sp, 1, None, 0, None))
            check(env.L.lash_sketch_sync(sk._h))
        e2e_step()
        env.barrier()
        w0 = time.perf_counter()
        e2e_step()
        torch.cuda.synchronize(env.device)
        wall = env.max_f64([time.perf_counter() - w0])[0]
        res["e2e_h2d"] = {"gbp_per_s": env.world * n_push * reads * READ_LEN / wall / 1e9, "wall_s": wall,
                          "h2d_bytes_per_gpu": n_push * nb, "note": "one pinned chunk of this rank's share re-pushed for every chunk of the share (throughput only)"}
        check(env.L.lash_host_free(pin))
    sk.close()
    del pieces, gath, merged
    torch.cuda.empty_cache()
    return res


# ----------------------------------------------------------------------------------------------------------------------
# C5: 100k x 100k ULL p=10 ML --dm, row ranges per rank, streamed to pinned host blocks
# ----------------------------------------------------------------------------------------------------------------------
def leg_c5(env: Env, n_total=100_000, length=100_000, batch=500, est=EST_ML):
    torch, dist = env.torch, env.dist
    P, K = 10, 16
    rb = 1 << P
    n_batches = n_total // batch
    bshards = shard.genome_shards([batch] * n_batches, env.world)   # whole batches per rank
    mine_b = bshards[env.rank]
    nb_max = max(len(s) for s in bshards)
    stride = padded_bytes(length)
    g = torch.Generator(device=env.device)
    g.manual_seed(SEED + 5)
    anc = torch.randint(0, 4, (length,), dtype=torch.uint8, device=env.device, generator=g)
    buf = torch.zeros(len(mine_b) * batch * stride + 64, dtype=torch.uint8, device=env.device)
    rows_v = buf[: len(mine_b) * batch * stride].view(len(mine_b) * batch, stride)
    for t, b in enumerate(mine_b):
        mutated_batch(torch, env.device, anc, b, batch, SEED, rows_v[t * batch:(t + 1) * batch])
    n_loc = len(mine_b) * batch
    spans = (Span * max(n_loc, 1))()
    for i in range(n_loc):
        spans[i] = Span(i, i * stride, length, 0, 1, 0)
    sk = ops.Sketcher(env.ctx, ALGO_ULL, P, K, SEED, nb_max * batch)
    sk.set_stream(env.sptr)
    regs_view = as_tensor(torch, sk.regs_dev(), nb_max * batch * rb, env.device)
    # global genome id of batch b, slot t: b * batch + t
    shards_g = [[b * batch + t for b in bs for t in range(batch)] for bs in bshards]
    perm_t = torch.from_numpy(np.asarray(shard.gather_permutation(shards_g, pad_to=nb_max * batch), dtype=np.int64)).to(env.device)
    env.barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record(env.stream)
    sk.push_raw(buf.data_ptr(), n_loc * stride, spans, n_loc, None, 0, dev=True)
    e1.record(env.stream)
    if env.world > 1:
        gath = torch.empty(env.world * nb_max * batch * rb, dtype=torch.uint8, device=env.device)
        dist.all_gather_into_tensor(gath, regs_view)
        regs_all = gath.view(-1, rb).index_select(0, perm_t)
        del gath
    else:
        regs_all = regs_view.view(-1, rb)[:n_total]
    e2.record(env.stream)
    host_regs_t = torch.empty((n_total, rb), dtype=torch.uint8, pin_memory=True)
    host_regs_t.copy_(regs_all)
    env.barrier()
    sk_ms, ga_ms = env.max_f64([e0.elapsed_time(e1), e1.elapsed_time(e2)])
    rhash = reg_hash(torch, regs_all)
    host_regs = host_regs_t.numpy()
    rows = shard.row_shard(n_total, env.rank, env.world, triangular=True)
    # cells to compare with the oracle (rank 0), captured from the pinned blocks as they are delivered
    rng = np.random.default_rng(5)
    want = sorted({(int(i), int(rng.integers(0, i + 1))) for i in rng.integers(rows[0], rows[1], size=64)}) if env.rank == 0 else []
    got = {}
    seen = {"blocks": 0, "rows": 0}
    lead = {"block": None}

    def _cb(user, row0, nrows, ptr):
        row0, nrows = int(row0), int(nrows)
        seen["blocks"] += 1
        seen["rows"] += nrows
        if want and row0 == rows[0] and lead["block"] is None:
            nb, nc = min(128, nrows), min(128, n_total)
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(nrows * n_total,))
            lead["block"] = arr.reshape(nrows, n_total)[:nb, :nc].copy()
        if want:
            dptr = C.cast(ptr, C.POINTER(C.c_double))
            for (i, j) in want:
                if row0 <= i < row0 + nrows:
                    got[(i, j)] = dptr[(i - row0) * n_total + j]
        return 0

    cb = capi.DIST_BLOCK_CB(_cb)
    hr = host_regs.ctypes.data_as(C.c_void_p)
    check(env.L.lash_dist_set_checksum(env.ctx.handle, 1))
    # warm the staging buffers (two pinned 256 MiB blocks, ML scratch) on the first rows of the range: page-locking is not
    # part of the path
    rpb = max(1, (256 << 20) // (n_total * 8))     # the library's default block: 256 MiB of f64 rows
    warm_end = min(rows[1], rows[0] + 2 * rpb + 1)
    want_saved, want = want, []
    check(env.L.lash_dist_stream_rows(env.ctx.handle, ALGO_ULL, P, K, est, 1, 0, hr, n_total, hr, n_total, 1, rows[0], warm_end, 0, cb, None))
    want = want_saved
    seen["blocks"] = seen["rows"] = 0
    env.barrier()
    w0 = time.perf_counter()
    check(env.L.lash_dist_stream_rows(env.ctx.handle, ALGO_ULL, P, K, est, 1, 0, hr, n_total, hr, n_total, 1, rows[0], rows[1], 0, cb, None))
    wall_rank = time.perf_counter() - w0
    k_ms, launches = ops.dist_stats(env.ctx)
    cs, cc = C.c_uint64(), C.c_uint64()
    check(env.L.lash_dist_checksum(env.ctx.handle, C.byref(cs), C.byref(cc)))
    check(env.L.lash_dist_set_checksum(env.ctx.handle, 0))
    wall, k_ms_max = env.max_f64([wall_rank, k_ms])
    dsum, dcells = env.sum_i64([cs.value, cc.value])
    n_pairs = n_total * (n_total + 1) // 2
    my_pairs = shard.pair_count(rows, n_total, True)
    parity = None
    if env.rank == 0:
        import oracle as O
        gsel = [0, 1]
        gen = [[unpack_to_ascii(rows_v[i, : (length + 3) // 4].cpu().numpy(), length)] for i in gsel]
        exp = O.sketch_genomes(O.ULL, P, K, SEED, gen, threads=2)
        gids = [mine_b[0] * batch + i for i in gsel]
        regs_ok = bool(np.array_equal(exp, host_regs[gids]))
        cells = sorted(got)
        st = spot_check_dist(O, O.ULL, P, K, O.ML if est == EST_ML else O.FGRA, lambda i: host_regs[i], cells, [got[c] for c in cells])
        blk = None
        if lead["block"] is not None:   # the first 128 rows of this rank x the first 128 columns, every cell
            nb = lead["block"].shape[0]
            exp_full = O.dist(O.ULL, P, K, O.ML if est == EST_ML else O.FGRA, O.POISSON, False, host_regs[rows[0]: rows[0] + nb],
                              host_regs[: lead["block"].shape[1]], threads=8)
            tri_ok = np.arange(lead["block"].shape[1])[None, :] <= (rows[0] + np.arange(nb))[:, None]
            blk = error_stats(lead["block"][tri_ok], exp_full[tri_ok], K)
        parity = {"ok": bool(regs_ok and st["ok"] and (blk is None or blk["ok"]) and len(cells) == len(want) and dcells == n_pairs),
                  "registers_bit_exact_2_genomes": regs_ok, "dist_64_random_cells": st, "dist_block": blk,
                  "cells_covered_once": dcells == n_pairs}
    est_name = "ML" if est == EST_ML else "FGRA"
    kern = "dist_ml_tab_kernel+ml_finish_kernel" if est == EST_ML else "dist_fgra_tab_kernel"
    res = {"config": f"configs[4]: all-vs-all dist of 100k x 100k ULL p=10 sketches, {est_name} estimator, output rows tiled across the GPUs "
                     "(--dm shape: lower triangle, f64, streamed to pinned host row blocks)",
           "scaling": "strong", "sketches": n_total, "pairs": n_pairs, "pairs_this_gpu": my_pairs, "rows_this_gpu": list(rows),
           "pairs_per_s": n_pairs / wall, "wall_s": wall, "dist_kernel_ms_max_rank": k_ms_max,
           "pairs_per_s_kernel_per_gpu": my_pairs / (k_ms * 1e-3), "register_merges_per_s": n_pairs * rb / wall,
           "d2h_bytes_this_gpu": int(sum(min(n_total, min(rows[1], r0 + rpb)) * (min(rows[1], r0 + rpb) - r0) * 8
                                         for r0 in range(rows[0], rows[1], rpb))),
           "blocks": seen["blocks"], "rows_delivered": seen["rows"], "launches": int(launches),
           "inputs": {"sketch_ms": sk_ms, "gather+permute_ms": ga_ms, "genomes": n_total, "genome_len": length,
                      "sketch_gbp_per_s": n_total * length / (sk_ms * 1e-3) / 1e9},
           f"roofline_{kern}": env.issue_frac(kern, my_pairs * rb / (k_ms * 1e-3)),
           "registers_hash": f"{rhash:016x}", "dist_checksum": f"{dsum:016x}", "dist_cells": dcells, "parity": parity}
    sk.close()
    del buf, rows_v, regs_all, host_regs_t
    torch.cuda.empty_cache()
    return res


EQUAL_ACROSS_N = ["c3.registers_hash", "c3.dist_checksum", "c3.dist_cells", "c4.registers_hash", "c5.registers_hash", "c5.dist_checksum",
                  "c5.dist_cells"]
