"""Host-side logic and the C-ABI surface, without a GPU (-m "not gpu")."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_loads_and_exports_every_declared_symbol():
    from lash_b200 import capi
    L = capi.lib()
    header = open(os.path.join(ROOT, "include", "lash_gpu.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(lash_[a-z0-9_]+)\s*\(", header))
    declared -= {"lash_dist_block_cb"}
    assert len(declared) >= 25
    for name in sorted(declared):
        assert hasattr(L, name), f"{name} declared in include/lash_gpu.h but not exported"
    assert declared == set(capi.PROTOTYPES), declared ^ set(capi.PROTOTYPES)
    out = subprocess.run(["nm", "-D", "--defined-only", capi.lib_path()], capture_output=True, text=True).stdout
    exported = set(re.findall(r" T (lash_[a-z0-9_]+)", out))
    assert declared <= exported
    assert L.lash_gpu_abi_version() == 2


def test_library_carries_sm100a_code_only():
    from lash_b200 import capi
    out = subprocess.run(["cuobjdump", "-lelf", capi.lib_path()], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_cpu_fallback_without_a_device():
    """Without a GPU the product must fail loudly, not compute on the CPU."""
    from lash_b200 import LashError, capi
    from lash_b200.ops import Context
    if capi.lib().lash_gpu_device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(LashError, match="no CUDA device"):
        Context(0)


def test_product_never_touches_the_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "lash_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")) or f == "Makefile":
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.replace("the oracle's", "").replace("oracle/", "ORACLEDIR") or "import oracle" not in src
                assert "import oracle" not in src and "liblash_oracle" not in src and "lo_" + "dist" not in src, f


def test_static_size_helpers():
    from lash_b200 import capi
    from lash_b200.pack import padded_bytes
    L = capi.lib()
    assert L.lash_sketch_reg_bytes(capi.ALGO_HMH, 0) == 32768
    assert L.lash_sketch_reg_bytes(capi.ALGO_ULL, 10) == 1024
    assert L.lash_sketch_reg_bytes(capi.ALGO_HLL, 14) == 16384
    assert L.lash_sketch_reg_bytes(capi.ALGO_ULL, 2) == 0 and L.lash_sketch_reg_bytes(capi.ALGO_HLL, 19) == 0
    for n in (0, 1, 3, 4, 63, 64, 65, 1000, 12345):
        assert L.lash_sketch_padded_bytes(n) == padded_bytes(n)
        assert padded_bytes(n) % 16 == 0 and padded_bytes(n) >= (n + 3) // 4 + 8


def test_packer_matches_reference_front_end(oracle):
    from lash_b200.pack import PackedBatch, encode_record, pack_codes
    seq = b"ACgtNNRYGT\nAC-*TTTGACCA"
    codes = encode_record(seq)
    assert bytes(b"ACGT"[c] for c in codes) == oracle.filter_out_n(seq)
    packed = pack_codes(codes)
    # first base in the two most significant bits of byte 0
    assert packed[0] >> 6 == codes[0] and (packed[0] >> 4) & 3 == codes[1] and packed[0] & 3 == codes[3]
    b = PackedBatch()
    b.add_genome(0, [b"ACGTACGTAC", b"", b"nnnn", b"GGGTTTAAACCC"])
    b.add_genome(1, [b"ACGT" * 100])
    buf = b.buffer()
    assert b.spans[0][1] == 0 and b.spans[1][1] % 16 == 0 and len(buf) == b.n_bytes
    assert b.spans[0][2] == 22 and b.spans[0][4] == 4 and b.spans[1][4] == 1
    assert b.rec_start == [0, 10, 10, 10, 22]
    # unpack and compare with the concatenated filtered records
    def unpack(buf, off, n):
        by = buf[off: off + (n + 3) // 4]
        c = np.stack([(by >> 6) & 3, (by >> 4) & 3, (by >> 2) & 3, by & 3], axis=1).reshape(-1)[:n]
        return bytes(b"ACGT"[x] for x in c)
    assert unpack(buf, 0, 22) == b"ACGTACGTACGGGTTTAAACCC"
    assert unpack(buf, b.spans[1][1], 400) == b"ACGT" * 100


def _shard_worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lash_b200 import shard
    import torch
    sizes = [5_000_000 + 1000 * (i % 7) for i in range(37)]
    mine = shard.genome_shard(sizes, rank, world)
    rows = shard.row_shard(1001, rank, world, triangular=True)
    t = torch.zeros(37 + 1001, dtype=torch.int64)
    for g in mine:
        t[g] += 1
    t[37 + rows[0]: 37 + rows[1]] += 1
    dist.all_reduce(t)
    load = torch.tensor([float(sum(sizes[g] for g in mine)), float(shard.pair_count(rows, 1001, True))], dtype=torch.float64)
    loads = [torch.zeros(2, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(loads, load)
    # the "broadcast of the reference sketch set" step: rank 0's registers reach everyone
    regs = torch.arange(64, dtype=torch.uint8) if rank == 0 else torch.zeros(64, dtype=torch.uint8)
    dist.broadcast(regs, src=0)
    if rank == 0:
        q.put((t.tolist(), [l.tolist() for l in loads], regs.tolist()))
    dist.destroy_process_group()


def test_sharding_covers_everything_once_world_size_2():
    """N>1 host logic on CPU (gloo, world_size 2): genome shards and distance row ranges are a
    partition, balanced, and the one collective (broadcast of the sketch set) delivers rank 0's data."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    cover, loads, regs = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert cover == [1] * (37 + 1001)
    g0, g1 = loads[0][0], loads[1][0]
    assert abs(g0 - g1) / max(g0, g1) < 0.06
    p0, p1 = loads[0][1], loads[1][1]
    assert p0 + p1 == 1001 * 1002 // 2 and abs(p0 - p1) / max(p0, p1) < 0.01
    assert regs == list(range(64))


def _sample_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist

    import oracle as O
    from lash_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(77)
    genome = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=60_000)].tobytes()
    starts = rng.integers(0, len(genome) - 150, size=3001)
    reads = [genome[s:s + 150] for s in starts]
    out = {}
    for name, algo, p in (("ull", O.ULL, 12), ("hll", O.HLL, 10), ("hmh", O.HMH, 14)):
        b, e = shard.read_shard(len(reads), rank, world)
        regs = O.sketch_genomes(algo, p, 21, 42, [reads[b:e]])[0]
        mine = torch.from_numpy(regs.view(np.uint8).copy())     # registers travel as bytes (gloo has no u16)
        parts = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(parts, mine)                      # the path's one exchange step for a single sample
        parts = [t.numpy().view(regs.dtype) for t in parts]
        acc = parts[0].copy()
        for part in parts[1:]:                             # what lash_sketch_merge_dev does on the GPU
            acc = O.ull_merge(acc, part, p) if algo == O.ULL else np.maximum(acc, part)
        whole = O.sketch_genomes(algo, p, 21, 42, [reads])[0]
        out[name] = bool(np.array_equal(acc, whole))
    if rank == 0:
        q.put(out)
    dist.destroy_process_group()


def test_single_sample_in_shares_world_size_2():
    """N>1 path for ONE sample (config 4) on CPU (gloo, world_size 2): reads are split by
    shard.read_shard, partial sketches all-gathered and folded with the sketch's own merge
    (ULL: pack(unpack|unpack), HLL/HMH: max) -- the result equals the sketch of all reads."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sample_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert out == {"ull": True, "hll": True, "hmh": True}


def _perm_worker(rank, world, port, q):
    import numpy as np
    import torch
    import torch.distributed as dist

    from lash_b200 import shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sizes = [5] * 11 + [9, 1, 7]                        # 14 genomes, uneven: the shards differ in length
    shards = shard.genome_shards(sizes, world)
    n_max = max(len(s) for s in shards)
    mine = torch.zeros((n_max, 4), dtype=torch.int64)    # "registers" of genome g = [g, g, g, g]; padding rows stay 0
    for t, g in enumerate(shards[rank]):
        mine[t] = g + 1
    gath = torch.empty((world * n_max, 4), dtype=torch.int64)
    dist.all_gather_into_tensor(gath, mine)             # what bench.py's legs do over NCCL
    perm = torch.tensor(shard.gather_permutation(shards), dtype=torch.int64)
    ordered = gath.index_select(0, perm)
    if rank == 0:
        q.put(ordered[:, 0].tolist())
    dist.destroy_process_group()


def test_gathered_sketches_return_to_list_order_world_size_2():
    """bench.py configs[2] / configs[4] legs: every rank sketches its genome shard, the all-gather returns the blocks rank
    after rank (padded), and shard.gather_permutation puts them back into list order (utils.rs:507: Vec<S> in file order)."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33000 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_perm_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got == list(range(1, 15))


def test_shard_functions_edge_cases():
    from lash_b200 import shard
    # link-aware shares: proportional to the measured rates, exact total, deterministic
    assert shard.weighted_counts(8000, [23.2] * 4 + [35.2] * 4) == [795] * 4 + [1205] * 4
    for n in (0, 1, 7, 1000):
        for w in ([1.0], [1, 1, 1], [3, 1], [0, 0], [5.5, 0.0, 2.25]):
            c = shard.weighted_counts(n, w)
            assert sum(c) == n and len(c) == len(w) and all(x >= 0 for x in c)
            if sum(w) > 0:
                assert all(abs(x - n * wi / sum(w)) < 1 for x, wi in zip(c, w))
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, 5, 8, 1000):
            got = sorted(g for r in range(world) for g in shard.genome_shard([10] * n, r, world))
            assert got == list(range(n))
            shards = shard.genome_shards([10] * n, world)
            assert shards == [list(range(r, n, world)) for r in range(world)]     # equal sizes: round robin
            perm = shard.gather_permutation(shards)
            pad = max(len(s) for s in shards)
            assert sorted(perm) == sorted(r * pad + t for r, s in enumerate(shards) for t in range(len(s)))
            cuts = [shard.read_shard(n, r, world) for r in range(world)]
            assert cuts[0][0] == 0 and cuts[-1][1] == n and all(cuts[i][1] == cuts[i + 1][0] for i in range(world - 1))
            for tri in (False, True):
                rows = [shard.row_shard(n, r, world, tri) for r in range(world)]
                assert rows[0][0] == 0 and rows[-1][1] == n
                assert all(rows[i][1] == rows[i + 1][0] for i in range(world - 1))
                assert sum(shard.pair_count(rw, n, tri) for rw in rows) == (n * (n + 1) // 2 if tri else n * n)


def test_bench_reference_arm_prints_the_contract_line(monkeypatch, capsys):
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours): one JSON line with the base contract's
    keys, `impl`, a `cpu_baseline` describing the run and an `e2e` that repeats the line's own value.  Run here on a
    shrunken genome length; ranks other than 0 must print nothing."""
    import argparse
    import json

    import bench
    monkeypatch.setattr(bench, "GENOME_LEN", 40_000)
    args = argparse.Namespace(gpus=1, steps=1, warmup=0)
    bench.run_reference(args)
    out = capsys.readouterr().out.strip().splitlines()
    assert len(out) == 1
    line = json.loads(out[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
                "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert key in line, key
    assert line["impl"] == "reference" and line["metric"] == "kmer_sketch_gbp_per_s" and line["unit"] == "Gbp/s"
    assert line["value"] > 0 and line["vs_baseline"] is None and line["higher_is_better"] is True
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in line["config"]
    monkeypatch.setenv("RANK", "1")
    bench.run_reference(args)
    assert capsys.readouterr().out == ""


def test_bench_clock_sampler_parses_nvidia_smi_lines():
    """clocks / throttle reasons: only samples after mark() count; median of the busy samples; reasons by column."""
    import bench
    s = bench.ClockSampler(0)
    s.lines = ["0, 210, 1965, 140.2, 0x0, Not Active, Not Active, Not Active, Not Active",      # idle, before the timed region
               "0, 1965, 1965, 820.0, 0x4, Not Active, Not Active, Not Active, Active",
               "0, 1950, 1965, 900.0, 0x4, Not Active, Not Active, Not Active, Active",
               "0, 1965, 1965, 880.0, 0x0, Not Active, Not Active, Not Active, Not Active",
               "garbage line"]
    s.mark_at = 1
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    c = s.stop()
    assert c["sm_mhz"] == 1965.0 and c["sm_max_mhz"] == 1965.0 and c["reasons"] == ["sw_power_cap"] and c["samples"] == 3


def test_pin_parity_table_readers(tmp_path):
    """tools/pin_parity.py compares the reference's and our `dist` outputs as unordered pairs (the reference emits one
    orientation of each pair, chosen by HashMap order)."""
    from tools import pin_parity as P
    a, b = tmp_path / "a.tsv", tmp_path / "b.dm"
    a.write_text("Reference\tQuery\tDistance\nx\ty\t0.100000\ny\ty\t0.000000\nx\tx\t0.000000\n")
    b.write_text("\tx\ty\nx\t0.000000\ny\t0.100000\t0.000000\n")
    n, bad = P.same_pairs(P.read_table(str(a)), P.read_table(str(b)))
    assert n == 3 and bad == []
    b.write_text("\tx\ty\nx\t0.000000\ny\t0.100001\t0.000000\n")
    n, bad = P.same_pairs(P.read_table(str(a)), P.read_table(str(b)))
    assert len(bad) == 1 and bad[0][0] == ("x", "y")


def test_pin_parity_self_test_names_every_flipped_convention(capsys):
    """tools/pin_parity.py --self-test: "reference" registers fabricated with each recalled convention flipped (HMH x / y
    halves, HLL index end) are diagnosed as exactly that switch, the current conventions as "current", a corrupted register
    as "unknown" -- so the first run against a real lash binary pins parity or says what to flip."""
    from tools import pin_parity
    assert pin_parity.self_test() == 0
    out = capsys.readouterr().out
    assert "hmh_x_low64" in out and "hll_index_high" in out and "SELF-TEST PASSED" in out
