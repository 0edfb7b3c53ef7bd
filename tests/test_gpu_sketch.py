"""GPU sketch kernel vs the CPU oracle: registers must be BIT-EXACT (integer work).

Mirrors what a test of the reference's sketch_files (src/utils.rs:439-510) would check: per-file
registers for each algorithm / k dispatch branch (k<=14 Kmer32bit, 16 Kmer16b32bit, else
Kmer64bit, utils.rs:466-502), the filter quirks (utils.rs:33-41) and the skip-short rule (:460).
All calls go through the C ABI (lash_b200.ops -> liblash_gpu.so).
"""
import numpy as np
import pytest

from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, LashError
from lash_b200.ops import Sketcher, sketch_genomes, sketch_genomes_text
from lash_b200.pack import PackedBatch
from tools import synth

pytestmark = pytest.mark.gpu

SEED = 42


def _check(oracle, gpu_ctx, algo, p, k, genomes, seed=SEED, **kw):
    got = sketch_genomes(gpu_ctx, algo, p, k, seed, genomes, **kw)
    exp = oracle.sketch_genomes(algo, p, k, seed, [list(g) for g in genomes], threads=4)
    assert got.dtype == exp.dtype and got.shape == exp.shape
    bad = np.argwhere(got != exp)
    assert bad.size == 0, f"{len(bad)} register mismatches, first at {bad[0]}: gpu={got[tuple(bad[0])]} cpu={exp[tuple(bad[0])]}"
    return got


@pytest.mark.parametrize("algo,p,k", [
    (ALGO_ULL, 10, 16),   # BASELINE config 2 / reference default k
    (ALGO_HLL, 14, 21),   # config 3
    (ALGO_HMH, 14, 16),   # config 1
    (ALGO_ULL, 14, 31),   # config 4 precision, wide k
    (ALGO_HLL, 10, 16),
    (ALGO_HMH, 14, 21),   # HMH hashes only the low 32 bits of a 42-bit k-mer (utils.rs:397)
])
def test_registers_bit_exact_configs(oracle, gpu_ctx, algo, p, k):
    genomes = synth.genomes(5, 300_000, seed=SEED)
    regs = _check(oracle, gpu_ctx, algo, p, k, genomes)
    assert (regs != 0).any()


@pytest.mark.parametrize("k", [1, 2, 3, 7, 13, 14, 15, 16, 17, 20, 24, 31, 32])
@pytest.mark.parametrize("algo,p", [(ALGO_ULL, 8), (ALGO_HLL, 8), (ALGO_HMH, 14)])
def test_every_k_dispatch_branch(oracle, gpu_ctx, algo, p, k):
    genomes = synth.genomes(2, 40_000, seed=k)
    _check(oracle, gpu_ctx, algo, p, k, genomes)


@pytest.mark.parametrize("algo,p", [(ALGO_ULL, 3), (ALGO_ULL, 5), (ALGO_ULL, 12), (ALGO_ULL, 17), (ALGO_HLL, 4),
                                    (ALGO_HLL, 16), (ALGO_HLL, 17)])
def test_precision_range_shared_memory_path(oracle, gpu_ctx, algo, p):
    genomes = synth.genomes(2, 150_000, seed=p)
    _check(oracle, gpu_ctx, algo, p, 16, genomes)


@pytest.mark.parametrize("algo,p", [(ALGO_ULL, 18), (ALGO_ULL, 20), (ALGO_HLL, 18)])
def test_precision_range_global_accumulator_path(oracle, gpu_ctx, algo, p):
    """2^p bytes no longer fit a CTA's shared memory: the kernel updates HBM/L2 directly."""
    genomes = synth.genomes(2, 200_000, seed=p)
    _check(oracle, gpu_ctx, algo, p, 21, genomes)


@pytest.mark.parametrize("algo,p,k", [(ALGO_ULL, 10, 16), (ALGO_HLL, 12, 21), (ALGO_HMH, 14, 16), (ALGO_ULL, 10, 5), (ALGO_ULL, 10, 32)])
def test_dirty_multi_record_genomes(oracle, gpu_ctx, algo, p, k):
    """lowercase / N / IUPAC deleted with flanks joined (utils.rs:36), k-mers never span records,
    records shorter than k skipped (:460), empty records, all-filtered records."""
    genomes = [synth.dirty_genome(60_000, k, seed=s) for s in range(4)]
    genomes.append([b""])                # file with one empty record
    genomes.append([])                   # file with no records at all
    genomes.append([b"ACGT" * 3, b"NNNN", b"acgt"])
    _check(oracle, gpu_ctx, algo, p, k, genomes)


def test_short_reads_many_records(oracle, gpu_ctx):
    """config 4 shape in miniature: 150 bp reads, every read its own record."""
    rng = np.random.default_rng(3)
    pool = synth.ancestor_codes(200_000, seed=9)
    reads = []
    for _ in range(4000):
        s = int(rng.integers(0, len(pool) - 150))
        reads.append(synth.to_ascii(pool[s:s + 150]))
    _check(oracle, gpu_ctx, ALGO_ULL, 14, 21, [reads, reads[:1000]])


def test_fixed_length_reads_without_boundary_table(oracle, gpu_ctx):
    """lash_span.rec_len: fixed-length reads need no rec_start[] table; the last read may be shorter."""
    import ctypes as C
    from lash_b200.capi import Span
    from lash_b200.pack import encode_record, pack_codes, padded_bytes
    rng = np.random.default_rng(5)
    for read_len, n_reads, tail, k in ((150, 3000, 0, 21), (100, 1001, 37, 31), (36, 500, 5, 16), (20, 64, 0, 21)):
        reads = [synth.to_ascii(rng.integers(0, 4, size=read_len, dtype=np.uint8)) for _ in range(n_reads)]
        if tail:
            reads.append(synth.to_ascii(rng.integers(0, 4, size=tail, dtype=np.uint8)))
        codes = np.concatenate([encode_record(r) for r in reads])
        buf = np.zeros(padded_bytes(len(codes)), dtype=np.uint8)
        pk = pack_codes(codes)
        buf[: len(pk)] = pk
        spans = (Span * 1)(Span(0, 0, len(codes), 0, len(reads), read_len))
        with Sketcher(gpu_ctx, ALGO_ULL, 12, k, SEED, 1) as sk:
            sk.push_raw(buf.ctypes.data, buf.nbytes, spans, 1, None, 0)
            got = sk.fetch()
        exp = oracle.sketch_genomes(ALGO_ULL, 12, k, SEED, [reads])
        assert np.array_equal(got, exp), (read_len, n_reads, tail, k)
    with Sketcher(gpu_ctx, ALGO_ULL, 12, 21, SEED, 1) as sk:   # n_rec must match ceil(n_bases / rec_len)
        spans = (Span * 1)(Span(0, 0, 1000, 0, 3, 150))
        with pytest.raises(LashError):
            sk.push_raw(buf.ctypes.data, buf.nbytes, spans, 1, None, 0)


def test_split_pushes_and_repeats_are_idempotent(oracle, gpu_ctx):
    """Register updates are commutative, associative and idempotent: the same genome pushed as one
    span, as many spans over several pushes, or twice, must give identical registers."""
    g = synth.dirty_genome(120_000, 16, seed=11)
    exp = oracle.sketch_genomes(ALGO_ULL, 12, 16, SEED, [g])
    for per_push in (1, 3, 1000):
        with Sketcher(gpu_ctx, ALGO_ULL, 12, 16, SEED, 1) as sk:
            for i in range(0, len(g), per_push):
                b = PackedBatch()
                b.add_genome(0, g[i:i + per_push])
                sk.push_batch(b)
            for i in range(0, len(g), per_push):  # and once more
                b = PackedBatch()
                b.add_genome(0, g[i:i + per_push])
                sk.push_batch(b)
            got = sk.fetch()
        assert np.array_equal(got, exp)


def test_seed_changes_sketch_and_matches(oracle, gpu_ctx):
    genomes = synth.genomes(1, 50_000, seed=5)
    a = _check(oracle, gpu_ctx, ALGO_ULL, 10, 16, genomes, seed=42)
    b = _check(oracle, gpu_ctx, ALGO_ULL, 10, 16, genomes, seed=(1 << 63) + 12345)
    assert not np.array_equal(a, b)


def test_many_small_genomes_one_push(oracle, gpu_ctx):
    genomes = synth.genomes(300, 3_000, seed=8)
    _check(oracle, gpu_ctx, ALGO_ULL, 10, 16, genomes, genomes_per_push=300)
    _check(oracle, gpu_ctx, ALGO_HMH, 14, 16, genomes[:40], genomes_per_push=7)


def test_large_genome_many_tiles(oracle, gpu_ctx):
    genomes = synth.genomes(2, 6_000_000, seed=21)
    _check(oracle, gpu_ctx, ALGO_ULL, 10, 16, genomes)
    _check(oracle, gpu_ctx, ALGO_HLL, 14, 21, genomes)


def test_error_behaviour(gpu_ctx):
    """k outside 1..32 is a hard error in the reference (utils.rs:500-502 panics)."""
    for k in (0, 33):
        with pytest.raises(LashError):
            Sketcher(gpu_ctx, ALGO_ULL, 10, k, SEED, 1)
    with pytest.raises(LashError):
        Sketcher(gpu_ctx, ALGO_ULL, 2, 16, SEED, 1)   # UltraLogLog::new rejects p < 3
    with pytest.raises(LashError):
        Sketcher(gpu_ctx, ALGO_ULL, 27, 16, SEED, 1)
    with pytest.raises(LashError):
        Sketcher(gpu_ctx, 7, 10, 16, SEED, 1)         # main.rs:245 "Algorithm must be either hmh, ull, or hll"
    with Sketcher(gpu_ctx, ALGO_ULL, 10, 16, SEED, 2) as sk:
        b = PackedBatch()
        b.add_genome(5, [b"ACGT" * 10])               # genome slot out of range
        with pytest.raises(LashError):
            sk.push_batch(b)


@pytest.mark.parametrize("algo,p", [(ALGO_ULL, 10), (ALGO_ULL, 14), (ALGO_ULL, 3), (ALGO_ULL, 20), (ALGO_HLL, 14), (ALGO_HLL, 4),
                                    (ALGO_HLL, 18)])
def test_adversarial_hashes_with_32_or_more_leading_zeros(oracle, gpu_ctx, algo, p):
    """The kernel's fast path looks at 32 bits of the hash; the remaining 2^-32 of hashes take an exact
    slow path that random genomes never reach.  XXH3 on 8 bytes is invertible, so build 32-mers whose
    hash has 32..(64-p) leading zeros where it matters and check registers bit for bit."""
    from tools.xxh3_invert import adversarial_32mers
    rng = np.random.default_rng(p)
    targets = []
    for _ in range(400):
        if algo == ALGO_ULL:       # idx | 32+ zeros | tail
            idx = int(rng.integers(0, 1 << p))
            tail_bits = 64 - p - 32
            tail = int(rng.integers(0, 1 << tail_bits)) >> int(rng.integers(0, tail_bits + 1))
            targets.append((idx << (64 - p)) | tail)
            # the fast path looks at the 32-p bits after the index only: hashes with 32-p .. 31 zeros there (the first
            # one lands in the p bits the fast path ignores) must take the exact path too
            # ... and so must those whose first one is among the last 4 of those 32-p bits (the fast path drops them
            # together with the xorshift term that reaches them)
            for z in (int(rng.integers(32 - p, 32)), int(rng.integers(max(28 - p, 0), 32 - p))):
                below = 63 - p - z
                targets.append((idx << (64 - p)) | (1 << below) | int(rng.integers(0, 1 << below)))
        else:                      # hi word zero; low p bits = index
            lo = int(rng.integers(0, 1 << 32)) >> int(rng.integers(0, 33 - p))
            targets.append((lo << p | int(rng.integers(0, 1 << p))) & 0xFFFFFFFF)
    kmers = adversarial_32mers(targets, SEED)
    assert len(kmers) > 100
    # hide them in ordinary sequence: own records (exactly one k-mer each) and inside long records
    filler = synth.genomes(1, 30_000, seed=p)[0][0]
    g1 = [filler] + kmers
    g2 = [filler[:5000] + b"N" + b"N".join(kmers[:50]) + b"N" + filler[5000:]]
    regs = _check(oracle, gpu_ctx, algo, p, 32, [g1, g2, kmers])
    if algo == ALGO_HLL:
        assert regs[2].max() >= 33     # rho beyond what 32 bits can see
    else:
        assert (regs[2] >> 2).max() >= 31 + p   # update value u = nlz + p - 1 with nlz >= 32


@pytest.mark.parametrize("algo,p", [(ALGO_ULL, 10), (ALGO_ULL, 14), (ALGO_ULL, 3), (ALGO_HLL, 12), (ALGO_HMH, 14)])
def test_merge_of_shares_equals_sketch_of_the_whole(oracle, gpu_ctx, algo, p):
    """One sample sketched in shares (several GPUs / passes) and folded with lash_sketch_merge must equal
    the sketch of all its records: UltraLogLog::merge (utils.rs:260-262; not a byte max),
    HyperLogLog::union and the hyperminhash register max."""
    from lash_b200 import ops
    k = 16
    parts_a = [synth.dirty_genome(40_000 + 1000 * g, k, seed=100 + g) for g in range(5)]
    parts_b = [synth.dirty_genome(30_000 + 1500 * g, k, seed=200 + g) for g in range(5)]
    parts_b[3] = []                                                  # an empty share
    ra = sketch_genomes(gpu_ctx, algo, p, k, SEED, parts_a)
    rb = sketch_genomes(gpu_ctx, algo, p, k, SEED, parts_b)
    merged = ops.merge(gpu_ctx, algo, p, ra, rb)
    whole = oracle.sketch_genomes(algo, p, k, SEED, [a + b for a, b in zip(parts_a, parts_b)], threads=4)
    assert np.array_equal(merged, whole)
    if algo == ALGO_ULL:
        exp = np.stack([oracle.ull_merge(x, y, p) for x, y in zip(ra, rb)])
        assert np.array_equal(merged, exp)
        if p >= 10:
            assert not np.array_equal(merged, np.maximum(ra, rb))   # ULL merge is not max
    else:
        assert np.array_equal(merged, np.maximum(ra, rb))
    assert np.array_equal(ops.merge(gpu_ctx, algo, p, merged, merged), merged)      # idempotent
    assert np.array_equal(ops.merge(gpu_ctx, algo, p, rb, ra), merged)              # commutative


# ---- lash_sketch_push_ascii: filter_out_n + 2-bit pack on the device (text_kernels.cu) ------------------------------
def _check_text(oracle, gpu_ctx, algo, p, k, genomes, **kw):
    got = sketch_genomes_text(gpu_ctx, algo, p, k, SEED, genomes, **kw)
    exp = oracle.sketch_genomes(algo, p, k, SEED, [list(g) for g in genomes], threads=4)
    bad = np.argwhere(got != exp)
    assert bad.size == 0, f"{len(bad)} register mismatches, first at {bad[0]}: gpu={got[tuple(bad[0])]} cpu={exp[tuple(bad[0])]}"
    return got


def _fasta_lines(seq: bytes, width: int = 80) -> bytes:
    """A FASTA body as it sits in the file: line breaks (and a stray CR) are just bytes the filter deletes."""
    return b"\r\n".join(seq[o:o + width] for o in range(0, len(seq), width)) + b"\n"


@pytest.mark.parametrize("algo,p,k", [(ALGO_ULL, 10, 16), (ALGO_HLL, 12, 21), (ALGO_HMH, 14, 16), (ALGO_ULL, 10, 5), (ALGO_ULL, 14, 32)])
def test_ascii_push_dirty_multi_record_genomes(oracle, gpu_ctx, algo, p, k):
    """The same dirty inputs as the host-packed path: lowercase / N / IUPAC deleted with flanks joined (utils.rs:36),
    k-mers never span records, records shorter than k skipped (:460), empty and all-filtered records."""
    genomes = [synth.dirty_genome(60_000, k, seed=s) for s in range(4)]
    genomes.append([b""])
    genomes.append([])
    genomes.append([b"ACGT" * 3, b"NNNN", b"acgt"])
    genomes.append([_fasta_lines(synth.genomes(1, 50_000, seed=5)[0][0])])      # one record with line breaks inside
    _check_text(oracle, gpu_ctx, algo, p, k, genomes)
    _check_text(oracle, gpu_ctx, algo, p, k, genomes, genomes_per_push=3)


@pytest.mark.parametrize("algo,p,k", [(ALGO_ULL, 10, 16), (ALGO_HLL, 14, 21)])
def test_ascii_push_spans_many_text_blocks(oracle, gpu_ctx, algo, p, k):
    """Genomes far larger than one 32 KiB text block, with long deleted runs, so kept positions drift away from byte
    positions: block prefixes, word sharing between neighbouring blocks and the tile clipping all matter."""
    rng = np.random.default_rng(11)
    genomes = []
    for g in range(3):
        seq = bytearray(synth.genomes(1, 1_500_000 + 77 * g, seed=20 + g)[0][0])
        for _ in range(40):
            a = int(rng.integers(0, len(seq) - 50_000))
            n = int(rng.integers(1, 40_000))
            seq[a:a + n] = (b"N" if rng.integers(0, 2) else b"a") * n
        genomes.append([_fasta_lines(bytes(seq), 60)])
    genomes.append([b"N" * 200_000])                                               # nothing survives
    genomes.append([b"N" * 100_000 + b"ACGTTGCAAGGCTTAACCGGTTAAACCCGGGTTTACGT" + b"n" * 70_000])   # one short island
    _check_text(oracle, gpu_ctx, algo, p, k, genomes)


def test_ascii_push_every_byte_value(oracle, gpu_ctx):
    """The device classifies four bytes per 32-bit word (PRMT-selected expected letter + zero-byte test): every byte value in
    every position of the word, mixed with bases at random, must be kept / deleted exactly like filter_out_n (utils.rs:33-41)."""
    rng = np.random.default_rng(12)
    every = np.array([v for v in range(256) if v != 1], dtype=np.uint8)          # 0x01 is the record separator of the text ABI
    recs = []
    for r in range(6):
        n = 40_000 + 13 * r
        body = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
        pos = rng.choice(n, size=n // 5, replace=False)
        body[pos] = every[rng.integers(0, len(every), size=len(pos))]
        body[r::997] = every[(np.arange(len(body[r::997])) + r) % len(every)]      # each value at every word offset over the records
        recs.append(body.tobytes())
    _check_text(oracle, gpu_ctx, ALGO_ULL, 12, 16, [recs, recs[:2], [recs[3][5:]], [recs[4][1:70_001]]])
    _check_text(oracle, gpu_ctx, ALGO_HLL, 10, 31, [recs[::-1]])


def test_ascii_push_short_reads(oracle, gpu_ctx):
    """config 4 shape: 150 bp reads, one record each, some with N, some shorter than k after filtering."""
    rng = np.random.default_rng(4)
    pool = synth.to_ascii(synth.ancestor_codes(300_000, seed=9))
    reads = []
    for i in range(20_000):
        a = int(rng.integers(0, len(pool) - 150))
        r = bytearray(pool[a:a + 150])
        if i % 17 == 0:
            r[40:45] = b"NNNNN"
        if i % 501 == 0:
            r = bytearray(b"N" * 140 + bytes(r[:10]))
        reads.append(bytes(r))
    _check_text(oracle, gpu_ctx, ALGO_ULL, 14, 21, [reads, reads[:7], reads[100:5000]])


def test_ascii_push_equals_packed_push_and_rejects_bad_spans(oracle, gpu_ctx):
    import ctypes as C

    from lash_b200 import capi
    genomes = synth.genomes(3, 200_000, seed=SEED)
    a = sketch_genomes(gpu_ctx, ALGO_ULL, 10, 16, SEED, genomes)
    b = sketch_genomes_text(gpu_ctx, ALGO_ULL, 10, 16, SEED, genomes)
    assert np.array_equal(a, b)
    with Sketcher(gpu_ctx, ALGO_ULL, 10, 16, SEED, 2) as sk:
        buf = np.zeros(64, dtype=np.uint8)
        for span in (capi.TextSpan(2, 0, 16, 1, 0), capi.TextSpan(0, 8, 16, 1, 0), capi.TextSpan(0, 0, 128, 1, 0)):
            arr = (capi.TextSpan * 1)(span)
            rc = capi.lib().lash_sketch_push_ascii(sk._h, buf.ctypes.data_as(C.c_void_p), buf.nbytes, arr, 1, None)
            assert rc == -1


@pytest.mark.parametrize("seed", range(6))
def test_ascii_push_fuzz_against_packed_push_and_oracle(oracle, gpu_ctx, seed):
    """Randomised text: record counts and lengths from 0 to beyond a text block, junk runs of every kind at random places
    (incl. right at 16-byte, 4 KiB-iteration and 32 KiB-block edges), random k / precision / algorithm; the device-side
    filter + pack, the host packer and the oracle must agree register for register."""
    rng = np.random.default_rng(1000 + seed)
    junk = np.array([v for v in range(256) if v not in (1, 65, 67, 71, 84)], dtype=np.uint8)
    genomes = []
    for g in range(int(rng.integers(3, 9))):
        recs = []
        for r in range(int(rng.integers(0, 7))):
            n = int(rng.choice([0, 1, 15, 16, 17, 31, 33, 150, 4095, 4096, 4097, 32767, 32768, 32769, 70_000, int(rng.integers(1, 200_000))]))
            body = np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, size=n)].copy()
            for _ in range(int(rng.integers(0, 6))):
                if n == 0:
                    break
                a = int(rng.choice([0, 15, 16, 4095, 4096, 32767, 32768, int(rng.integers(0, n))])) % n
                m = int(rng.choice([1, 2, 3, 16, 17, 100, 5000]))
                body[a:a + m] = junk[rng.integers(0, len(junk), size=len(body[a:a + m]))]
            recs.append(body.tobytes())
        genomes.append(recs)
    algo, p = [(ALGO_ULL, 10), (ALGO_ULL, 13), (ALGO_HLL, 11), (ALGO_HMH, 14), (ALGO_ULL, 16), (ALGO_HLL, 14)][seed % 6]
    k = int(rng.choice([4, 11, 16, 21, 32]))
    exp = oracle.sketch_genomes(algo, p, k, SEED, genomes, threads=4)
    per = int(rng.integers(1, 5))
    assert np.array_equal(sketch_genomes_text(gpu_ctx, algo, p, k, SEED, genomes, genomes_per_push=per), exp)
    assert np.array_equal(sketch_genomes(gpu_ctx, algo, p, k, SEED, genomes, genomes_per_push=per), exp)
