"""End-to-end parity of the C++ host layer on a GPU: FASTA/FASTQ files -> sketch_files<S> ->
{out}_sketches.bin / _files.json / _parameters.json -> `dist` -> TSV / --dm text, against the CPU
oracle fed with the same records.  Mirrors what an integration test of the reference's
`lash sketch` + `lash dist` (src/main.rs:179-613) would check.

Bar: registers bit-exact; distances as text equal to the oracle's `{:.6}` text except where the
1e-12 relative difference between CUDA and glibc log/pow straddles a rounding boundary of the 6th
decimal (at most 1 unit in the last printed place, and rare)."""
import gzip
import os

import numpy as np
import pytest

from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, hostapi
from tools import synth

pytestmark = pytest.mark.gpu


def _write_fasta(path, records, width=70, gz=False):
    opener = gzip.open if gz else open
    with opener(path, "wb") as f:
        for i, s in enumerate(records):
            f.write(b">rec%d some description\n" % i)
            for o in range(0, len(s), width):
                f.write(s[o:o + width] + b"\n")


def _write_fastq(path, records):
    with open(path, "wb") as f:
        for i, s in enumerate(records):
            f.write(b"@read%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))


@pytest.fixture(params=["packed", "ascii"])
def ingest(request):
    """sketch_files with the host packer (lash_sketch_push) and with the device-side filter + pack (lash_sketch_push_ascii)."""
    hostapi.check(hostapi.lib().lash_host_set_ingest_mode({"packed": 1, "ascii": 2}[request.param]))
    yield request.param
    hostapi.check(hostapi.lib().lash_host_set_ingest_mode(0))


def _oname(algo):
    return {ALGO_HMH: "HMH", ALGO_HLL: "HLL", ALGO_ULL: "ULL"}[algo]


@pytest.mark.parametrize("algo,p,k", [(ALGO_ULL, 10, 16), (ALGO_HLL, 12, 21), (ALGO_HMH, 14, 16), (ALGO_ULL, 14, 31), (ALGO_ULL, 8, 5)])
def test_sketch_files_matches_oracle_on_dirty_fasta(oracle, gpu_ctx, tmp_path, algo, p, k, ingest):
    """Per-file record loop of utils.rs:457-503 on multi-record, multi-line, dirty, gzipped and
    short files -- several files per staging chunk."""
    genomes = [synth.dirty_genome(60_000 + 7919 * g, k, seed=g + 1) for g in range(6)]
    genomes.append([b"ACGT" * 3])                       # one tiny record (shorter than k for most k)
    genomes.append([b"NNNN", b"acgt"])                  # nothing survives the filter: an empty sketch
    genomes += synth.genomes(2, 300_000, seed=9)
    files = []
    for g, recs in enumerate(genomes):
        path = str(tmp_path / f"g{g}.fa{'.gz' if g % 3 == 1 else ''}")
        _write_fasta(path, recs, width=60 + g, gz=(g % 3 == 1))
        files.append(path)
    regs, st = hostapi.sketch_files_regs(gpu_ctx, algo, p, k, 42, files, threads=3)
    exp = oracle.sketch_genomes(getattr(oracle, _oname(algo)), p, k, 42, genomes, threads=4)
    assert np.array_equal(regs, exp)
    assert st.n_records == sum(len(g) for g in genomes)
    assert not regs[7].any()


@pytest.mark.parametrize("k", [16, 21, 32])
def test_large_records_split_across_staging_chunks(oracle, gpu_ctx, tmp_path, k, ingest):
    """A record larger than a staging chunk is cut with a (k-1)-base overlap: every k-mer start
    exactly once (chunk = 1 MiB = 4 Mbp; records of 9 and 5 Mbp, plus small ones around them)."""
    rng = np.random.default_rng(k)
    recs = [synth.to_ascii(rng.integers(0, 4, size=n, dtype=np.uint8)) for n in (1000, 9_000_000, k - 1, 5_000_000, 3 * k)]
    recs[1] = recs[1][:4_100_000] + b"N" * 50 + recs[1][4_100_050:]     # dirt right where a chunk boundary falls
    path = str(tmp_path / "big.fa")
    _write_fasta(path, recs, width=80)
    other = str(tmp_path / "small.fa")
    _write_fasta(other, [recs[0]])
    for algo, p in ((ALGO_ULL, 12), (ALGO_HMH, 14)):
        regs, st = hostapi.sketch_files_regs(gpu_ctx, algo, p, k, 7, [other, path, other], threads=2, chunk_bytes=1 << 20)
        exp = oracle.sketch_genomes(getattr(oracle, _oname(algo)), p, k, 7, [[recs[0]], recs, [recs[0]]], threads=3)
        assert st.n_pushes >= 4
        assert np.array_equal(regs, exp)


def test_fasta_read_straight_into_pinned_chunks(oracle, gpu_ctx, tmp_path, ingest):
    """Plain FASTA in device-filter mode is read() directly into the pinned chunk and fixed up in place: headers blanked
    (even when they are made of base letters), '>' inside a sequence line, CRLF line ends, a separator byte in the data, a
    header that straddles a chunk boundary (chunk = 4 MiB here), no line end at the end of the file."""
    rng = np.random.default_rng(31)
    seq = lambda n: synth.to_ascii(rng.integers(0, 4, size=n, dtype=np.uint8))
    recs = [seq((4 << 20) - 150_000), seq(300_000), seq(20), seq(2_000_000)]
    recs[1] = recs[1][:1000] + b">" + recs[1][1000:2000] + b"\x01" + recs[1][2000:]      # '>' and the separator byte inside a line
    path = str(tmp_path / "odd.fa")
    with open(path, "wb") as f:
        f.write(b">first\r\n" + b"\r\n".join(recs[0][o:o + 70] for o in range(0, len(recs[0]), 70)) + b"\r\n")
        f.write(b">" + b"ACGT" * 75_000 + b" a header made of bases, 300 kB long\n")         # straddles the 4 MiB chunk end
        f.write(b"\n".join(recs[1][o:o + 61] for o in range(0, len(recs[1]), 61)) + b"\n")
        f.write(b">tiny\n" + recs[2] + b"\n>last\n" + recs[3])                           # no trailing line end
    # what needletail hands over: line ends removed, everything else (the in-line '>', the 0x01) is sequence content
    want = [recs[0], recs[1], recs[2], recs[3]]
    for algo, p, k in ((ALGO_ULL, 12, 21), (ALGO_HMH, 14, 16)):
        regs, st = hostapi.sketch_files_regs(gpu_ctx, algo, p, k, 42, [path, path], threads=2, chunk_bytes=1 << 20)
        exp = oracle.sketch_genomes(getattr(oracle, _oname(algo)), p, k, 42, [want, want], threads=2)
        assert np.array_equal(regs, exp)
        assert st.n_records == 8


def test_fastq_reads_uniform_and_ragged(oracle, gpu_ctx, tmp_path, ingest):
    """config 4 shape: 150 bp reads of one sample.  Equal-length reads take the table-free
    rec_len path, ragged reads (and reads that lose bases to the filter) the rec_start table."""
    rng = np.random.default_rng(4)
    genome = synth.to_ascii(rng.integers(0, 4, size=400_000, dtype=np.uint8))
    starts = rng.integers(0, len(genome) - 150, size=20_000)
    uniform = [genome[s:s + 150] for s in starts]
    ragged = [genome[s:s + int(n)] for s, n in zip(starts, rng.integers(1, 151, size=len(starts)))]
    dirty = [r[:40] + b"N" + r[41:] if i % 7 == 0 else r for i, r in enumerate(uniform)]
    cases = {"uniform": uniform, "ragged": ragged, "dirty": dirty}
    files = []
    for name, reads in cases.items():
        path = str(tmp_path / f"{name}.fq")
        _write_fastq(path, reads)
        files.append(path)
    regs, st = hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 14, 21, 42, files, threads=2)
    exp = oracle.sketch_genomes(oracle.ULL, 14, 21, 42, list(cases.values()), threads=3)
    assert np.array_equal(regs, exp)
    assert st.n_records == 3 * len(starts)


def test_many_small_files_share_chunks(oracle, gpu_ctx, tmp_path, ingest):
    genomes = synth.genomes(300, 20_000, seed=5)
    files = []
    for g, recs in enumerate(genomes):
        path = str(tmp_path / f"s{g}.fa")
        _write_fasta(path, recs)
        files.append(path)
    regs, st = hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 10, 16, 42, files, threads=4)
    assert st.n_pushes <= 8          # not one push per file
    assert np.array_equal(regs, oracle.sketch_genomes(oracle.ULL, 10, 16, 42, genomes, threads=4))


def test_sketch_files_error_behaviour(gpu_ctx, tmp_path):
    good = str(tmp_path / "ok.fa")
    _write_fasta(good, [b"ACGT" * 100])
    bad = str(tmp_path / "bad.fa")
    open(bad, "wb").write(b"this is not a sequence file\n")
    with pytest.raises(hostapi.HostError, match="Invalid input file"):      # utils.rs:453
        hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 10, 16, 42, [good, bad])
    with pytest.raises(hostapi.HostError, match="Invalid input file"):
        hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 10, 16, 42, [str(tmp_path / "missing.fa")])
    with pytest.raises(hostapi.HostError, match="k-mer length must be 1-32"):  # utils.rs:500-502
        hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 10, 33, 42, [good])
    with pytest.raises(hostapi.HostError):                                   # UltraLogLog::new(p) Err
        hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 2, 16, 42, [good])
    regs, _ = hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 10, 16, 42, [])
    assert regs.shape == (0, 1024)


def _parse_list(path):
    lines = open(path).read().split("\n")
    assert lines[0] == "Reference\tQuery\tDistance" and lines[-1] == ""
    return [tuple(ln.split("\t")) for ln in lines[1:-1]]


def _expected_text(oracle, algo, p, k, est, model, fp32, regs_r, regs_q, names_r, names_q, tri):
    d = oracle.dist(getattr(oracle, _oname(algo)), p, k, est, model, fp32, regs_r, regs_q)
    rows = []
    for i, rn in enumerate(names_r):
        for j, qn in enumerate(names_q):
            if tri and j > i:
                continue
            v = 0.0 if rn == qn else float(d[i, j])
            rows.append((rn, qn, v))
    return rows


def _assert_text_rows(got, exp):
    assert len(got) == len(exp)
    off = 0
    for (gr, gq, gd), (er, eq, ev) in zip(got, exp):
        assert (gr, gq) == (er, eq)
        et = "%.6f" % ev
        if gd != et:
            assert abs(float(gd) - ev) <= 1.0000001e-6, (gr, gq, gd, et)
            off += 1
    assert off <= max(1, len(exp) // 200)


@pytest.mark.parametrize("algo,p,k,estimator", [(ALGO_ULL, 10, 16, "fgra"), (ALGO_ULL, 10, 16, "ml"), (ALGO_HLL, 12, 21, "fgra"),
                                                (ALGO_HMH, 14, 16, "fgra")])
@pytest.mark.parametrize("fused", [True, False])
def test_sketch_then_dist_end_to_end(oracle, gpu_ctx, tmp_path, algo, p, k, estimator, fused):
    """`lash sketch` + `lash dist` on files: same prefix on both sides (same_files: lower triangle
    incl. diagonal, main.rs:404 + utils.rs:158-160) and a different query set (full matrix)."""
    genomes = synth.genomes(9, 150_000, seed=21)
    files = []
    for g, recs in enumerate(genomes):
        path = str(tmp_path / f"genome_{g}.fasta")
        _write_fasta(path, recs)
        files.append(path)
    ref_prefix, qry_prefix = str(tmp_path / "refs"), str(tmp_path / "qrys")
    hostapi.sketch_files(gpu_ctx, algo, p, k, 42, files, ref_prefix, threads=2)
    hostapi.sketch_files(gpu_ctx, algo, p, k, 42, files[2:6], qry_prefix, threads=2)
    for suffix in ("_sketches.bin", "_files.json", "_parameters.json"):
        assert os.path.exists(ref_prefix + suffix)
    O = getattr(oracle, _oname(algo))
    regs = oracle.sketch_genomes(O, p, k, 42, genomes, threads=4)
    on_disk, _ = hostapi.read_sketches(ref_prefix + "_sketches.bin", algo, len(files))
    assert np.array_equal(on_disk, regs)
    est = 1 if estimator == "ml" else 0
    for model in (1, 0):
        for fp32 in (False, True):
            out = str(tmp_path / f"self_{model}_{int(fp32)}.tsv")
            hostapi.dist(gpu_ctx, ref_prefix, ref_prefix, out, estimator, model, dm=False, fp32=fp32, threads=3, fused=fused)
            exp = _expected_text(oracle, algo, p, k, est, model, fp32, regs, regs, files, files, tri=True)
            _assert_text_rows([(r, q, d) for r, q, d in _parse_list(out)], exp)
    out = str(tmp_path / "cross.tsv")
    hostapi.dist(gpu_ctx, ref_prefix, qry_prefix, out, estimator, 1, dm=False, fp32=False, threads=2, fused=fused)
    exp = _expected_text(oracle, algo, p, k, est, 1, False, regs, regs[2:6], files, files[2:6], tri=False)
    _assert_text_rows(_parse_list(out), exp)


@pytest.mark.parametrize("fused", [True, False])
def test_dm_matrix_layout(oracle, gpu_ctx, tmp_path, fused):
    """--dm (main.rs:438-467): header row of "\\t{query}", then "\\n{ref}" + "\\t{:.6}" cells, no final newline."""
    genomes = synth.genomes(5, 100_000, seed=3)
    files = []
    for g, recs in enumerate(genomes):
        path = str(tmp_path / f"m{g}.fa")
        _write_fasta(path, recs)
        files.append(path)
    a, b = str(tmp_path / "A"), str(tmp_path / "B")
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 16, 42, files, a)
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 16, 42, files[1:4], b)
    regs = oracle.sketch_genomes(oracle.ULL, 10, 16, 42, genomes, threads=4)
    out = str(tmp_path / "dm.txt")
    hostapi.dist(gpu_ctx, a, b, out, "fgra", 1, dm=True, fused=fused)
    text = open(out).read()
    lines = text.split("\n")
    assert not text.endswith("\n")
    assert lines[0] == "".join("\t" + q for q in files[1:4])
    exp = _expected_text(oracle, ALGO_ULL, 10, 16, 0, 1, False, regs, regs[1:4], files, files[1:4], tri=False)
    got = []
    for ln, rn in zip(lines[1:], files):
        cells = ln.split("\t")
        assert cells[0] == rn and len(cells) == 4
        got += [(rn, qn, c) for qn, c in zip(files[1:4], cells[1:])]
    _assert_text_rows(got, exp)
    # same files: lower-triangular rows
    hostapi.dist(gpu_ctx, a, a, out, "fgra", 1, dm=True, fused=fused)
    lines = open(out).read().split("\n")
    assert [len(ln.split("\t")) for ln in lines[1:]] == [2, 3, 4, 5, 6]
    assert lines[1].split("\t")[1] == "0.000000"


def test_dist_command_error_behaviour_and_name_rules(oracle, gpu_ctx, tmp_path):
    genomes = synth.genomes(3, 60_000, seed=8)
    files = []
    for g, recs in enumerate(genomes):
        path = str(tmp_path / f"e{g}.fa")
        _write_fasta(path, recs)
        files.append(path)
    u16, u21, h16 = str(tmp_path / "u16"), str(tmp_path / "u21"), str(tmp_path / "h16")
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 16, 42, files, u16)
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 21, 42, files, u21)
    hostapi.sketch_files(gpu_ctx, ALGO_HLL, 10, 16, 42, files, h16)
    out = str(tmp_path / "o.tsv")
    with pytest.raises(hostapi.HostError, match="same k"):                 # main.rs:363
        hostapi.dist(gpu_ctx, u16, u21, out)
    with pytest.raises(hostapi.HostError, match="Algorithms do not match"):  # main.rs:366
        hostapi.dist(gpu_ctx, u16, h16, out)
    with pytest.raises(hostapi.HostError, match="fgra or ml"):              # utils.rs:217
        hostapi.dist(gpu_ctx, u16, u16, out, estimator="bogus")
    with pytest.raises(hostapi.HostError, match="model needs to be 0 or 1"):  # main.rs:421
        hostapi.dist(gpu_ctx, u16, u16, out, model=2)
    with pytest.raises(hostapi.HostError, match="There should be 3 files"):  # main.rs:330-336
        hostapi.dist(gpu_ctx, str(tmp_path / "nothing_here"), u16, out)
    # a name listed twice collapses to one entry holding the LAST sketch (HashMap insert, utils.rs:219);
    # an identical sketch under a DIFFERENT name is not forced to 0 (main.rs:452 compares names): -ln(1)/k = -0
    dup = str(tmp_path / "dup")
    alias = str(tmp_path / "alias.fa")
    _write_fasta(alias, genomes[1])
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 16, 42, [files[0], files[1], files[0], alias], dup)
    hostapi.dist(gpu_ctx, dup, dup, out, fused=False)
    rows = _parse_list(out)
    assert [(r, q) for r, q, _ in rows] == [(files[0], files[0]), (files[1], files[0]), (files[1], files[1]),
                                           (alias, files[0]), (alias, files[1]), (alias, alias)]
    assert rows[4][2] == "-0.000000" and rows[5][2] == "0.000000"


def test_cli_binary_mirrors_lash_sketch_and_dist(oracle, tmp_path):
    """`lash-b200 sketch` / `lash-b200 dist`: the reference's flags and defaults (main.rs:26-177), its three
    output files, its TSV -- driven as a user would, from a list file in the working directory."""
    import json
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lash_b200", "_lib", "lash-b200")
    genomes = synth.genomes(5, 80_000, seed=13)
    names = []
    for g, recs in enumerate(genomes):
        _write_fasta(str(tmp_path / f"c{g}.fna"), recs)
        names.append(f"c{g}.fna")
    (tmp_path / "list.txt").write_text("\n".join(names[:3]) + "\n\n   \n" + "\n".join(names[3:]) + "\n")   # blank lines are skipped
    run = lambda *a: subprocess.run([exe, *a], cwd=tmp_path, capture_output=True, text=True)
    r = run("sketch", "-f", "list.txt", "-o", "db", "-a", "ull", "-p", "10", "-t", "2")       # k, seed: defaults 16 / 42
    assert r.returncode == 0, r.stderr
    assert json.load(open(tmp_path / "db_files.json")) == names
    assert json.load(open(tmp_path / "db_parameters.json")) == {"algorithm": "ull", "k": "16", "molecule": "nucleotide", "precision": "10", "seed": "42"}
    regs = oracle.sketch_genomes(oracle.ULL, 10, 16, 42, genomes, threads=4)
    on_disk, p = hostapi.read_sketches(str(tmp_path / "db_sketches.bin"), ALGO_ULL, 5)
    assert p == 10 and np.array_equal(on_disk, regs)
    for extra, fused in ((), True), (("--mirror",), False):
        r = run("dist", "-q", "db", "-r", "db", "-o", "d.tsv", *extra)                       # estimator fgra, model 1: defaults
        assert r.returncode == 0, r.stderr
        exp = _expected_text(oracle, ALGO_ULL, 10, 16, 0, 1, False, regs, regs, names, names, tri=True)
        _assert_text_rows(_parse_list(str(tmp_path / "d.tsv")), exp)
    r = run("dist", "--query", "db", "--reference", "db", "--output_file", "m.txt", "--dm", "--fp32", "-e", "ml", "-m", "0")
    assert r.returncode == 0, r.stderr
    lines = open(tmp_path / "m.txt").read().split("\n")
    assert lines[0] == "".join("\t" + n for n in names) and [ln.split("\t")[0] for ln in lines[1:]] == names
    d = oracle.dist(oracle.ULL, 10, 16, oracle.ML, 0, True, regs, regs)
    assert lines[3].split("\t")[1:] == ["%.6f" % float(d[2, j]) if j != 2 else "0.000000" for j in range(3)]
    r = run("sketch", "-f", "list.txt", "-a", "minhash")
    assert r.returncode != 0 and "Algorithm must be either hmh, ull, or hll" in r.stderr      # main.rs:245
    r = run("sketch", "-f", "list.txt", "-a", "hll", "-k", "40")
    assert r.returncode != 0 and "k-mer length must be 1-32" in r.stderr                      # utils.rs:501
    r = run("dist", "-q", "db", "-r", "nothing")
    assert r.returncode != 0 and "There should be 3 files" in r.stderr                         # main.rs:330


def test_cli_hll_bias_regime_fails_loudly_and_same_files_by_inode(oracle, tmp_path):
    """(1) HLL sketches of tiny inputs land in the HLL++ bias-table regime, which this build cannot reproduce: `dist` must fail
    and write nothing unless --allow-hll-bias-regime is given.  (2) `-r sub/a -q ./sub/a` is the same file set:
    lower triangle, not the full matrix."""
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "lash_b200", "_lib", "lash-b200")
    (tmp_path / "sub").mkdir()
    rng = np.random.default_rng(21)
    names = []
    for g in range(3):
        _write_fasta(str(tmp_path / f"v{g}.fa"), [synth.to_ascii(rng.integers(0, 4, size=30_000, dtype=np.uint8))])   # 30k k-mers < 5 * 2^14
        names.append(f"v{g}.fa")
    (tmp_path / "list.txt").write_text("\n".join(names) + "\n")
    run = lambda *a: subprocess.run([exe, *a], cwd=tmp_path, capture_output=True, text=True)
    assert run("sketch", "-f", "list.txt", "-o", "sub/h", "-a", "hll", "-p", "14", "-k", "21").returncode == 0
    r = run("dist", "-q", "sub/h", "-r", "sub/h", "-o", "d.tsv")
    assert r.returncode != 0 and "bias" in r.stderr and not (tmp_path / "d.tsv").exists()
    r = run("dist", "-q", "sub/h", "-r", "sub/h", "-o", "d.tsv", "--allow-hll-bias-regime")
    assert r.returncode == 0 and "warning" in r.stderr
    rows = _parse_list(str(tmp_path / "d.tsv"))
    assert len(rows) == 6 and all(d == "1.000000" for a, b, d in rows if a != b)
    # the same sketches as ULL: no regime, and the two spellings of the prefix are recognised as one file set
    assert run("sketch", "-f", "list.txt", "-o", "sub/u", "-a", "ull", "-p", "10").returncode == 0
    r = run("dist", "-q", "./sub/u", "-r", "sub/u", "-o", "u.tsv")
    assert r.returncode == 0, r.stderr
    assert len(_parse_list(str(tmp_path / "u.tsv"))) == 6           # 3 * 4 / 2 pairs, not 9


def test_damaged_record_keeps_the_sketch_so_far(oracle, gpu_ctx, tmp_path):
    """utils.rs:458 `if let Ok(seqrec) = res`: a record that fails to parse is skipped and the file keeps its sketch; only a
    file that cannot be opened / recognised at all is "Invalid input file" (utils.rs:453)."""
    rng = np.random.default_rng(2)
    reads = [synth.to_ascii(rng.integers(0, 4, size=150, dtype=np.uint8)) for _ in range(300)]
    good, bad = str(tmp_path / "good.fq"), str(tmp_path / "bad.fq")
    _write_fastq(good, reads)
    body = open(good, "rb").read()
    cut = body.index(b"@", len(body) // 2)                                     # a record boundary in the middle
    n_ok = body[:cut].count(b"\n") // 4
    open(bad, "wb").write(body[:cut] + b"@broken\nACGTACGTACGTACGTACGTACGTAC\n+\nIII\n" + body[cut:])   # seq / qual lengths differ
    regs, st = hostapi.sketch_files_regs(gpu_ctx, ALGO_ULL, 12, 21, 42, [good, bad, good], threads=2)
    exp = oracle.sketch_genomes(oracle.ULL, 12, 21, 42, [reads, reads[:n_ok], reads], threads=2)
    assert np.array_equal(regs, exp)


@pytest.mark.parametrize("dm", [False, True])
@pytest.mark.parametrize("same", [True, False])
def test_row_sharded_dist_parts_concatenate_to_the_single_file(gpu_ctx, tmp_path, dm, same):
    """One process per GPU: rank r of world w writes its reference-row range to <out>.part000r; the parts in
    rank order are byte for byte the single-process file (triangular cuts balance pairs, not rows)."""
    genomes = synth.genomes(23, 30_000, seed=17)
    files = []
    for g, recs in enumerate(genomes):
        path = str(tmp_path / f"r{g}.fa")
        _write_fasta(path, recs)
        files.append(path)
    a, b = str(tmp_path / "A"), str(tmp_path / "B")
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 16, 42, files, a)
    hostapi.sketch_files(gpu_ctx, ALGO_ULL, 10, 16, 42, files[5:12], b)
    q = a if same else b
    single = str(tmp_path / "single.out")
    hostapi.dist(gpu_ctx, a, q, single, dm=dm, threads=2, fused=True)
    for world in (2, 3, 8):
        out = str(tmp_path / f"sharded{world}.out")
        for rank in range(world):
            hostapi.dist_rows(gpu_ctx, a, q, out, rank, world, dm=dm, threads=2)
        joined = b"".join(open(f"{out}.part{r:04d}", "rb").read() for r in range(world))
        assert joined == open(single, "rb").read()
        sizes = [os.path.getsize(f"{out}.part{r:04d}") for r in range(world)]
        if same and world == 2 and not dm:
            assert abs(sizes[0] - sizes[1]) < 0.25 * max(sizes)       # pairs, not rows, are balanced
