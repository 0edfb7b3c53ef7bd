"""CPU tests of the C++ host layer (liblash_host.so): reader, filter + packer, file formats, `{:.6}`.
No GPU and no sketch arithmetic here -- the oracle is used only as the checker for filter_out_n."""
import bz2
import ctypes as C
import gzip
import json
import lzma
import os
import random
import struct
import subprocess

import numpy as np
import pytest

from lash_b200 import ALGO_HLL, ALGO_HMH, ALGO_ULL, capi, hostapi
from lash_b200.pack import encode_record, pack_codes

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_host_library_loads_and_exports_every_declared_symbol():
    import re
    hdr = open(os.path.join(ROOT, "include", "lash_host.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(lash_(?:host|fastx)_\w+)\s*\(", hdr))
    assert declared and declared == set(hostapi.PROTOTYPES)
    L = hostapi.lib()
    for name in declared:
        assert getattr(L, name) is not None
    out = subprocess.run(["nm", "-D", "--defined-only", hostapi.lib_path()], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    assert declared <= exported


def _dirty(rng, n):
    alphabet = b"ACGT" * 12 + b"acgtNnRYKM\n\r >@-*" + bytes([0, 255, 0x81, 0x41 + 16, 0x54 - 16, 0x47 + 128])
    return bytes(rng.choice(alphabet) for _ in range(n))


def _need_isa(simd):
    if simd in (1, 2) and not hostapi.lib().lash_host_pack_has_simd():
        pytest.skip("no AVX2+BMI2 on this CPU")
    if simd == 3 and hostapi.lib().lash_host_pack_isa() < 2:
        pytest.skip("no AVX-512 VBMI2 on this CPU")


@pytest.mark.parametrize("simd", [0, 2, 3, 1], ids=["scalar", "avx2", "avx512", "default"])
def test_filter_pack_matches_reference_front_end(oracle, simd):
    """filter_out_n (utils.rs:33-41) + KSeq 2-bit codes: C++ packer == oracle filter + numpy packing,
    on clean and dirty input, at every length / alignment around the 32- and 64-byte SIMD blocks."""
    _need_isa(simd)
    rng = random.Random(5)
    for n in list(range(0, 100)) + [127, 128, 129, 1000, 4097, 65536 + 31]:
        for maker in (lambda m: bytes(rng.choice(b"ACGT") for _ in range(m)), lambda m: _dirty(rng, m)):
            s = maker(n)
            kept = oracle.filter_out_n(s)
            want = pack_codes(encode_record(kept))
            got, nb = hostapi.filter_pack(s, simd=simd)
            assert nb == len(kept)
            assert np.array_equal(got, want), (n, simd)


@pytest.mark.parametrize("simd", [0, 2, 3, 1], ids=["scalar", "avx2", "avx512", "default"])
def test_filter_pack_appends_at_any_base_offset(oracle, simd):
    _need_isa(simd)
    rng = random.Random(6)
    whole = b""
    packed, nb = None, 0
    for _ in range(80):
        piece = _dirty(rng, rng.randrange(0, 200))       # pieces on both sides of the 64-byte block, every bit offset
        whole += piece
        packed, nb = hostapi.filter_pack(piece, packed, nb, simd=simd)
        kept = oracle.filter_out_n(whole)
        assert nb == len(kept)
        assert np.array_equal(packed, pack_codes(encode_record(kept)))


def test_all_256_byte_values_classified_like_filter_out_n(oracle):
    s = bytes(range(256)) * 3
    got, nb = hostapi.filter_pack(s, simd=True)
    got2, nb2 = hostapi.filter_pack(s, simd=False)
    got3, nb3 = hostapi.filter_pack(s, simd=2)
    got4, nb4 = hostapi.filter_pack(s, simd=3)       # falls back to AVX2 / scalar where there is no AVX-512
    assert nb == nb2 == nb3 == nb4 == 12 and np.array_equal(got, got2) and np.array_equal(got, got3) and np.array_equal(got, got4)
    assert np.array_equal(got, pack_codes(encode_record(b"ACGT" * 3)))


FASTA = b">s1 first record\nACGTNN\nacgtACGT\n\n>s2\n>s3 empty above\r\nAC>GT\r\nTTTT\n>s4 no newline at end\nGGGG"
FASTA_RECS = [(b"s1 first record", b"ACGTNNacgtACGT"), (b"s2", b""), (b"s3 empty above", b"AC>GTTTTT"), (b"s4 no newline at end", b"GGGG")]
FASTQ = b"@r1 x\nACGTN\n+\nIIIII\n@r2\nAC\n+r2\n@@\n@r3\r\nGGG\r\n+\r\nIII\r\n"
FASTQ_RECS = [(b"r1 x", b"ACGTN"), (b"r2", b"AC"), (b"r3", b"GGG")]


def _zstd_compress(data: bytes) -> bytes:
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_compressBound.restype = C.c_size_t
    z.ZSTD_compressBound.argtypes = [C.c_size_t]
    z.ZSTD_compress.restype = C.c_size_t
    z.ZSTD_compress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_int]
    cap = z.ZSTD_compressBound(len(data))
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_compress(buf, cap, data, len(data), 3)
    return buf.raw[:n]


def _zstd_decompress(data: bytes, cap: int) -> bytes:
    z = C.CDLL("libzstd.so.1")
    z.ZSTD_decompress.restype = C.c_size_t
    z.ZSTD_decompress.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t]
    buf = C.create_string_buffer(cap)
    n = z.ZSTD_decompress(buf, cap, data, len(data))
    assert n <= cap
    return buf.raw[:n]


@pytest.mark.parametrize("codec", ["plain", "gz", "gz2", "bz2", "xz", "zst"])
@pytest.mark.parametrize("kind", ["fasta", "fastq"])
def test_fastx_reader_formats_and_compression(tmp_path, codec, kind):
    """needletail behaviour the reference relies on (utils.rs:453-458): format and compression
    are sniffed from content; seq() has line breaks stripped; multi-member gzip is read through."""
    data, recs = (FASTA, FASTA_RECS) if kind == "fasta" else (FASTQ, FASTQ_RECS)
    enc = {"plain": lambda b: b, "gz": gzip.compress, "gz2": lambda b: gzip.compress(b[:20]) + gzip.compress(b[20:]),
           "bz2": bz2.compress, "xz": lzma.compress, "zst": _zstd_compress}[codec]
    p = tmp_path / "in.dat"   # the extension is deliberately meaningless
    p.write_bytes(enc(data))
    assert list(hostapi.read_fastx(str(p))) == recs


def test_fastx_reader_large_records_cross_buffer_boundaries(tmp_path):
    rng = np.random.default_rng(3)
    seqs = [bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)) for n in (5_000_000, 1, 70, 9_000_000)]
    with open(tmp_path / "big.fa", "wb") as f:
        for i, s in enumerate(seqs):
            f.write(b">g%d\n" % i)
            for o in range(0, len(s), 70):
                f.write(s[o:o + 70] + b"\n")
    got = list(hostapi.read_fastx(str(tmp_path / "big.fa")))
    assert [g[1] for g in got] == seqs
    with gzip.open(tmp_path / "reads.fq.gz", "wb", compresslevel=1) as f:
        for i in range(30000):
            f.write(b"@r%d\n%s\n+\n%s\n" % (i, seqs[0][i * 150:(i + 1) * 150], b"I" * 150))
    got = list(hostapi.read_fastx(str(tmp_path / "reads.fq.gz")))
    assert len(got) == 30000 and got[29999] == (b"r29999", seqs[0][29999 * 150:30000 * 150])


@pytest.mark.parametrize("content", [b"", b"ACGT\n", b"@r1\nACGT\n+\nII\n", b"@r1\nACGT\nIIII\nIIII\n", b"@r1\nACGT\n+\n"])
def test_fastx_reader_rejects_invalid_input(tmp_path, content):
    """`parse_fastx_file(..).expect("Invalid input file")` (utils.rs:453): empty / unrecognisable / malformed."""
    p = tmp_path / "bad"
    p.write_bytes(content)
    with pytest.raises(hostapi.HostError):
        list(hostapi.read_fastx(str(p)))
    with pytest.raises(hostapi.HostError):
        list(hostapi.read_fastx(str(tmp_path / "does_not_exist")))


def test_fixed6_is_rusts_format(oracle):
    """`{:.6}` (main.rs:459,465) prints the exactly-rounded decimal; CPython's %.6f does the same."""
    rng = np.random.default_rng(11)
    vals = [0.0, -0.0, 1.0, 0.5, 0.0000005, 0.00000049999999, 0.0000015, 0.0000025, 1e-300, 5e-324, 0.9999995, 0.99999949999,
            0.1234565, 0.1234575, 2.5e-7, 7.5e-7, 123456.7890125, 1048575.9999996, 1e15, 1e300]
    vals += list(rng.random(20000)) + list(rng.random(2000) * 1e-5) + [float(i) / 2**21 for i in range(0, 2**21, 997)]
    for v in vals:
        assert hostapi.format_fixed6(v) == "%.6f" % v, v
    for v in list(rng.random(5000).astype(np.float32)) + [np.float32(0.1), np.float32(1e-7), np.float32(0.0000005)]:
        assert hostapi.format_fixed6(float(v), fp32=True) == "%.6f" % float(v)
    # the bulk writer's fast path gives the same text, including values a hair from a rounding boundary
    half = (np.arange(0, 4000, dtype=np.float64) + 0.5) * 1e-6
    near = np.concatenate([half, np.nextafter(half, 0), np.nextafter(half, 1), half * (1 + 3e-13), half * (1 - 3e-13)])
    bulk_in = np.concatenate([np.array(vals, dtype=np.float64), near, rng.random(200000), -rng.random(100), rng.random(1000) * 1500,
                              np.array([np.nan, np.inf, -np.inf, -0.0, 1023.9999995, 1024.0, 1e300])])
    got = hostapi.format_fixed6_bulk(bulk_in)
    exp = ["NaN" if np.isnan(v) else "%.6f" % v for v in bulk_in]
    assert got == exp
    assert hostapi.format_fixed6(float("nan")) == "NaN" and hostapi.format_fixed6(float("inf")) == "inf"
    assert hostapi.format_fixed6(float("-inf")) == "-inf" and hostapi.format_fixed6(-0.0) == "-0.000000"


@pytest.mark.parametrize("algo,p", [(ALGO_HMH, 14), (ALGO_ULL, 10), (ALGO_ULL, 3), (ALGO_HLL, 12), (ALGO_HLL, 4)])
def test_sketch_file_layout_and_round_trip(tmp_path, algo, p):
    """`_sketches.bin` = one zstd stream of S::save records, byte for byte."""
    rng = np.random.default_rng(p)
    n = 5
    if algo == ALGO_HMH:
        regs = rng.integers(0, 65536, size=(n, 16384), dtype=np.uint16)
    else:
        regs = rng.integers(0, 40, size=(n, 1 << p), dtype=np.uint8)
        regs[0, :] = 0
    path = str(tmp_path / "x_sketches.bin")
    hostapi.write_sketches(path, algo, p, regs, threads=2)
    raw = _zstd_decompress(open(path, "rb").read(), n * (regs[0].nbytes + 64))
    off = 0
    for i in range(n):
        if algo == ALGO_HMH:
            body = raw[off:off + 32768]
            assert np.array_equal(np.frombuffer(body, dtype="<u2"), regs[i])
            off += 32768
            continue
        m = 1 << p
        if algo == ALGO_HLL:
            alpha, zero, s, pp = struct.unpack_from("<dQdB", raw, off)
            assert pp == p and zero == int((regs[i] == 0).sum())
            assert s == float(np.sum(np.ldexp(1.0, -regs[i].astype(np.int64))))
            assert alpha == ({4: 0.673, 5: 0.697, 6: 0.709}.get(p) or 0.7213 / (1.0 + 1.079 / m))
            off += 25
        (length,) = struct.unpack_from("<Q", raw, off)
        assert length == m
        assert raw[off + 8:off + 8 + m] == regs[i].tobytes()
        off += 8 + m
    assert off == len(raw)
    back, got_p = hostapi.read_sketches(path, algo, n)
    assert np.array_equal(back, regs) and (algo == ALGO_HMH or got_p == p)
    with pytest.raises(hostapi.HostError):   # reading more sketches than the stream holds: read_exact fails
        hostapi.read_sketches(path, algo, n + 1, p)
    if algo != ALGO_HMH:
        with pytest.raises(hostapi.HostError):
            hostapi.read_sketches(path, algo, n, p + 1)


def test_parameters_json_is_what_serde_json_writes(tmp_path):
    """main.rs:249-276: string values, BTreeMap key order, two-space pretty printing."""
    out = str(tmp_path / "sk")
    hostapi.check(hostapi.lib().lash_host_write_parameters(out.encode(), ALGO_ULL, 10, 16, 42))
    assert open(out + "_parameters.json").read() == (
        '{\n  "algorithm": "ull",\n  "k": "16",\n  "molecule": "nucleotide",\n  "precision": "10",\n  "seed": "42"\n}')
    hostapi.check(hostapi.lib().lash_host_write_parameters(out.encode(), ALGO_HMH, 14, 21, 7))
    assert json.load(open(out + "_parameters.json")) == {"algorithm": "hmh", "k": "21", "molecule": "nucleotide", "seed": "7"}


def test_filter_pack_simd_paths_agree_with_scalar_on_long_dirty_input():
    """Differential check of the 32- and 64-byte SIMD packers against the scalar table on long inputs: FASTA-like text
    (80-column lines), runs of N / lower case, and uniformly random bytes, appended in pieces of random size."""
    rng = np.random.default_rng(11)
    text = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=300_000)
    text[80::81] = 10                                                     # newlines
    for o in rng.integers(0, len(text) - 500, size=40):
        text[o:o + int(rng.integers(1, 400))] = rng.choice(np.frombuffer(b"Nacgtn", dtype=np.uint8))
    noise = rng.integers(0, 256, size=100_000, dtype=np.uint8)
    data = np.concatenate([text, noise, text[::-1]]).tobytes()
    cuts = sorted(set([0, len(data)] + [int(x) for x in rng.integers(0, len(data), size=25)]))
    ref = None
    for simd in (0, 2, 3, 1):
        packed, nb = None, 0
        for a, b in zip(cuts[:-1], cuts[1:]):
            packed, nb = hostapi.filter_pack(data[a:b], packed, nb, simd=simd)
        if ref is None:
            ref = (packed, nb)
            assert nb > 500_000
        else:
            assert nb == ref[1] and np.array_equal(packed, ref[0]), simd


def test_cli_surface_without_a_gpu():
    """`lash-b200`: the reference's command surface (main.rs:26-177) -- help, per-command help with every flag and
    default of the reference, version, and clap-style errors for unknown / missing arguments, all before any GPU is
    touched.  (With sketches in hand and no GPU the commands must fail loudly: there is no CPU fallback.)"""
    import subprocess
    exe = os.path.join(os.path.dirname(hostapi.lib_path()), "lash-b200")
    run = lambda *a: subprocess.run([exe, *a], capture_output=True, text=True)  # noqa: E731
    r = run("-V")
    assert r.returncode == 0 and r.stdout.startswith("lash-b200 ")
    r = run("sketch", "-h")
    assert r.returncode == 0
    for flag in ("-f, --file", "-o, --output", "-k, --kmer", "-t, --threads", "-a, --algorithm", "-p, --precision", "-s, --seed",
                 "[default: sketch]", "[default: 16]", "[default: hmh]", "[default: 10]", "[default: 42]"):
        assert flag in r.stdout, flag
    r = run("help", "dist")
    assert r.returncode == 0
    for flag in ("-q, --query", "-r, --reference", "-o, --output_file", "-e, --estimator", "-m, --model", "--fp32", "--dm",
                 "[default: fgra]", "[default: 1]"):
        assert flag in r.stdout, flag
    assert run("dist", "--help").stdout == r.stdout
    r = run("dist", "--bogus")
    assert r.returncode != 0 and "unexpected argument '--bogus'" in r.stderr
    r = run("sketch")
    assert r.returncode != 0 and "--file <file>" in r.stderr
    r = run("dist", "-q", "a")
    assert r.returncode != 0 and "--reference <reference>" in r.stderr
    assert run().returncode != 0 and "Usage: lash-b200 <COMMAND>" in run().stderr
