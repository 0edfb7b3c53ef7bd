/*
 * lash_host.h -- C ABI of the native (C++) host layer that sits ABOVE lash_gpu.h.
 *
 * The reference's host is Rust; no Rust toolchain exists in this image, so the host side of the hot
 * paths (the callers and data formats either side of the two GPU paths, SURVEY.md section 8f rows
 * 1-3) is C++17 (lash_b200/host/, built into lash_b200/_lib/liblash_host.so).  The C++ API in
 * lash_b200/host/lash_host.hpp mirrors the reference's generics one to one
 *     sketch_files<S>(precision, files, kmer_length, output_name, threads, seed)   src/utils.rs:439-510,566-580
 *     {hmh,ull,hll}_distance<T>(names, sketch files, [estimator], create_matrix, same_files, emit)
 *                                                                                   src/utils.rs:84-373
 *     compute_distance<T>, print_dist<T>, the `dist` sub-command body               src/main.rs:279-613
 * and this header exports the same operations with plain pointers for bindings and tests.
 *
 * Every sketch / distance number is computed by liblash_gpu.so (CUDA).  The host layer only parses,
 * filters + 2-bit packs (utils.rs:33-41,464), stages, serialises and formats; none of it falls back
 * to a CPU sketch or estimator.
 *
 * Return convention: 0 ok, <0 error (lash_host_last_error()), >0 warning passed through from
 * lash_dist (LASH_W_HLL_BIAS_REGIME).
 */
#ifndef LASH_HOST_H
#define LASH_HOST_H

#include <stddef.h>
#include <stdint.h>

#include "lash_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

#define LASH_HOST_E_IO -10      /* open / read / write failure, or a (de)compression library is missing */
#define LASH_HOST_E_FORMAT -11  /* malformed FASTA/FASTQ ("Invalid input file"), sketch or JSON file */
#define LASH_HOST_E_PARAMS -12  /* the reference's parameter-mismatch panics (main.rs:362-377) */

const char* lash_host_last_error(void);

/* ---- front end: needletail-style record reader (utils.rs:453-456) --------------------------------
 * Format (FASTA '>' / FASTQ '@') and compression (gzip, bzip2, xz, zstd, none) are sniffed from the
 * content, not the file name.  Sequences are returned with line breaks removed (needletail's seq()). */
typedef struct lash_fastx lash_fastx;
int lash_fastx_open(const char* path, lash_fastx** out);
/* 1: a record was returned, 0: end of file, <0: error.  Pointers stay valid until the next call. */
int lash_fastx_next(lash_fastx* r, const char** id, size_t* id_len, const char** seq, size_t* seq_len);
int lash_fastx_close(lash_fastx* r);

/* ---- filter_out_n (utils.rs:33-41) fused with the 2-bit packing of KSeq::new(&seq, 2) (utils.rs:464) ----
 * Appends the bases of seq[0..n) that are one of "ACGT" to a packed stream that already holds
 * *n_bases bases (lash_gpu.h format: A0 C1 G2 T3, first base in the high bits of each byte);
 * `packed` must have room for (*n_bases + n + 3) / 4 + 16 bytes.  use_simd: 0 scalar table, 1 the default SIMD
 * path (the 64-byte AVX-512 VBMI2 path where the CPU has it, else AVX2+BMI2 with 32 bytes per step; LASH_PACK_ISA=avx2 forces
 * the latter), 2 AVX2, 3 AVX-512. */
int lash_host_filter_pack(const uint8_t* seq, size_t n, uint8_t* packed, uint64_t* n_bases, int use_simd);
/* 1 when the AVX2+BMI2 packer is usable on this CPU */
int lash_host_pack_has_simd(void);
/* 0 = scalar only, 1 = AVX2+BMI2, 2 = AVX-512 (F/BW/VL/VBMI2) */
int lash_host_pack_isa(void);

/* ---- sketch_files<S> (utils.rs:439-510): FASTA/FASTQ files -> registers ----------------------------
 * One host worker per file at a time ("parallel by sample"), each parsing + packing into pinned
 * chunks that are pushed to the GPU (lash_sketch_push) while the next chunk is parsed.
 * regs_out: n_files * lash_sketch_reg_bytes(algo, p) bytes, in list order.
 * chunk_bytes: pinned staging chunk per worker buffer (0 = sized from the input, 1..16 MiB). */
typedef struct lash_sketch_files_stats {
    uint64_t n_records;      /* records seen (all files) */
    uint64_t n_bases_in;     /* sequence bytes read, before filter_out_n */
    uint64_t n_bases_kept;   /* bases that survive filter_out_n (each counted once) */
    uint64_t n_pushes;       /* lash_sketch_push calls */
    double seconds_total;    /* wall time of the call */
    double gpu_kernel_ms;    /* sum of sketch kernel time (lash_sketch_stats) */
    double seconds_open;     /* lash_sketch_open + first pinned staging chunks (page-locking when the pool is cold) */
    double seconds_workers;  /* parse + pack + push, all workers, until the last one joined */
    double seconds_drain;    /* lash_sketch_fetch (waits for the last kernels) + close */
} lash_sketch_files_stats;

int lash_host_sketch_files_regs(lash_ctx* ctx, int algo, int p, int k, uint64_t seed, const char* const* files,
                                uint64_t n_files, int threads, uint64_t chunk_bytes, void* regs_out,
                                lash_sketch_files_stats* stats);
/* Same, then writes {output_name}_sketches.bin (one zstd stream, level 3, records in list order,
 * utils.rs:566-575) and {output_name}_files.json (utils.rs:577-580). */
int lash_host_sketch_files(lash_ctx* ctx, int algo, int p, int k, uint64_t seed, const char* const* files,
                           uint64_t n_files, const char* output_name, int threads, lash_sketch_files_stats* stats);
/* Measurement aid: the same parse + filter + pack over the files, chunks dropped instead of pushed
 * (no GPU work, no registers) -- the host ingest ceiling bench.py reports next to the end-to-end rate. */
int lash_host_pack_files_dry(const char* const* files, uint64_t n_files, int k, int threads, uint64_t chunk_bytes,
                             lash_sketch_files_stats* stats);
/* Who runs filter_out_n + the 2-bit packing for sketch_files: 0 = auto (the host packer when the CPU has its SIMD path, else
 * the device), 1 = the host packer (lash_sketch_push: 0.25 B/base over PCIe), 2 = the device (lash_sketch_push_ascii: the
 * host only moves sequence bytes into pinned chunks -- plain FASTA is read() straight into them -- 1 B/base over PCIe).
 * Process-wide;
 * the environment variable LASH_INGEST=packed|ascii overrides it.  Registers are identical either way.
 * In device mode lash_sketch_files_stats.n_bases_kept is 0 (the host never looks at the bases). */
int lash_host_set_ingest_mode(int mode);
/* sketch_files keeps its pinned staging blocks (up to 1 GiB) for the next call, because page-locking costs more
 * than sketching a small batch; this returns them to the driver. */
int lash_host_release_pinned(void);
/* {output_name}_parameters.json as the `sketch` sub-command writes it (main.rs:249-276). */
int lash_host_write_parameters(const char* output_name, int algo, int p, int k, uint64_t seed);

/* ---- sketch file format (S::save / S::load, utils.rs:95-105,202-222,303-319,400-433) ----------------
 * HMH: u16 LE x 16384.  ULL: bincode {state: Vec<u8>} = u64 LE length + bytes.
 * HLL: bincode {alpha f64, zero u64, sum f64, p u8, m: u64 length + bytes}. */
int lash_host_write_sketches(const char* path, int algo, int p, const void* regs, uint64_t n, int threads);
/* Reads exactly n sketches.  *p_inout: expected precision, or 0 to take it from the file (HLL/ULL). */
int lash_host_read_sketches(const char* path, int algo, int* p_inout, uint64_t n, void* regs_out);

/* ---- `lash dist` body (main.rs:279-613) ------------------------------------------------------------
 * ref_prefix / query_prefix: the three files "<prefix>*parameters.json", "*files.json", "*.bin" are
 * found by basename prefix in the prefix's directory (the reference looks in the CWD only).
 * estimator: "fgra" | "ml" (ULL only); model 0|1; dm: --dm matrix output; fp32: --fp32.
 * fused != 0: compute_distance runs inside the GPU kernel and the writer only formats (fast path);
 * fused == 0: the kernel returns `frac`, and compute_distance + the name rule run on the host in T
 * exactly as print_dist does (main.rs:452-456).  Rows are written in list order (the reference's
 * order is hashbrown iteration order; only the set of pairs is contractual). */
int lash_host_dist(lash_ctx* ctx, const char* ref_prefix, const char* query_prefix, const char* output_file,
                   const char* estimator, int model, int dm, int fp32, int threads, int fused);

/* One process per GPU: process `rank` of `world` computes and writes only its range of reference rows (cut so that
 * pair counts are equal) into "<output_file>.part<rank, 4 digits>"; rank 0's part carries the header line; the parts
 * concatenated in rank order are byte for byte the file lash_host_dist(fused) writes.  world == 1: plain output_file. */
int lash_host_dist_rows(lash_ctx* ctx, const char* ref_prefix, const char* query_prefix, const char* output_file,
                        const char* estimator, int model, int dm, int fp32, int threads, int rank, int world);

/* Rust's `{:.6}` for f64 / f32 (main.rs:459,465): exact decimal expansion, round-half-even.
 * Writes at most 32 bytes for |v| < 2^20, at most 330 otherwise (no terminator); returns the length. */
int lash_host_format_fixed6_f64(double v, char* out);
int lash_host_format_fixed6_f32(float v, char* out);
/* The bulk writer's formatter (fast path + exact fallback; same text): n values, '\n' after each; `out` needs 341
 * bytes per value in the worst case.  Returns the bytes written. */
size_t lash_host_format_fixed6_bulk(const double* v, size_t n, char* out);

#ifdef __cplusplus
}
#endif
#endif /* LASH_HOST_H */
