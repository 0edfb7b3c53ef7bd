/*
 * lash_gpu.h -- C ABI of the B200 (sm_100a) implementation of lash's two hot paths.
 *
 * This is the drop-in boundary a host (the Rust `lash` binary through a bindgen/cc FFI crate,
 * or the C++/Python hosts in this repo) binds.  The reference has no FFI layer of its own; the
 * seams it replaces are Rust generics (file:line under the reference tree):
 *
 *   sketch side  src/utils.rs:377-386  trait KmerSketch { new, add_kmer, save }
 *                src/utils.rs:457-503  the per-file record loop (filter -> 2-bit -> k-mers ->
 *                                      canonical -> mask -> add_kmer).  A per-k-mer call is far too
 *                                      fine for a device boundary, so the ABI lifts the seam to
 *                                      "packed bases of whole records in, registers out".
 *   dist side    src/utils.rs:150-180 (hmh_distance), :248-285 (ull_distance), :342-370
 *                (hll_distance) par_iter bodies + src/main.rs:415-423 compute_distance.
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (lash_gpu_last_error() gives the
 *     thread-local message); no exceptions cross the boundary; nothing falls back to the CPU.
 *   - the caller owns every host buffer; the library owns device memory behind opaque handles.
 *   - a handle is not thread-safe; distinct handles are (one sketcher per host worker, like
 *     "one sketch object per rayon task", utils.rs:454).
 *   - register layout == what `S::save` serialises (utils.rs:400-433): HMH u16[16384] (host
 *     endianness, LE on every supported host), HLL/ULL u8[2^p].
 */
#ifndef LASH_GPU_H
#define LASH_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LASH_GPU_ABI_VERSION 2

/* algorithm ids (main.rs:210-246: "hmh" | "hll" | "ull") */
#define LASH_ALGO_HMH 0
#define LASH_ALGO_HLL 1
#define LASH_ALGO_ULL 2
/* ULL estimator (main.rs:144-149, utils.rs:214-218) */
#define LASH_EST_FGRA 0
#define LASH_EST_ML 1
/* Mash distance model (main.rs:415-423): 0 binomial, 1 poisson */
#define LASH_MODEL_BINOMIAL 0
#define LASH_MODEL_POISSON 1
/* not a model: return `frac = 2s/(1+s)` itself (cast to T), i.e. exactly the value the reference's
 * *_distance functions hand to emit() (utils.rs:176,277,364) -- lets a host keep print_dist /
 * compute_distance (main.rs:415-471) unchanged */
#define LASH_MODEL_FRAC 2

/* error codes */
#define LASH_OK 0
#define LASH_E_INVALID -1     /* bad argument (k outside 1..32: utils.rs:500-502; p out of range; ...) */
#define LASH_E_CUDA -2        /* CUDA runtime failure, message has the cudaError string */
#define LASH_E_NOMEM -3       /* device or pinned-host allocation failed */
#define LASH_E_STATE -4       /* call sequence error (e.g. fetch before sync) */
/* lash_dist* warning (positive): some pair hit the HLL++ bias-table regime (estimate <= 5m outside
 * linear counting).  Google's empirical bias tables are not reproducible offline; those cells are
 * computed with union = NaN (=> s = max(NaN,0) = 0, distance 1) and counted here. */
#define LASH_W_HLL_BIAS_REGIME 1

typedef struct lash_ctx lash_ctx;
typedef struct lash_sketcher lash_sketcher;

const char* lash_gpu_last_error(void);
int lash_gpu_abi_version(void);
int lash_gpu_device_count(void);

/* One context per GPU (one host process/thread per GPU is the intended deployment). */
int lash_ctx_create(int device, lash_ctx** out);
int lash_ctx_destroy(lash_ctx* ctx);
int lash_ctx_device(const lash_ctx* ctx);
/* Multi-GPU hosts (one process or thread per GPU): pin the CALLING thread to the CPUs next to `device` (the sysfs
 * local_cpulist of its PCI function, intersected with the CPUs the process may use).  Call it before lash_host_alloc /
 * lash_ctx_create: pinned staging memory lands on the NUMA node of the allocating thread, and a rank whose staging
 * memory sits on the far socket pays the inter-socket link for every H2D byte.  Returns the number of CPUs in the new
 * mask (> 0), or < 0 when the box exposes no such information (single-socket boxes: harmless to ignore). */
int lash_bind_thread_to_device(int device);

/* Pinned host memory for the packer (so lash_sketch_push copies are truly asynchronous). */
int lash_host_alloc(size_t bytes, void** out);
int lash_host_free(void* p);

/* register bytes of one sketch: HMH 32768, HLL/ULL 2^p */
size_t lash_sketch_reg_bytes(int algo, int p);

/* ------------------------------------------------------------------------------------------------
 * Sketching.  Replaces utils.rs:454-507 for a whole batch of files.
 *
 * Packed base format (what kmerutils `Sequence::new(&seq, 2)` holds, utils.rs:464): A=0 C=1 G=2 T=3,
 * four bases per byte, FIRST base in the MOST significant two bits of the byte.  Only bases that
 * survive filter_out_n (utils.rs:33-41: uppercase A/C/G/T) are packed; records shorter than k may be
 * passed or dropped by the host (the kernel produces no k-mer for them, utils.rs:460-462).
 *
 * A push carries one contiguous buffer holding any number of *spans*.  A span is a run of records
 * of ONE genome (input file), packed densely back to back (a record may start in the middle of a
 * byte); each span starts at a 16-byte aligned offset of the buffer and the buffer must be readable
 * up to a multiple of 16 bytes past the last span (lash_sketch_padded_bytes()).  k-mers never cross
 * record boundaries (utils.rs:457-464).
 * ---------------------------------------------------------------------------------------------- */
typedef struct lash_span {
    uint64_t genome;     /* accumulator slot in [0, n_genomes) */
    uint64_t byte_off;   /* offset of the span's first base in the push buffer, multiple of 16 */
    uint64_t n_bases;    /* bases in the span (all records together) */
    uint64_t rec_first;  /* index of the span's first entry in rec_start[] (ignored when n_rec <= 1) */
    uint32_t n_rec;      /* records in the span; 0 or 1: the whole span is one record */
    uint32_t rec_len;    /* 0: boundaries come from rec_start[]; != 0: every record has rec_len bases (the last
                            may be shorter), n_rec = ceil(n_bases / rec_len), rec_start[] is not read -- fixed-length
                            reads (config "150 bp FASTQ") then cost no 8 B/read boundary table over PCIe */
} lash_span;

/* bytes a span of n_bases occupies in a push buffer, including alignment padding */
uint64_t lash_sketch_padded_bytes(uint64_t n_bases);

/* algo: LASH_ALGO_*; p: precision for HLL (4..18) / ULL (3..26), ignored for HMH (fixed 14);
 * k in [1,32]; seed: xxh3 seed (main.rs:88-94, default 42); n_genomes accumulators, zeroed. */
int lash_sketch_open(lash_ctx* ctx, int algo, int p, int k, uint64_t seed, uint64_t n_genomes, lash_sketcher** out);

/* Enqueue H2D copy + kernels for one buffer.  rec_start: for spans with n_rec > 1, the n_rec+1
 * ascending base offsets (relative to the span's first base; first = 0, last = n_bases) stored at
 * rec_start[rec_first .. rec_first+n_rec].  n_rec_entries = length of rec_start (may be 0 / NULL).
 * Returns immediately when `packed` is pinned; *ticket (optional) identifies the push. */
int lash_sketch_push(lash_sketcher* s, const uint8_t* packed, uint64_t n_bytes, const lash_span* spans,
                     uint32_t n_spans, const uint64_t* rec_start, uint64_t n_rec_entries, uint64_t* ticket);
/* Same, but `packed` already lives in device memory of the context's GPU (no copy). */
int lash_sketch_push_dev(lash_sketcher* s, const void* packed_dev, uint64_t n_bytes, const lash_span* spans,
                         uint32_t n_spans, const uint64_t* rec_start, uint64_t n_rec_entries, uint64_t* ticket);
/* ---- raw sequence text in: filter_out_n (utils.rs:33-41) + 2-bit packing (utils.rs:464) on the device -------------------
 * For hosts with few cores per GPU: the host only finds the records and copies their sequence bytes into a pinned chunk
 * (1 B/base over PCIe instead of 0.25, but no per-base CPU work); the device deletes every byte that is not one of "ACGT"
 * (line breaks, N, lowercase, IUPAC codes, anything -- flanks are joined, exactly filter_out_n), packs, and sketches.
 * A text span = the sequence bytes of the records of ONE genome, back to back; records are separated IN BAND by one
 * LASH_TEXT_RECORD_SEP byte (k-mers never cross it, utils.rs:457-464; a record that keeps fewer than k bases yields
 * nothing, utils.rs:460-462).  The separator may appear before the first or after the last record and repeatedly.  A host
 * must not pass a sequence byte equal to the separator (replace it by any other non-ACGT byte: the filter deletes both).
 * byte_off: multiple of 16; the buffer must be readable up to the next multiple of 16 past each span.
 * n_rec: 0 or 1 = the span holds no separator (whole genomes: saves the boundary bitmask), otherwise any value > 1. */
#define LASH_TEXT_RECORD_SEP 0x01
typedef struct lash_text_span {
    uint64_t genome;   /* accumulator slot in [0, n_genomes) */
    uint64_t byte_off; /* offset of the span's first byte in the push buffer, multiple of 16 */
    uint64_t n_bytes;  /* raw bytes in the span */
    uint32_t n_rec;    /* see above */
    uint32_t reserved; /* 0 */
} lash_text_span;
int lash_sketch_push_ascii(lash_sketcher* s, const uint8_t* text, uint64_t n_bytes, const lash_text_span* spans,
                           uint32_t n_spans, uint64_t* ticket);
/* Same, `text` already in device memory of the context's GPU (16-byte aligned). */
int lash_sketch_push_ascii_dev(lash_sketcher* s, const void* text_dev, uint64_t n_bytes, const lash_text_span* spans,
                               uint32_t n_spans, uint64_t* ticket);
/* Block until the H2D copy of push `ticket` has completed (its host buffer may be reused). */
int lash_sketch_wait_copied(lash_sketcher* s, uint64_t ticket);
/* Block until every enqueued push has been folded into the accumulators. */
int lash_sketch_sync(lash_sketcher* s);
/* Copy registers of genomes [first, first+n) to host: n * lash_sketch_reg_bytes bytes. Implies sync. */
int lash_sketch_fetch(lash_sketcher* s, uint64_t first, uint64_t n, void* regs_out);
/* Device pointer to the accumulator array [n_genomes][reg_bytes] (valid until close; after sync). */
int lash_sketch_regs_dev(lash_sketcher* s, void** regs_dev);
/* Zero all accumulators (reuse the sketcher for another batch).  Synchronises first, unless a
 * caller stream is set (lash_sketch_set_stream), in which case the clear is enqueued on it. */
int lash_sketch_reset(lash_sketcher* s);
/* Run all further pushes on the caller's stream (a cudaStream_t passed as void*; NULL restores the
 * sketcher's own double-buffered streams).  Lets a host that already owns a stream (e.g. the one
 * its NCCL collectives run on) order sketching against its other work and time it with its events. */
int lash_sketch_set_stream(lash_sketcher* s, void* stream);
/* GPU time (ms, CUDA events on the sketcher's streams) spent in sketch kernels since open/reset,
 * and kernel launches issued; for bench.py's roofline accounting. */
int lash_sketch_stats(lash_sketcher* s, double* kernel_ms, uint64_t* launches);
int lash_sketch_close(lash_sketcher* s);

/* Fold sketches: dst[i] = merge(dst[i], src[i]) for n_sketches register arrays of the same (algo, p) --
 * register-wise max for HLL / HMH (HyperLogLog::union), pack(unpack(a) | unpack(b)) for ULL
 * (UltraLogLog::merge, utils.rs:260-262; NOT a byte max).  This is the one exchange step when a SINGLE
 * sample is sketched in shares (config "100 Gbp of reads -> one sketch": every GPU sketches its share of the
 * reads, the world x reg_bytes accumulators are all-gathered and folded).  _dev: device pointers, enqueued on
 * `stream` (NULL = the context's stream), not synchronised.  Host variant: dst_regs is updated in place. */
int lash_sketch_merge_dev(lash_ctx* ctx, int algo, int p, void* dst_dev, const void* src_dev, uint64_t n_sketches,
                          void* stream);
int lash_sketch_merge(lash_ctx* ctx, int algo, int p, void* dst_regs, const void* src_regs, uint64_t n_sketches);

/* ------------------------------------------------------------------------------------------------
 * Distance.  Replaces the par_iter bodies of utils.rs:150-180 / 248-285 / 342-370 and
 * compute_distance (main.rs:415-423): for every (reference i, query j) pair
 *     union -> estimator -> s = max(0,(a+b-U)/U)   [HMH: similarity()]  -> frac = 2s/(1+s)
 *     -> cast to T (f32 when fp32) -> model 1: min(1, -ln(frac)/k) | model 0: 1 - frac^(1/k)
 * out[i * n_qry + j] (row-major, f64 or f32).  triangular != 0 (the reference's same_files rule,
 * utils.rs:158-160,256-258,350-352, with idx = array position): requires the same set on both
 * sides; only j <= i is computed and `out` is the PACKED lower triangle, row i at i*(i+1)/2.
 * The name-equality => 0 rule (main.rs:452-453) stays with the host, which owns the names.
 * estimator is used for ULL only.  Returns LASH_OK, an error, or LASH_W_HLL_BIAS_REGIME.
 * ---------------------------------------------------------------------------------------------- */
int lash_dist(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
              uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, void* out);

/* Device-resident variant: register arrays and `out_dev` are device pointers on the context's GPU;
 * work is enqueued on `stream` (a cudaStream_t passed as void*, NULL = the context's stream) and
 * NOT synchronised.  Computes reference rows [row_begin, row_end) only (output tiling across GPUs:
 * each rank takes a row range); out_dev is indexed like `out` above (full-matrix indexing).
 * card_ref_dev / card_qry_dev: per-sketch cardinalities from lash_cardinality_dev (ignored for HMH).
 * flags_dev: optional uint32 counter incremented for each pair in the HLL bias regime.
 * Register arrays should be 16-byte aligned (anything cudaMalloc returns is): the HLL / HMH tile kernels stage with 16-byte loads
 * and fall back to the slower generic kernel for unaligned pointers.  ULL ML allocates a per-context scratch (60 B per output cell
 * at p <= 11) for its two-kernel form and falls back to the fused kernel above LASH_ML_SCRATCH_MAX_MB (default 16 GiB).
 * HLL and ULL ML first make one pass over both register arrays for every sketch's smallest / largest register (4 B per sketch of
 * context scratch; the tile kernels anchor their fixed-point windows on them).  HMH: when the row range and the query set both
 * hold sketches of <= 2^19 k-mers, the call counts them (an 8-byte device-to-host copy and ONE stream synchronisation -- the
 * only one in the device-resident API, so such a call cannot be captured in a CUDA graph) and precomputes hyperminhash's
 * expected-collision loop for all pairs of them in context scratch (336 KB per small sketch + 8 B per small pair, capped by
 * LASH_HMH_EC_MAX_MB, default 24 GiB; beyond the cap, or with LASH_HMH_EC=loop, the loop runs per pair as in the reference).
 * One context's scratch serves one call at a time: concurrent lash_dist* calls need one context each. */
int lash_dist_dev(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_dev,
                  uint64_t n_ref, const void* qry_dev, uint64_t n_qry, const double* card_ref_dev,
                  const double* card_qry_dev, int triangular, uint64_t row_begin, uint64_t row_end, void* out_dev,
                  uint32_t* flags_dev, void* stream);

/* Per-sketch cardinality (utils.rs:213-219 ULL, :314-316 HLL, hyperminhash cardinality()):
 * card_dev[i] for n sketches in device memory; enqueued on `stream`, not synchronised. */
int lash_cardinality_dev(lash_ctx* ctx, int algo, int p, int estimator, const void* regs_dev, uint64_t n,
                         double* card_dev, void* stream);
/* Host-buffer convenience: cardinalities of n sketches. */
int lash_cardinality(lash_ctx* ctx, int algo, int p, int estimator, const void* regs, uint64_t n, double* card_out);

/* Streaming all-vs-all for matrices larger than host memory (config "100k x 100k --dm"):
 * computes row blocks of `rows_per_block` references, copies each to a pinned buffer and calls
 * cb(user, row_begin, n_rows, block) from the calling thread while the next block computes.
 * block is dense [n_rows][n_qry] (triangular: cells j > i are unspecified). */
typedef int (*lash_dist_block_cb)(void* user, uint64_t row_begin, uint64_t n_rows, const void* block);
int lash_dist_stream(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
                     uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, uint64_t rows_per_block,
                     lash_dist_block_cb cb, void* user);

/* Same for reference rows [row_begin, row_end) only: the output tiling of the N x M matrix across GPUs / processes
 * (each takes a row range, e.g. lash_b200/shard.py::row_shard; every one uploads both register arrays). */
int lash_dist_stream_rows(lash_ctx* ctx, int algo, int p, int k, int estimator, int model, int fp32, const void* ref_regs,
                          uint64_t n_ref, const void* qry_regs, uint64_t n_qry, int triangular, uint64_t row_begin,
                          uint64_t row_end, uint64_t rows_per_block, lash_dist_block_cb cb, void* user);

/* Size-independent check of a distance run ("checksum of checksums"): when enabled, every lash_dist / lash_dist_stream*
 * call also sums, on the device, the bit patterns of the cells it defines (f64 as u64, f32 zero-extended; wrapping) and
 * counts them.  The sum is order-free: the totals of a matrix computed in row blocks, or in row ranges on several GPUs
 * (added up by the caller), equal those of one lash_dist call -- the multi-GPU equality test of the path (every pair is
 * computed by the same arithmetic whatever tile, block or rank it lands in). */
int lash_dist_set_checksum(lash_ctx* ctx, int enable);
int lash_dist_checksum(lash_ctx* ctx, uint64_t* sum_bits, uint64_t* n_cells);
/* Device-resident variant for outputs of lash_dist_dev (full-matrix indexing, packed lower triangle when triangular):
 * sums_dev[0] += sum of bit patterns, sums_dev[1] += cells over reference rows [row_begin, row_end); enqueued on `stream`. */
int lash_dist_checksum_dev(lash_ctx* ctx, int fp32, const void* out_dev, uint64_t n_qry, int triangular, uint64_t row_begin,
                           uint64_t row_end, uint64_t* sums_dev, void* stream);

/* Kernel time (ms) and kernel launches (cardinality, register minimum, distance tiles) of the last lash_dist /
 * lash_dist_stream call on this ctx; the launch count also accumulates over lash_cardinality_dev / lash_dist_dev calls
 * (those are not timed by the library: the caller owns the stream). */
int lash_dist_stats(lash_ctx* ctx, double* kernel_ms, uint64_t* launches);

#ifdef __cplusplus
}
#endif
#endif /* LASH_GPU_H */
