// Internal launcher interface between the C ABI (api.cu) and the kernels.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "hash.cuh"

namespace lash {

// One CTA's share of a span: k-mer start positions [begin, end) of the span whose first base sits
// at 32-bit word `word_off` of the staged buffer.  begin is a multiple of 64.
struct SketchTile {
    uint64_t word_off;      // span start, in 32-bit words (multiple of 4)
    uint64_t mask_word_off; // offset (32-bit words) of the span's invalid-start bitmask, or ~0ull
    uint64_t begin, end;    // k-mer start range inside the span
    uint32_t genome;
    uint32_t clip;          // 0xffffffff, or index into span_kept[]: the span holds only that many bases (device-side
                            // filter: the host planned the tiles on the raw byte count), so `end` is clipped to kept - k + 1
};

// A multi-record span, for the boundary-mask builder.
struct SpanRecs {
    uint64_t mask_word_off;  // into the mask buffer
    uint64_t rec_first;      // into rec_start[] (unused when uniform_len != 0)
    uint64_t n_bases;        // bases in the span
    uint32_t n_rec;
    uint32_t uniform_len;    // != 0: every record has this many bases (the last may be shorter)
};

// ---- device-side filter + 2-bit pack (text_kernels.cu) ----------------------------------------------------------
constexpr uint8_t kTextRecordSep = 0x01;            // == LASH_TEXT_RECORD_SEP
constexpr uint32_t kTextBlockBytes = 32u * 1024u;    // text one CTA compacts at a time
struct TextBlock {
    uint64_t byte_begin;  // offset of the block in the staged text (span start + multiple of kTextBlockBytes)
    uint32_t n_bytes;     // <= kTextBlockBytes
    uint32_t span;        // index into TextSpanDev[]
};
struct TextSpanDev {
    uint64_t out_word_off;   // packed output of the span, in 32-bit words (multiple of 4)
    uint64_t mask_word_off;  // invalid-start bitmask of the span (32-bit words), or ~0ull: single record
    uint32_t first_block, n_blocks;
};
// count -> scan -> compact: text spans -> packed spans (+ invalid-start bits at record separators), kept bases per span.
// block_cnt_dev: n_blocks + 1 words, block_prefix_dev: n_blocks, span_kept_dev: n_spans (all uint64); packed_out_dev and
// mask_dev must be zeroed by the caller.
cudaError_t launch_text_pack(const uint8_t* text_dev, const TextBlock* blocks_dev, uint32_t n_blocks, const TextSpanDev* spans_dev,
                             uint32_t n_spans, uint64_t* block_cnt_dev, uint64_t* block_prefix_dev, uint64_t* span_kept_dev,
                             uint32_t* packed_out_dev, uint32_t* mask_dev, int k, int n_sm, cudaStream_t st);

constexpr int kStartsPerThread = 64;  // one 16-byte load of packed bases per thread per iteration

struct SketchParams {
    int algo, p, k;
    HashConsts hc;
    uint32_t cell_words;  // 32-bit words per sketch in the register domain (global accumulator)
    uint32_t n_cells;     // registers per sketch
    uint32_t smem_bytes;  // private accumulator size (ULL 8 B, HLL/HMH 4 B per register)
    int n_sm;             // SMs of the device (persistent grid sizing)
    uint32_t threads;     // CTA size: grows with smem_bytes so the SM keeps >= 32 warps resident
    bool global_acc;      // accumulator too large for shared memory: update global memory directly
};

// max dynamic shared memory we are willing to use for a private accumulator
constexpr uint32_t kMaxSmemAccBytes = 128 * 1024;
// fills n_cells / smem_bytes / threads / global_acc from algo and p
void plan_sketch(SketchParams& sp);

cudaError_t launch_build_invalid_mask(const SpanRecs* spans_dev, uint32_t n_spans, uint64_t n_rec_total,
                                      const uint64_t* rec_start_dev, uint32_t* mask_dev, int k, int n_sm, cudaStream_t st);
cudaError_t launch_sketch(const SketchParams& sp, const uint32_t* packed_dev, const uint32_t* mask_dev,
                          const SketchTile* tiles_dev, uint32_t n_tiles, uint32_t* acc_dev, cudaStream_t st,
                          const uint64_t* span_kept_dev = nullptr);

// dst[i] = merge(dst[i], src[i]) over 32-bit words of register arrays
cudaError_t launch_merge(int algo, uint32_t* dst, const uint32_t* src, uint64_t n_words, int n_sm, cudaStream_t st);

struct DistParams {
    int algo, p, k, estimator, model, fp32, triangular;
    const void* ref;
    const void* qry;
    uint64_t n_ref, n_qry;
    const double* card_ref;
    const double* card_qry;
    uint64_t row_begin, row_end;
    void* out;
    // output addressing: packed_tri ? out[i*(i+1)/2 + j] : out[(i - out_row0) * n_qry + j]
    int packed_tri;
    uint64_t out_row0;
    uint32_t* flags;  // optional bias-regime counter
    // optional (ULL FGRA pair-table kernel): device word holding the smallest non-empty register of both sets
    const uint32_t* regmin = nullptr;
    // optional (pair-table kernels): device word the persistent CTAs draw tile indices from (zeroed by the launcher);
    // nullptr: static round-robin over the grid
    uint32_t* tile_counter = nullptr;
    int n_sm = 148;
    // optional (ULL ML pair-table kernel): scratch for the two-kernel form -- the tile kernel stores each pair's integer
    // statistics (S, exact-path flag, bit planes of b[]) and ml_finish_kernel runs the per-pair solver at full occupancy.
    // Word w of cell c lives at ml_scratch[w * ml_cells + c], c = output index - ml_o_base.  nullptr: fused epilogue.
    uint32_t* ml_scratch = nullptr;
    uint64_t ml_cells = 0;
    uint64_t ml_o_base = 0;
    // optional (HLL, ULL ML): per-sketch smallest | largest << 8 register byte (launch_hll_minmax), indexed by sketch;
    // nullptr: K4h runs instead of K4i / the ML tile kernel takes every S from its table
    const uint32_t* reg_mm_ref = nullptr;
    const uint32_t* reg_mm_qry = nullptr;
    // optional (HMH): expected-collision sums of pairs of SMALL sketches (cardinality <= 2^19), computed before the tile kernel:
    // hmh_slot_ref[i - hmh_row0] / hmh_slot_qry[j] = slot of the sketch's term vector (or -1), hmh_ec[slot_r * hmh_ec_ld + slot_q]
    // = the loop sum x of expectedCollision.  nullptr: every such pair runs the 41 x 1024 loop itself.
    const int32_t* hmh_slot_ref = nullptr;
    const int32_t* hmh_slot_qry = nullptr;
    const double* hmh_ec = nullptr;
    uint64_t hmh_row0 = 0;
    uint32_t hmh_ec_ld = 0;
};

// mm[i] = min | max << 8 over the register bytes of (HLL or ULL) sketch i (cell_bytes a multiple of 16, 16-byte aligned array)
cudaError_t launch_hll_minmax(const void* regs, uint64_t n, uint32_t cell_bytes, uint32_t* mm, cudaStream_t st);

// ---- HMH small-sketch path (dist_kernels.cu) ---------------------------------------------------------------------------------
// number of sketches among card[begin, end) whose expectedCollision runs the double loop (cardinality <= 2^19): *count_dev += n
cudaError_t launch_hmh_count_small(const double* card, uint64_t begin, uint64_t end, uint32_t* count_dev, cudaStream_t st);
// slot[i - begin] = rank of sketch i among the small ones of [begin, end) (capped at `cap`: beyond it -1), src[slot] = i,
// *count_dev = min(number of small sketches, cap)
cudaError_t launch_hmh_slots(const double* card, uint64_t begin, uint64_t end, uint32_t cap, int32_t* slot, uint32_t* src,
                             uint32_t* count_dev, cudaStream_t st);
// terms[slot][kHmhEcTermsPerSketch] of the first *count_dev slots: 41 x 1024 terms, then the largest |term| of each row
cudaError_t launch_hmh_ec_fill(const double* card, const uint32_t* src, const uint32_t* count_dev, uint32_t cap, double* terms,
                               cudaStream_t st);
// ec[r * ld + q] = sum over (i, j) in the reference's order of terms_r[r][ij] * terms_q[q][ij]; triangular: tiles whose largest
// reference sketch index is below their smallest query sketch index are skipped
cudaError_t launch_hmh_ec_gemm(const double* terms_r, const uint32_t* src_r, const uint32_t* count_r, uint32_t cap_r,
                               const double* terms_q, const uint32_t* src_q, const uint32_t* count_q, uint32_t cap_q, int triangular,
                               double* ec, uint32_t ld, cudaStream_t st);
constexpr uint32_t kHmhEcTermsPerSketch = 41u * 1024u + 64u;   // the terms, then the 41 row maxima (padded)
// 32-bit words of ML scratch per pair for sketches of precision p (S lo/hi, flag, bit planes)
uint32_t ml_scratch_words(int p);
// atomicMin of the smallest non-zero register byte into *out_dev (preset to 0xffffffff by the caller)
cudaError_t launch_regmin(const void* regs, uint64_t n_bytes, uint32_t* out_dev, int n_sm, cudaStream_t st);
cudaError_t launch_dist(const DistParams& dp, cudaStream_t st, uint32_t* n_launches);
// sums_dev[0] += wrapping sum of the output cells' bit patterns over rows [row_begin, row_end), sums_dev[1] += their count
cudaError_t launch_out_checksum(const DistParams& dp, unsigned long long* sums_dev, cudaStream_t st);
cudaError_t launch_cardinality(int algo, int p, int estimator, const void* regs, uint64_t n, double* card,
                               uint32_t* flags, cudaStream_t st);
// host-side one-time upload of estimator tables into constant memory (idempotent, per device)
cudaError_t ensure_tables();

}  // namespace lash
