// K0: filter_out_n + 2-bit packing on the device (sm_100a, HBM-bound byte work).
//
// Replaces, for hosts that ship raw sequence text, the reference's per-record front end
//   src/utils.rs:33-41   filter_out_n: keep bytes that are exactly one of "ACGT", delete everything else, join the flanks
//   src/utils.rs:464     KSeq::new(&seq, 2): A0 C1 G2 T3, four bases per byte, first base in the high bits
// and the "k-mers never span records" rule (utils.rs:457-464), which the sketch kernel gets as an invalid-start bitmask.
//
// Input: a span of raw sequence bytes of ONE genome (line breaks, N, lowercase, IUPAC codes ... anything), records
// separated IN BAND by the byte LASH_TEXT_RECORD_SEP.  Output: exactly the packed span + invalid-start mask that
// lash_sketch_push takes from a host packer, so K1 (sketch_kernel) runs unchanged on it.
//
// Three small kernels per push (all streaming; the text is read twice = 2 B/base of HBM traffic against the 1 B/base
// that crossed PCIe at 1/100 of the bandwidth to get here):
//   text_count_kernel     kept bases per 32 KiB block of text
//   text_scan_kernel      exclusive prefix of the block counts inside each span (one CTA), kept bases per span
//   text_compact_kernel   re-reads the block: classify, CTA-wide prefix, append the 2-bit codes to a shared-memory bit
//                         stream aligned with the output words, flush (first / last word with atomicOr: neighbouring
//                         blocks share them), and turn every separator into invalid-start bits [e-k+1, e)
#include <algorithm>

#include "kernels.h"

namespace lash {

constexpr int kTextThreads = 256;
constexpr uint32_t kTextIterBytes = kTextThreads * 16;   // one 16-byte load per thread per iteration
static_assert(kTextBlockBytes % kTextIterBytes == 0, "a text block is a whole number of CTA iterations");

// byte -> (is one of "ACGT", 2-bit code).  (c >> 1) & 3 separates the four letters (A 0, C 1, T 2, G 3); the byte is a
// base iff it equals the letter its own index selects; code = idx ^ (idx >> 1) gives A0 C1 G2 T3.
__device__ __forceinline__ bool base_code(uint32_t c, uint32_t& code) {
    const uint32_t idx = (c >> 1) & 3u;
    code = idx ^ (idx >> 1);
    return c == ((0x47544341u >> (8u * idx)) & 0xffu);
}

// Four bytes at a time (SIMD in a 32-bit word; byte j of the little-endian word is the j-th byte of the text):
//   idx    = (w >> 1) & 0x03030303                                   per byte: which letter it could be
//   expect = PRMT("ACTG", nibbles(idx))                              the letter each byte would have to equal
//   valid  = bit 7 of every byte of  ~(((w ^ expect) & 0x7f..) + 0x7f.. | (w ^ expect))     (zero-byte test, no carries between bytes)
//   vm     = (valid * 0x00204081) >> 28                              the four bit-7s gathered into a 4-bit mask
// -- 12 instructions per word instead of ~18 per byte; the first version of these kernels classified byte by byte and was
// ALU-bound at 37 instructions per text byte over its two passes (0.55 TB/s of text).
__device__ __forceinline__ uint32_t valid_bits4(uint32_t w, uint32_t& idx) {
    idx = (w >> 1) & 0x03030303u;
    const uint32_t t = idx | (idx >> 4);
    const uint32_t expect = __byte_perm(0x47544341u, 0u, __byte_perm(t, 0u, 0x4420));
    const uint32_t x = w ^ expect;
    return ~(((x & 0x7f7f7f7fu) + 0x7f7f7f7fu) | x) & 0x80808080u;
}
__device__ __forceinline__ uint32_t gather_bit7s(uint32_t v) { return (v * 0x00204081u) >> 28; }   // bit 7 of byte j -> bit j
// bytes beyond the live part of a thread's 16 (only the last thread of a span's last block): make them a non-base, non-separator
__device__ __forceinline__ uint32_t live_word(uint32_t w, uint32_t n_live, int wi) {
    const int lim = (int)n_live - 4 * wi;
    return lim >= 4 ? w : lim <= 0 ? 0u : (w & ((1u << (8 * lim)) - 1u));
}

// kept bases among the 16 bytes a thread owns (count pass)
__device__ __forceinline__ uint32_t count16(const uint4& q, uint32_t n_live) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        uint32_t idx;
        cnt += __popc(valid_bits4(n_live >= 16u ? w[i] : live_word(w[i], n_live, i), idx));
    }
    return cnt;
}

// s_lut[vm] (shared, built by compact_lut_init): low 16 bits = PRMT selector that moves the valid bytes of a word to the
// front in order (the rest become zero bytes), bits 16.. = 2 * number of valid bytes
__device__ __forceinline__ void compact_lut_init(uint32_t* s_lut) {
    if (threadIdx.x < 16) {
        uint32_t sel = 0, n = 0;
        for (uint32_t j = 0; j < 4; ++j)
            if ((threadIdx.x >> j) & 1u) sel |= j << (4 * n++);
        for (uint32_t k = n; k < 4; ++k) sel |= 4u << (4 * k);   // selector 4 = byte 0 of the second PRMT operand (zero)
        s_lut[threadIdx.x] = sel | ((2u * n) << 16);
    }
}

// the 16 bytes a thread owns (compact pass): number of bases kept, their codes MSB-first (first kept base in the top two
// bits), and -- only when WANT_SEPS -- which of the 16 bytes are record separators
template <bool WANT_SEPS>
__device__ __forceinline__ void classify16(const uint4& q, uint32_t n_live, const uint32_t* s_lut, uint32_t& cnt, uint32_t& bits, uint32_t& seps) {
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
    uint32_t sh_total = 0;
    bits = 0;
    seps = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t wi = n_live >= 16u ? w[i] : live_word(w[i], n_live, i);
        uint32_t idx;
        const uint32_t vm = gather_bit7s(valid_bits4(wi, idx));
        const uint32_t codes = (idx ^ (idx >> 1)) & 0x03030303u;                     // A0 C1 G2 T3 per byte
        const uint32_t lut = s_lut[vm];
        const uint32_t front = __byte_perm(codes, 0u, lut);                          // valid codes first, in order
        // c0 << 30 | c1 << 28 | c2 << 26 | c3 << 24 in the top byte of the product (the lower bits are cross terms); the
        // funnel shift takes exactly the 2 * n top bits that hold codes
        const uint32_t sh = lut >> 16;
        bits = __funnelshift_l(front * 0x40100401u, bits, sh);
        sh_total += sh;
        if (WANT_SEPS) {
            const uint32_t y = wi ^ (0x01010101u * (uint32_t)kTextRecordSep);
            const uint32_t z = ~(((y & 0x7f7f7f7fu) + 0x7f7f7f7fu) | y) & 0x80808080u;
            seps |= gather_bit7s(z) << (4 * i);
        }
    }
    cnt = sh_total >> 1;
    bits = cnt ? (bits << (32u - sh_total)) : 0u;
}

__global__ void __launch_bounds__(kTextThreads) text_count_kernel(const uint8_t* __restrict__ text, const TextBlock* __restrict__ blocks,
                                                                  uint32_t n_blocks, uint64_t* __restrict__ block_cnt) {
    __shared__ uint32_t s_sum[kTextThreads / 32];
    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const TextBlock tb = blocks[b];
        uint32_t mine = 0;
        for (uint32_t off = threadIdx.x * 16u; off < tb.n_bytes; off += kTextIterBytes) {
            const uint4 q = __ldg(reinterpret_cast<const uint4*>(text + tb.byte_begin + off));
            mine += count16(q, min(16u, tb.n_bytes - off));
        }
        mine = __reduce_add_sync(0xffffffffu, mine);
        if ((threadIdx.x & 31u) == 0) s_sum[threadIdx.x >> 5] = mine;
        __syncthreads();
        if (threadIdx.x == 0) {
            uint32_t t = 0;
#pragma unroll
            for (int i = 0; i < kTextThreads / 32; ++i) t += s_sum[i];
            block_cnt[b] = t;
        }
        __syncthreads();
    }
}

// block_cnt[0..n) -> G (in place, n+1 entries: exclusive prefix over ALL blocks); block_prefix[b] = kept bases of the
// SAME span before block b = G[b] - G[first block of the span]; span_kept[s] = kept bases of span s.  One CTA.
__global__ void __launch_bounds__(1024) text_scan_kernel(uint64_t* __restrict__ block_cnt, uint64_t* __restrict__ block_prefix, uint32_t n_blocks,
                                                         const TextBlock* __restrict__ blocks, const TextSpanDev* __restrict__ spans,
                                                         uint32_t n_spans, uint64_t* __restrict__ span_kept) {
    __shared__ uint64_t s_warp[32];
    __shared__ uint64_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t base = 0; base < n_blocks; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint64_t v = i < n_blocks ? block_cnt[i] : 0ull;
        uint64_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31u) >= (uint32_t)d) x += y;
        }
        if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint64_t w = s_warp[threadIdx.x];
            uint64_t xw = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint64_t y = __shfl_up_sync(0xffffffffu, xw, d);
                if (threadIdx.x >= (uint32_t)d) xw += y;
            }
            s_warp[threadIdx.x] = xw - w;  // exclusive over warps
        }
        __syncthreads();
        const uint64_t excl = s_carry + s_warp[threadIdx.x >> 5] + (x - v);
        if (i < n_blocks) block_cnt[i] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) block_cnt[n_blocks] = s_carry;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n_blocks; i += 1024) block_prefix[i] = block_cnt[i] - block_cnt[spans[blocks[i].span].first_block];
    for (uint32_t s = threadIdx.x; s < n_spans; s += 1024) {
        const TextSpanDev sp = spans[s];
        span_kept[s] = block_cnt[sp.first_block + sp.n_blocks] - block_cnt[sp.first_block];
    }
}

// big-endian base order inside a 32-bit word <-> the little-endian bytes of the lash_gpu.h format
__device__ __forceinline__ uint32_t bswap32(uint32_t v) { return __byte_perm(v, 0, 0x0123); }

__global__ void __launch_bounds__(kTextThreads) text_compact_kernel(const uint8_t* __restrict__ text, const TextBlock* __restrict__ blocks,
                                                                    uint32_t n_blocks, const uint64_t* __restrict__ block_prefix,
                                                                    const TextSpanDev* __restrict__ spans, const uint64_t* __restrict__ span_kept,
                                                                    uint32_t* __restrict__ packed_out, uint32_t* __restrict__ inv_mask, int k) {
    constexpr uint32_t kStageWords = kTextIterBytes / 16 + 2;
    __shared__ uint32_t s_stage[kStageWords];
    __shared__ uint32_t s_warp[kTextThreads / 32];
    __shared__ uint32_t s_lut[16];
    const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
    for (uint32_t i = threadIdx.x; i < kStageWords; i += kTextThreads) s_stage[i] = 0u;
    compact_lut_init(s_lut);
    __syncthreads();
    // starts [e-k+1, e) cannot begin a k-mer when a record ends at kept position e
    auto mark_boundary = [&](uint32_t* mask, uint64_t e) {
        uint64_t lo = e >= (uint64_t)(k - 1) ? e - (uint64_t)(k - 1) : 0;
        for (uint64_t x = lo; x < e;) {
            const uint32_t bit = (uint32_t)(x & 31);
            const uint64_t n = min((uint64_t)(32 - bit), e - x);
            const uint32_t m = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << bit;
            atomicOr(mask + (x >> 5), m);
            x += n;
        }
    };
    for (uint32_t b = blockIdx.x; b < n_blocks; b += gridDim.x) {
        const TextBlock tb = blocks[b];
        const TextSpanDev sp = spans[tb.span];
        uint32_t* out = packed_out + sp.out_word_off;
        uint32_t* mask = sp.mask_word_off != ~0ull ? inv_mask + sp.mask_word_off : nullptr;
        uint64_t P = block_prefix[b];  // kept bases of this span before the block
        for (uint32_t off0 = 0; off0 < tb.n_bytes; off0 += kTextIterBytes) {
            const uint32_t off = off0 + threadIdx.x * 16u;
            uint32_t cnt = 0, bits = 0, seps = 0;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            if (off < tb.n_bytes) {
                q = __ldg(reinterpret_cast<const uint4*>(text + tb.byte_begin + off));
                if (mask) classify16<true>(q, min(16u, tb.n_bytes - off), s_lut, cnt, bits, seps);    // block-uniform
                else classify16<false>(q, min(16u, tb.n_bytes - off), s_lut, cnt, bits, seps);
            }
            // CTA-wide exclusive prefix of cnt
            uint32_t x = cnt;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
                if (lane >= (uint32_t)d) x += y;
            }
            if (lane == 31u) s_warp[warp] = x;
            __syncthreads();
            uint32_t wbase = 0, total = 0;
#pragma unroll
            for (int i = 0; i < kTextThreads / 32; ++i) {
                const uint32_t v = s_warp[i];
                wbase += (uint32_t)i < warp ? v : 0u;
                total += v;
            }
            const uint32_t excl = wbase + x - cnt;
            // append this thread's codes to the staged bit stream, aligned with the output words
            const uint32_t lead = (uint32_t)(P & 15u);
            if (cnt) {
                const uint32_t o = 2u * (lead + excl);
                const uint32_t w = o >> 5, sh = o & 31u;
                atomicOr(&s_stage[w], bits >> sh);
                if (sh && ((bits << (32u - sh)) != 0u)) atomicOr(&s_stage[w + 1], bits << (32u - sh));
            }
            // record separators (rare: one per record): invalid-start bits relative to the span
            if (seps && mask) {
                const uint32_t wq[4] = {q.x, q.y, q.z, q.w};
                uint32_t before = 0;
                for (int i = 0; i < 16; ++i) {
                    const uint32_t c = (wq[i >> 2] >> (8 * (i & 3))) & 0xffu;
                    if ((seps >> i) & 1u) mark_boundary(mask, P + excl + before);
                    uint32_t code;
                    before += base_code(c, code) ? 1u : 0u;
                }
            }
            __syncthreads();
            // flush: the words this iteration touched; the first and the last are shared with the neighbours
            const uint32_t n_words = (lead + total + 15u) >> 4;
            const uint64_t w0 = P >> 4;
            for (uint32_t w = threadIdx.x; w < n_words; w += kTextThreads) {
                const uint32_t v = s_stage[w];
                s_stage[w] = 0u;
                if (w == 0 || w + 1 == n_words) {
                    if (v) atomicOr(out + w0 + w, bswap32(v));
                } else {
                    out[w0 + w] = bswap32(v);
                }
            }
            P += total;
            __syncthreads();
        }
        // the end of the span ends its last record
        if (mask && threadIdx.x == 0 && b + 1 == sp.first_block + sp.n_blocks) mark_boundary(mask, span_kept[tb.span]);
    }
}

cudaError_t launch_text_pack(const uint8_t* text_dev, const TextBlock* blocks_dev, uint32_t n_blocks, const TextSpanDev* spans_dev,
                             uint32_t n_spans, uint64_t* block_cnt_dev, uint64_t* block_prefix_dev, uint64_t* span_kept_dev,
                             uint32_t* packed_out_dev, uint32_t* mask_dev, int k, int n_sm, cudaStream_t st) {
    if (n_blocks == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<uint64_t>(n_blocks, (uint64_t)n_sm * 8);
    text_count_kernel<<<grid, kTextThreads, 0, st>>>(text_dev, blocks_dev, n_blocks, block_cnt_dev);
    text_scan_kernel<<<1, 1024, 0, st>>>(block_cnt_dev, block_prefix_dev, n_blocks, blocks_dev, spans_dev, n_spans, span_kept_dev);
    text_compact_kernel<<<grid, kTextThreads, 0, st>>>(text_dev, blocks_dev, n_blocks, block_prefix_dev, spans_dev, span_kept_dev,
                                                       packed_out_dev, mask_dev, k);
    return cudaGetLastError();
}

}  // namespace lash
