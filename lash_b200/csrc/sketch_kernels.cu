// K1: k-mer sketching kernel (sm_100a, integer pipe).
//
// Replaces the body of the reference's per-file closure, src/utils.rs:457-503:
//   for each record: KmerSeqIterator -> min(kmer, reverse_complement) -> mask_bits -> add_kmer
// Design (not a port of the serial iterator):
//   * a CTA owns a chunk of ONE genome's k-mer start positions and a PRIVATE shared-memory
//     accumulator in the final register domain (u8 / u16), so registers never round-trip HBM
//     per k-mer; at the end the non-zero words are merged into the genome's global accumulator
//     with word CAS (max for HLL/HMH, packed-domain OR-merge for ULL -- ULL is not a max sketch).
//   * a thread owns 64 consecutive start positions = one coalesced 16-byte load (+8 bytes halo).
//     k-mers are NOT rolled serially: forward words are funnel-shift windows of the big-endian
//     base stream, reverse-complement words are funnel-shift windows of the per-word
//     reverse-complemented stream (brev + pair swap + not), so all 64 hashes are independent (ILP).
//   * register update = byte/halfword load + "would it change?" filter; the CAS loop runs only
//     for the (rare, after warm-up) k-mers that actually raise a register.
//   * record boundaries (k-mers never span records, utils.rs:457-464) come from an
//     "invalid start" bitmask built on device from rec_start[] by build_invalid_mask().
#include "kernels.h"
#include "registers.cuh"

namespace lash {

__device__ __forceinline__ uint32_t rc16(uint32_t f) {
    // reverse-complement of 16 bases held big-endian (first base in the top 2 bits)
    uint32_t y = __brev(f);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return ~y;
}

template <typename CellT, bool GLOBAL>
__device__ __forceinline__ uint32_t load_cell(const uint32_t* acc, uint32_t idx) {
    const CellT* a = reinterpret_cast<const CellT*>(acc);
    if (GLOBAL) {
        return (uint32_t)(*reinterpret_cast<const volatile CellT*>(a + idx));
    } else {
        return (uint32_t)a[idx];
    }
}

// CAS loop on the containing 32-bit word; only reached when the filter saw a change.
template <int ALGO>
__device__ __noinline__ void cell_cas(uint32_t* acc, uint32_t idx, uint32_t val) {
    using C = Cell<ALGO>;
    constexpr uint32_t per = 4 / C::kBytes;
    constexpr uint32_t cmask = C::kBytes == 1 ? 0xffu : 0xffffu;
    uint32_t* wp = acc + idx / per;
    uint32_t sh = (idx % per) * (8 * C::kBytes);
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
    for (;;) {
        uint32_t r = (old >> sh) & cmask;
        uint32_t nw = C::update(r, val);
        if (nw == r) break;
        uint32_t assumed = old;
        old = atomicCAS(wp, assumed, (assumed & ~(cmask << sh)) | (nw << sh));
        if (old == assumed) break;
    }
}

template <int ALGO, bool WIDE, bool GLOBAL>
__global__ void __launch_bounds__(kSketchThreads)
    sketch_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ inv_mask,
                  const SketchTile* __restrict__ tiles, uint32_t* __restrict__ acc_global, int p, int k, HashConsts hc,
                  uint32_t cell_words) {
    using C = Cell<ALGO>;
    extern __shared__ uint32_t sacc[];
    const SketchTile t = tiles[blockIdx.x];
    uint32_t* gacc = acc_global + (size_t)t.genome * cell_words;
    uint32_t* acc = GLOBAL ? gacc : sacc;
    if (!GLOBAL) {
        for (uint32_t i = threadIdx.x; i < cell_words; i += kSketchThreads) sacc[i] = 0u;
        __syncthreads();
    }
    const uint32_t* base = packed + t.word_off;
    const uint32_t* mbase = (t.mask_word_off != ~0ull) ? inv_mask + t.mask_word_off : nullptr;

    // uniform shift amounts / masks
    const uint32_t narrow_shr = WIDE ? 0u : (uint32_t)(32 - 2 * k);             // fwd >> (32-2k)
    const uint32_t narrow_mask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    const uint32_t wide_shr = WIDE ? (uint32_t)(64 - 2 * k) : 0u;               // in [0,30]
    const uint32_t wide_mask_hi = (k >= 32) ? 0xffffffffu : ((1u << ((2 * k - 32) & 31)) - 1u);

    for (uint64_t s0 = t.begin + (uint64_t)threadIdx.x * kStartsPerThread; s0 < t.end;
         s0 += (uint64_t)kStartsPerIter) {
        const uint32_t* wp = base + (s0 >> 4);
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(wp));
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(wp + 4));
        // validity of the 64 starts of this thread
        uint64_t remain = t.end - s0;
        uint64_t valid = remain >= 64 ? ~0ull : ((1ull << remain) - 1ull);
        if (mbase) {
            const uint2 mv = __ldg(reinterpret_cast<const uint2*>(mbase + (s0 >> 5)));
            valid &= ~(((uint64_t)mv.y << 32) | mv.x);
        }
        // big-endian base order inside each word: first base in the top bits
        uint32_t f0 = __byte_perm(q.x, 0, 0x0123), f1 = __byte_perm(q.y, 0, 0x0123);
        uint32_t f2 = __byte_perm(q.z, 0, 0x0123), f3 = __byte_perm(q.w, 0, 0x0123);
        uint32_t f4 = __byte_perm(h.x, 0, 0x0123), f5 = __byte_perm(h.y, 0, 0x0123);

#pragma unroll 1
        for (int w = 0; w < 4; ++w) {
            const uint32_t v16 = (uint32_t)(valid >> (16 * w)) & 0xffffu;
            if (v16) {
                const uint32_t A = f0, B = f1, Cw = f2;
                const uint32_t Ar = rc16(A), Br = rc16(B), Cr = WIDE ? rc16(Cw) : 0u;
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    uint32_t klo, khi;
                    if (!WIDE) {
                        uint32_t fw = __funnelshift_l(B, A, 2 * i) >> narrow_shr;
                        uint32_t rc = __funnelshift_r(Ar, Br, 2 * i) & narrow_mask;
                        klo = min(fw, rc);
                        khi = 0u;
                    } else {
                        uint32_t fhi = __funnelshift_l(B, A, 2 * i), flo = __funnelshift_l(Cw, B, 2 * i);
                        flo = __funnelshift_r(flo, fhi, wide_shr);
                        fhi >>= wide_shr;
                        uint32_t rlo = __funnelshift_r(Ar, Br, 2 * i);
                        uint32_t rhi = __funnelshift_r(Br, Cr, 2 * i) & wide_mask_hi;
                        uint64_t f64 = mk64(flo, fhi), r64 = mk64(rlo, rhi);
                        uint64_t c64 = f64 < r64 ? f64 : r64;
                        klo = (uint32_t)c64;
                        khi = (uint32_t)(c64 >> 32);
                    }
                    uint32_t idx, val;
                    C::from_kmer(klo, khi, hc, p, idx, val);
                    if (v16 & (1u << i)) {
                        uint32_t r = load_cell<typename C::T, GLOBAL>(acc, idx);
                        if (C::update(r, val) != r) cell_cas<ALGO>(acc, idx, val);
                    }
                }
            }
            f0 = f1; f1 = f2; f2 = f3; f3 = f4; f4 = f5;
        }
    }

    if (!GLOBAL) {
        __syncthreads();
        for (uint32_t i = threadIdx.x; i < cell_words; i += kSketchThreads) {
            const uint32_t v = sacc[i];
            if (v == 0u) continue;
            uint32_t* gp = gacc + i;
            uint32_t old = *reinterpret_cast<volatile uint32_t*>(gp);
            for (;;) {
                uint32_t nw = C::merge_word(old, v);
                if (nw == old) break;
                uint32_t assumed = old;
                old = atomicCAS(gp, assumed, nw);
                if (old == assumed) break;
            }
        }
    }
}

// One CTA per multi-record span: mark every k-mer start that would cross the END of a record
// (or belongs to a record shorter than k).  Bit b of the span's mask <=> start position b.
__global__ void build_invalid_mask_kernel(const SpanRecs* __restrict__ spans, const uint64_t* __restrict__ rec_start,
                                          uint32_t* __restrict__ mask, int k) {
    const SpanRecs s = spans[blockIdx.x];
    const uint64_t* rs = rec_start + s.rec_first;
    uint32_t* m = mask + s.mask_word_off;
    for (uint32_t r = threadIdx.x; r < s.n_rec; r += blockDim.x) {
        const uint64_t b = rs[r], e = rs[r + 1];
        // starts in [max(b, e-k+1), e) are invalid
        uint64_t lo = (e >= (uint64_t)(k - 1)) ? e - (uint64_t)(k - 1) : 0;
        if (lo < b) lo = b;
        for (uint64_t x = lo; x < e;) {
            const uint32_t bit = (uint32_t)(x & 31);
            const uint64_t n = min((uint64_t)(32 - bit), e - x);
            const uint32_t bits = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << bit;
            atomicOr(m + (x >> 5), bits);
            x += n;
        }
    }
}

cudaError_t launch_build_invalid_mask(const SpanRecs* spans_dev, uint32_t n_spans, const uint64_t* rec_start_dev,
                                      uint32_t* mask_dev, int k, cudaStream_t st) {
    if (n_spans == 0) return cudaSuccess;
    build_invalid_mask_kernel<<<n_spans, 256, 0, st>>>(spans_dev, rec_start_dev, mask_dev, k);
    return cudaGetLastError();
}

template <int ALGO, bool WIDE, bool GLOBAL>
static cudaError_t launch_one(const SketchParams& sp, const uint32_t* packed, const uint32_t* mask,
                              const SketchTile* tiles, uint32_t n_tiles, uint32_t* acc, cudaStream_t st) {
    auto kern = sketch_kernel<ALGO, WIDE, GLOBAL>;
    size_t smem = GLOBAL ? 0 : (size_t)sp.cell_words * 4;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    kern<<<n_tiles, kSketchThreads, smem, st>>>(packed, mask, tiles, acc, sp.p, sp.k, sp.hc, sp.cell_words);
    return cudaGetLastError();
}

template <int ALGO>
static cudaError_t launch_algo(const SketchParams& sp, const uint32_t* packed, const uint32_t* mask,
                               const SketchTile* tiles, uint32_t n_tiles, uint32_t* acc, cudaStream_t st) {
    const bool wide = sp.k > 16;
    if (sp.global_acc) {
        return wide ? launch_one<ALGO, true, true>(sp, packed, mask, tiles, n_tiles, acc, st)
                    : launch_one<ALGO, false, true>(sp, packed, mask, tiles, n_tiles, acc, st);
    }
    return wide ? launch_one<ALGO, true, false>(sp, packed, mask, tiles, n_tiles, acc, st)
                : launch_one<ALGO, false, false>(sp, packed, mask, tiles, n_tiles, acc, st);
}

cudaError_t launch_sketch(const SketchParams& sp, const uint32_t* packed_dev, const uint32_t* mask_dev,
                          const SketchTile* tiles_dev, uint32_t n_tiles, uint32_t* acc_dev, cudaStream_t st) {
    if (n_tiles == 0) return cudaSuccess;
    switch (sp.algo) {
        case HMH: return launch_algo<HMH>(sp, packed_dev, mask_dev, tiles_dev, n_tiles, acc_dev, st);
        case HLL: return launch_algo<HLL>(sp, packed_dev, mask_dev, tiles_dev, n_tiles, acc_dev, st);
        case ULL: return launch_algo<ULL>(sp, packed_dev, mask_dev, tiles_dev, n_tiles, acc_dev, st);
    }
    return cudaErrorInvalidValue;
}

}  // namespace lash
