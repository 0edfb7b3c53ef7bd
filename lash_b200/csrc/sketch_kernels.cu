// K1: k-mer sketching kernel (sm_100a, integer pipe).
//
// Replaces the body of the reference's per-file closure, src/utils.rs:457-503:
//   for each record: KmerSeqIterator -> min(kmer, reverse_complement) -> mask_bits -> add_kmer
// Design (not a port of the serial iterator):
//   * a CTA owns a chunk of ONE genome's k-mer start positions and a PRIVATE shared-memory
//     accumulator, so registers never round-trip HBM per k-mer; at the end the cells are converted
//     to the register domain (u8 / u16) and the non-zero words merged into the genome's global
//     accumulator with word CAS (max for HLL/HMH, packed-domain OR-merge for ULL -- ULL is not a
//     max sketch).
//   * a thread owns 64 consecutive start positions = one coalesced 16-byte load (+8 bytes halo).
//     k-mers are NOT rolled serially: forward words are funnel-shift windows of the big-endian
//     base stream, reverse-complement words are funnel-shift windows of the per-word
//     reverse-complemented stream (brev + pair swap + not), so all 64 hashes are independent (ILP).
//   * register update = one shared-memory load as a "would it change?" filter + one native
//     32-bit shared atomic (OR of a seen-value bit for ULL, max for HLL/HMH) only when it would.
//   * record boundaries (k-mers never span records, utils.rs:457-464) come from an
//     "invalid start" bitmask built on device from rec_start[] by build_invalid_mask().
#include <algorithm>
#include <type_traits>

#include "kernels.h"
#include "registers.cuh"
#include "kmer_windows.cuh"

namespace lash {

// ------------------------------------------------------------------------------------------------
// Private (shared-memory) accumulators.  They are NOT kept in the byte/halfword register domain:
// a sub-word read-modify-write needs a CAS loop, and a CAS loop inside the k-mer loop makes lanes
// diverge for good.  Instead every cell is laid out so that ONE native 32-bit shared-memory atomic
// is the whole update, and a plain shared load filters out the (vast majority of) k-mers that
// would not change it:
//   ULL  cell = two 32-bit words of "seen nlz" bits -> atomicOr.  Word 0 bit j <=> nlz = 31-j was
//        seen, word 1 bit j <=> nlz = 63-j (the bit index is the raw FLO result, no subtraction).
//        At flush: nlz-mask M = brev(w0) | brev(w1) << 32, unpacked hash prefix = M << (p-1),
//        register = ultraloglog pack(prefix)  (exact: sequential add()s == pack(OR of 1 << u)).
//   HLL  cell = u32 rho                -> atomicMax
//   HMH  cell = u32 (lz << 10 | sig)   -> atomicMax
// Each algorithm has a FAST preparation that is exact whenever the top 32 bits it looks at are
// non-zero (all but ~2^-32 of the hashes) and otherwise requests nothing wrong (ULL: no bit;
// HLL/HMH: a lower bound of the true value), plus an EXACT update used for those rare hashes, for
// blocks with invalid starts, and for the global-memory fallback.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lds_u32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void red_or(uint32_t saddr, uint32_t bits) {
    asm volatile("red.shared.or.b32 [%0], %1;" ::"r"(saddr), "r"(bits) : "memory");
}
__device__ __forceinline__ void red_max(uint32_t saddr, uint32_t v) {
    asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

template <int ALGO>
struct SmemAcc;

// LASH_ULL_PLANES=1: ULL cells as TWO PLANES (all word-0s, then all word-1s) instead of interleaved pairs.  The fast path only
// ever loads word 0, and with 8-byte cells those loads touch the even banks only (ncu r01: 4.7 wavefronts per warp-level
// LDS).  Measured in round 2 (tools/bench_configs.py, B200): the planes are 1-1.7 % SLOWER (C2 836.9 vs 845.9 Gbp/s, k = 12
// 769.5 vs 782.6, p = 14 542 vs 547) -- the shared-memory pipe is not what binds this kernel -- so the default stays the
// interleaved cell.
#ifndef LASH_ULL_PLANES
#define LASH_ULL_PLANES 0
#endif
template <>
struct SmemAcc<ULL> {
    static constexpr uint32_t kWordsPerCell = 2;
    static constexpr uint32_t kCellStride = LASH_ULL_PLANES ? 4u : 8u;   // bytes between the word-0s of neighbouring cells
    // fast: (address of word 0, bit to set [0 if the hash is a rare one], rare indicator word)
    // g = hash BEFORE its last step h = g ^ (g >> 28).  Because p <= 26 that xorshift cannot reach the index bits
    // (idx = g.hi >> (32-p)).  The fast path looks only at the HIGH word of g: the 32-p hash bits after the index are
    //     t = (g.hi ^ (g.hi >> 28)) & (2^(32-p) - 1)            (g.lo only reaches bits further down),
    // so the second 64-bit multiply of the hash needs no low word (IMAD.HI + 2 IMAD) and nlz = 31 - p - bfind(t):
    // the bit to set is (1 << p) << bfind(t), in the cell layout described above.  The xorshift term only reaches the
    // low 4 bits of t, so for small p the fast path drops it AND those 4 bits: t' = g.hi & (2^(32-p) - 16) is one LOP3
    // and has the same highest set bit as t whenever it is non-zero.  A hash with t' == 0 (2^-(28-p) of them; 2^-18 at
    // p = 10) is "rare": it sets no bit here and sends its 16-k-mer group through the exact path (~2 instructions per
    // 512 k-mers at p = 10).
    // DROP4 is set for the small-accumulator instantiation (p <= 11); at p = 14 one group in 30 would hold a "rare"
    // hash (measured: -3 %), so larger precisions keep the exact t (one more SHF).
    template <bool NARROW, bool DROP4>
    __device__ static __forceinline__ void prep(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t sbase,
                                                uint32_t& saddr, uint32_t& v, uint32_t& rare_word) {
        const uint32_t hi = NARROW ? xxh3_64_narrow_pre_hi(klo, hc) : xxh3_64_wide_pre_hi(klo, khi, hc);
        const uint32_t t = DROP4 ? hi & ((0xffffffffu >> p) & ~15u) : (hi ^ (hi >> 28)) & (0xffffffffu >> p);
#ifdef LASH_SADDR_IMAD  // tuning switch (tools/variant_sweep): no difference measured, ptxas picks LEA or IMAD itself
        saddr = mad32_opaque(__umulhi(hi, 1u << p), kCellStride, sbase);
#else
        saddr = sbase + __umulhi(hi, 1u << p) * kCellStride;       // (hi >> (32-p)) * stride on the FMA pipe
#endif
        v = shl_clamp(1u << p, bfind32(t));              // t == 0 -> bfind = 0xffffffff -> v = 0
        rare_word = t;
    }
    __device__ static __forceinline__ uint32_t need(uint32_t cur, uint32_t v) { return ~cur & v; }
    __device__ static __forceinline__ void apply(uint32_t saddr, uint32_t needv) { red_or(saddr, needv); }
    __device__ static __forceinline__ void exact(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t sbase) {
        const uint64_t h = xxh3_64_le64(klo, khi, hc);
        const uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
        const uint32_t yh = __funnelshift_l(lo, hi, p), yl = (lo << p) | ((1u << p) - 1u);
        const uint32_t w1_off = LASH_ULL_PLANES ? (4u << p) : 4u;   // word 1: the second plane, or the next word
        const uint32_t saddr = sbase + (hi >> (32 - p)) * kCellStride + (yh ? 0u : w1_off);
        const uint32_t bit = 1u << bfind32(yh ? yh : yl);
        if (~lds_u32(saddr) & bit) red_or(saddr, bit);
    }
    // word index of word z of a cell in the accumulator array
    __device__ static __forceinline__ uint32_t word_of(uint32_t cell, uint32_t z, uint32_t n_cells) {
        return LASH_ULL_PLANES ? z * n_cells + cell : 2u * cell + z;
    }
    __device__ static __forceinline__ uint32_t to_reg(const uint32_t* acc, uint32_t cell, int p, uint32_t n_cells) {
        return ull_cell_to_reg(acc[word_of(cell, 0, n_cells)], acc[word_of(cell, 1, n_cells)], p);
    }
};
template <>
struct SmemAcc<HLL> {
    static constexpr uint32_t kWordsPerCell = 1;
    template <bool NARROW, bool DROP4>
    __device__ static __forceinline__ void prep(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t sbase,
                                                uint32_t& saddr, uint32_t& v, uint32_t& rare_word) {
        const uint64_t g = NARROW ? xxh3_64_narrow_pre(klo, hc) : xxh3_64_wide_pre(klo, khi, hc);
        const uint32_t glo = (uint32_t)g, ghi = (uint32_t)(g >> 32);
        // h = g ^ (g >> 28): the index needs the low p bits of h.lo.  rho = clz(h.hi) + 1 = clz(g.hi) + 1: the xorshift
        // folds the top 4 bits of g.hi into its low 4 bits, which moves the highest set bit of neither a g.hi >= 2^28
        // nor (g.hi >> 28 == 0) a smaller one -- so h.hi is never materialised.
        const uint32_t idx = (glo ^ __funnelshift_r(glo, ghi, 28)) & ((1u << p) - 1u);
        // wide k: the cell address as IMAD by a run-time 4 (FMA pipe) instead of LEA (ALU pipe, 88 % busy there)
        saddr = (!NARROW && LASH_WIDE_FMA) ? mad32_opaque(idx, hc.two29 >> 27, sbase) : sbase + idx * 4u;
        v = 32u - bfind32(ghi);  // rho when g.hi != 0; g.hi == 0 (<=> h.hi == 0) -> 33 <= true rho (p <= 18)
        rare_word = ghi;
    }
    __device__ static __forceinline__ uint32_t need(uint32_t cur, uint32_t v) { return cur < v ? v : 0u; }
    __device__ static __forceinline__ void apply(uint32_t saddr, uint32_t needv) { red_max(saddr, needv); }
    __device__ static __forceinline__ void exact(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t sbase) {
        uint32_t idx, rho;
        Cell<HLL>::from_kmer(klo, khi, hc, p, idx, rho);
        const uint32_t saddr = sbase + idx * 4u;
        if (lds_u32(saddr) < rho) red_max(saddr, rho);
    }
    __device__ static __forceinline__ uint32_t word_of(uint32_t cell, uint32_t, uint32_t) { return cell; }
    __device__ static __forceinline__ uint32_t to_reg(const uint32_t* acc, uint32_t cell, int, uint32_t) { return acc[cell]; }
};
template <>
struct SmemAcc<HMH> {
    static constexpr uint32_t kWordsPerCell = 1;
    template <bool NARROW, bool DROP4>
    __device__ static __forceinline__ void prep(uint32_t klo, uint32_t /*khi*/, const HashConsts& hc, int /*p*/, uint32_t sbase,
                                                uint32_t& saddr, uint32_t& v, uint32_t& rare_word) {
        uint64_t hlo, hhi;
        xxh3_128_le32(klo, hc, hlo, hhi);
#if LASH_HMH_X_IS_HIGH64
        const uint64_t x = hhi, y = hlo;
#else
        const uint64_t x = hlo, y = hhi;
#endif
        const uint32_t xlo = (uint32_t)x, xhi = (uint32_t)(x >> 32);
        saddr = sbase + (xhi >> 18) * 4u;                     // x >> 50
        const uint32_t th = __funnelshift_l(xlo, xhi, 14);     // top 32 bits of (x << 14) | 0x3fff
        v = ((32u - bfind32(th)) << 10) | ((uint32_t)y & 1023u);  // th == 0 -> lz = 33 <= true lz
        rare_word = th;
    }
    __device__ static __forceinline__ uint32_t need(uint32_t cur, uint32_t v) { return cur < v ? v : 0u; }
    __device__ static __forceinline__ void apply(uint32_t saddr, uint32_t needv) { red_max(saddr, needv); }
    __device__ static __forceinline__ void exact(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t sbase) {
        uint32_t idx, val;
        Cell<HMH>::from_kmer(klo, khi, hc, p, idx, val);
        const uint32_t saddr = sbase + idx * 4u;
        if (lds_u32(saddr) < val) red_max(saddr, val);
    }
    __device__ static __forceinline__ uint32_t word_of(uint32_t cell, uint32_t, uint32_t) { return cell; }
    __device__ static __forceinline__ uint32_t to_reg(const uint32_t* acc, uint32_t cell, int, uint32_t) { return acc[cell]; }
};

// Global-accumulator fallback (2^p too large for shared memory): byte / halfword cells of the
// genome's accumulator in HBM/L2, volatile-load filter + CAS on the containing word.
template <int ALGO>
__device__ __forceinline__ void global_update(uint32_t* gacc, uint32_t idx, uint32_t val) {
    using C = Cell<ALGO>;
    const uint32_t r = (uint32_t)(*reinterpret_cast<const volatile typename C::T*>(reinterpret_cast<typename C::T*>(gacc) + idx));
    if (C::update(r, val) == r) return;
    constexpr uint32_t per = 4 / C::kBytes;
    constexpr uint32_t cmask = C::kBytes == 1 ? 0xffu : 0xffffu;
    uint32_t* wp = gacc + idx / per;
    const uint32_t sh = (idx % per) * (8 * C::kBytes);
    uint32_t old = *reinterpret_cast<volatile uint32_t*>(wp);
    for (;;) {
        const uint32_t cur = (old >> sh) & cmask;
        const uint32_t nw = C::update(cur, val);
        if (nw == cur) break;
        const uint32_t assumed = old;
        old = atomicCAS(wp, assumed, (assumed & ~(cmask << sh)) | (nw << sh));
        if (old == assumed) break;
    }
}

// k-mer width classes: the reference's own dispatch is k<=14 / 16 / else (utils.rs:466-502); here the
// split is by what fits one 32-bit word, with k == 16 (lash's default) special-cased because the
// window IS the word (no shift, no mask).
constexpr int kGroup = 16;  // k-mers whose atomics are deferred together (one 32-bit word of bases)
#ifndef LASH_DEFER_QUARTERS
#define LASH_DEFER_QUARTERS 1
#endif
#ifndef LASH_W_UNROLL
#define LASH_W_UNROLL 1
#endif
// the four word steps of a thread's 64 starts stay rolled: unrolled twice ULL k16 +0.3 %, HLL k21 -1.6 %, HMH +0.9 %; four
// times -6 % everywhere (instruction cache: one straight-line block is ~6.5 KB) -- tools/variant_sweep, B200
constexpr int kWordUnroll = LASH_W_UNROLL;

// CTA size is a template parameter so that the register budget follows it.  Measured on B200 (tools/variant_sweep):
// ~80 registers with 24 resident warps per SM beats 64 registers with 32 warps (+4 % at C2) and everything
// tighter (48 / 40 registers spill) -- the 16-k-mer deferred group wants the registers more than the SM wants warps.
//   accumulator <= LASH_SMALL_SMEM_KB (24 KiB) : 256-thread CTAs, LASH_MINB_256 (3) of them per SM
//   larger                                      : ONE CTA of LASH_TB_BIG (768) threads per SM -- every resident CTA
//                                                 is one more private accumulator to warm up and flush, which costs
//                                                 more than it hides once the accumulator is tens of KiB
#ifndef LASH_MINB_256
#define LASH_MINB_256 3
#endif
#ifndef LASH_TB_BIG
#define LASH_TB_BIG 768
#endif
#ifndef LASH_SMALL_SMEM_KB
#define LASH_SMALL_SMEM_KB 24
#endif
constexpr int kTbSmall = 256, kTbBig = LASH_TB_BIG;
template <int TB>
struct MinBlocks { static constexpr int value = TB == kTbSmall ? LASH_MINB_256 : 1; };

template <int ALGO, int KM, bool GLOBAL, int TB>
__global__ void __launch_bounds__(TB, MinBlocks<TB>::value)
    sketch_kernel(const uint32_t* __restrict__ packed, const uint32_t* __restrict__ inv_mask,
                  const SketchTile* __restrict__ tiles, uint32_t n_tiles, uint32_t* __restrict__ acc_global, int p, int k,
                  HashConsts hc, uint32_t cell_words, uint32_t n_cells, const uint64_t* __restrict__ span_kept) {
    using C = Cell<ALGO>;
    using A = SmemAcc<ALGO>;
    constexpr bool WIDE = KM == KWIDE;
    extern __shared__ uint32_t sacc[];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(sacc);
    if (!GLOBAL) {
        for (uint32_t i = threadIdx.x; i < n_cells * A::kWordsPerCell; i += blockDim.x) sacc[i] = 0u;
        __syncthreads();
    }
    // cells -> register bytes/halfwords -> merge the non-zero words into the genome's global
    // accumulator (max for HLL/HMH, packed-domain OR-merge for ULL); leaves the cells zeroed
    auto flush = [&](uint32_t genome) {
        __syncthreads();
        uint32_t* gacc = acc_global + (size_t)genome * cell_words;
        constexpr uint32_t per = 4 / C::kBytes;
        for (uint32_t i = threadIdx.x; i < cell_words; i += blockDim.x) {
            uint32_t v = 0u;
#pragma unroll
            for (uint32_t j = 0; j < per; ++j) {
                const uint32_t c = i * per + j;
                if (c < n_cells) {
                    v |= A::to_reg(sacc, c, p, n_cells) << (j * 8 * C::kBytes);
#pragma unroll
                    for (uint32_t z = 0; z < A::kWordsPerCell; ++z) sacc[A::word_of(c, z, n_cells)] = 0u;
                }
            }
            if (v == 0u) continue;
            uint32_t* gp = gacc + i;
            uint32_t old = *reinterpret_cast<volatile uint32_t*>(gp);
            for (;;) {
                const uint32_t nw = C::merge_word(old, v);
                if (nw == old) break;
                const uint32_t assumed = old;
                old = atomicCAS(gp, assumed, nw);
                if (old == assumed) break;
            }
        }
        __syncthreads();
    };

    // uniform shift amounts / masks
    const uint32_t narrow_shr = WIDE ? 0u : (uint32_t)(32 - 2 * k);             // fwd >> (32-2k)
    const uint32_t narrow_mask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    const uint32_t wide_shr = WIDE ? (uint32_t)(64 - 2 * k) : 0u;               // in [0,30]
    const uint32_t wide_mask_hi = (k >= 32) ? 0xffffffffu : ((1u << ((2 * k - 32) & 31)) - 1u);
    const uint32_t wide_mul = (WIDE && k < 32) ? (hc.two29 >> 29) << ((32u - wide_shr) & 31u) : 0u;   // 2^(32 - wide_shr), run-time

    // Persistent CTAs: CTA c owns the contiguous tile range [c*T/G, (c+1)*T/G) of the (genome-ordered)
    // tile list, so equal-cost tiles are balanced statically with no tail, and the private accumulator
    // is carried across consecutive tiles of the same genome (one warm-up + one flush per genome a CTA
    // touches instead of one per tile).
    const uint32_t t_begin = (uint32_t)(((uint64_t)blockIdx.x * n_tiles) / gridDim.x);
    const uint32_t t_end = (uint32_t)(((uint64_t)(blockIdx.x + 1) * n_tiles) / gridDim.x);
    uint32_t cur_genome = 0xffffffffu;
    for (uint32_t ti = t_begin; ti < t_end; ++ti) {
        SketchTile t = tiles[ti];
        if (t.clip != 0xffffffffu) {
            // span packed on the device (text_kernels.cu): the tile plan used the raw byte count as the base count
            const uint64_t kept = span_kept[t.clip];
            const uint64_t starts = kept >= (uint64_t)k ? kept - (uint64_t)k + 1 : 0;
            t.end = min(t.end, starts);
            if (t.begin >= t.end) continue;   // CTA-uniform
        }
        if (t.genome != cur_genome) {
            if (!GLOBAL && cur_genome != 0xffffffffu) flush(cur_genome);
            cur_genome = t.genome;
        }
        uint32_t* gacc = acc_global + (size_t)t.genome * cell_words;
        const uint32_t* base = packed + t.word_off;
        const uint32_t* mbase = (t.mask_word_off != ~0ull) ? inv_mask + t.mask_word_off : nullptr;

        // The trip count is uniform over the CTA (out-of-range threads carry valid == 0), so the warp can
        // be re-converged explicitly at the end of every iteration.
        const uint64_t per_iter = (uint64_t)blockDim.x * kStartsPerThread;
        const uint32_t n_iter = (uint32_t)((t.end - t.begin + per_iter - 1) / per_iter);
        for (uint32_t it = 0; it < n_iter; ++it) {
            const uint64_t s0 = t.begin + (uint64_t)it * per_iter + (uint64_t)threadIdx.x * kStartsPerThread;
            uint4 q = make_uint4(0u, 0u, 0u, 0u);
            uint2 h = make_uint2(0u, 0u);
            uint64_t valid = 0ull;
            if (s0 < t.end) {
                const uint32_t* wp = base + (s0 >> 4);
                q = __ldg(reinterpret_cast<const uint4*>(wp));
                h = __ldg(reinterpret_cast<const uint2*>(wp + 4));
                const uint64_t remain = t.end - s0;
                valid = remain >= 64 ? ~0ull : ((1ull << remain) - 1ull);
                if (mbase) {
                    const uint2 mv = __ldg(reinterpret_cast<const uint2*>(mbase + (s0 >> 5)));
                    valid &= ~(((uint64_t)mv.y << 32) | mv.x);
                }
            }
            // big-endian base order inside each word: first base in the top bits
            uint32_t f0 = __byte_perm(q.x, 0, 0x0123), f1 = __byte_perm(q.y, 0, 0x0123);
            uint32_t f2 = __byte_perm(q.z, 0, 0x0123), f3 = __byte_perm(q.w, 0, 0x0123);
            uint32_t f4 = __byte_perm(h.x, 0, 0x0123), f5 = __byte_perm(h.y, 0, 0x0123);

            // reverse complements of the words are carried across the four word steps (each word is the "B" of one
            // step and the "A" of the next; wide k-mers also look one word further)
            uint32_t rcA = rc16(f0), rcB = WIDE ? rc16(f1) : 0u;
    #pragma unroll kWordUnroll
            for (int w = 0; w < 4; ++w) {
                const uint32_t v16 = (uint32_t)(valid >> (16 * w)) & 0xffffu;
                const uint32_t A0 = f0, B0 = f1, C0 = f2;
                const uint32_t Ar = rcA, Br = WIDE ? rcB : rc16(B0), Cr = WIDE ? rc16(C0) : 0u;
                // canonical masked k-mer starting at base i of this word (sh = 2*i): funnel-shift windows
                // of the forward stream and of the reverse-complemented stream, then min
                auto kmer = [&](const int sh, uint32_t& klo, uint32_t& khi) {
                    canonical_kmer<KM>(A0, B0, C0, Ar, Br, Cr, sh, narrow_shr, narrow_mask, wide_shr, wide_mask_hi, wide_mul, klo, khi);
                };
                // exact, checked, rolled path: partial validity, rare hashes, global accumulators
                auto exact_block = [&](const uint32_t mask16) {
    #pragma unroll 1
                    for (int i = 0; i < 16; ++i) {
                        if (!((mask16 >> i) & 1u)) continue;
                        uint32_t klo, khi;
                        kmer(2 * i, klo, khi);
                        if (GLOBAL) {
                            uint32_t idx, val;
                            C::from_kmer(klo, khi, hc, p, idx, val);
                            global_update<ALGO>(gacc, idx, val);
                        } else {
                            A::exact(klo, khi, hc, p, sbase);
                        }
                    }
                };
                // straight-line group: the shared-memory atomics of 16 k-mers are deferred behind ONE
                // branch, so the hot path has no divergence.  CHECKED masks out starts that are invalid
                // (record boundary inside the word, tile tail) -- every ~150 bases for short reads.
                auto fast_block = [&](auto checked, const uint32_t mask16) {
                    constexpr bool CHECKED = decltype(checked)::value;
                    uint32_t addr[kGroup], need[kGroup];
                    uint32_t rare = 0xffffffffu, rw_prev = 0xffffffffu;
                    // The deferred atomics run in four quarters of four: late in a genome a warp's 512 k-mers hold one or two
                    // updates, and walking all 16 predicated atomics for them costs more than four extra tests.  Measured
                    // (tools/variant_sweep, B200): ULL k16 +1.0 %, HLL k21 +0.7 ... 1.8 %, 150 bp reads +1.4 %, HMH -0.7 % (its
                    // block is entered more often: every new 10-bit signature is an update) -- so not for HMH.
                    constexpr bool kQuarters = LASH_DEFER_QUARTERS && ALGO != HMH;
                    uint32_t anyq[4] = {0u, 0u, 0u, 0u};
#pragma unroll
                    for (int i = 0; i < kGroup; ++i) {
                        uint32_t klo, khi, v, rw;
                        kmer(2 * i, klo, khi);
                        A::template prep<!WIDE, TB == kTbSmall>(klo, khi, hc, p, sbase, addr[i], v, rw);
                        need[i] = A::need(lds_u32(addr[i]), v);
                        if (CHECKED) need[i] = (mask16 & (1u << i)) ? need[i] : 0u;
                        anyq[kQuarters ? i >> 2 : 0] |= need[i];
                        if (i & 1) rare = __vimin3_u32(rare, rw_prev, rw);  // one VIMNMX3 per two k-mers
                        else rw_prev = rw;
                    }
                    if (anyq[0] | anyq[1] | anyq[2] | anyq[3]) {
                        if constexpr (kQuarters) {
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                if (anyq[q]) {
#pragma unroll
                                    for (int i = 4 * q; i < 4 * q + 4; ++i)
                                        if (need[i]) A::apply(addr[i], need[i]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int i = 0; i < kGroup; ++i)
                                if (need[i]) A::apply(addr[i], need[i]);
                        }
                    }
                    if (rare == 0u) exact_block(mask16);  // some hash had 32 leading zeros where it matters
                };
                // The choice between the two straight-line blocks is made per WARP, not per lane: with short reads
                // nearly every warp holds some lane with a record boundary in its 16 starts, and a per-lane choice
                // would run both ~500-instruction blocks back to back with half the lanes idle in each (ncu: 15 of 32
                // active lanes, 252 Gbp/s on 150 bp reads).  Lanes without a valid start ride along with mask 0.
                if (GLOBAL) {
                    if (v16 != 0u) exact_block(v16);
                } else if (__all_sync(0xffffffffu, v16 == 0xffffu)) {
                    fast_block(std::false_type{}, 0xffffu);
                } else if (__any_sync(0xffffffffu, v16 != 0u)) {
                    fast_block(std::true_type{}, v16);
                }
                rcA = Br;
                rcB = Cr;
                f0 = f1; f1 = f2; f2 = f3; f3 = f4; f4 = f5;
            }
            __syncwarp();
        }

    }
    if (!GLOBAL && cur_genome != 0xffffffffu) flush(cur_genome);
}

// Mark every k-mer start that would cross the END of a record (or belongs to a record shorter than
// k).  Bit b of a span's mask <=> start position b.  All threads of the grid stride over the records
// of every multi-record span (a metagenome sample is ONE span with ~10^8 records).
__global__ void build_invalid_mask_kernel(const SpanRecs* __restrict__ spans, uint32_t n_spans,
                                          const uint64_t* __restrict__ rec_start, uint32_t* __restrict__ mask, int k) {
    const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t nthreads = (uint64_t)gridDim.x * blockDim.x;
    for (uint32_t si = 0; si < n_spans; ++si) {
        const SpanRecs s = spans[si];
        const uint64_t* rs = rec_start + s.rec_first;
        uint32_t* m = mask + s.mask_word_off;
        for (uint64_t r = tid; r < s.n_rec; r += nthreads) {
            uint64_t b, e;
            if (s.uniform_len) {
                b = r * (uint64_t)s.uniform_len;
                e = min(b + (uint64_t)s.uniform_len, s.n_bases);
            } else {
                b = rs[r];
                e = rs[r + 1];
            }
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
}

cudaError_t launch_build_invalid_mask(const SpanRecs* spans_dev, uint32_t n_spans, uint64_t n_rec_total,
                                      const uint64_t* rec_start_dev, uint32_t* mask_dev, int k, int n_sm, cudaStream_t st) {
    if (n_spans == 0) return cudaSuccess;
    uint64_t blocks = (n_rec_total + 255) / 256;
    const uint64_t cap = (uint64_t)n_sm * 8;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    build_invalid_mask_kernel<<<(unsigned)blocks, 256, 0, st>>>(spans_dev, n_spans, rec_start_dev, mask_dev, k);
    return cudaGetLastError();
}

// Fold `src` sketches into `dst` register-wise (same algorithm and precision): max for HLL / HMH
// (HyperLogLog::union, hyperminhash merge), pack(unpack(a) | unpack(b)) for ULL (UltraLogLog::merge,
// utils.rs:260-262).  Used when ONE sample is sketched in shares (several GPUs, several passes).
template <int ALGO>
__global__ void merge_kernel(uint32_t* __restrict__ dst, const uint32_t* __restrict__ src, uint64_t n_words) {
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride)
        dst[i] = Cell<ALGO>::merge_word(dst[i], __ldg(src + i));
}
cudaError_t launch_merge(int algo, uint32_t* dst, const uint32_t* src, uint64_t n_words, int n_sm, cudaStream_t st) {
    if (n_words == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<uint64_t>((n_words + 255) / 256, (uint64_t)n_sm * 8);
    switch (algo) {
        case HMH: merge_kernel<HMH><<<grid, 256, 0, st>>>(dst, src, n_words); break;
        case HLL: merge_kernel<HLL><<<grid, 256, 0, st>>>(dst, src, n_words); break;
        case ULL: merge_kernel<ULL><<<grid, 256, 0, st>>>(dst, src, n_words); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

void plan_sketch(SketchParams& sp) {
    sp.n_cells = sp.algo == HMH ? 16384u : (1u << sp.p);
    const uint32_t cell_bytes = sp.algo == HMH ? 2u : 1u;
    sp.cell_words = (sp.n_cells * cell_bytes + 3u) / 4u;
    const uint64_t smem = (uint64_t)sp.n_cells * (sp.algo == ULL ? 8u : 4u);
    sp.global_acc = smem > kMaxSmemAccBytes;
    sp.smem_bytes = sp.global_acc ? 0u : (uint32_t)smem;
    sp.threads = sp.smem_bytes <= LASH_SMALL_SMEM_KB * 1024u ? (uint32_t)kTbSmall : (uint32_t)kTbBig;
}

template <int ALGO, int KM, bool GLOBAL, int TB>
static cudaError_t launch_tb(const SketchParams& sp, const uint32_t* packed, const uint32_t* mask,
                             const SketchTile* tiles, uint32_t n_tiles, uint32_t* acc, cudaStream_t st, const uint64_t* span_kept) {
    auto kern = sketch_kernel<ALGO, KM, GLOBAL, TB>;
    size_t smem = GLOBAL ? 0 : (size_t)sp.smem_bytes;
    if (smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int occ = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, (int)sp.threads, smem);
    if (e != cudaSuccess) return e;
    const uint32_t resident = (uint32_t)std::max(occ, 1) * (uint32_t)sp.n_sm;
    const uint32_t grid = n_tiles < resident ? n_tiles : resident;  // one persistent CTA per resident slot
    kern<<<grid, sp.threads, smem, st>>>(packed, mask, tiles, n_tiles, acc, sp.p, sp.k, sp.hc, sp.cell_words, sp.n_cells, span_kept);
    return cudaGetLastError();
}
template <int ALGO, int KM, bool GLOBAL>
static cudaError_t launch_one(const SketchParams& sp, const uint32_t* packed, const uint32_t* mask,
                              const SketchTile* tiles, uint32_t n_tiles, uint32_t* acc, cudaStream_t st, const uint64_t* span_kept) {
    if (GLOBAL || sp.threads == (uint32_t)kTbSmall) return launch_tb<ALGO, KM, GLOBAL, kTbSmall>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept);
    return launch_tb<ALGO, KM, false, kTbBig>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept);
}

template <int ALGO>
static cudaError_t launch_algo(const SketchParams& sp, const uint32_t* packed, const uint32_t* mask,
                               const SketchTile* tiles, uint32_t n_tiles, uint32_t* acc, cudaStream_t st, const uint64_t* span_kept) {
    if (sp.global_acc) {
        return sp.k > 16 ? launch_one<ALGO, KWIDE, true>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept)
                         : launch_one<ALGO, KNARROW, true>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept);
    }
    if (sp.k > 16) return launch_one<ALGO, KWIDE, false>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept);
    if (sp.k == 16) return launch_one<ALGO, K16, false>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept);
    return launch_one<ALGO, KNARROW, false>(sp, packed, mask, tiles, n_tiles, acc, st, span_kept);
}

cudaError_t launch_sketch(const SketchParams& sp, const uint32_t* packed_dev, const uint32_t* mask_dev,
                          const SketchTile* tiles_dev, uint32_t n_tiles, uint32_t* acc_dev, cudaStream_t st,
                          const uint64_t* span_kept_dev) {
    if (n_tiles == 0) return cudaSuccess;
    switch (sp.algo) {
        case HMH: return launch_algo<HMH>(sp, packed_dev, mask_dev, tiles_dev, n_tiles, acc_dev, st, span_kept_dev);
        case HLL: return launch_algo<HLL>(sp, packed_dev, mask_dev, tiles_dev, n_tiles, acc_dev, st, span_kept_dev);
        case ULL: return launch_algo<ULL>(sp, packed_dev, mask_dev, tiles_dev, n_tiles, acc_dev, st, span_kept_dev);
    }
    return cudaErrorInvalidValue;
}

}  // namespace lash
