// Pure arithmetic of the pair-table distance kernels (K4b / K4c in dist_kernels.cu): register recoding, table entries,
// hash4j's ML contributions.  Kept in a header of its own so that tests/host_shim can compile it with g++ and check every
// table entry against the register-level definitions on a machine without a GPU (tests/test_device_math.py).
#pragma once
#include <cstdint>

#include "registers.cuh"

namespace lash {

// sentinel contribution of a register outside FGRA's table range [4p+4, 252): makes the running sum
// explode (>= 2^600) so the pair is re-done by the exact path; normal sums are <= 2^26 * 0.85
#define LASH_FGRA_SENTINEL 0x1p600
#define LASH_FGRA_SENTINEL_TEST 0x1p500

// ------------------------------------------------------------------------------------------------
// register-pair primitives (plain 32-bit integer ops; the byte-SIMD "video" intrinsics are emulated
// with ~10 instructions each on sm_100 and were the bottleneck of the first version)
// ------------------------------------------------------------------------------------------------
// ULL union of two register bytes == pack(unpack(a) | unpack(b))  (ultraloglog merge, utils.rs:260-262)
__device__ __forceinline__ uint32_t ull_merge_fast(uint32_t a, uint32_t b) {
    const uint32_t hi = max(a, b), lo = min(a, b);
    const uint32_t d = min((hi >> 2) - (lo >> 2), 3u);
    // 3-bit window (1,w1,w0) of the smaller register, 0 if it is empty (valid non-empty registers are >= 8)
    const uint32_t x = (lo & 3u) | (min(lo, 4u) & 4u);
    return hi | ((x >> d) & 3u);
}

// hash4j contribute(): alpha contribution (scaled by 2^64) of one register byte, and the bit pattern
// W it adds to the b[] statistics: b[j] += bit j of W  (W = unpack(r) >> (p-1) for valid registers)
__device__ __forceinline__ uint64_t ml_ret_of(uint32_t r, int p) {
    const int r2 = (int)r - 4 * p - 4;
    if (r2 < 0) {
        uint64_t ret = 4;
        if (r2 == -2 || r2 == -8) ret -= 2;
        if (r2 == -2 || r2 == -4) ret -= 1;
        return ret << (62 - p);
    }
    const int k = r2 >> 2;
    uint64_t ret = 0xE000000000000000ULL;
    ret -= (uint64_t)(r & 1u) << 63;
    ret -= (uint64_t)((r >> 1) & 1u) << 62;
    return ret >> (k + p);
}
__device__ __forceinline__ uint64_t ml_w_of(uint32_t r, int p) {
    const int r2 = (int)r - 4 * p - 4;
    if (r2 < 0) {
        uint64_t w = 0;
        if (r2 == -2 || r2 == -8) w |= 1;
        if (r2 == -2 || r2 == -4) w |= 2;
        return w;
    }
    return (uint64_t)(4u | (r & 3u)) << (r2 >> 2);  // b[k] += y0, b[k+1] += y1, b[k+2] += 1
}

// ---- pair tables ------------------------------------------------------------------------------------------------
// Register bytes are recoded at staging time to c = 0 (empty) or r - base + 1 (1..126; 127 = "outside the table").
// Contract: base <= every non-empty register present (it is 4p-4 -- whose predecessor is not a valid register -- or the
// smallest register of the two sets rounded down to a multiple of 4), so r == base - 1, which would alias code 0, cannot occur.
__device__ __forceinline__ uint32_t fgra_code(uint32_t r, uint32_t base) {
    const uint32_t c = min(r - base + 1u, 127u);  // r < base wraps to a huge value -> 127
    return r ? c : 0u;
}

// FGRA table entry for the code pair (ca, cb): the contribution of merge(ra, rb), or the sentinel when a code is 127 or the
// merged register needs the small- / large-range treatment (off = 4p+4; reg = REGISTER_CONTRIBUTIONS).
__device__ __forceinline__ double fgra_tab_entry(uint32_t ca, uint32_t cb, uint32_t base, uint32_t off, const double* reg) {
    double v = LASH_FGRA_SENTINEL;
    if (ca != 127u && cb != 127u) {
        const uint32_t ra = ca ? ca + base - 1u : 0u, rb = cb ? cb + base - 1u : 0u;
        const uint32_t m = ull_merge1(ra, rb);
        if (m >= off && m < 252u) v = reg[m - off];
    }
    return v;
}
// ML tables: the merged register of the code pair (code 127 is staged like an empty register; its sketch is flagged and the
// pair redone exactly, so the entry only has to be harmless)
__device__ __forceinline__ uint32_t ml_tab_merged(uint32_t ca, uint32_t cb, uint32_t base) {
    const uint32_t ra = (ca && ca != 127u) ? ca + base - 1u : 0u, rb = (cb && cb != 127u) ? cb + base - 1u : 0u;
    return ull_merge1(ra, rb);
}

// ---- K4c, "G-sum" form of S ------------------------------------------------------------------------------------------------
// For a register with r2 = r - 4p - 4 >= 0 (top level k = r2 >> 2, sub-bits y1 y0) hash4j's contribute() returns
//     ret = (7 - 4 y0 - 2 y1) * 2^(61-k-p) = g(k+2) + (1 - y1) g(k+1) + (1 - y0) g(k),      g(j) = 2^(63-j-p),
// and adds  W = 1 << (k+2) | y1 << (k+1) | y0 << k  to the counts b[].  ret plus the register's share of  sum_j b[j] g(j)
// is  2 g(k+2) + g(k+1) + g(k) = 2 g(k), so over a whole (merged) sketch, mod 2^64,
//     S = sum_regs 2^(64-k-p)  -  sum_j b[j] * 2^(63-j-p):
// the table lookup of the 64-bit contribution (two LDS + a 64-bit add per register pair) becomes a sum of powers of two of
// the merged TOP LEVEL -- and the top level of a merge is the larger of the two, so with G(r) = 2^(28-(k-k0)), k0 the
// smallest top level among the tile's sketches, the term is min(G(ra), G(rb)): VIMNMX + IMAD, K4i's arithmetic, and
//     sum_regs 2^(64-k-p) = (sum of the minima) << (36 - p - k0).
// Valid while every register of both sketches has r2 >= 0 (no empty or "small-range" registers: the tile's smallest
// register decides), k - k0 <= 27 (sketches above it are flagged for the tile, as for W's 32-bit limit k <= 29) and
// 36 - p - k0 >= 0; exact because k + p <= 61 always holds on this path.  Eight terms fit a 32-bit batch (8 * 2^28).
// n = k - k0 + 4 in 4..31 is what the staged query word carries (5 bits); G = 2^32 >> n.
__device__ __forceinline__ uint32_t ml_gs_n(uint32_t r, int p, uint32_t k0) { return (((r - 4u * (uint32_t)p + 12u) >> 2) - k0) & 31u; }
__device__ __forceinline__ uint32_t ml_gs_term(uint32_t n) { return __funnelshift_r(0u, 1u, n); }   // uses n & 31
constexpr uint32_t kMlGsSpan = 27;
// top level of a register with r2 >= 0
__device__ __forceinline__ uint32_t ml_gs_k(uint32_t r, int p) { return (r - 4u * (uint32_t)p - 4u) >> 2; }
// largest register a sketch may hold on a G-sum tile anchored at k0
__device__ __forceinline__ uint32_t ml_gs_max_reg(int p, uint32_t k0) { return 4u * (uint32_t)p + 4u + 4u * (k0 + kMlGsSpan) + 3u; }
// two query registers in one staged word: q = code << 2 (byte offset in a 128-word table row) and n, one instruction each
//     bits 0..8 q0 | 9..13 n0 | 14..18 n1 | 23..31 q1
__device__ __forceinline__ uint32_t ml_pack_b(uint32_t c0, uint32_t n0, uint32_t c1, uint32_t n1) {
    return (c0 << 2) | (n0 << 9) | (n1 << 14) | (c1 << 25);
}
__device__ __forceinline__ uint32_t ml_b_q0(uint32_t w) { return w & 0x1fcu; }
__device__ __forceinline__ uint32_t ml_b_q1(uint32_t w) { return w >> 23; }
__device__ __forceinline__ uint32_t ml_b_n0(uint32_t w) { return w >> 9; }     // low 5 bits
__device__ __forceinline__ uint32_t ml_b_n1(uint32_t w) { return w >> 14; }    // low 5 bits
// S from the G-sum and the counts (b[j], j < 32: W fits 32 bits on this path)
__device__ __forceinline__ uint64_t ml_gs_S(uint64_t gsum, const int* b, int p, uint32_t k0) {
    uint64_t s = gsum << (36 - p - (int)k0);
    for (int j = 0; j < 32; ++j) s -= (uint64_t)(uint32_t)b[j] << (63 - j - p);
    return s;
}
// marker in MlAccT::mmax: S holds the G-sum of a tile anchored at k0 = mmax & 0xff (a real merged register is a byte)
constexpr uint32_t kMlGsMarker = 0x100u;

// ---- K4 / K4c: tables in shared memory and the ML accumulator -------------------------------------------------------------
struct SharedTables {
    const double* fgra_tab;    // [256] contribution of a merged register byte (sentinel outside [4p+4, 252))
    const uint64_t* ml_ret;    // [256]
    const uint32_t* ml_wlo;    // [256] low 32 bits of W
    const double* hll_pow;     // [256] 2^-r
};

// carry-save adder on 32-bit bit-planes: (a + b + c) -> sum (weight 1) and carry (weight 2); 2 LOP3
__device__ __forceinline__ void csa(uint32_t& carry, uint32_t& sum, uint32_t a, uint32_t b, uint32_t c) {
    const uint32_t u = a ^ b;
    carry = (a & b) | (u & c);
    sum = u ^ c;
}


constexpr int kMlPlanes = 27;  // counts up to 2^26 registers
// NPL = bit planes of the vertical counters: a count never exceeds the 2^p registers of a sketch, so p + 1 planes are
// enough; the pair-table kernel is instantiated for 12 / 16 / 27 planes (15 fewer planes = 30 fewer live registers and
// 30 fewer LOP3 per ripple for the two accumulators of a thread at p <= 11).
template <int NPL>
struct MlAccT {
    using CT = uint8_t;
    static constexpr int RM = 1, QM = 2;
    static constexpr int kTableBytes = 256 * 8 + 256 * 4;
    uint64_t S;
    uint32_t mmax;                 // largest merged register seen (decides whether W fits 32 bits)
    uint32_t pl[NPL];              // vertical (bit-sliced) counters of the low 32 bits of W
    __device__ __forceinline__ void init() {
        S = 0;
        mmax = 0;
#pragma unroll
        for (int i = 0; i < NPL; ++i) pl[i] = 0u;
    }
    // add a bit-plane of weight 2^L into the vertical counter (ripple carry, nplanes is CTA-uniform)
    template <int L>
    __device__ __forceinline__ void ripple(uint32_t x, int nplanes) {
#pragma unroll
        for (int l = L; l < NPL; ++l) {
            if (l < nplanes) {
                const uint32_t c = pl[l] & x;
                pl[l] ^= x;
                x = c;
            }
        }
    }
    template <int N>
    __device__ __forceinline__ void add_group(const uint32_t* a, const uint32_t* b, const SharedTables& t, int nplanes) {
        uint32_t w[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t m = ull_merge_fast(a[i], b[i]);
            mmax = max(mmax, m);
            S += t.ml_ret[m];
            w[i] = t.ml_wlo[m];
        }
        add_w<N>(w, nplanes);
    }
    // unconditional ripple through every plane from L up (no per-plane range test: 2 LOP3 per plane)
    template <int L>
    __device__ __forceinline__ void ripple_all(uint32_t x) {
#pragma unroll
        for (int l = L; l < NPL; ++l) {
            const uint32_t c = pl[l] & x;
            pl[l] ^= x;
            x = c;
        }
    }
    // Harley-Seal step over 8 patterns: planes 0..2 absorb them, the weight-8 carry is returned
    __device__ __forceinline__ uint32_t csa8(const uint32_t* w) {
        uint32_t t2a, t2b, t4a, t4b, t8;
        csa(t2a, pl[0], pl[0], w[0], w[1]);
        csa(t2b, pl[0], pl[0], w[2], w[3]);
        csa(t4a, pl[1], pl[1], t2a, t2b);
        csa(t2a, pl[0], pl[0], w[4], w[5]);
        csa(t2b, pl[0], pl[0], w[6], w[7]);
        csa(t4b, pl[1], pl[1], t2a, t2b);
        csa(t8, pl[2], pl[2], t4a, t4b);
        return t8;
    }
    // eight weight-8 carries (64 patterns) -> planes 3..5, then ONE ripple from plane 6: the ripple, which costs
    // 2 ops per plane, runs once per 64 registers instead of once per 8 or 16
    __device__ __forceinline__ void fold64(const uint32_t* t8) {
        uint32_t t16a, t16b, t32a, t32b, t64;
        csa(t16a, pl[3], pl[3], t8[0], t8[1]);
        csa(t16b, pl[3], pl[3], t8[2], t8[3]);
        csa(t32a, pl[4], pl[4], t16a, t16b);
        csa(t16a, pl[3], pl[3], t8[4], t8[5]);
        csa(t16b, pl[3], pl[3], t8[6], t8[7]);
        csa(t32b, pl[4], pl[4], t16a, t16b);
        csa(t64, pl[5], pl[5], t32a, t32b);
        ripple_all<6>(t64);
    }
    // fold N bit patterns W (b[j] += bit j of W) into the vertical counters
    template <int N>
    __device__ __forceinline__ void add_w(const uint32_t* w, int nplanes) {
        // Harley-Seal: N one-bit words -> weights 1,2,4,(8) planes, then one ripple of the top carry
        uint32_t t2a, t2b, t4a, t4b;
        csa(t2a, pl[0], pl[0], w[0], w[1]);
        csa(t2b, pl[0], pl[0], w[2], w[3]);
        csa(t4a, pl[1], pl[1], t2a, t2b);
        csa(t2a, pl[0], pl[0], w[4], w[5]);
        csa(t2b, pl[0], pl[0], w[6], w[7]);
        csa(t4b, pl[1], pl[1], t2a, t2b);
        uint32_t t8a;
        csa(t8a, pl[2], pl[2], t4a, t4b);
        if (N == 8) {
            ripple<3>(t8a, nplanes);
        } else {
            csa(t2a, pl[0], pl[0], w[8 % N], w[9 % N]);
            csa(t2b, pl[0], pl[0], w[10 % N], w[11 % N]);
            csa(t4a, pl[1], pl[1], t2a, t2b);
            csa(t2a, pl[0], pl[0], w[12 % N], w[13 % N]);
            csa(t2b, pl[0], pl[0], w[14 % N], w[15 % N]);
            csa(t4b, pl[1], pl[1], t2a, t2b);
            uint32_t t8b, t16;
            csa(t8b, pl[2], pl[2], t4a, t4b);
            csa(t16, pl[3], pl[3], t8a, t8b);
            ripple<4>(t16, nplanes);
        }
    }
};
using MlAcc = MlAccT<kMlPlanes>;

// counts b[j] (j = 0..65) out of the vertical counters: bit l of b[j] = bit j of plane l
template <int NPL>
__device__ __forceinline__ void ml_counts_from_planes(const MlAccT<NPL>& a, int* bb) {
#pragma unroll 1
    for (int j = 0; j < 66; ++j) bb[j] = 0;
    uint32_t any = 0;
#pragma unroll
    for (int l = 0; l < NPL; ++l) any |= a.pl[l];
    if (any) {
        const int jlo = __ffs((int)any) - 1, jhi = 31 - __clz((int)any);
        for (int j = jlo; j <= jhi; ++j) {
            uint32_t c = 0;
#pragma unroll
            for (int l = 0; l < NPL; ++l) c |= ((a.pl[l] >> j) & 1u) << l;
            bb[j] = (int)c;
        }
    }
}

// ---- K4h: HLL registers recoded to the high word of the double 2^-r ---------------------------------------------------
// (the low word of a power of two is 0, so 2^-max(ra, rb) = hiloint2double(min(va, vb), 0) with va = kHllOne - (ra << 20))
constexpr uint32_t kHllOne = 0x3FF00000u;                                 // high word of 2^-0
__device__ __forceinline__ uint4 hll_recode(uint32_t w) {
    return make_uint4(kHllOne - (w & 0xffu) * 0x100000u, kHllOne - ((w >> 8) & 0xffu) * 0x100000u,
                      kHllOne - ((w >> 16) & 0xffu) * 0x100000u, kHllOne - (w >> 24) * 0x100000u);
}
// any zero byte in w?  (exact for all byte values)
__device__ __forceinline__ bool has_zero_byte(uint32_t w) { return ((w - 0x01010101u) & ~w & 0x80808080u) != 0u; }

// ---- K4i: HLL registers as 32-bit fixed-point terms -------------------------------------------------------------------
// With `lo` = the smallest register of the sketches a tile touches, register r becomes the INTEGER
//     v(r) = 2^(kHllIntW - (r - lo))   for lo <= r <= lo + kHllIntW,   0 for larger r   ("out of window"),
// so min(v(ra), v(rb)) = v(max(ra, rb)) and the pair's  sum 2^-max  is  2^-(lo + kHllIntW) * (sum of the minima), an integer
// sum.  When no register of either sketch is out of the window, every term of the reference's sequential f64 loop is a
// multiple of u = 2^-(lo + 28) and every PARTIAL sum is n * u with n <= 2^p * 2^28 <= 2^46 < 2^53: the f64 loop never
// rounds, so its result IS the exact sum, in any order, and equals the integer sum scaled by a power of two
// (tests/test_device_math.py checks this against sequential double sums).  Sketches with an out-of-window register are
// flagged per tile (per-sketch min / max, hll_minmax_kernel) and their pairs take the sequential f64 path.
// kHllIntBatch minima are added in a 32-bit register before they go into the 64-bit sum: 8 * 2^28 = 2^31 cannot overflow.
constexpr int kHllIntW = 28;
constexpr int kHllIntBatch = 8;
// bytes of w minus lo (no byte of w is below lo, so there is no borrow between the bytes)
__device__ __forceinline__ uint32_t hll_int_rebase(uint32_t w, uint32_t lo) { return w - lo * 0x01010101u; }
// one rebased register byte -> v:  2^28 >> min(d, 32)
__device__ __forceinline__ uint32_t hll_int_term(uint32_t d) { return __funnelshift_rc(1u << kHllIntW, 0u, d); }
__device__ __forceinline__ uint4 hll_int_recode(uint32_t w, uint32_t lo) {
    const uint32_t d = hll_int_rebase(w, lo);
    return make_uint4(hll_int_term(d & 0xffu), hll_int_term((d >> 8) & 0xffu), hll_int_term((d >> 16) & 0xffu), hll_int_term(d >> 24));
}

}  // namespace lash
