// Test infrastructure: compiles the DEVICE arithmetic headers (lash_b200/csrc/hash.cuh, registers.cuh) with g++ so that the
// hash paths and register algebra the kernels use can be checked on a machine without a GPU (tests/test_device_math.py).
// The shim below stands in for the CUDA intrinsics those headers call; LASH_HOST_SHIM swaps their four inline-PTX helpers
// for plain C.  Nothing here is part of the product.
#include <algorithm>
#include <cstdint>

#define LASH_HOST_SHIM 1
#define __CUDACC__ 1
#define __device__
#define __host__
#define __forceinline__ inline

using std::max;
using std::min;

struct uint4 { uint32_t x, y, z, w; };
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t s) {  // high word of (hi:lo) << (s & 31)
    s &= 31u;
    return s ? (hi << s) | (lo >> (32u - s)) : hi;
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t s) {  // low word of (hi:lo) >> (s & 31)
    s &= 31u;
    return s ? (lo >> s) | (hi << (32u - s)) : lo;
}
static inline uint32_t __funnelshift_rc(uint32_t lo, uint32_t hi, uint32_t s) {  // low word of (hi:lo) >> min(s, 32)
    s = s > 32u ? 32u : s;
    return s == 32u ? hi : s ? (lo >> s) | (hi << (32u - s)) : lo;
}
static inline uint32_t __brev(uint32_t x) {
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0f0f0f0fu) | ((x & 0x0f0f0f0fu) << 4);
    return __builtin_bswap32(x);
}
static inline uint32_t __umulhi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
static inline uint64_t __umul64hi(uint64_t a, uint64_t b) { return (uint64_t)(((unsigned __int128)a * b) >> 64); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
static inline int __clzll(long long x) { return x ? __builtin_clzll((unsigned long long)x) : 64; }
static inline uint32_t bytewise(uint32_t a, uint32_t b, uint32_t (*f)(uint32_t, uint32_t)) {
    uint32_t r = 0;
    for (int i = 0; i < 4; ++i) r |= (f((a >> (8 * i)) & 0xffu, (b >> (8 * i)) & 0xffu) & 0xffu) << (8 * i);
    return r;
}
static inline uint32_t __vmaxu4(uint32_t a, uint32_t b) { return bytewise(a, b, [](uint32_t x, uint32_t y) { return x > y ? x : y; }); }
static inline uint32_t __vminu4(uint32_t a, uint32_t b) { return bytewise(a, b, [](uint32_t x, uint32_t y) { return x < y ? x : y; }); }
static inline uint32_t __vcmpne4(uint32_t a, uint32_t b) { return bytewise(a, b, [](uint32_t x, uint32_t y) { return x != y ? 0xffu : 0u; }); }
static inline uint32_t __vmaxu2(uint32_t a, uint32_t b) {
    const uint32_t lo = std::max(a & 0xffffu, b & 0xffffu), hi = std::max(a >> 16, b >> 16);
    return (hi << 16) | lo;
}

#include "../../lash_b200/csrc/registers.cuh"
#include "../../lash_b200/csrc/dist_tables.cuh"
#include "../../lash_b200/csrc/kmer_windows.cuh"

using namespace lash;

// ---- ML bit-sliced counters (MlAccT in dist_tables.cuh) ---------------------------------------------------------------
// feeds n patterns W (b[j] += bit j of W) through the accumulator exactly as dist_ml_tab_kernel does -- 8 at a time through
// csa8, eight carries through fold64 (or ripple_all<3> per 8 when fewer than 64 are left in a "chunk") -- and extracts b[]
template <int NPL>
static void run_ml_counters(const uint32_t* w, uint64_t n, uint64_t chunk, int* bb) {
    MlAccT<NPL> acc;
    acc.init();
    for (uint64_t c0 = 0; c0 < n; c0 += chunk) {
        if (chunk >= 64) {
            for (uint64_t e = c0; e < c0 + chunk; e += 64) {
                uint32_t carry[8];
                for (int g = 0; g < 8; ++g) carry[g] = acc.csa8(w + e + 8 * g);
                acc.fold64(carry);
            }
        } else {
            for (uint64_t e = c0; e < c0 + chunk; e += 8) acc.template ripple_all<3>(acc.csa8(w + e));
        }
    }
    ml_counts_from_planes(acc, bb);
}

extern "C" {

// ---- hashes (arrays in, arrays out) -----------------------------------------------------------------------------
void dm_xxh3_64(const uint64_t* v, uint64_t n, uint64_t seed, uint64_t* out) {
    const HashConsts hc = make_hash_consts(seed);
    for (uint64_t i = 0; i < n; ++i) out[i] = xxh3_64_le64((uint32_t)v[i], (uint32_t)(v[i] >> 32), hc);
}
// pre-xorshift forms used by the kernels' fast paths: out[0] = narrow/wide_pre (64 bit), out_hi = the *_hi variant
void dm_pre(const uint64_t* v, uint64_t n, uint64_t seed, int narrow, uint64_t* out, uint32_t* out_hi) {
    const HashConsts hc = make_hash_consts(seed);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
        out[i] = narrow ? xxh3_64_narrow_pre(lo, hc) : xxh3_64_wide_pre(lo, hi, hc);
        out_hi[i] = narrow ? xxh3_64_narrow_pre_hi(lo, hc) : xxh3_64_wide_pre_hi(lo, hi, hc);
    }
}
void dm_xxh3_128(const uint32_t* w, uint64_t n, uint64_t seed, uint64_t* out_lo, uint64_t* out_hi) {
    const HashConsts hc = make_hash_consts(seed);
    for (uint64_t i = 0; i < n; ++i) xxh3_128_le32(w[i], hc, out_lo[i], out_hi[i]);
}

// ---- (index, value) a k-mer contributes, in the register domain (Cell<ALGO>::from_kmer) -------------------------------
void dm_cell(int algo, const uint64_t* v, uint64_t n, uint64_t seed, int p, uint32_t* idx, uint32_t* val) {
    const HashConsts hc = make_hash_consts(seed);
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t lo = (uint32_t)v[i], hi = (uint32_t)(v[i] >> 32);
        if (algo == HLL) Cell<HLL>::from_kmer(lo, hi, hc, p, idx[i], val[i]);
        else if (algo == ULL) Cell<ULL>::from_kmer(lo, hi, hc, p, idx[i], val[i]);
        else Cell<HMH>::from_kmer(lo, hi, hc, p, idx[i], val[i]);
    }
}

// ---- what the ULL / HLL fast paths of sketch_kernels.cu derive from the pre-xorshift high word -------------------------
// (restated from SmemAcc<ULL>::prep / SmemAcc<HLL>::prep: index, the word whose highest set bit gives nlz / rho, and whether
// the group would be sent to the exact path)
void dm_ull_fast(const uint32_t* ghi, uint64_t n, int p, int drop4, uint32_t* idx, uint32_t* nlz, uint32_t* rare) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t hi = ghi[i];
        const uint32_t t = drop4 ? hi & ((0xffffffffu >> p) & ~15u) : (hi ^ (hi >> 28)) & (0xffffffffu >> p);
        idx[i] = __umulhi(hi, 1u << p);
        const uint32_t v = shl_clamp(1u << p, bfind32(t));  // bit j of cell word 0 <=> nlz = 31 - j
        rare[i] = t == 0u;
        nlz[i] = v ? 31u - bfind32(v) : 0xffffffffu;
    }
}

// ---- canonical k-mers as the sketch kernel extracts them (kmer_windows.cuh) --------------------------------------------
// packed: 2-bit bases, first base in the high bits of byte 0 (the ABI layout); out[s] = canonical k-mer starting at base s
void dm_kmers(const uint8_t* packed, uint64_t n_bytes, uint64_t n_bases, int k, uint64_t* out) {
    auto word = [&](uint64_t j) -> uint32_t {  // 16 bases, first base in the top bits (what the kernel holds after its byte swap)
        uint32_t w = 0;
        for (int b = 0; b < 4; ++b) w = (w << 8) | (4 * j + b < n_bytes ? packed[4 * j + b] : 0u);
        return w;
    };
    const bool wide = k > 16;
    const uint32_t narrow_shr = wide ? 0u : (uint32_t)(32 - 2 * k);
    const uint32_t narrow_mask = (k >= 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    const uint32_t wide_shr = wide ? (uint32_t)(64 - 2 * k) : 0u;
    const uint32_t wide_mask_hi = (k >= 32) ? 0xffffffffu : ((1u << ((2 * k - 32) & 31)) - 1u);
    const uint32_t wide_mul = (wide && k < 32) ? 1u << ((32u - wide_shr) & 31u) : 0u;
    for (uint64_t s0 = 0; s0 + (uint64_t)k <= n_bases; ++s0) {
        const uint64_t j = s0 >> 4;
        const int sh = 2 * (int)(s0 & 15);
        const uint32_t A0 = word(j), B0 = word(j + 1), C0 = word(j + 2);
        const uint32_t Ar = rc16(A0), Br = rc16(B0), Cr = rc16(C0);
        uint32_t klo, khi;
        if (k == 16) canonical_kmer<K16>(A0, B0, C0, Ar, Br, Cr, sh, narrow_shr, narrow_mask, wide_shr, wide_mask_hi, wide_mul, klo, khi);
        else if (!wide) canonical_kmer<KNARROW>(A0, B0, C0, Ar, Br, Cr, sh, narrow_shr, narrow_mask, wide_shr, wide_mask_hi, wide_mul, klo, khi);
        else canonical_kmer<KWIDE>(A0, B0, C0, Ar, Br, Cr, sh, narrow_shr, narrow_mask, wide_shr, wide_mask_hi, wide_mul, klo, khi);
        out[s0] = ((uint64_t)khi << 32) | klo;
    }
}

// HLL fast path (SmemAcc<HLL>::prep): index from the low p bits of h.lo, rho from the pre-xorshift high word
void dm_hll_fast(const uint64_t* g, uint64_t n, int p, uint32_t* idx, uint32_t* rho, uint32_t* rare) {
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t glo = (uint32_t)g[i], ghi = (uint32_t)(g[i] >> 32);
        idx[i] = (glo ^ __funnelshift_r(glo, ghi, 28)) & ((1u << p) - 1u);
        rho[i] = 32u - bfind32(ghi);
        rare[i] = ghi == 0u;
    }
}

// ---- shared-memory ULL cell (two words of seen-nlz bits) and its conversion at flush ---------------------------------------
// adds the hashes h[0..n) to an empty cell with the EXACT path's rule (SmemAcc<ULL>::exact: word 0 takes nlz < 32, word 1
// the rest, bit index = raw bfind result) and converts the cell with ull_cell_to_reg; idx_out = the register index of h[0]
uint32_t dm_ull_cell(const uint64_t* h, uint64_t n, int p, uint32_t* idx_out, uint32_t* w_out) {
    uint32_t w[2] = {0u, 0u};
    for (uint64_t i = 0; i < n; ++i) {
        const uint32_t lo = (uint32_t)h[i], hi = (uint32_t)(h[i] >> 32);
        const uint32_t yh = __funnelshift_l(lo, hi, p), yl = (lo << p) | ((1u << p) - 1u);
        w[yh ? 0 : 1] |= 1u << bfind32(yh ? yh : yl);
    }
    *idx_out = (uint32_t)(h[0] >> (64 - p));
    w_out[0] = w[0];
    w_out[1] = w[1];
    return ull_cell_to_reg(w[0], w[1], p);
}

// ---- ML bit-sliced counters (run_ml_counters above) ---------------------------------------------------------------------
void dm_ml_counters(int npl, const uint32_t* w, uint64_t n, uint64_t chunk, int* bb) {
    if (npl == 12) run_ml_counters<12>(w, n, chunk, bb);
    else if (npl == 16) run_ml_counters<16>(w, n, chunk, bb);
    else run_ml_counters<kMlPlanes>(w, n, chunk, bb);
}
// the generic kernels' path: add_w<16> / add_w<8> with the run-time plane limit
void dm_ml_counters_generic(const uint32_t* w, uint64_t n, int group, int nplanes, int* bb) {
    MlAcc acc;
    acc.init();
    for (uint64_t e = 0; e < n; e += (uint64_t)group) {
        if (group == 16) acc.add_w<16>(w + e, nplanes);
        else acc.add_w<8>(w + e, nplanes);
    }
    ml_counts_from_planes(acc, bb);
}

// the G-sum form of S (dist_tables.cuh): two sketches without empty / small-range registers through the staged formats of
// dist_ml_tab_kernel's G-sum tiles -- reference side code << 9 and G, query side two registers per packed word -- one W lookup
// and min(G_a, G_b) per register pair, eight terms per 32-bit batch; returns S rebuilt by ml_gs_S, the counts in bits[32]
uint64_t dm_ml_gs_pair(const uint8_t* a, const uint8_t* b, uint32_t m, int p, uint32_t k0, uint64_t* gsum_out, int* bits) {
    const uint32_t base = (uint32_t)(4 * p - 4);
    uint64_t gsum = 0;
    for (int j = 0; j < 32; ++j) bits[j] = 0;
    for (uint32_t e = 0; e < m; e += 8) {
        uint32_t batch = 0;
        for (uint32_t i = 0; i < 8; i += 2) {
            const uint32_t wb = ml_pack_b(fgra_code(b[e + i], base), ml_gs_n(b[e + i], p, k0), fgra_code(b[e + i + 1], base), ml_gs_n(b[e + i + 1], p, k0));
            for (uint32_t h = 0; h < 2; ++h) {
                const uint32_t ca = fgra_code(a[e + i + h], base);
                const uint32_t ga = ml_gs_term(ml_gs_n(a[e + i + h], p, k0));
                const uint32_t q = h ? ml_b_q1(wb) : ml_b_q0(wb);
                const uint32_t gb = ml_gs_term(h ? ml_b_n1(wb) : ml_b_n0(wb));
                batch += min(ga, gb);
                const uint32_t w = (uint32_t)ml_w_of(ml_tab_merged(ca, q >> 2, base), p);
                for (int j = 0; j < 32; ++j) bits[j] += (w >> j) & 1u;
            }
        }
        gsum += batch;
    }
    *gsum_out = gsum;
    return ml_gs_S(gsum, bits, p, k0);
}
uint32_t dm_ml_gs_max_reg(int p, uint32_t k0) { return ml_gs_max_reg(p, k0); }
uint32_t dm_ml_gs_k(uint32_t r, int p) { return ml_gs_k(r, p); }
uint64_t dm_ml_ret_of(uint32_t r, int p) { return ml_ret_of(r, p); }

// ---- K4h: HLL registers as high words of 2^-r ---------------------------------------------------------------------------
void dm_hll_recode(uint32_t w, uint32_t* out4, int* zero_byte) {
    const uint4 v = hll_recode(w);
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
    *zero_byte = has_zero_byte(w) ? 1 : 0;
}

// ---- K4i: HLL registers as 32-bit fixed-point terms; the pair sum as the tile kernel forms it --------------------------
void dm_hll_int_recode(uint32_t w, uint32_t lo, uint32_t* out4) {
    const uint4 v = hll_int_recode(w, lo);
    out4[0] = v.x; out4[1] = v.y; out4[2] = v.z; out4[3] = v.w;
}
// sum of 2^-max(a[i], b[i]) over n registers (n a multiple of 8) through recode -> min -> 32-bit batches of kHllIntBatch ->
// 64-bit sum -> double scaled by 2^-(lo + kHllIntW); *zero = both-empty count the way the COUNT_ZERO loop takes it (lo == 0)
double dm_hll_int_pair_sum(const uint8_t* a, const uint8_t* b, uint32_t n, uint32_t lo, uint32_t* zero) {
    uint64_t sum = 0;
    uint32_t z = 0;
    for (uint32_t e = 0; e < n; e += kHllIntBatch) {
        uint32_t acc = 0;
        for (uint32_t i = 0; i < (uint32_t)kHllIntBatch; i += 4) {
            uint32_t wa, wb;
            __builtin_memcpy(&wa, a + e + i, 4);
            __builtin_memcpy(&wb, b + e + i, 4);
            const uint4 va = hll_int_recode(wa, lo), vb = hll_int_recode(wb, lo);
            const uint32_t m[4] = {min(va.x, vb.x), min(va.y, vb.y), min(va.z, vb.z), min(va.w, vb.w)};
            for (int j = 0; j < 4; ++j) {
                acc += m[j];   // 32-bit, as the kernel's batch register
                z += m[j] >> kHllIntW;
            }
        }
        sum += acc;
    }
    *zero = z;
    return (double)sum * __builtin_ldexp(1.0, -(int)(lo + kHllIntW));
}

uint32_t dm_ull_cell_to_reg(uint32_t w0, uint32_t w1, int p) { return ull_cell_to_reg(w0, w1, p); }

// ---- register algebra -----------------------------------------------------------------------------------------------
uint32_t dm_ull_update(uint32_t r, uint32_t u) { return ull_update(r, u); }
uint32_t dm_ull_merge1(uint32_t a, uint32_t b) { return ull_merge1(a, b); }
uint32_t dm_ull_merge4(uint32_t a, uint32_t b) { return ull_merge4(a, b); }
uint32_t dm_ull_merge_fast(uint32_t a, uint32_t b) { return ull_merge_fast(a, b); }

// ---- pair tables of the distance kernels (dist_tables.cuh) ----------------------------------------------------------------
uint32_t dm_fgra_code(uint32_t r, uint32_t base) { return fgra_code(r, base); }
// the whole 128 x 128 FGRA table exactly as dist_fgra_tab_kernel builds it; returns the sentinel value
double dm_fgra_table(uint32_t base, int p, const double* reg, double* out) {
    for (uint32_t e = 0; e < 128u * 128u; ++e) out[e] = fgra_tab_entry(e >> 7, e & 127u, base, (uint32_t)(4 * p + 4), reg);
    return LASH_FGRA_SENTINEL;
}
// the ML tables exactly as dist_ml_tab_kernel builds them (R = contribution to S, W = bit pattern added to b[])
void dm_ml_tables(int p, uint64_t* R, uint32_t* W) {
    const uint32_t base = (uint32_t)(4 * p - 4);
    for (uint32_t e = 0; e < 128u * 128u; ++e) {
        const uint32_t m = ml_tab_merged(e >> 7, e & 127u, base);
        R[e] = ml_ret_of(m, p);
        W[e] = (uint32_t)ml_w_of(m, p);
    }
}
}
