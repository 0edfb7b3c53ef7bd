// Register arithmetic of the three sketches, in the on-disk register domain (u8 / u16), so an
// accumulator cell IS the value `S::save` writes (reference utils.rs:400-433).
//
//   HLL  streaming_algorithms 0.3.3 push_hash64: idx = low p bits, rho = clz(h >> p) - p + 1, max
//   ULL  ultraloglog 0.1.6 add():  idx = top p bits, u = nlz + p - 1, reg = pack(unpack(reg) | 1<<u)
//        -- NOT a max; done here in the packed domain (no 64-bit unpack), see ull_update().
//   HMH  hyperminhash 0.1.4 add_hash(x,y): idx = x >> 50, reg = (lz << 10) | (y & 1023), max (u16)
#pragma once
#include <cstdint>

#include "hash.cuh"

namespace lash {

enum Algo : int { HMH = 0, HLL = 1, ULL = 2 };

// hyperminhash: which half of the 128-bit hash is `x` (index + leading zeros) and which is `y`
// (signature bits).  Recalled convention, isolated here.
#ifndef LASH_HMH_X_IS_HIGH64
#define LASH_HMH_X_IS_HIGH64 1
#endif

#ifdef __CUDACC__

// ---- ULL packed-domain update: new register after OR-ing bit u into the unpacked prefix --------
// register r = 4*t + 2*bit(t-1) + bit(t-2), t = highest set bit position (0 => empty).
__device__ __forceinline__ uint32_t ull_update(uint32_t r, uint32_t u) {
    int d = (int)u - (int)(r >> 2);
    uint32_t x = r ? (4u | (r & 3u)) : 0u;
    uint32_t up = (u << 2) | ((x >> min(d, 3)) & 3u);   // d > 0: u becomes the top bit
    uint32_t dn = r | ((d == -1) ? 2u : (d == -2) ? 1u : 0u);  // u inside / below the 2-bit window
    return d > 0 ? up : dn;
}

// ULL merge of two registers == pack(unpack(a) | unpack(b)) (ultraloglog merge, utils.rs:260-262)
__device__ __forceinline__ uint32_t ull_merge1(uint32_t a, uint32_t b) {
    uint32_t hi = max(a, b), lo = min(a, b);
    uint32_t d = (hi >> 2) - (lo >> 2);
    uint32_t x = lo ? (4u | (lo & 3u)) : 0u;
    return hi | ((x >> min(d, 3u)) & 3u);
}

// Four registers at once (SIMD-in-word).
__device__ __forceinline__ uint32_t ull_merge4(uint32_t a, uint32_t b) {
    uint32_t hi = __vmaxu4(a, b), lo = __vminu4(a, b);
    uint32_t d = ((hi >> 2) & 0x3f3f3f3fu) - ((lo >> 2) & 0x3f3f3f3fu);  // per byte, no borrow (hi>=lo)
    d = __vminu4(d, 0x03030303u);
    uint32_t nz = __vcmpne4(lo, 0u);                       // 0xff where lo != 0
    uint32_t x = ((lo & 0x03030303u) | 0x04040404u) & nz;  // 3-bit window of the smaller register
    uint32_t m1 = (d & 0x01010101u) * 0xffu;               // bytes with shift bit0
    uint32_t m2 = ((d >> 1) & 0x01010101u) * 0xffu;        // bytes with shift bit1
    x = (x & ~m1) | ((x >> 1) & 0x03030303u & m1);
    x = (x & ~m2) | ((x >> 2) & 0x01010101u & m2);
    return hi | (x & 0x03030303u);
}

// Shared-memory ULL cell of the sketch kernel -> register.  The cell is two words of "seen nlz" bits: word 0 bit j <=>
// nlz = 31-j was seen, word 1 bit j <=> nlz = 63-j.  nlz-mask M = brev(w0) | brev(w1) << 32, unpacked hash prefix =
// M << (p-1), register = ultraloglog pack(prefix) -- exact because sequential add()s equal pack(OR of 1 << u).
__device__ __forceinline__ uint32_t ull_cell_to_reg(uint32_t w0, uint32_t w1, int p) {
    if ((w0 | w1) == 0u) return 0u;
    const uint64_t m = mk64(__brev(w0), __brev(w1)) << (p - 1);  // unpacked hash prefix
    const uint32_t u = 63u - (uint32_t)__clzll((long long)m);   // u >= p-1 >= 2
    return (u << 2) | ((uint32_t)(m >> (u - 2u)) & 3u);          // ultraloglog pack()
}

template <int ALGO>
struct Cell;

template <>
struct Cell<HLL> {
    using T = uint8_t;
    static constexpr int kBytes = 1;
    __device__ static __forceinline__ void from_kmer(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t& idx,
                                                     uint32_t& val) {
        uint64_t h = xxh3_64_le64(klo, khi, hc);
        uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
        idx = lo & ((1u << p) - 1u);
        uint32_t wl = __funnelshift_r(lo, hi, p), wh = hi >> p;
        val = (uint32_t)(clz64_parts(wl, wh) - p + 1);
    }
    // shared-memory accumulator cells hold the same rho
    __device__ static __forceinline__ void from_kmer_smem(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t& idx,
                                                          uint32_t& val) {
        from_kmer(klo, khi, hc, p, idx, val);
    }
    __device__ static __forceinline__ uint32_t update(uint32_t r, uint32_t v) { return max(r, v); }
    __device__ static __forceinline__ uint32_t merge_word(uint32_t a, uint32_t b) { return __vmaxu4(a, b); }
};

template <>
struct Cell<ULL> {
    using T = uint8_t;
    static constexpr int kBytes = 1;
    __device__ static __forceinline__ void from_kmer(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t& idx,
                                                     uint32_t& val) {
        uint64_t h = xxh3_64_le64(klo, khi, hc);
        uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
        idx = hi >> (32 - p);
        // nlz = clz64(~(~h << p)) = clz64((h << p) | (2^p - 1))
        uint32_t yh = __funnelshift_l(lo, hi, p), yl = (lo << p) | ((1u << p) - 1u);
        val = (uint32_t)(clz64_parts(yl, yh) + p - 1);
    }
    // shared-memory accumulator cells are indexed by nlz itself (u = nlz + p - 1 is applied at flush)
    __device__ static __forceinline__ void from_kmer_smem(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t& idx,
                                                          uint32_t& nlz) {
        uint64_t h = xxh3_64_le64(klo, khi, hc);
        uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
        idx = hi >> (32 - p);
        uint32_t yh = __funnelshift_l(lo, hi, p), yl = (lo << p) | ((1u << p) - 1u);
        nlz = (uint32_t)clz64_parts(yl, yh);
    }
    __device__ static __forceinline__ uint32_t update(uint32_t r, uint32_t v) { return ull_update(r, v); }
    __device__ static __forceinline__ uint32_t merge_word(uint32_t a, uint32_t b) { return ull_merge4(a, b); }
};

template <>
struct Cell<HMH> {
    using T = uint16_t;
    static constexpr int kBytes = 2;
    __device__ static __forceinline__ void from_kmer(uint32_t klo, uint32_t /*khi*/, const HashConsts& hc, int /*p*/, uint32_t& idx,
                                                     uint32_t& val) {
        uint64_t hlo, hhi;
        xxh3_128_le32(klo, hc, hlo, hhi);  // utils.rs:397: only the low 32 bits of the k-mer are hashed
#if LASH_HMH_X_IS_HIGH64
        uint64_t x = hhi, y = hlo;
#else
        uint64_t x = hlo, y = hhi;
#endif
        idx = (uint32_t)(x >> 50);
        uint64_t t = (x << 14) | 0x3fffULL;
        uint32_t lz = (uint32_t)__clzll((long long)t) + 1u;
        val = (lz << 10) | ((uint32_t)y & 1023u);
    }
    __device__ static __forceinline__ void from_kmer_smem(uint32_t klo, uint32_t khi, const HashConsts& hc, int p, uint32_t& idx,
                                                          uint32_t& val) {
        from_kmer(klo, khi, hc, p, idx, val);
    }
    __device__ static __forceinline__ uint32_t update(uint32_t r, uint32_t v) { return max(r, v); }
    __device__ static __forceinline__ uint32_t merge_word(uint32_t a, uint32_t b) { return __vmaxu2(a, b); }
};

#endif  // __CUDACC__
}  // namespace lash
