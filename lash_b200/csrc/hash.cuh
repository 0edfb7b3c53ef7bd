// XXH3 short-input paths on the integer pipe (sm_100a).
//
// What the reference hashes (file:line under the reference tree):
//   HLL/ULL  utils.rs:412,428  xxh3_64_with_seed(&masked.to_le_bytes(), seed)     -> 8-byte input
//   HMH      utils.rs:397      add_bytes_with_seed(&(masked as u32).to_le_bytes()) -> 4-byte input,
//                              128-bit XXH3 inside hyperminhash
// Both are XXH3 "len 4..8" short paths with the default secret (XXH3 v0.8 spec); the seed-dependent
// constants (bitflip words) are folded on the host once per sketcher (HashConsts).
#pragma once
#include <cstdint>

namespace lash {

struct HashConsts {
    // 64-bit path: bitflip = (secret[8..16] ^ secret[16..24]) - seed'
    uint32_t bf64_lo, bf64_hi;
    // 128-bit path: bitflip = (secret[16..24] ^ secret[24..32]) + seed'
    uint32_t bf128_lo, bf128_hi;
    // 64-bit path for k-mers that fit 32 bits (k <= 16): the keyed word is (lo = bf64_lo,
    // hi = kmer ^ bf64_hi), so the rotate-xor stage  h ^= rotl(h,49) ^ rotl(h,24)  splits into shifts
    // of the k-mer alone plus seed-only constants (shifts distribute over xor):
    //   new_lo = (kmer << 17) ^ (kmer >> 8) ^ nar_cl
    //   new_hi = kmer ^ ((kmer >> 15) | (nar_ch & 0xfffe0000)) ^ ((kmer << 24) + (nar_ch & 0x1ffff))
    uint32_t nar_cl, nar_ch;
    // 2^29 as a RUN-TIME value: (h >> 35) + 8 of the wide-k hash runs as one IMAD.HI on the FMA pipe (the wide kernels are
    // ALU-bound); with a literal ptxas strength-reduces it back to SHF + IADD
    uint32_t two29;
};

constexpr uint64_t kSecretX_8_16 = 0xc73ab174c5ecd5a2ULL;   // readLE64(kSecret+8) ^ readLE64(kSecret+16)
constexpr uint64_t kSecretX_16_24 = 0xc4f023344dc994acULL;  // readLE64(kSecret+16) ^ readLE64(kSecret+24)
constexpr uint64_t kPrimeMX1 = 0x165667919E3779F9ULL;
constexpr uint64_t kPrimeMX2 = 0x9FB21C651E98DF25ULL;
constexpr uint64_t kPrime64_1 = 0x9E3779B185EBCA87ULL;

inline HashConsts make_hash_consts(uint64_t seed) {
    uint32_t s32 = (uint32_t)seed;
    uint32_t sw = (s32 >> 24) | ((s32 >> 8) & 0xff00u) | ((s32 << 8) & 0xff0000u) | (s32 << 24);
    uint64_t sp = seed ^ ((uint64_t)sw << 32);
    uint64_t b64 = kSecretX_8_16 - sp;
    uint64_t b128 = kSecretX_16_24 + sp;
    HashConsts c;
    c.two29 = 1u << 29;
    c.bf64_lo = (uint32_t)b64;
    c.bf64_hi = (uint32_t)(b64 >> 32);
    c.bf128_lo = (uint32_t)b128;
    c.bf128_hi = (uint32_t)(b128 >> 32);
    const uint32_t L = c.bf64_lo, B = c.bf64_hi;
    c.nar_cl = L ^ (L >> 15) ^ (L << 24) ^ (B << 17) ^ (B >> 8);
    c.nar_ch = B ^ (B >> 15) ^ (B << 24) ^ (L << 17) ^ (L >> 8);
    return c;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint64_t mk64(uint32_t lo, uint32_t hi) { return ((uint64_t)hi << 32) | lo; }

// XXH3_rrmxmx(h, len = 8)
__device__ __forceinline__ uint64_t xxh3_rrmxmx8(uint64_t h) {
    uint32_t lo = (uint32_t)h, hi = (uint32_t)(h >> 32);
    // rotl64(h,49) = rotr64(h,15); rotl64(h,24)
    uint32_t r49_lo = __funnelshift_r(lo, hi, 15), r49_hi = __funnelshift_r(hi, lo, 15);
    uint32_t r24_lo = __funnelshift_l(hi, lo, 24), r24_hi = __funnelshift_l(lo, hi, 24);
    lo ^= r49_lo ^ r24_lo;
    hi ^= r49_hi ^ r24_hi;
    h = mk64(lo, hi) * kPrimeMX2;
    // h ^= (h >> 35) + len: (h >> 35) < 2^29, so the +8 cannot carry into the high word
    lo = (uint32_t)h;
    hi = (uint32_t)(h >> 32);
    lo ^= (hi >> 3) + 8u;
    h = mk64(lo, hi) * kPrimeMX2;
    return h ^ (h >> 28);
}

// 32-bit add the optimiser cannot widen: written in C, `lo ^ ((hi >> 3) + 8u)` next to mk64() is re-fused by
// the front end into a 64-bit add and lowered with a carry into the high word (LEA.HI.P + IMAD.X + LOP3:
// two wasted instructions per hash) although (h >> 35) + 8 < 2^30 can never carry.
// (LASH_HOST_SHIM: tests/host_shim compiles these headers with g++ to check the device arithmetic on a CPU; it is never
// defined in a product build, where the PTX forms below are the code.)
__device__ __forceinline__ uint32_t add32_opaque(uint32_t a, uint32_t b) {
#ifdef LASH_HOST_SHIM
    return a + b;
#else
    uint32_t r;
    asm("add.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
    return r;
#endif
}

// mad.lo.u32 the optimiser cannot re-associate: a 64x64->64 multiply is IMAD.WIDE (lo*M.lo) and two IMADs that add the
// cross terms ON TOP of the wide product's high word -- 3 instructions.  Written in C the cross terms are summed first
// and added to the high word afterwards (4 instructions: the extra IADD is pure issue-slot cost in this kernel).
__device__ __forceinline__ uint32_t mad32_opaque(uint32_t a, uint32_t b, uint32_t c) {
#ifdef LASH_HOST_SHIM
    return a * b + c;
#else
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
#endif
}
// (h1 >> 3) + 8, the high-word part of  h ^= (h >> 35) + len.  FMA = true: one IMAD.HI (h1 * 2^29 >> 32, + 8) on the FMA-heavy
// pipe instead of SHF + IADD on the ALU pipe.  LASH_WIDE_FMA=1 applies it -- together with the k-mer window's  >> wide_shr  as
// IMAD.HI and HLL's cell address as IMAD -- to the wide-k kernels, whose ALU pipe is 84-88 % busy with the FMA pipe at 26 %.
// Measured in round 2 (tools/variant_sweep, B200): 2.4 fewer ALU instructions per k-mer and 3 % SLOWER (HLL p14 k21 545 -> 530,
// ULL p10 k31 545 -> 527 Gbp/s) -- as for k <= 16 (LASH_SHR8_IMAD, -1.7 %) IMAD.HI costs the FMA pipe more than the SHF it
// replaces costs the ALU.  Default off.
#ifndef LASH_WIDE_FMA
#define LASH_WIDE_FMA 0
#endif
template <bool FMA>
__device__ __forceinline__ uint32_t shr3_add8(uint32_t h1, const struct HashConsts& c) {
    if (!FMA) return add32_opaque(h1 >> 3, 8u);
#ifdef LASH_HOST_SHIM
    return (uint32_t)(((uint64_t)h1 * c.two29) >> 32) + 8u;
#else
    uint32_t r;
    asm("mad.hi.u32 %0, %1, %2, 8;" : "=r"(r) : "r"(h1), "r"(c.two29));
    return r;
#endif
}
// (lo, hi) * kPrimeMX2 mod 2^64 in three instructions
__device__ __forceinline__ void mul_mx2(uint32_t& lo, uint32_t& hi) {
    constexpr uint32_t m_lo = (uint32_t)kPrimeMX2, m_hi = (uint32_t)(kPrimeMX2 >> 32);
    const uint64_t w = (uint64_t)lo * m_lo;
    hi = mad32_opaque(hi, m_lo, mad32_opaque(lo, m_hi, (uint32_t)(w >> 32)));
    lo = (uint32_t)w;
}
// tail of XXH3_rrmxmx after the rotate-xor stage, on halves
template <bool FMA = false>
__device__ __forceinline__ uint64_t xxh3_rrmxmx8_tail(uint32_t lo, uint32_t hi, const HashConsts& c) {
    mul_mx2(lo, hi);
    lo ^= shr3_add8<FMA>(hi, c);      // h ^= (h >> 35) + len, no carry into the high word
    mul_mx2(lo, hi);
    return mk64(lo, hi);               // caller applies the final h ^= h >> 28 (or only the part it needs)
}
template <bool FMA = false>
__device__ __forceinline__ uint32_t xxh3_rrmxmx8_tail_hi(uint32_t lo, uint32_t hi, const HashConsts& c) {
    constexpr uint32_t m_lo = (uint32_t)kPrimeMX2, m_hi = (uint32_t)(kPrimeMX2 >> 32);
    const uint64_t w = (uint64_t)lo * m_lo;                                   // IMAD.WIDE
    uint32_t h1 = mad32_opaque(lo, m_hi, (uint32_t)(w >> 32));                // + lo * M.hi
    h1 = mad32_opaque(hi, m_lo, h1);                                          // + hi * M.lo
    const uint32_t l1 = (uint32_t)w ^ shr3_add8<FMA>(h1, c);                  // h ^= (h >> 35) + 8
    return mad32_opaque(h1, m_lo, mad32_opaque(l1, m_hi, __umulhi(l1, m_lo)));  // high word of the second product
}
// pre-xorshift hash (h before `h ^= h >> 28`) of a k-mer that fits 32 bits; see HashConsts::nar_*
__device__ __forceinline__ uint64_t xxh3_64_narrow_pre(uint32_t kmer, const HashConsts& c) {
    const uint32_t t2 = __funnelshift_r(kmer, c.nar_ch >> 17, 15);      // (kmer >> 15) | (nar_ch & 0xfffe0000)
    const uint32_t t3 = kmer * (1u << 24) + (c.nar_ch & 0x1ffffu);      // disjoint bits: + is ^
    const uint32_t hi = kmer ^ t2 ^ t3;
    const uint32_t lo = (kmer * (1u << 17)) ^ (kmer >> 8) ^ c.nar_cl;
    return xxh3_rrmxmx8_tail(lo, hi, c);
}
__device__ __forceinline__ uint32_t xxh3_64_narrow_pre_hi(uint32_t kmer, const HashConsts& c) {
    const uint32_t t2 = __funnelshift_r(kmer, c.nar_ch >> 17, 15);
    const uint32_t t3 = kmer * (1u << 24) + (c.nar_ch & 0x1ffffu);
    const uint32_t hi = kmer ^ t2 ^ t3;
#ifdef LASH_SHR8_IMAD  // tuning switch: kmer >> 8 as IMAD.HI on the FMA pipe -- measured 1.7 % SLOWER (826 -> 812 Gbp/s)
    const uint32_t lo = (kmer * (1u << 17)) ^ __umulhi(kmer, 1u << 24) ^ c.nar_cl;
#else
    const uint32_t lo = (kmer * (1u << 17)) ^ (kmer >> 8) ^ c.nar_cl;
#endif
    return xxh3_rrmxmx8_tail_hi(lo, hi, c);
}
__device__ __forceinline__ uint32_t xxh3_64_wide_pre_hi(uint32_t v_lo, uint32_t v_hi, const HashConsts& c) {
    uint32_t lo = v_hi ^ c.bf64_lo, hi = v_lo ^ c.bf64_hi;
    const uint32_t r49_lo = __funnelshift_r(lo, hi, 15), r49_hi = __funnelshift_r(hi, lo, 15);
    const uint32_t r24_lo = __funnelshift_l(hi, lo, 24), r24_hi = __funnelshift_l(lo, hi, 24);
    lo ^= r49_lo ^ r24_lo;
    hi ^= r49_hi ^ r24_hi;
    return xxh3_rrmxmx8_tail_hi<LASH_WIDE_FMA>(lo, hi, c);
}
// pre-xorshift hash of a general 64-bit k-mer value
__device__ __forceinline__ uint64_t xxh3_64_wide_pre(uint32_t v_lo, uint32_t v_hi, const HashConsts& c) {
    uint32_t lo = v_hi ^ c.bf64_lo, hi = v_lo ^ c.bf64_hi;
    const uint32_t r49_lo = __funnelshift_r(lo, hi, 15), r49_hi = __funnelshift_r(hi, lo, 15);
    const uint32_t r24_lo = __funnelshift_l(hi, lo, 24), r24_hi = __funnelshift_l(lo, hi, 24);
    lo ^= r49_lo ^ r24_lo;
    hi ^= r49_hi ^ r24_hi;
    return xxh3_rrmxmx8_tail<LASH_WIDE_FMA>(lo, hi, c);
}

// xxh3_64_with_seed(le64(v), seed); v given as two 32-bit halves.
// input64 = in2 + (in1 << 32) with in1 = low half, in2 = high half  => (lo,hi) swapped.
__device__ __forceinline__ uint64_t xxh3_64_le64(uint32_t v_lo, uint32_t v_hi, const HashConsts& c) {
    return xxh3_rrmxmx8(mk64(v_hi ^ c.bf64_lo, v_lo ^ c.bf64_hi));
}

// xxh3_128_with_seed(le32(w), seed) -> (lo64, hi64)
__device__ __forceinline__ void xxh3_128_le32(uint32_t w, const HashConsts& c, uint64_t& out_lo, uint64_t& out_hi) {
    uint64_t keyed = mk64(w ^ c.bf128_lo, w ^ c.bf128_hi);
    constexpr uint64_t mult = kPrime64_1 + (4u << 2);
    uint64_t lo = keyed * mult;
    uint64_t hi = __umul64hi(keyed, mult);
    hi += lo << 1;
    lo ^= hi >> 3;
    lo ^= lo >> 35;
    lo *= kPrimeMX2;
    lo ^= lo >> 28;
    hi ^= hi >> 37;
    hi *= kPrimeMX1;
    hi ^= hi >> 32;
    out_lo = lo;
    out_hi = hi;
}

// bfind.u32: bit position of the most significant 1, 0xffffffff for 0 (SASS FLO, XU pipe)
__device__ __forceinline__ uint32_t bfind32(uint32_t x) {
#ifdef LASH_HOST_SHIM
    return x ? 31u - (uint32_t)__builtin_clz(x) : 0xffffffffu;
#else
    uint32_t r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
#endif
}
// shl.b32 clamps: shift amounts >= 32 give 0 (C's << would be undefined)
__device__ __forceinline__ uint32_t shl_clamp(uint32_t x, uint32_t n) {
#ifdef LASH_HOST_SHIM
    return n >= 32u ? 0u : x << n;
#else
    uint32_t r;
    asm("shl.b32 %0, %1, %2;" : "=r"(r) : "r"(x), "r"(n));
    return r;
#endif
}
__device__ __forceinline__ int clz64_parts(uint32_t lo, uint32_t hi) {
    return hi ? __clz(hi) : 32 + __clz(lo);  // __clz(0) == 32
}
#endif

}  // namespace lash
