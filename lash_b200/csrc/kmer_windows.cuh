// Canonical k-mers as funnel-shift windows (K1, sketch_kernels.cu): pure bit arithmetic, in a header of its own so that
// tests/host_shim can compile it with g++ and compare every k against the oracle's string-level k-mers on a CPU.
//
// The packed stream holds 16 bases per 32-bit word, first base in the TOP bits (after the kernel's byte swap).  For a
// k-mer starting at base i of word A (sh = 2*i): the forward word is a left funnel-shift window of (A, B[, C]); the
// reverse complement is a right funnel-shift window of the per-word reverse-complemented stream (Ar = rc16(A), ...),
// whose words come in the opposite order.  canonical = min(forward, reverse complement), masked to 2k bits
// (utils.rs:57-64, 470).
#pragma once
#include <cstdint>

#include "hash.cuh"

namespace lash {

// k-mer width classes: the reference's own dispatch is k<=14 / 16 / else (utils.rs:466-502); here the
// split is by what fits one 32-bit word, with k == 16 (lash's default) special-cased because the
// window IS the word (no shift, no mask).
enum KMode : int { K16 = 0, KNARROW = 1, KWIDE = 2 };

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t rc16(uint32_t f) {
    // reverse-complement of 16 bases held big-endian (first base in the top 2 bits)
    uint32_t y = __brev(f);
    y = ((y >> 1) & 0x55555555u) | ((y & 0x55555555u) << 1);
    return ~y;
}

// narrow_shr = 32 - 2k, narrow_mask = 2^(2k) - 1 (k <= 16);  wide_shr = 64 - 2k, wide_mask_hi = 2^(2k-32) - 1 (k > 16);
// wide_mul = 2^(32 - wide_shr) as a run-time value (0 for k = 32, where nothing is shifted): the high word's  >> wide_shr
// runs as IMAD.HI on the FMA pipe (LASH_WIDE_FMA, hash.cuh)
template <int KM>
__device__ __forceinline__ void canonical_kmer(uint32_t A0, uint32_t B0, uint32_t C0, uint32_t Ar, uint32_t Br, uint32_t Cr, int sh,
                                               uint32_t narrow_shr, uint32_t narrow_mask, uint32_t wide_shr, uint32_t wide_mask_hi,
                                               uint32_t wide_mul, uint32_t& klo, uint32_t& khi) {
    if (KM == K16) {
        klo = min(__funnelshift_l(B0, A0, sh), __funnelshift_r(Ar, Br, sh));
        khi = 0u;
    } else if (KM == KNARROW) {
        const uint32_t fw = __funnelshift_l(B0, A0, sh) >> narrow_shr;
        const uint32_t rc = __funnelshift_r(Ar, Br, sh) & narrow_mask;
        klo = min(fw, rc);
        khi = 0u;
    } else {
        uint32_t fhi = __funnelshift_l(B0, A0, sh), flo = __funnelshift_l(C0, B0, sh);
        flo = __funnelshift_r(flo, fhi, wide_shr);
        if (LASH_WIDE_FMA)
            fhi = wide_shr ? __umulhi(fhi, wide_mul) : fhi;   // uniform predicate
        else
            fhi >>= wide_shr;
        const uint32_t rlo = __funnelshift_r(Ar, Br, sh);
        const uint32_t rhi = __funnelshift_r(Br, Cr, sh) & wide_mask_hi;
        const uint64_t f64 = mk64(flo, fhi), r64 = mk64(rlo, rhi);
        const uint64_t c64 = f64 < r64 ? f64 : r64;
        klo = (uint32_t)c64;
        khi = (uint32_t)(c64 >> 32);
    }
}
#endif  // __CUDACC__

}  // namespace lash
