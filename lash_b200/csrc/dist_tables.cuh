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

}  // namespace lash
