// See pack.hpp.  Two implementations of the same byte -> (keep?, 2-bit code) map:
//   scalar: 256-entry table (any x86-64 / any CPU)
//   AVX2 + BMI2: 32 bytes per step -- pshufb classifies, shifts/xor code, pext gathers the 2-bit
//   codes, pdep/pext squeezes out the deleted bytes (newlines, N, lowercase, IUPAC ...), so dirty
//   input costs the same as clean input and there is no per-byte branch.
#include "pack.hpp"

#include <cstdlib>
#include <cstring>
#include <string>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

namespace lashhost {

namespace {

struct Lut {
    uint8_t code[256];  // 0..3, or 0xff when filter_out_n deletes the byte (utils.rs:36: only "ACTG" survive)
    Lut() {
        memset(code, 0xff, sizeof(code));
        code[(unsigned)'A'] = 0;
        code[(unsigned)'C'] = 1;
        code[(unsigned)'G'] = 2;
        code[(unsigned)'T'] = 3;
    }
};
const Lut g_lut;

uint64_t append_scalar(BaseStream& bs, const uint8_t* s, size_t n) {
    uint64_t kept = 0;
    size_t i = 0;
    while (i < n) {
        uint64_t v = 0;
        unsigned cnt = 0;
        const size_t e = (n - i) < 32 ? n : i + 32;
        for (; i < e; ++i) {
            const uint8_t c = g_lut.code[s[i]];
            if (c != 0xff) {
                v |= (uint64_t)c << (2 * cnt);
                ++cnt;
            }
        }
        if (cnt) bs.append64(v, cnt);
        kept += cnt;
    }
    return kept;
}

#if defined(__x86_64__)
__attribute__((target("avx2,bmi2,popcnt"))) uint64_t append_avx2(BaseStream& bs, const uint8_t* s, size_t n) {
    // pshufb table indexed by the low nibble: the only byte with that nibble that survives the filter
    // ('A' 0x41, 'C' 0x43, 'T' 0x54, 'G' 0x47); entry 0 is 0xff so that byte 0x00 cannot match itself,
    // and bytes >= 0x80 make pshufb return 0, which never equals them.
    const __m256i lut = _mm256_setr_epi8((char)0xff, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0,
                                         (char)0xff, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0);
    const __m256i three = _mm256_set1_epi8(3);
    const __m256i w14 = _mm256_set1_epi16(0x0401);       // byte pairs  -> c0 + 4 c1
    const __m256i w116 = _mm256_set1_epi32(0x00100001);  // word pairs  -> (c0 + 4 c1) + 16 (c2 + 4 c3): 4 bases per byte
    const __m256i low_bytes = _mm256_setr_epi8(0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1,
                                               0, 4, 8, 12, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1);
    // the partially filled last word lives in a register for the whole call
    uint64_t* w = bs.words();
    const uint64_t n0 = bs.size();
    uint64_t idx = n0 >> 5;
    unsigned bit = (unsigned)(n0 & 31) * 2;
    uint64_t acc = bit ? w[idx] : 0;
    uint64_t kept = 0;
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        const __m256i c = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(s + i));
        const uint32_t m = (uint32_t)_mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(lut, c), c));
        // code = ((c >> 1) ^ (c >> 2)) & 3 : A 0, C 1, G 2, T 3 (bits leaking in from the neighbour byte are masked)
        const __m256i code = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(c, 1), _mm256_srli_epi16(c, 2)), three);
        const __m256i g = _mm256_shuffle_epi8(_mm256_madd_epi16(_mm256_maddubs_epi16(code, w14), w116), low_bytes);
        const uint64_t all = (uint64_t)(uint32_t)_mm256_cvtsi256_si32(g) |
                             ((uint64_t)(uint32_t)_mm_cvtsi128_si32(_mm256_extracti128_si256(g, 1)) << 32);
        const uint64_t keep2 = _pdep_u64((uint64_t)m, 0x5555555555555555ull) * 3ull;  // both bits of every kept base
        const uint64_t v = _pext_u64(all, keep2);                                       // deleted bytes squeezed out
        const unsigned cnt = (unsigned)_mm_popcnt_u32(m);
        acc |= v << bit;
        const unsigned nbits = bit + 2 * cnt;
        w[idx] = acc;
        const bool carry = nbits >= 64;
        idx += carry;
        acc = carry ? (v >> 1) >> (63 - bit) : acc;  // == v >> (64 - bit), and 0 when bit == 0
        bit = nbits & 63;
        kept += cnt;
    }
    if (bit) w[idx] = acc;
    bs.set_size(n0 + kept);
    if (i < n) kept += append_scalar(bs, s + i, n - i);
    return kept;
}

// AVX-512 (VBMI2): 64 bytes per step.  vpcompressb squeezes the deleted bytes out of the CODE vector directly (no
// pdep / pext round trip through general registers), vpmovdb gathers the four-codes-per-byte result, and a masked load
// takes the tail, so there is no scalar remainder.  Same stream layout and same results as the two paths above.
__attribute__((target("avx512f,avx512bw,avx512vl,avx512vbmi2,popcnt"))) uint64_t append_avx512(BaseStream& bs, const uint8_t* s, size_t n) {
    const __m512i lut = _mm512_broadcast_i32x4(_mm_setr_epi8((char)0xff, 'A', 0, 'C', 'T', 0, 0, 'G', 0, 0, 0, 0, 0, 0, 0, 0));
    const __m512i three = _mm512_set1_epi8(3);
    const __m512i w14 = _mm512_set1_epi16(0x0401);
    const __m512i w116 = _mm512_set1_epi32(0x00100001);
    uint64_t* w = bs.words();
    const uint64_t n0 = bs.size();
    uint64_t idx = n0 >> 5;
    unsigned bit = (unsigned)(n0 & 31) * 2;
    uint64_t acc = bit ? w[idx] : 0;
    uint64_t kept = 0;
    for (size_t i = 0; i < n; i += 64) {
        // the last, partial block is a masked load: bytes past the end read as 0x00, which the filter deletes
        const __m512i c = (i + 64 <= n) ? _mm512_loadu_si512(reinterpret_cast<const void*>(s + i))
                                        : _mm512_maskz_loadu_epi8((1ull << (n - i)) - 1ull, s + i);
        const __mmask64 k = _mm512_cmpeq_epi8_mask(_mm512_shuffle_epi8(lut, c), c);
        // code = ((c >> 1) ^ (c >> 2)) & 3 : A 0, C 1, G 2, T 3
        const __m512i code = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(c, 1), _mm512_srli_epi16(c, 2)), three);
        const __m512i dense = _mm512_maskz_compress_epi8(k, code);   // kept codes first, zeros behind
        const __m128i g = _mm512_cvtepi32_epi8(_mm512_madd_epi16(_mm512_maddubs_epi16(dense, w14), w116));  // 4 codes per byte
        const uint64_t lo = (uint64_t)_mm_cvtsi128_si64(g), hi = (uint64_t)_mm_extract_epi64(g, 1);
        const unsigned cnt = (unsigned)_mm_popcnt_u64((uint64_t)k);
        // append 2*cnt bits (lo, hi) at bit offset `bit`: up to two words complete.  (A mask-select instead of the
        // ternaries below is slower: 3.0 vs 4.1 GB/s -- it lengthens the loop-carried chain through acc, and on FASTA
        // text the word count per step is periodic enough to predict.)
        const uint64_t t0 = acc | (lo << bit);
        const uint64_t t1 = ((lo >> 1) >> (63 - bit)) | (hi << bit);   // (x >> 1) >> (63 - bit) == x >> (64 - bit), 0 for bit == 0
        const uint64_t t2 = (hi >> 1) >> (63 - bit);
        w[idx] = t0;
        w[idx + 1] = t1;                                               // the caller keeps slack past room()
        const unsigned nbits = bit + 2 * cnt;
        const unsigned full = nbits >> 6;
        idx += full;
        acc = full == 0 ? t0 : full == 1 ? t1 : t2;
        bit = nbits & 63;
        kept += cnt;
    }
    if (bit) w[idx] = acc;
    bs.set_size(n0 + kept);
    return kept;
}
#endif

}  // namespace

bool pack_has_avx512() {
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx512f") && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vl") &&
                           __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("popcnt");
    return ok;
#else
    return false;
#endif
}

bool pack_has_simd() {
#if defined(__x86_64__)
    static const bool ok = __builtin_cpu_supports("avx2") && __builtin_cpu_supports("bmi2") && __builtin_cpu_supports("popcnt");
    return ok;
#else
    return false;
#endif
}

uint64_t BaseStream::append_filtered(const uint8_t* s, size_t n, int use_simd) {
#if defined(__x86_64__)
    // Default: the 64-byte AVX-512 VBMI2 path where the CPU has it, else AVX2 + BMI2.  Round 1 measured the two the same
    // (the reader's mmap traffic hid the packer); with the buffered reader of round 2 the 64-byte path is ahead, modestly, on the GPU
    // boxes: pack-only 47.9 -> 52.5 Gbp/s on 16 workers and 15.6 -> 16.8 on 4 on one box, FASTA -> registers 49.2 -> 50.9 Gbp/s on
    // another (tools/ingest_probe.py; in cache 5.8 vs 7.7 GB/s per core).  LASH_PACK_ISA=avx2 forces the 32-byte path (A/B), =avx512 is accepted for symmetry.
    static const bool want_avx2 = [] { const char* v = getenv("LASH_PACK_ISA"); return v && std::string(v) == "avx2"; }();
    if ((use_simd == 3 || (use_simd == 1 && !want_avx2)) && pack_has_avx512()) return append_avx512(*this, s, n);
    if (use_simd && pack_has_simd()) return append_avx2(*this, s, n);
#endif
    (void)use_simd;
    return append_scalar(*this, s, n);
}

uint64_t BaseStream::finalize() {
    const uint64_t padded = padded_bytes(n_);
    const uint64_t words_used = (n_ + 31) / 32;
    const uint64_t words_total = padded / 8;
    // reverse the four 2-bit groups of every byte: LSB-first stream -> "first base in the high bits"
    for (uint64_t i = 0; i < words_used; ++i) {
        uint64_t x = w_[i];
        x = ((x >> 4) & 0x0f0f0f0f0f0f0f0full) | ((x & 0x0f0f0f0f0f0f0f0full) << 4);
        x = ((x >> 2) & 0x3333333333333333ull) | ((x & 0x3333333333333333ull) << 2);
        w_[i] = x;
    }
    for (uint64_t i = words_used; i < words_total; ++i) w_[i] = 0;
    return padded;
}

}  // namespace lashhost
