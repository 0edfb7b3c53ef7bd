// filter_out_n (reference src/utils.rs:33-41) fused with the 2-bit packing that
// kmerutils' `Sequence::new(&seq, 2)` performs (utils.rs:464): keep bytes that are exactly one of
// "ACGT", code A0 C1 G2 T3, four bases per byte.
//
// While a span is being built the stream is kept LSB-first in 64-bit little-endian words (base i at
// bits [2(i%32), 2(i%32)+1] of word i/32): appending a block of compressed codes is then one shift
// + or.  finalize() flips the 2-bit groups inside every byte so the bytes match the lash_gpu.h
// format (first base in the two most significant bits).
#pragma once
#include <cstddef>
#include <cstdint>

namespace lashhost {

class BaseStream {
  public:
    // buf: 16-byte aligned; cap_bytes: writable bytes (the caller keeps 32 bytes of slack past room())
    void attach(uint8_t* buf, size_t cap_bytes) {
        w_ = reinterpret_cast<uint64_t*>(buf);
        cap_ = cap_bytes;
        n_ = 0;
    }
    uint64_t size() const { return n_; }
    uint64_t* words() { return w_; }
    void set_size(uint64_t n) { n_ = n; }  // after the caller wrote words() itself, keeping the invariants of append64
    // bases that can still be appended (leaves room for the spill word and the 16-byte ABI pad)
    uint64_t room() const {
        const uint64_t usable = cap_ > 48 ? (cap_ - 48) * 4 : 0;
        return usable > n_ ? usable - n_ : 0;
    }
    // v: cnt 2-bit codes, first base in the low bits, bits above 2*cnt zero; cnt <= 32
    inline void append64(uint64_t v, unsigned cnt) {
        const uint64_t idx = n_ >> 5;
        const unsigned bit = (unsigned)(n_ & 31) * 2;
        if (bit == 0) {
            w_[idx] = v;
        } else {
            w_[idx] |= v << bit;
            w_[idx + 1] = v >> (64 - bit);
        }
        n_ += cnt;
    }
    inline void push_base(unsigned code) { append64(code & 3u, 1); }
    unsigned base_at(uint64_t i) const { return (unsigned)(w_[i >> 5] >> ((i & 31) * 2)) & 3u; }
    void truncate(uint64_t n) {
        n_ = n;
        const unsigned bit = (unsigned)(n & 31) * 2;
        if (bit) w_[n >> 5] &= (1ull << bit) - 1ull;
    }
    // filter + code + pack; the caller guarantees room() >= n.  Returns the number of bases kept.
    // use_simd: 0 scalar table, 1 default (AVX-512 VBMI2 where present, else AVX2+BMI2; LASH_PACK_ISA=avx2 forces the latter), 2 AVX2+BMI2,
    // 3 AVX-512 VBMI2 where present
    uint64_t append_filtered(const uint8_t* s, size_t n, int use_simd = 1);
    // LSB-first words -> ABI bytes, zero padding up to padded_bytes(); returns padded_bytes(size())
    uint64_t finalize();

    static uint64_t padded_bytes(uint64_t n_bases) {  // == lash_sketch_padded_bytes
        const uint64_t b = (n_bases + 3) / 4;
        return ((b + 15) / 16) * 16 + 16;
    }

  private:
    uint64_t* w_ = nullptr;
    size_t cap_ = 0;
    uint64_t n_ = 0;
};

bool pack_has_simd();
bool pack_has_avx512();  // AVX-512 F/BW/VL/VBMI2: the 64-byte path

}  // namespace lashhost
