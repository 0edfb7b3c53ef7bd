// K3/K4: per-sketch cardinality and all-vs-all distance tiles (sm_100a; integer + FP64 pipes,
// no tensor cores -- none of this is a dense contraction).
//
// Replaces the reference's par_iter bodies src/utils.rs:150-180 (HMH), :248-285 (ULL),
// :342-370 (HLL) and compute_distance (src/main.rs:415-423).
//
// Design: a CTA owns a TR x TQ tile of (reference, query) pairs.  The two register tiles are staged
// through shared memory in chunks of <= 1 KiB per sketch (row stride padded by one word, so the 16
// distinct query rows a warp touches fall in 16 distinct banks and the reference row is a
// broadcast).  A thread owns an RM x QM micro-tile of pairs and walks the registers IN INDEX ORDER
// with one scalar accumulator per pair: the FP64 sums therefore see exactly the addition sequence
// of the scalar CPU code (register order), which makes the estimator inputs bit-identical to the
// oracle's; only pow/log in the per-pair epilogue can differ from a host libm by an ulp.
// The union itself never exists in memory: HLL = __vmaxu4, ULL = packed-domain OR-merge of four
// registers per 32-bit word (ull_merge4), HMH = SIMD halfword equality / non-zero counts.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <string>
#include <type_traits>

#include "estimators.cuh"
#include "dist_tables.cuh"
#include "kernels.h"
#include "registers.cuh"

namespace lash {


constexpr int kDistThreads = 256;
constexpr int kChunkBytes = 1024;  // bytes of registers per sketch per stage


// ------------------------------------------------------------------------------------------------
// per-pair accumulators.  add_group<N>() consumes N consecutive registers of both sketches (N = 16,
// or 8 for the 8-register ULL p=3), ALWAYS in index order with one scalar FP64 accumulator.
// ------------------------------------------------------------------------------------------------
struct HllAcc {
    using CT = uint8_t;
    static constexpr int RM = 2, QM = 2;
    static constexpr int kTableBytes = 256 * 8;
    double sum;
    uint32_t zero;
    __device__ __forceinline__ void init() { sum = 0.0; zero = 0; }
    template <int N>
    __device__ __forceinline__ void add_group(const uint32_t* a, const uint32_t* b, const SharedTables& t) {
#pragma unroll
        for (int i = 0; i < N; ++i) {
            const uint32_t m = max(a[i], b[i]);
            zero += (m == 0u);
            sum += t.hll_pow[m];
        }
    }
};

struct FgraAcc {
    using CT = uint8_t;
    static constexpr int RM = 2, QM = 2;
    static constexpr int kTableBytes = 256 * 8;
    double sum;
    __device__ __forceinline__ void init() { sum = 0.0; }
    template <int N>
    __device__ __forceinline__ void add_group(const uint32_t* a, const uint32_t* b, const SharedTables& t) {
#pragma unroll
        for (int i = 0; i < N; ++i) sum += t.fgra_tab[ull_merge_fast(a[i], b[i])];
    }
};


struct HmhAcc {
    using CT = uint16_t;
    static constexpr int RM = 2, QM = 2;
    static constexpr int kTableBytes = 0;
    uint32_t C, N;
    __device__ __forceinline__ void init() { C = N = 0; }
    template <int G>
    __device__ __forceinline__ void add_group(const uint32_t* a, const uint32_t* b, const SharedTables&) {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            C += (a[i] == b[i]) & (a[i] != 0u);   // hyperminhash similarity(): equal and non-empty
            N += ((a[i] | b[i]) != 0u);
        }
    }
};

template <class ACC>
__device__ __forceinline__ void build_tables(SharedTables& t, unsigned char* smem_tab, int p) {
    double* d = reinterpret_cast<double*>(smem_tab);
    uint64_t* u = reinterpret_cast<uint64_t*>(smem_tab);
    uint32_t* w = reinterpret_cast<uint32_t*>(smem_tab + 256 * 8);
    t.fgra_tab = d;
    t.hll_pow = d;
    t.ml_ret = u;
    t.ml_wlo = w;
    if (ACC::kTableBytes == 0) return;
    const int off = 4 * p + 4;
    for (int r = threadIdx.x; r < 256; r += blockDim.x) {
        if constexpr (std::is_same<ACC, FgraAcc>::value) {
            d[r] = (r >= off && r < 252) ? c_ull.reg[r - off] : LASH_FGRA_SENTINEL;
        } else if constexpr (std::is_same<ACC, MlAcc>::value) {
            u[r] = ml_ret_of((uint32_t)r, p);
            w[r] = (uint32_t)ml_w_of((uint32_t)r, p);
        } else {
            d[r] = pow2neg((uint32_t)r);
        }
    }
}

// ---- exact per-pair fallbacks (read the two sketches from global memory, index order) -----------
// FGRA with small/large-range registers: counts + table sum, as get_distinct_count_estimate does
__device__ __noinline__ double fgra_exact_pair(const uint8_t* a, const uint8_t* b, int p) {
    const uint32_t m = 1u << p;
    const int off = 4 * p + 4;
    uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double sum = 0.0;
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t r = ull_merge1(a[i], b[i]);
        const int r2 = (int)r - off;
        if (r2 < 0) {
            cnt[0] += (r2 < -8);
            cnt[1] += (r2 == -8);
            cnt[2] += (r2 == -4);
            cnt[3] += (r2 == -2);
        } else if (r < 252u) {
            sum += c_ull.reg[r2];
        } else {
            cnt[4 + (r - 252u)] += 1;
        }
    }
    return ull_fgra_finalize(sum, cnt, p);
}
// ML statistics with direct increments (used when W does not fit 32 bits: astronomically large sketches)
__device__ __noinline__ double ml_exact_pair(const uint8_t* a, const uint8_t* b, int p) {
    const uint32_t m = 1u << p;
    int bb[66];
    for (int i = 0; i < 66; ++i) bb[i] = 0;
    uint64_t S = 0;
    for (uint32_t i = 0; i < m; ++i) {
        const uint32_t r = ull_merge1(a[i], b[i]);
        S += ml_ret_of(r, p);
        uint64_t w = ml_w_of(r, p);
        while (w) {
            const int j = __ffsll((long long)w) - 1;
            bb[j] += 1;
            w &= w - 1;
        }
    }
    return ull_ml_finalize(S, bb, p, ull_merge1(a[0], b[0]));
}

// union cardinality from a finished accumulator; ga/gb = the two sketches in global memory
__device__ __forceinline__ double finish_union(HllAcc& a, int p, const void*, const void*, bool* bias) {
    return hll_len(a.sum, a.zero, p, bias);
}
__device__ __forceinline__ double finish_union(FgraAcc& a, int p, const void* ga, const void* gb, bool* bias) {
    *bias = false;
    if (a.sum >= LASH_FGRA_SENTINEL_TEST)  // some merged register outside [4p+4, 252): NaN compares false
        return fgra_exact_pair(reinterpret_cast<const uint8_t*>(ga), reinterpret_cast<const uint8_t*>(gb), p);
    const uint32_t cnt[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    return ull_fgra_finalize(a.sum, cnt, p);
}
template <int NPL>
__device__ __forceinline__ double finish_union(MlAccT<NPL>& a, int p, const void* ga, const void* gb, bool* bias) {
    *bias = false;
    // W = (4|w) << k fits 32 bits iff k <= 29, i.e. merged register < 4p+4 + 4*30 (mmax >= 256: a G-sum tile's marker)
    if (a.mmax >= (uint32_t)(4 * p + 4 + 120) && a.mmax < kMlGsMarker)
        return ml_exact_pair(reinterpret_cast<const uint8_t*>(ga), reinterpret_cast<const uint8_t*>(gb), p);
    int bb[66];
    ml_counts_from_planes(a, bb);
    const uint8_t* pa = reinterpret_cast<const uint8_t*>(ga);
    const uint8_t* pb = reinterpret_cast<const uint8_t*>(gb);
    const uint64_t S = a.mmax >= kMlGsMarker ? ml_gs_S(a.S, bb, p, a.mmax & 0xffu) : a.S;   // G-sum tiles of dist_ml_tab_kernel
    return ull_ml_finalize(S, bb, p, ull_merge1(pa[0], pb[0]));
}

template <class ACC>
struct IsHmh { static constexpr bool value = false; };
template <>
struct IsHmh<HmhAcc> { static constexpr bool value = true; };

// the precomputed expected-collision loop sum of pair (i, j), when both sketches are small and have a stored term vector
__device__ __forceinline__ const double* hmh_ec_of(const DistParams& dp, uint64_t i, uint64_t j) {
    if (!dp.hmh_ec) return nullptr;
    const int32_t sr = dp.hmh_slot_ref[i - dp.hmh_row0], sq = dp.hmh_slot_qry[j];
    return (sr >= 0 && sq >= 0) ? dp.hmh_ec + (size_t)sr * dp.hmh_ec_ld + (size_t)sq : nullptr;
}

template <class ACC, int G>
__device__ __forceinline__ void acc_add(ACC& acc, const uint32_t* a, const uint32_t* b, const SharedTables& t, int nplanes) {
    if constexpr (std::is_same<ACC, MlAcc>::value)
        acc.template add_group<G>(a, b, t, nplanes);
    else
        acc.template add_group<G>(a, b, t);
}

// ------------------------------------------------------------------------------------------------
// K4: distance tiles
// ------------------------------------------------------------------------------------------------
template <class ACC, int G>
__global__ void __launch_bounds__(kDistThreads) dist_kernel(DistParams dp, uint32_t cell_bytes, uint32_t chunk_bytes) {
    using CT = typename ACC::CT;
    constexpr int RM = ACC::RM, QM = ACC::QM;
    constexpr int TR = 16 * RM, TQ = 16 * QM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables tabs;
    build_tables<ACC>(tabs, smem_raw, dp.p);
    unsigned char* sref = smem_raw + ACC::kTableBytes;
    const uint32_t stride = chunk_bytes + 4;  // one pad word per row: 16 rows -> 16 banks
    unsigned char* sqry = sref + TR * stride;
    const int nplanes = dp.p + 1;

    const uint64_t row0 = dp.row_begin + (uint64_t)blockIdx.y * TR;
    const uint64_t col0 = (uint64_t)blockIdx.x * TQ;
    if (row0 >= dp.row_end) return;
    const uint64_t row_hi = min(row0 + TR, dp.row_end);  // exclusive
    if (dp.triangular && col0 > row_hi - 1) return;      // tile entirely above the diagonal

    const int ty = threadIdx.x >> 4, tx = threadIdx.x & 15;
    ACC acc[RM][QM];
#pragma unroll
    for (int a = 0; a < RM; ++a)
#pragma unroll
        for (int b = 0; b < QM; ++b) acc[a][b].init();

    const unsigned char* gref = reinterpret_cast<const unsigned char*>(dp.ref);
    const unsigned char* gqry = reinterpret_cast<const unsigned char*>(dp.qry);
    const uint32_t chunk_words = chunk_bytes / 4;

    for (uint32_t c0 = 0; c0 < cell_bytes; c0 += chunk_bytes) {
        __syncthreads();  // previous chunk fully consumed (also orders the table build)
        for (uint32_t e = threadIdx.x; e < (uint32_t)TR * chunk_words; e += kDistThreads) {
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gi = row0 + r;
            const uint32_t v = gi < dp.row_end ? __ldg(reinterpret_cast<const uint32_t*>(gref + gi * cell_bytes + c0) + w) : 0u;
            *reinterpret_cast<uint32_t*>(sref + r * stride + 4 * w) = v;
        }
        for (uint32_t e = threadIdx.x; e < (uint32_t)TQ * chunk_words; e += kDistThreads) {
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gj = col0 + r;
            const uint32_t v = gj < dp.n_qry ? __ldg(reinterpret_cast<const uint32_t*>(gqry + gj * cell_bytes + c0) + w) : 0u;
            *reinterpret_cast<uint32_t*>(sqry + r * stride + 4 * w) = v;
        }
        __syncthreads();
        const CT* pr = reinterpret_cast<const CT*>(sref + ty * stride);
        const CT* pq = reinterpret_cast<const CT*>(sqry + tx * stride);
        const uint32_t row16 = 16u * stride / sizeof(CT);
        const uint32_t n_el = chunk_bytes / sizeof(CT);
#pragma unroll 1
        for (uint32_t e = 0; e < n_el; e += G) {
            uint32_t ra[RM][G], qb[QM][G];
#pragma unroll
            for (int a = 0; a < RM; ++a)
#pragma unroll
                for (int i = 0; i < G; ++i) ra[a][i] = pr[a * row16 + e + i];
#pragma unroll
            for (int b = 0; b < QM; ++b)
#pragma unroll
                for (int i = 0; i < G; ++i) qb[b][i] = pq[b * row16 + e + i];
#pragma unroll
            for (int a = 0; a < RM; ++a)
#pragma unroll
                for (int b = 0; b < QM; ++b) acc_add<ACC, G>(acc[a][b], ra[a], qb[b], tabs, nplanes);
        }
    }

    // epilogue: estimator -> Jaccard -> 2s/(1+s) -> Mash distance  (utils.rs:164-167,273-278,362-364)
#pragma unroll
    for (int a = 0; a < RM; ++a) {
#pragma unroll
        for (int b = 0; b < QM; ++b) {
            const uint64_t i = row0 + ty + 16 * a, j = col0 + tx + 16 * b;
            if (i >= dp.row_end || j >= dp.n_qry) continue;
            if (dp.triangular && j > i) continue;
            double s;
            if constexpr (IsHmh<ACC>::value) {
                double sim = hmh_similarity_from(acc[a][b].C, acc[a][b].N, dp.card_qry[j], dp.card_ref[i], hmh_ec_of(dp, i, j));
                s = fmax(sim, 0.0);
            } else {
                bool bias;
                const double U = finish_union(acc[a][b], dp.p, gref + i * cell_bytes, gqry + j * cell_bytes, &bias);
                if (bias && dp.flags) atomicAdd(dp.flags, 1u);
                const double ca = dp.card_ref[i], cb = dp.card_qry[j];
                const double sim = (ca + cb - U) / U;
                if constexpr (std::is_same<ACC, HllAcc>::value)
                    s = fmax(sim, 0.0);  // f64::max: NaN -> 0 (utils.rs:362)
                else
                    s = sim < 0.0 ? 0.0 : sim;  // utils.rs:274: NaN propagates
            }
            const double frac = 2.0 * s / (1.0 + s);
            const uint64_t o = dp.packed_tri ? (i * (i + 1) / 2 + j) : ((i - dp.out_row0) * dp.n_qry + j);
            if (dp.fp32)
                reinterpret_cast<float*>(dp.out)[o] = mash_distance_f32((float)frac, dp.k, dp.model);
            else
                reinterpret_cast<double*>(dp.out)[o] = mash_distance_f64(frac, dp.k, dp.model);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4h: HLL distance tiles without a table or a per-register zero test.
//
// dist_kernel<HllAcc> spends, per register pair, a byte extract, a max, a zero test + add, an LDS.64 of 2^-r and the
// DADD (ncu: shared-memory and issue bound, 2.0 T register pairs/s).  Here the registers are recoded ONCE, at staging
// time, into the high word of the double 2^-r  (v = 0x3FF00000 - (r << 20); the low word of a power of two is 0), so
//     2^-max(ra, rb) = hiloint2double(min(va, vb), 0)      -- one VIMNMX, then the DADD, per register pair,
// the same doubles added in the same (register index) order as dist_kernel / the scalar CPU loop -> bit-identical sums.
// The zero count (HLL++ linear counting, streaming_algorithms len()) only matters where BOTH sketches have an empty
// register at the same index; staging records, per chunk, whether any reference row and any query column of the tile
// holds an empty register at all, and only such chunks run the loop variant that also counts (v == 0x3FF00000).
// A warp owns 4 reference rows (warp-uniform -> broadcast LDS.128) x 64 query columns, two per lane.
// ------------------------------------------------------------------------------------------------
constexpr int kHllThreads = 256;
constexpr int kHllRM = 4, kHllQM = 2;
constexpr int kHllTR = (kHllThreads / 32) * kHllRM, kHllTQ = 32 * kHllQM;  // 32 x 64 pairs per CTA
constexpr int kHllChunk = 128;                                            // registers per sketch per stage

template <bool COUNT_ZERO>
__device__ __forceinline__ void hll_chunk(double (&sum)[kHllRM][kHllQM], uint32_t (&zero)[kHllRM][kHllQM], const uint32_t* pa,
                                          const uint32_t* pb, uint32_t a_row, uint32_t b_row32, uint32_t chunk) {
#pragma unroll 2
    for (uint32_t e = 0; e < chunk; e += 4) {
        uint4 a[kHllRM], b[kHllQM];
#pragma unroll
        for (int r = 0; r < kHllRM; ++r) a[r] = *reinterpret_cast<const uint4*>(pa + r * a_row + e);
#pragma unroll
        for (int c = 0; c < kHllQM; ++c) b[c] = *reinterpret_cast<const uint4*>(pb + c * b_row32 + e);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int r = 0; r < kHllRM; ++r) {
                const uint32_t av = j == 0 ? a[r].x : j == 1 ? a[r].y : j == 2 ? a[r].z : a[r].w;
#pragma unroll
                for (int c = 0; c < kHllQM; ++c) {
                    const uint32_t bv = j == 0 ? b[c].x : j == 1 ? b[c].y : j == 2 ? b[c].z : b[c].w;
                    const uint32_t m = min(av, bv);
                    if (COUNT_ZERO) zero[r][c] += (m == kHllOne);
                    sum[r][c] += __hiloint2double((int)m, 0);
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kHllThreads, 3) dist_hll_fast_kernel(DistParams dp, uint32_t cell_bytes, uint32_t chunk) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_zero[2][2];                    // [chunk parity][ref, qry]: an empty register was staged
    const uint32_t stride = chunk + 4;                   // u32 per staged row (+16 B pad: conflict-free LDS.128)
    uint32_t* sa = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* sb = sa + (size_t)kHllTR * stride;

    const uint64_t row0 = dp.row_begin + (uint64_t)blockIdx.y * kHllTR;
    const uint64_t col0 = (uint64_t)blockIdx.x * kHllTQ;
    if (row0 >= dp.row_end) return;
    const uint64_t row_hi = min(row0 + kHllTR, dp.row_end);
    if (dp.triangular && col0 > row_hi - 1) return;      // tile entirely above the diagonal

    const uint32_t wy = threadIdx.x >> 5, tx = threadIdx.x & 31u;
    const unsigned char* gref = reinterpret_cast<const unsigned char*>(dp.ref);
    const unsigned char* gqry = reinterpret_cast<const unsigned char*>(dp.qry);
    const uint32_t chunk_words = chunk / 4;                                                // powers of two, both
    const uint32_t g_shift = 31u - __clz(chunk_words) - 2u, cell_shift = 31u - __clz(cell_bytes);  // 16-register groups per row

    double sum[kHllRM][kHllQM];
    uint32_t zero[kHllRM][kHllQM];
#pragma unroll
    for (int r = 0; r < kHllRM; ++r)
#pragma unroll
        for (int c = 0; c < kHllQM; ++c) sum[r][c] = 0.0, zero[r][c] = 0u;
    if (threadIdx.x < 2) s_zero[0][threadIdx.x] = 0u;

    uint32_t par = 0;
    for (uint32_t c0 = 0; c0 < cell_bytes; c0 += chunk, par ^= 1u) {
        __syncthreads();  // previous chunk consumed; s_zero[par] was cleared during the previous staging pass (or above)
        if (threadIdx.x < 2) s_zero[par ^ 1u][threadIdx.x] = 0u;
        bool za = false, zb = false;
        // 16 registers per step: one LDG.128, four recoded STS.128 (register arrays are 16-byte aligned: the launcher checks)
        for (uint32_t e = threadIdx.x; e < ((uint32_t)kHllTR << g_shift); e += kHllThreads) {
            const uint32_t r = e >> g_shift, g = e & ((1u << g_shift) - 1u);
            const uint64_t gi = row0 + r;
            // rows past the end are staged as "register 255" (never empty, never read back)
            const uint4 v = gi < dp.row_end ? __ldg(reinterpret_cast<const uint4*>(gref + (gi << cell_shift) + c0) + g)
                                            : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            za |= has_zero_byte(v.x) | has_zero_byte(v.y) | has_zero_byte(v.z) | has_zero_byte(v.w);
            uint4* dst = reinterpret_cast<uint4*>(sa + r * stride + 16 * g);
            dst[0] = hll_recode(v.x);
            dst[1] = hll_recode(v.y);
            dst[2] = hll_recode(v.z);
            dst[3] = hll_recode(v.w);
        }
        for (uint32_t e = threadIdx.x; e < ((uint32_t)kHllTQ << g_shift); e += kHllThreads) {
            const uint32_t r = e >> g_shift, g = e & ((1u << g_shift) - 1u);
            const uint64_t gj = col0 + r;
            const uint4 v = gj < dp.n_qry ? __ldg(reinterpret_cast<const uint4*>(gqry + (gj << cell_shift) + c0) + g)
                                          : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            zb |= has_zero_byte(v.x) | has_zero_byte(v.y) | has_zero_byte(v.z) | has_zero_byte(v.w);
            uint4* dst = reinterpret_cast<uint4*>(sb + r * stride + 16 * g);
            dst[0] = hll_recode(v.x);
            dst[1] = hll_recode(v.y);
            dst[2] = hll_recode(v.z);
            dst[3] = hll_recode(v.w);
        }
        if (za) s_zero[par][0] = 1u;
        if (zb) s_zero[par][1] = 1u;
        __syncthreads();
        const uint32_t* pa = sa + (wy * kHllRM) * stride;
        const uint32_t* pb = sb + tx * stride;
        if (s_zero[par][0] & s_zero[par][1])  // CTA-uniform
            hll_chunk<true>(sum, zero, pa, pb, stride, 32u * stride, chunk);
        else
            hll_chunk<false>(sum, zero, pa, pb, stride, 32u * stride, chunk);
    }

    // epilogue: identical to dist_kernel<HllAcc>
#pragma unroll
    for (int a = 0; a < kHllRM; ++a) {
#pragma unroll
        for (int b = 0; b < kHllQM; ++b) {
            const uint64_t i = row0 + wy * kHllRM + a, j = col0 + tx + 32 * b;
            if (i >= dp.row_end || j >= dp.n_qry) continue;
            if (dp.triangular && j > i) continue;
            bool bias;
            const double U = hll_len(sum[a][b], zero[a][b], dp.p, &bias);
            if (bias && dp.flags) atomicAdd(dp.flags, 1u);
            const double ca = dp.card_ref[i], cb = dp.card_qry[j];
            const double sim = (ca + cb - U) / U;
            const double s = fmax(sim, 0.0);  // f64::max: NaN -> 0 (utils.rs:362)
            const double frac = 2.0 * s / (1.0 + s);
            const uint64_t o = dp.packed_tri ? (i * (i + 1) / 2 + j) : ((i - dp.out_row0) * dp.n_qry + j);
            if (dp.fp32)
                reinterpret_cast<float*>(dp.out)[o] = mash_distance_f32((float)frac, dp.k, dp.model);
            else
                reinterpret_cast<double*>(dp.out)[o] = mash_distance_f64(frac, dp.k, dp.model);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4i: HLL distance tiles in 32-bit fixed point (dist_tables.cuh: hll_int_recode).
//
// K4h pays a VIMNMX, a DADD and (ptxas) 0.6 MOV per register pair: 3.5 issue slots, 7.3 T register pairs/s.  The sum of a
// pair whose registers all lie within 29 levels of the tile's smallest register is exact in f64 -- no partial sum of the
// reference's loop ever rounds -- so it can be formed in ANY arithmetic: here min (VIMNMX, ALU pipe) + a 32-bit
// multiply-add by a run-time 1 (IMAD, FMA-heavy pipe; with a literal 1 ptxas turns it back into an ALU add), eight terms
// per 32-bit batch, one IMAD.WIDE per batch into the 64-bit sum: 2.1 issue slots per register pair split evenly over the
// two integer pipes.  Per tile: lo = min over the tile's sketches of their smallest register (zero registers included:
// then lo = 0 and an empty register is the term 2^28).  A tile that holds a sketch whose largest register exceeds lo + 28
// (6 * 10^-4 of the sketches of 5 Mbp genomes at p = 14: 7 % of the tiles; every tile of a set that mixes tiny and huge
// genomes) runs K4h's arithmetic instead -- registers recoded to the high word of 2^-r, VIMNMX + DADD in register order --
// in the same staged layout: 1.5x slower than a fixed-point tile, never slower than K4h.  (The first version redid the
// flagged pairs one by one from global memory: a flagged query column cost one lane per warp 8 x 2^p dependent adds, 0.5 ms
// per tile, a cliff on mixed sets.)  The zero count is only needed where both sketches hold an empty register, as in K4h.
// ------------------------------------------------------------------------------------------------
__global__ void hll_minmax_kernel(const unsigned char* __restrict__ regs, uint64_t n, uint32_t cell_bytes, uint32_t* __restrict__ mm) {
    const uint64_t s = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (s >= n) return;
    const uint32_t lane = threadIdx.x & 31u;
    const uint4* src = reinterpret_cast<const uint4*>(regs + s * cell_bytes);
    uint32_t mn = 0xffffffffu, mx = 0u;
    for (uint32_t e = lane; e < cell_bytes / 16; e += 32) {
        const uint4 v = __ldg(src + e);
        mn = __vminu4(__vminu4(mn, v.x), __vminu4(__vminu4(v.y, v.z), v.w));
        mx = __vmaxu4(__vmaxu4(mx, v.x), __vmaxu4(__vmaxu4(v.y, v.z), v.w));
    }
    mn = __vminu4(mn, mn >> 16), mn = __vminu4(mn, mn >> 8) & 0xffu;
    mx = __vmaxu4(mx, mx >> 16), mx = __vmaxu4(mx, mx >> 8) & 0xffu;
    mn = __reduce_min_sync(0xffffffffu, mn);
    mx = __reduce_max_sync(0xffffffffu, mx);
    if (lane == 0) mm[s] = mn | (mx << 8);
}
cudaError_t launch_hll_minmax(const void* regs, uint64_t n, uint32_t cell_bytes, uint32_t* mm, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    hll_minmax_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>(reinterpret_cast<const unsigned char*>(regs), n, cell_bytes, mm);
    return cudaGetLastError();
}

__device__ __forceinline__ uint32_t mad_one(uint32_t a, uint32_t one, uint32_t c) {
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(one), "r"(c));
    return d;
}

// Staged rows are laid out in 16-register groups of kHllIntGroup = 20 words (16 terms + 4 words of padding) and a row stride of
// groups * 20 + 4 words: a staging thread recodes one LDG.128 (16 registers) into four consecutive STS.128, and with 16-word
// groups the eight threads of a row hit only two bank quads (ncu: 4-way conflicts, the shared-memory pipe at 79 %); with
// 20-word groups they cover all eight.  The query rows of a warp's lanes still start 4 banks apart (stride = 4 mod 32).
// STRIDE: u32 per staged row at compile time (the usual 128-register chunk: every LDS offset is an immediate), 0 = run time
// Two tile shapes: a thread owns RM reference rows x 2 query columns.  RM = 8 (64 x 64 pairs per CTA, 128 registers, two CTAs per
// SM) stages 1/32 sketch row per pair instead of 3/64 and reads fewer shared-memory wavefronts per pair: +7 % at n = 6000;
// RM = 4 (32 x 64, 80 registers, three CTAs per SM) has half-size tiles for grids of only a few waves.
constexpr int kHiQM = 2;
constexpr int kHiTQ = 32 * kHiQM;
constexpr int kHiChunk = 128;                                             // registers per sketch per stage
template <int RM>
struct HiShape {
    static constexpr int kTR = (kHllThreads / 32) * RM;                   // reference rows per CTA
    static constexpr int kMinBlocks = RM == 8 ? 2 : 3;
};
constexpr uint32_t kHllIntGroup = 20;
__host__ __device__ constexpr uint32_t hll_int_stride(uint32_t chunk) { return chunk / 16 * kHllIntGroup + 4; }

template <bool COUNT_ZERO, int STRIDE, int kHiRM>
__device__ __forceinline__ void hll_int_chunk(uint64_t (&sum)[kHiRM][kHiQM], uint32_t (&zero)[kHiRM][kHiQM], const uint32_t* pa,
                                              const uint32_t* pb, uint32_t stride_rt, uint32_t chunk, uint32_t one) {
    static_assert(kHllIntBatch == 8, "two 4-register steps per 32-bit batch");
    const uint32_t a_row = STRIDE ? (uint32_t)STRIDE : stride_rt, b_row32 = 32u * a_row;
    const uint32_t n_groups = chunk / 16;
#pragma unroll 1
    for (uint32_t g = 0; g < n_groups; ++g) {
        const uint32_t* ga = pa + g * kHllIntGroup;
        const uint32_t* gb = pb + g * kHllIntGroup;
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            uint32_t acc[kHiRM][kHiQM];
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint4 a[kHiRM], b[kHiQM];
#pragma unroll
                for (int r = 0; r < kHiRM; ++r) a[r] = *reinterpret_cast<const uint4*>(ga + r * a_row + 8 * half + 4 * h);
#pragma unroll
                for (int c = 0; c < kHiQM; ++c) b[c] = *reinterpret_cast<const uint4*>(gb + c * b_row32 + 8 * half + 4 * h);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
#pragma unroll
                    for (int r = 0; r < kHiRM; ++r) {
                        const uint32_t av = j == 0 ? a[r].x : j == 1 ? a[r].y : j == 2 ? a[r].z : a[r].w;
#pragma unroll
                        for (int c = 0; c < kHiQM; ++c) {
                            const uint32_t bv = j == 0 ? b[c].x : j == 1 ? b[c].y : j == 2 ? b[c].z : b[c].w;
                            const uint32_t m = min(av, bv);
                            if (COUNT_ZERO) zero[r][c] += m >> kHllIntW;            // lo == 0 here: 2^28 <=> both registers empty
                            acc[r][c] = (h == 0 && j == 0) ? m : mad_one(m, one, acc[r][c]);
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < kHiRM; ++r)
#pragma unroll
                for (int c = 0; c < kHiQM; ++c) {
                    sum[r][c] += (uint64_t)acc[r][c];
                }
        }
    }
}

// the same tile in K4h's arithmetic (staged words = high words of 2^-r): f64 adds in register order; the 64-bit accumulators
// hold the doubles' bit patterns
template <bool COUNT_ZERO, int kHiRM>
__device__ __forceinline__ void hll_flt_chunk(uint64_t (&sum)[kHiRM][kHiQM], uint32_t (&zero)[kHiRM][kHiQM], const uint32_t* pa,
                                              const uint32_t* pb, uint32_t a_row, uint32_t chunk) {
    const uint32_t b_row32 = 32u * a_row, n_groups = chunk / 16;
    double d[kHiRM][kHiQM];
#pragma unroll
    for (int r = 0; r < kHiRM; ++r)
#pragma unroll
        for (int c = 0; c < kHiQM; ++c) d[r][c] = __longlong_as_double((long long)sum[r][c]);
#pragma unroll 1
    for (uint32_t g = 0; g < n_groups; ++g) {
#pragma unroll 1
        for (uint32_t q = 0; q < 16; q += 4) {
            uint4 a[kHiRM], b[kHiQM];
#pragma unroll
            for (int r = 0; r < kHiRM; ++r) a[r] = *reinterpret_cast<const uint4*>(pa + g * kHllIntGroup + r * a_row + q);
#pragma unroll
            for (int c = 0; c < kHiQM; ++c) b[c] = *reinterpret_cast<const uint4*>(pb + g * kHllIntGroup + c * b_row32 + q);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
#pragma unroll
                for (int r = 0; r < kHiRM; ++r) {
                    const uint32_t av = j == 0 ? a[r].x : j == 1 ? a[r].y : j == 2 ? a[r].z : a[r].w;
#pragma unroll
                    for (int c = 0; c < kHiQM; ++c) {
                        const uint32_t bv = j == 0 ? b[c].x : j == 1 ? b[c].y : j == 2 ? b[c].z : b[c].w;
                        const uint32_t m = min(av, bv);
                        if (COUNT_ZERO) zero[r][c] += (m == kHllOne);
                        d[r][c] += __hiloint2double((int)m, 0);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < kHiRM; ++r)
#pragma unroll
        for (int c = 0; c < kHiQM; ++c) sum[r][c] = (uint64_t)__double_as_longlong(d[r][c]);
}

template <int kHiRM>
__global__ void __launch_bounds__(kHllThreads, HiShape<kHiRM>::kMinBlocks) dist_hll_int_kernel(DistParams dp, uint32_t cell_bytes, uint32_t chunk, uint32_t one) {
    constexpr int kHiTR = HiShape<kHiRM>::kTR;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint32_t s_zero[2][2];                    // [chunk parity][ref, qry]: an empty register was staged
    __shared__ uint32_t s_lo;
    const uint32_t stride = hll_int_stride(chunk);        // u32 per staged row (20-word groups + 16 B pad)
    uint32_t* sa = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* sb = sa + (size_t)kHiTR * stride;

    const uint64_t row0 = dp.row_begin + (uint64_t)blockIdx.y * kHiTR;
    const uint64_t col0 = (uint64_t)blockIdx.x * kHiTQ;
    if (row0 >= dp.row_end) return;
    const uint64_t row_hi = min(row0 + kHiTR, dp.row_end);
    if (dp.triangular && col0 > row_hi - 1) return;      // tile entirely above the diagonal

    const uint32_t wy = threadIdx.x >> 5, tx = threadIdx.x & 31u;
    const unsigned char* gref = reinterpret_cast<const unsigned char*>(dp.ref);
    const unsigned char* gqry = reinterpret_cast<const unsigned char*>(dp.qry);
    const uint32_t chunk_words = chunk / 4;
    const uint32_t g_shift = 31u - __clz(chunk_words) - 2u, cell_shift = 31u - __clz(cell_bytes);

    // the tile's window: lo = smallest register of its sketches; flag the sketches that reach above lo + 28
    if (threadIdx.x == 0) s_lo = 0xffu;
    if (threadIdx.x < 2) s_zero[0][threadIdx.x] = 0u;
    __syncthreads();
    uint32_t my_mm = 0xff00ffu;   // min 255, max 0 ... rows past the end: never flagged, never lower the minimum
    if (threadIdx.x < kHiTR) {
        if (row0 + threadIdx.x < dp.row_end) my_mm = dp.reg_mm_ref[row0 + threadIdx.x];
    } else if (threadIdx.x < kHiTR + kHiTQ) {
        if (col0 + (threadIdx.x - kHiTR) < dp.n_qry) my_mm = dp.reg_mm_qry[col0 + (threadIdx.x - kHiTR)];
    }
    if (threadIdx.x < kHiTR + kHiTQ) {
        const uint32_t wmin = __reduce_min_sync(0xffffffffu, my_mm & 0xffu);
        if (tx == 0) atomicMin(&s_lo, wmin);
    }
    __syncthreads();
    const uint32_t lo = s_lo;
    // a sketch above the window: the whole tile runs in f64 (CTA-uniform)
    const bool flt = __syncthreads_or(threadIdx.x < kHiTR + kHiTQ && ((my_mm >> 8) & 0xffu) > lo + (uint32_t)kHllIntW) != 0;

    uint64_t sum[kHiRM][kHiQM];
    uint32_t zero[kHiRM][kHiQM];
#pragma unroll
    for (int r = 0; r < kHiRM; ++r)
#pragma unroll
        for (int c = 0; c < kHiQM; ++c) sum[r][c] = 0ull, zero[r][c] = 0u;

    uint32_t par = 0;
    for (uint32_t c0 = 0; c0 < cell_bytes; c0 += chunk, par ^= 1u) {
        __syncthreads();  // previous chunk consumed; s_zero[par] was cleared during the previous staging pass (or above)
        if (threadIdx.x < 2) s_zero[par ^ 1u][threadIdx.x] = 0u;
        bool za = false, zb = false;
        for (uint32_t e = threadIdx.x; e < ((uint32_t)kHiTR << g_shift); e += kHllThreads) {
            const uint32_t r = e >> g_shift, g = e & ((1u << g_shift) - 1u);
            const uint64_t gi = row0 + r;
            const uint4 v = gi < dp.row_end ? __ldg(reinterpret_cast<const uint4*>(gref + (gi << cell_shift) + c0) + g)
                                            : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            za |= has_zero_byte(v.x) | has_zero_byte(v.y) | has_zero_byte(v.z) | has_zero_byte(v.w);
            uint4* dst = reinterpret_cast<uint4*>(sa + r * stride + kHllIntGroup * g);
            dst[0] = flt ? hll_recode(v.x) : hll_int_recode(v.x, lo);
            dst[1] = flt ? hll_recode(v.y) : hll_int_recode(v.y, lo);
            dst[2] = flt ? hll_recode(v.z) : hll_int_recode(v.z, lo);
            dst[3] = flt ? hll_recode(v.w) : hll_int_recode(v.w, lo);
        }
        for (uint32_t e = threadIdx.x; e < ((uint32_t)kHiTQ << g_shift); e += kHllThreads) {
            const uint32_t r = e >> g_shift, g = e & ((1u << g_shift) - 1u);
            const uint64_t gj = col0 + r;
            const uint4 v = gj < dp.n_qry ? __ldg(reinterpret_cast<const uint4*>(gqry + (gj << cell_shift) + c0) + g)
                                          : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            zb |= has_zero_byte(v.x) | has_zero_byte(v.y) | has_zero_byte(v.z) | has_zero_byte(v.w);
            uint4* dst = reinterpret_cast<uint4*>(sb + r * stride + kHllIntGroup * g);
            dst[0] = flt ? hll_recode(v.x) : hll_int_recode(v.x, lo);
            dst[1] = flt ? hll_recode(v.y) : hll_int_recode(v.y, lo);
            dst[2] = flt ? hll_recode(v.z) : hll_int_recode(v.z, lo);
            dst[3] = flt ? hll_recode(v.w) : hll_int_recode(v.w, lo);
        }
        if (za) s_zero[par][0] = 1u;
        if (zb) s_zero[par][1] = 1u;
        __syncthreads();
        const uint32_t* pa = sa + (wy * kHiRM) * stride;
        const uint32_t* pb = sb + tx * stride;
        if (flt) {
            if (s_zero[par][0] & s_zero[par][1])
                hll_flt_chunk<true, kHiRM>(sum, zero, pa, pb, stride, chunk);
            else
                hll_flt_chunk<false, kHiRM>(sum, zero, pa, pb, stride, chunk);
        } else if (s_zero[par][0] & s_zero[par][1])  // CTA-uniform; an empty register anywhere means lo == 0
            hll_int_chunk<true, 0, kHiRM>(sum, zero, pa, pb, stride, chunk, one);
        else if (chunk == (uint32_t)kHiChunk)
            hll_int_chunk<false, (int)hll_int_stride(kHiChunk), kHiRM>(sum, zero, pa, pb, stride, chunk, one);
        else
            hll_int_chunk<false, 0, kHiRM>(sum, zero, pa, pb, stride, chunk, one);
    }

    // epilogue: identical to dist_kernel<HllAcc> once the sum is back in f64 (exact: < 2^53, times a power of two)
    const double scale = __hiloint2double((int)((1023u - (uint32_t)kHllIntW - lo) << 20), 0);
#pragma unroll
    for (int a = 0; a < kHiRM; ++a) {
#pragma unroll
        for (int b = 0; b < kHiQM; ++b) {
            const uint64_t i = row0 + wy * kHiRM + a, j = col0 + tx + 32 * b;
            if (i >= dp.row_end || j >= dp.n_qry) continue;
            if (dp.triangular && j > i) continue;
            const double sm = flt ? __longlong_as_double((long long)sum[a][b]) : (double)sum[a][b] * scale;
            bool bias;
            const double U = hll_len(sm, zero[a][b], dp.p, &bias);
            if (bias && dp.flags) atomicAdd(dp.flags, 1u);
            const double ca = dp.card_ref[i], cb = dp.card_qry[j];
            const double sim = (ca + cb - U) / U;
            const double s = fmax(sim, 0.0);  // f64::max: NaN -> 0 (utils.rs:362)
            const double frac = 2.0 * s / (1.0 + s);
            const uint64_t o = dp.packed_tri ? (i * (i + 1) / 2 + j) : ((i - dp.out_row0) * dp.n_qry + j);
            if (dp.fp32)
                reinterpret_cast<float*>(dp.out)[o] = mash_distance_f32((float)frac, dp.k, dp.model);
            else
                reinterpret_cast<double*>(dp.out)[o] = mash_distance_f64(frac, dp.k, dp.model);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4m: HyperMinHash distance tiles (hmh_distance inner loop, utils.rs:150-180; hyperminhash similarity()).
//
// Per pair, over the 16384 u16 registers:  C = #{a == b and a != 0},  N = #{a != 0 or b != 0}  -- integers, so any
// evaluation order is exact.  dist_kernel<HmhAcc> spends ~8 scalar instructions per register on halfword extracts and
// compares.  Here two registers are handled per 32-bit word with the packed-halfword minimum of sm_90+ (VIMNMX.U16x2):
//     m = min.u16x2(a ^ b, 0x00010001)          each 16-bit lane: 1 <=> the registers differ
//     nz = m * one + nz                         both lane counters at once, on the FMA pipe (`one` is a run-time 1:
//                                               with a literal ptxas turns the IMAD back into an ALU add)
// = LOP3 + VIMNMX + IMAD per two register pairs, and  C = N - NZ.  N itself only deviates from "all registers" where
// BOTH sketches hold an empty register at the same index; staging records per chunk whether any reference row and any
// query column has an empty register at all, and only such chunks run the loop variant that also counts
// min.u16x2(a | b, 0x00010001) (sketches of >= ~10^5 k-mers never do).
// A warp owns 4 reference rows (warp-uniform -> broadcast LDS.128) x 64 query columns, two per lane; 3 CTAs per SM.
// ------------------------------------------------------------------------------------------------
constexpr int kHmhThreads = 256;
constexpr int kHmhRM = 4, kHmhQM = 2;
constexpr int kHmhTR = (kHmhThreads / 32) * kHmhRM, kHmhTQ = 32 * kHmhQM;  // 32 x 64 pairs per CTA
constexpr int kHmhChunkWords = 64;                                        // 128 registers per sketch per stage
constexpr uint32_t kHmhWords = 8192;                                      // 16384 u16 registers

__device__ __forceinline__ bool has_empty_halfword(uint32_t x) { return __vminu2(x, 0x00010001u) != 0x00010001u; }

template <bool COUNT_N>
__device__ __forceinline__ void hmh_chunk(uint32_t (&nz)[kHmhRM][kHmhQM], uint32_t (&nn)[kHmhRM][kHmhQM], const uint32_t* pa,
                                          const uint32_t* pb, uint32_t a_row, uint32_t b_row32, uint32_t one) {
#pragma unroll 2
    for (uint32_t e = 0; e < (uint32_t)kHmhChunkWords; e += 4) {
        uint4 a[kHmhRM], b[kHmhQM];
#pragma unroll
        for (int r = 0; r < kHmhRM; ++r) a[r] = *reinterpret_cast<const uint4*>(pa + r * a_row + e);
#pragma unroll
        for (int c = 0; c < kHmhQM; ++c) b[c] = *reinterpret_cast<const uint4*>(pb + c * b_row32 + e);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
            for (int r = 0; r < kHmhRM; ++r) {
                const uint32_t av = j == 0 ? a[r].x : j == 1 ? a[r].y : j == 2 ? a[r].z : a[r].w;
#pragma unroll
                for (int c = 0; c < kHmhQM; ++c) {
                    const uint32_t bv = j == 0 ? b[c].x : j == 1 ? b[c].y : j == 2 ? b[c].z : b[c].w;
                    nz[r][c] = __vminu2(av ^ bv, 0x00010001u) * one + nz[r][c];
                    if (COUNT_N) nn[r][c] = __vminu2(av | bv, 0x00010001u) * one + nn[r][c];
                }
            }
        }
    }
}

__global__ void __launch_bounds__(kHmhThreads, 3) dist_hmh_fast_kernel(DistParams dp) {
    __shared__ __align__(16) uint32_t sa[kHmhTR * (kHmhChunkWords + 4)];
    __shared__ __align__(16) uint32_t sb[kHmhTQ * (kHmhChunkWords + 4)];
    __shared__ uint32_t s_zero[2][2];                    // [chunk parity][ref, qry]: an empty register was staged
    constexpr uint32_t stride = kHmhChunkWords + 4;      // +16 B pad: conflict-free LDS.128 across the 32 query rows of a warp

    const uint64_t row0 = dp.row_begin + (uint64_t)blockIdx.y * kHmhTR;
    const uint64_t col0 = (uint64_t)blockIdx.x * kHmhTQ;
    if (row0 >= dp.row_end) return;
    const uint64_t row_hi = min(row0 + kHmhTR, dp.row_end);
    if (dp.triangular && col0 > row_hi - 1) return;      // tile entirely above the diagonal

    const uint32_t wy = threadIdx.x >> 5, tx = threadIdx.x & 31u;
    const uint32_t* gref = reinterpret_cast<const uint32_t*>(dp.ref);
    const uint32_t* gqry = reinterpret_cast<const uint32_t*>(dp.qry);
    const uint32_t one = blockDim.x / kHmhThreads;       // 1, opaque to ptxas (see above)

    // per-lane pair counters: two 16-bit lanes per word (low / high halfword of the register words); a lane counts at
    // most 8192, so neither overflows
    uint32_t nz[kHmhRM][kHmhQM], nn[kHmhRM][kHmhQM];
    uint32_t full = 0;                                   // registers of chunks where N needed no counting (same for every pair)
#pragma unroll
    for (int r = 0; r < kHmhRM; ++r)
#pragma unroll
        for (int c = 0; c < kHmhQM; ++c) nz[r][c] = nn[r][c] = 0u;
    if (threadIdx.x < 2) s_zero[0][threadIdx.x] = 0u;

    uint32_t par = 0;
    for (uint32_t c0 = 0; c0 < kHmhWords; c0 += kHmhChunkWords, par ^= 1u) {
        __syncthreads();  // previous chunk consumed; s_zero[par] was cleared during the previous staging pass (or above)
        if (threadIdx.x < 2) s_zero[par ^ 1u][threadIdx.x] = 0u;
        bool za = false, zb = false;
        // 16-byte loads: 16 per staged row (register arrays are 16-byte aligned: the launcher checks)
        for (uint32_t e = threadIdx.x; e < (uint32_t)kHmhTR * (kHmhChunkWords / 4); e += kHmhThreads) {
            const uint32_t r = e / (kHmhChunkWords / 4), g = e % (kHmhChunkWords / 4);
            const uint64_t gi = row0 + r;
            // rows past the end are staged as all-ones registers (never empty, never read back)
            const uint4 v = gi < dp.row_end ? __ldg(reinterpret_cast<const uint4*>(gref + gi * kHmhWords + c0) + g)
                                            : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            za |= has_empty_halfword(v.x) | has_empty_halfword(v.y) | has_empty_halfword(v.z) | has_empty_halfword(v.w);
            *reinterpret_cast<uint4*>(sa + r * stride + 4 * g) = v;
        }
        for (uint32_t e = threadIdx.x; e < (uint32_t)kHmhTQ * (kHmhChunkWords / 4); e += kHmhThreads) {
            const uint32_t r = e / (kHmhChunkWords / 4), g = e % (kHmhChunkWords / 4);
            const uint64_t gj = col0 + r;
            const uint4 v = gj < dp.n_qry ? __ldg(reinterpret_cast<const uint4*>(gqry + gj * kHmhWords + c0) + g)
                                          : make_uint4(0xffffffffu, 0xffffffffu, 0xffffffffu, 0xffffffffu);
            zb |= has_empty_halfword(v.x) | has_empty_halfword(v.y) | has_empty_halfword(v.z) | has_empty_halfword(v.w);
            *reinterpret_cast<uint4*>(sb + r * stride + 4 * g) = v;
        }
        if (za) s_zero[par][0] = 1u;
        if (zb) s_zero[par][1] = 1u;
        __syncthreads();
        const uint32_t* pa = sa + (wy * kHmhRM) * stride;
        const uint32_t* pb = sb + tx * stride;
        if (s_zero[par][0] & s_zero[par][1]) {  // CTA-uniform
            hmh_chunk<true>(nz, nn, pa, pb, stride, 32u * stride, one);
        } else {
            hmh_chunk<false>(nz, nn, pa, pb, stride, 32u * stride, one);
            full += 2u * kHmhChunkWords;
        }
    }

    // epilogue: identical to dist_kernel<HmhAcc>
#pragma unroll
    for (int a = 0; a < kHmhRM; ++a) {
#pragma unroll
        for (int b = 0; b < kHmhQM; ++b) {
            const uint64_t i = row0 + wy * kHmhRM + a, j = col0 + tx + 32 * b;
            if (i >= dp.row_end || j >= dp.n_qry) continue;
            if (dp.triangular && j > i) continue;
            const uint32_t NZ = (nz[a][b] & 0xffffu) + (nz[a][b] >> 16);
            const uint32_t N = (nn[a][b] & 0xffffu) + (nn[a][b] >> 16) + full;
            const uint32_t Cc = N - NZ;   // equal and non-empty = (equal) - (both empty) = (16384 - NZ) - (16384 - N)
            const double sim = hmh_similarity_from(Cc, N, dp.card_qry[j], dp.card_ref[i], hmh_ec_of(dp, i, j));
            const double s = fmax(sim, 0.0);
            const double frac = 2.0 * s / (1.0 + s);
            const uint64_t o = dp.packed_tri ? (i * (i + 1) / 2 + j) : ((i - dp.out_row0) * dp.n_qry + j);
            if (dp.fp32)
                reinterpret_cast<float*>(dp.out)[o] = mash_distance_f32((float)frac, dp.k, dp.model);
            else
                reinterpret_cast<double*>(dp.out)[o] = mash_distance_f64(frac, dp.k, dp.model);
        }
    }
}

// ------------------------------------------------------------------------------------------------
// K4b: FGRA distance tiles through a PAIR table.
//
// dist_kernel<FgraAcc> spends ~12 ALU instructions per register pair on the packed-domain merge
// (ncu: ALU pipe 91 % busy, everything else idle).  Here the merge is looked up instead: register bytes
// are recoded at staging time to c = 0 (empty) or r - (4p-4) + 1 (1..126; 127 = "outside the table"),
// and a CTA-private table T[ca][cb] = REGISTER_CONTRIBUTIONS[merge(ra, rb)] (128 x 128 doubles =
// 128 KiB of shared memory) turns one register pair into  address add + LDS.64 + DADD.  Entries whose
// merged register needs FGRA's small/large-range treatment, and code 127, hold the same 2^600 sentinel
// as dist_kernel, so those pairs are redone by the exact per-pair path; every other pair adds exactly the
// same doubles in exactly the same (register index) order as before -- results are bit-identical.
// A warp owns 2 reference rows (uniform across lanes -> shared loads broadcast) x 64 query columns.
// ------------------------------------------------------------------------------------------------
constexpr int kTabN = 128;
constexpr int kTabThreads = 512;
constexpr int kTabTR = 32, kTabTQ = 64;   // pairs tile of a CTA
constexpr int kTabChunk = 256;            // registers per sketch per stage


// one table entry: low word from the first plane, high word from the second (two conflict-free LDS.32)
__device__ __forceinline__ double tab_lookup(uint32_t saddr, uint32_t plane_bytes) {
    uint32_t lo, hi;
    asm("ld.shared.u32 %0, [%1];" : "=r"(lo) : "r"(saddr));
    asm("ld.shared.u32 %0, [%1];" : "=r"(hi) : "r"(saddr + plane_bytes));
    return __hiloint2double((int)hi, (int)lo);
}

// Persistent CTAs (one per SM: the table fills most of its shared memory) fetch tiles from a global counter, so the
// table is built once per SM instead of once per tile and tiles above the diagonal cost one atomic.
__device__ __forceinline__ bool next_tile(const DistParams& dp, uint32_t* s_tile, uint64_t n_tiles, uint64_t iter, uint64_t& t) {
    __syncthreads();  // the previous tile is finished everywhere (its staging buffers and flags may be reused)
    if (threadIdx.x == 0) *s_tile = dp.tile_counter ? atomicAdd(dp.tile_counter, 1u) : (uint32_t)(blockIdx.x + iter * gridDim.x);
    __syncthreads();
    t = *s_tile;
    return t < n_tiles;
}

__global__ void __launch_bounds__(kTabThreads, 1) dist_fgra_tab_kernel(DistParams dp, uint32_t cell_bytes, uint32_t chunk,
                                                                       uint32_t tiles_x, uint32_t tiles_y) {
    __shared__ uint32_t s_tile;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // The table is stored as two planes of 32-bit words (low halves, then high halves) and read with two LDS.32: a warp
    // reads ONE table row (the reference register is warp-uniform), so with 4-byte entries the bank is the query code
    // mod 32 and the ~28 codes a sketch set actually uses (7-8 register levels x 4) never collide.  As one plane of
    // doubles (LDS.64: 16 banks per half-warp) codes 16 apart -- same sub-bits, 4 levels apart -- collided: 12 % extra
    // wavefronts on sketches of related genomes, ~80 % on unrelated ones (lanes then hold many distinct codes).
    uint32_t* T = reinterpret_cast<uint32_t*>(smem_raw);
    constexpr uint32_t kPlane = (uint32_t)kTabN * kTabN * 4u;  // bytes per plane
    const uint32_t a_stride = chunk + 4;                 // u32 elements per reference row (+16 B pad)
    const uint32_t b_stride = chunk + 8;                 // u16 elements per query row (+16 B pad)
    uint32_t* sa = reinterpret_cast<uint32_t*>(smem_raw + (size_t)kTabN * kTabN * 8);
    uint16_t* sb = reinterpret_cast<uint16_t*>(sa + (size_t)kTabTR * a_stride);

    // ---- pair table -------------------------------------------------------------------------------
    // The 126 register values the table covers start at the smallest non-empty register of the two sets
    // (regmin_kernel), not at the smallest possible one: sketches of large inputs (10^11 k-mers at p = 14)
    // have no register anywhere near 4p-4 but plenty above 4p-4+126, and would all take the exact path.
    const int p = dp.p;
    const uint32_t off = (uint32_t)(4 * p + 4);
    uint32_t base = (uint32_t)(4 * p - 4);
    if (dp.regmin) {
        const uint32_t mn = __ldg(dp.regmin);
        if (mn != 0xffffffffu) base = max(base, mn & ~3u);
    }
    for (uint32_t e = threadIdx.x; e < (uint32_t)(kTabN * kTabN); e += kTabThreads) {
        const double v = fgra_tab_entry(e >> 7, e & 127u, base, off, c_ull.reg);
        T[e] = (uint32_t)__double2loint(v);
        T[e + kTabN * kTabN] = (uint32_t)__double2hiint(v);
    }

    const uint32_t ty = threadIdx.x >> 5, tx = threadIdx.x & 31u;  // ty is warp-uniform
    const unsigned char* gref = reinterpret_cast<const unsigned char*>(dp.ref);
    const unsigned char* gqry = reinterpret_cast<const unsigned char*>(dp.qry);
    const uint32_t chunk_words = chunk / 4;
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(T);
    const uint64_t n_tiles = (uint64_t)tiles_x * tiles_y;

    uint64_t tile;
    for (uint64_t iter = 0; next_tile(dp, &s_tile, n_tiles, iter, tile); ++iter) {
    const uint64_t row0 = dp.row_begin + (tile / tiles_x) * kTabTR;
    const uint64_t col0 = (tile % tiles_x) * kTabTQ;
    const uint64_t row_hi = min(row0 + kTabTR, dp.row_end);
    if (dp.triangular && col0 > row_hi - 1) continue;  // tile entirely above the diagonal (CTA-uniform)
    double acc[2][2] = {{0.0, 0.0}, {0.0, 0.0}};

    // The registers of the NEXT chunk are fetched into registers before the current chunk is consumed, so the global-load
    // latency hides behind ~17 us of table lookups instead of standing between two barriers (one CTA per SM: nothing else
    // would cover it).  kPreA / kPreB words per thread at the largest chunk (256 registers = 64 words per row).
    constexpr int kPreA = kTabTR * (kTabChunk / 4) / kTabThreads, kPreB = kTabTQ * (kTabChunk / 4) / kTabThreads;
    uint32_t pre_a[kPreA], pre_b[kPreB];
    auto fetch = [&](uint32_t c0) {
#pragma unroll
        for (int u = 0; u < kPreA; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kTabThreads;
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gi = row0 + r;
            pre_a[u] = (e < (uint32_t)kTabTR * chunk_words && gi < dp.row_end) ? __ldg(reinterpret_cast<const uint32_t*>(gref + gi * cell_bytes + c0) + w) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kPreB; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kTabThreads;
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gj = col0 + r;
            pre_b[u] = (e < (uint32_t)kTabTQ * chunk_words && gj < dp.n_qry) ? __ldg(reinterpret_cast<const uint32_t*>(gqry + gj * cell_bytes + c0) + w) : 0u;
        }
    };
    fetch(0);
    for (uint32_t c0 = 0; c0 < cell_bytes; c0 += chunk) {
        __syncthreads();  // previous chunk consumed (first pass: orders the table build)
        // stage + recode: reference side as byte offsets of the table ROW (code << 9: 128 words), query side as byte
        // offsets inside a row (code << 2)
#pragma unroll
        for (int u = 0; u < kPreA; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kTabThreads;
            if (e < (uint32_t)kTabTR * chunk_words) {
                const uint32_t r = e / chunk_words, w = e % chunk_words;
                const uint32_t v = pre_a[u];
                uint4 o;
                o.x = fgra_code(v & 0xffu, base) << 9;
                o.y = fgra_code((v >> 8) & 0xffu, base) << 9;
                o.z = fgra_code((v >> 16) & 0xffu, base) << 9;
                o.w = fgra_code(v >> 24, base) << 9;
                *reinterpret_cast<uint4*>(sa + r * a_stride + 4 * w) = o;
            }
        }
#pragma unroll
        for (int u = 0; u < kPreB; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kTabThreads;
            if (e < (uint32_t)kTabTQ * chunk_words) {
                const uint32_t r = e / chunk_words, w = e % chunk_words;
                const uint32_t v = pre_b[u];
                uint2 o;
                o.x = (fgra_code(v & 0xffu, base) << 2) | (fgra_code((v >> 8) & 0xffu, base) << 18);
                o.y = (fgra_code((v >> 16) & 0xffu, base) << 2) | (fgra_code(v >> 24, base) << 18);
                *reinterpret_cast<uint2*>(sb + r * b_stride + 4 * w) = o;
            }
        }
        __syncthreads();
        if (c0 + chunk < cell_bytes) fetch(c0 + chunk);
        const uint32_t* pa0 = sa + ty * a_stride;
        const uint32_t* pa1 = sa + (ty + 16) * a_stride;
        const uint16_t* pb0 = sb + tx * b_stride;
        const uint16_t* pb1 = sb + (tx + 32) * b_stride;
#pragma unroll 1
        for (uint32_t e = 0; e < chunk; e += 8) {
            const uint4 a0l = *reinterpret_cast<const uint4*>(pa0 + e), a0h = *reinterpret_cast<const uint4*>(pa0 + e + 4);
            const uint4 a1l = *reinterpret_cast<const uint4*>(pa1 + e), a1h = *reinterpret_cast<const uint4*>(pa1 + e + 4);
            const uint4 b0 = *reinterpret_cast<const uint4*>(pb0 + e);
            const uint4 b1 = *reinterpret_cast<const uint4*>(pb1 + e);
            const uint32_t a0[8] = {a0l.x, a0l.y, a0l.z, a0l.w, a0h.x, a0h.y, a0h.z, a0h.w};
            const uint32_t a1[8] = {a1l.x, a1l.y, a1l.z, a1l.w, a1h.x, a1h.y, a1h.z, a1h.w};
            const uint32_t b0w[4] = {b0.x, b0.y, b0.z, b0.w}, b1w[4] = {b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t q0 = (i & 1) ? (b0w[i >> 1] >> 16) : (b0w[i >> 1] & 0xffffu);
                const uint32_t q1 = (i & 1) ? (b1w[i >> 1] >> 16) : (b1w[i >> 1] & 0xffffu);
                const uint32_t r0 = tbase + a0[i], r1 = tbase + a1[i];
                acc[0][0] += tab_lookup(r0 + q0, kPlane);
                acc[0][1] += tab_lookup(r0 + q1, kPlane);
                acc[1][0] += tab_lookup(r1 + q0, kPlane);
                acc[1][1] += tab_lookup(r1 + q1, kPlane);
            }
        }
    }

    // epilogue: identical to dist_kernel<FgraAcc>
#pragma unroll
    for (int a = 0; a < 2; ++a) {
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            const uint64_t i = row0 + ty + 16 * a, j = col0 + tx + 32 * b;
            if (i >= dp.row_end || j >= dp.n_qry) continue;
            if (dp.triangular && j > i) continue;
            FgraAcc fa;
            fa.sum = acc[a][b];
            bool bias;
            const double U = finish_union(fa, dp.p, gref + i * cell_bytes, gqry + j * cell_bytes, &bias);
            const double ca = dp.card_ref[i], cb = dp.card_qry[j];
            const double sim = (ca + cb - U) / U;
            const double s = sim < 0.0 ? 0.0 : sim;  // utils.rs:274: NaN propagates
            const double frac = 2.0 * s / (1.0 + s);
            const uint64_t o = dp.packed_tri ? (i * (i + 1) / 2 + j) : ((i - dp.out_row0) * dp.n_qry + j);
            if (dp.fp32)
                reinterpret_cast<float*>(dp.out)[o] = mash_distance_f32((float)frac, dp.k, dp.model);
            else
                reinterpret_cast<double*>(dp.out)[o] = mash_distance_f64(frac, dp.k, dp.model);
        }
    }
    }  // tiles
}

// ------------------------------------------------------------------------------------------------
// K4c: ML distance tiles through pair tables (same idea as K4b; everything here is integer, so any
// evaluation order is exact).  R[ca][cb] = contribution of merge(ra, rb) to the wrapping 64-bit sum S,
// W[ca][cb] = the bit pattern it adds to the b[] statistics (fits 32 bits for every code the table covers:
// codes 1..126 <-> registers 4p-4 .. 4p+121 <-> bit positions <= 31).  A register outside the table marks its
// whole sketch (a flag per staged row, set at recoding time) and pairs with a marked sketch take the exact
// per-pair path, so the inner loop carries no range check at all:
//     2 address ops + LDS.64 + LDS.32 + 64-bit add + ~2.7 carry-save ops per register pair  (K4: ~21, of which the
//     per-16-registers ripple through all counter planes was the largest part; here it runs once per 64 registers).
// A warp owns TWO reference rows (uniform -> broadcast loads) x 32 query columns, one per lane: what is pulled out of a
// query word (table column offset, G term) serves both rows.  (Round 1: one row x two columns per lane; on G-sum tiles that
// kernel was ALU-bound at 86 % with three of its seven ALU instructions per register pair spent on the query side.)
// ------------------------------------------------------------------------------------------------
constexpr int kMlTabThreads = 512;
constexpr int kMlTabTR = 32, kMlTabTQ = 32;
constexpr int kMlTabChunk = 64;

template <int NPL>
__device__ __forceinline__ void ml_pair_epilogue(const DistParams& dp, MlAccT<NPL>& acc, uint64_t i, uint64_t j, uint64_t o, uint32_t cell_bytes) {
    const unsigned char* gref = reinterpret_cast<const unsigned char*>(dp.ref);
    const unsigned char* gqry = reinterpret_cast<const unsigned char*>(dp.qry);
    bool bias;
    const double U = finish_union(acc, dp.p, gref + i * cell_bytes, gqry + j * cell_bytes, &bias);
    const double ca = dp.card_ref[i], cb = dp.card_qry[j];
    const double sim = (ca + cb - U) / U;
    const double s = sim < 0.0 ? 0.0 : sim;  // utils.rs:274: NaN propagates
    const double frac = 2.0 * s / (1.0 + s);
    if (dp.fp32)
        reinterpret_cast<float*>(dp.out)[o] = mash_distance_f32((float)frac, dp.k, dp.model);
    else
        reinterpret_cast<double*>(dp.out)[o] = mash_distance_f64<false>(frac, dp.k, dp.model);
}

// K4c, second kernel: one thread per output cell reads the statistics dist_ml_tab_kernel<NPL, true> stored and runs
// finish_union (count extraction + Ertl's solver) and the distance epilogue -- the same code as the fused form, but with
// 64 resident warps per SM to cover the divide chains.
template <int NPL>
__global__ void __launch_bounds__(256, 4) ml_finish_kernel(DistParams dp, uint32_t cell_bytes) {
    const uint64_t n = dp.ml_cells;
    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n; c += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t o = c + dp.ml_o_base;
        uint64_t i, j;
        if (dp.packed_tri) {
            i = (uint64_t)((sqrt(8.0 * (double)o + 1.0) - 1.0) * 0.5);
            while ((i + 1) * (i + 2) / 2 <= o) ++i;
            while (i * (i + 1) / 2 > o) --i;
            j = o - i * (i + 1) / 2;
        } else {
            i = o / dp.n_qry + dp.out_row0;
            j = o % dp.n_qry;
        }
        if (i < dp.row_begin || i >= dp.row_end || j >= dp.n_qry) continue;
        if (dp.triangular && j > i) continue;
        const uint32_t* sc = dp.ml_scratch + c;
        MlAccT<NPL> acc;
        acc.S = ((uint64_t)sc[n] << 32) | sc[0];
        acc.mmax = sc[2 * n];
#pragma unroll
        for (int l = 0; l < NPL; ++l) acc.pl[l] = sc[(3 + l) * n];
        ml_pair_epilogue(dp, acc, i, j, o, cell_bytes);
    }
}

// SPLIT: store the pair statistics for ml_finish_kernel instead of running the solver here.  The fused epilogue is 35 %
// of the instructions but 40 % of the time of the fused kernel: one CTA of 16 warps per SM (the tables fill its shared
// memory) cannot cover the dependent FP64 divide chains of the secant solver, and the main loop cannot overlap them.
template <int NPL, bool SPLIT>
__global__ void __launch_bounds__(kMlTabThreads, 1) dist_ml_tab_kernel(DistParams dp, uint32_t cell_bytes, uint32_t chunk,
                                                                       uint32_t tiles_x, uint32_t tiles_y, uint32_t one) {
    __shared__ uint32_t s_tile;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // three planes of 32-bit words (R low, R high, W), all indexed by the same (row, column) byte offset and read with
    // LDS.32: a warp reads one table row, so the bank is the query code mod 32 -- no conflicts (see K4b; with R as one
    // plane of 64-bit words 36 % of this kernel's shared-memory wavefronts were bank conflicts on unrelated sketches)
    uint32_t* R = reinterpret_cast<uint32_t*>(smem_raw);
    uint32_t* W = R + 2 * kTabN * kTabN;
    constexpr uint32_t kPlane = (uint32_t)kTabN * kTabN * 4u;
    const uint32_t a_stride = chunk + 4;   // u32 per reference row
    const uint32_t b_stride = chunk + 8;   // u16 per query row
    uint32_t* sa = W + kTabN * kTabN;
    uint32_t* sga = sa + (size_t)kMlTabTR * a_stride;                                 // G terms of the reference rows (G-sum tiles)
    uint16_t* sb = reinterpret_cast<uint16_t*>(sga + (size_t)kMlTabTR * a_stride);
    uint32_t* sflag = reinterpret_cast<uint32_t*>(sb + (size_t)kMlTabTQ * b_stride);  // [TR + TQ] "has a register outside the table"
    __shared__ uint32_t s_lo;                                                         // smallest register of the tile's sketches

    const int p = dp.p;
    const uint32_t base = (uint32_t)(4 * p - 4);
    for (uint32_t e = threadIdx.x; e < (uint32_t)(kTabN * kTabN); e += kMlTabThreads) {
        const uint32_t m = ml_tab_merged(e >> 7, e & 127u, base);
        const uint64_t ret = ml_ret_of(m, p);
        R[e] = (uint32_t)ret;
        R[e + kTabN * kTabN] = (uint32_t)(ret >> 32);
        W[e] = (uint32_t)ml_w_of(m, p);
    }
    const uint32_t ty = threadIdx.x >> 5, tx = threadIdx.x & 31u;  // the warp's reference rows: ty and ty + 16; the lane's column: tx
    const unsigned char* gref = reinterpret_cast<const unsigned char*>(dp.ref);
    const unsigned char* gqry = reinterpret_cast<const unsigned char*>(dp.qry);
    const uint32_t chunk_words = chunk / 4;
    const uint32_t rbase = (uint32_t)__cvta_generic_to_shared(R);
    const uint64_t n_tiles = (uint64_t)tiles_x * tiles_y;

    uint64_t tile;
    for (uint64_t iter = 0; next_tile(dp, &s_tile, n_tiles, iter, tile); ++iter) {
    const uint64_t row0 = dp.row_begin + (tile / tiles_x) * kMlTabTR;
    const uint64_t col0 = (tile % tiles_x) * kMlTabTQ;
    const uint64_t row_hi = min(row0 + kMlTabTR, dp.row_end);
    if (dp.triangular && col0 > row_hi - 1) continue;  // tile entirely above the diagonal (CTA-uniform)
    // G-sum tile (dist_tables.cuh): every register of every sketch of the tile has r2 >= 0, i.e. the smallest one is >= 4p + 4;
    // sketches reaching above k = 27 are flagged like those outside the table
    bool gs = false;
    uint32_t k0 = 0;
    if (dp.reg_mm_ref) {
        __syncthreads();                      // the previous tile's epilogue has read sflag
        if (threadIdx.x == 0) s_lo = 0xffu;
        __syncthreads();
        uint32_t mm = 0x00ffu;            // past the end (and threads beyond the tile's sketches): min 255, max 0
        if (threadIdx.x < kMlTabTR) {
            if (row0 + threadIdx.x < dp.row_end) mm = dp.reg_mm_ref[row0 + threadIdx.x];
        } else if (threadIdx.x < kMlTabTR + kMlTabTQ) {
            if (col0 + (threadIdx.x - kMlTabTR) < dp.n_qry) mm = dp.reg_mm_qry[col0 + (threadIdx.x - kMlTabTR)];
        }
        if (threadIdx.x < (uint32_t)((kMlTabTR + kMlTabTQ + 31) / 32 * 32)) {   // whole warps cover the tile's sketches
            const uint32_t wmin = __reduce_min_sync(0xffffffffu, mm & 0xffu);
            if (tx == 0) atomicMin(&s_lo, wmin);
        }
        __syncthreads();
        const uint32_t lo = s_lo;
        gs = lo >= (uint32_t)(4 * p + 4) && ml_gs_k(lo, p) + (uint32_t)p <= 36u;
        k0 = gs ? ml_gs_k(lo, p) : 0u;
        if (threadIdx.x < kMlTabTR + kMlTabTQ) sflag[threadIdx.x] = (gs && (mm >> 8) > ml_gs_max_reg(p, k0)) ? 1u : 0u;   // table tile: only code 127 flags
    } else if (threadIdx.x < kMlTabTR + kMlTabTQ) {
        sflag[threadIdx.x] = 0u;  // ordered before the staging by its first barrier
    }
    MlAccT<NPL> acc[2];
    acc[0].init();
    acc[1].init();

    // the next chunk's registers are fetched before the current chunk is consumed (as in K4b: one CTA per SM, nothing else
    // would cover the global-load latency between the two barriers)
    constexpr int kPreA = kMlTabTR * (kMlTabChunk / 4) / kMlTabThreads, kPreB = kMlTabTQ * (kMlTabChunk / 4) / kMlTabThreads;
    static_assert(kPreA >= 1 && kPreB >= 1, "prefetch registers per thread");
    uint32_t pre_a[kPreA], pre_b[kPreB];
    auto fetch = [&](uint32_t c0) {
#pragma unroll
        for (int u = 0; u < kPreA; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kMlTabThreads;
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gi = row0 + r;
            pre_a[u] = (e < (uint32_t)kMlTabTR * chunk_words && gi < dp.row_end) ? __ldg(reinterpret_cast<const uint32_t*>(gref + gi * cell_bytes + c0) + w) : 0u;
        }
#pragma unroll
        for (int u = 0; u < kPreB; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kMlTabThreads;
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gj = col0 + r;
            pre_b[u] = (e < (uint32_t)kMlTabTQ * chunk_words && gj < dp.n_qry) ? __ldg(reinterpret_cast<const uint32_t*>(gqry + gj * cell_bytes + c0) + w) : 0u;
        }
    };
    fetch(0);
    for (uint32_t c0 = 0; c0 < cell_bytes; c0 += chunk) {
        __syncthreads();
        // stage + recode: reference side code << 9 (byte offset of the 128-word table row), query side code << 2
#pragma unroll
        for (int u = 0; u < kPreA; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kMlTabThreads;
            if (e < (uint32_t)kMlTabTR * chunk_words) {
                const uint32_t r = e / chunk_words, w = e % chunk_words;
                const uint32_t v = pre_a[u];
                const uint32_t c0_ = fgra_code(v & 0xffu, base), c1_ = fgra_code((v >> 8) & 0xffu, base);
                const uint32_t c2_ = fgra_code((v >> 16) & 0xffu, base), c3_ = fgra_code(v >> 24, base);
                if (max(max(c0_, c1_), max(c2_, c3_)) == 127u) atomicOr(&sflag[r], 1u);
                // G-sum tiles stage the absolute shared address of the W-plane row (one add per register saved in the loop)
                const uint32_t a_off = gs ? rbase + 2u * kPlane : 0u;
                *reinterpret_cast<uint4*>(sa + r * a_stride + 4 * w) = make_uint4((c0_ << 9) + a_off, (c1_ << 9) + a_off, (c2_ << 9) + a_off, (c3_ << 9) + a_off);
                if (gs)
                    *reinterpret_cast<uint4*>(sga + r * a_stride + 4 * w) = make_uint4(ml_gs_term(ml_gs_n(v & 0xffu, p, k0)), ml_gs_term(ml_gs_n((v >> 8) & 0xffu, p, k0)),
                                                                                       ml_gs_term(ml_gs_n((v >> 16) & 0xffu, p, k0)), ml_gs_term(ml_gs_n(v >> 24, p, k0)));
            }
        }
#pragma unroll
        for (int u = 0; u < kPreB; ++u) {
            const uint32_t e = threadIdx.x + (uint32_t)u * kMlTabThreads;
            if (e < (uint32_t)kMlTabTQ * chunk_words) {
                const uint32_t r = e / chunk_words, w = e % chunk_words;
                const uint32_t v = pre_b[u];
                const uint32_t c0_ = fgra_code(v & 0xffu, base), c1_ = fgra_code((v >> 8) & 0xffu, base);
                const uint32_t c2_ = fgra_code((v >> 16) & 0xffu, base), c3_ = fgra_code(v >> 24, base);
                if (max(max(c0_, c1_), max(c2_, c3_)) == 127u) atomicOr(&sflag[kMlTabTR + r], 1u);
                *reinterpret_cast<uint2*>(sb + r * b_stride + 4 * w) =
                    make_uint2(ml_pack_b(c0_, ml_gs_n(v & 0xffu, p, k0), c1_, ml_gs_n((v >> 8) & 0xffu, p, k0)),
                               ml_pack_b(c2_, ml_gs_n((v >> 16) & 0xffu, p, k0), c3_, ml_gs_n(v >> 24, p, k0)));
            }
        }
        __syncthreads();
        if (c0 + chunk < cell_bytes) fetch(c0 + chunk);
        const uint32_t* pa0 = sa + ty * a_stride;
        const uint32_t* pa1 = sa + (ty + 16) * a_stride;
        const uint16_t* pb = sb + tx * b_stride;
        // 8 registers x 2 pairs (two reference rows, one query column): table lookups, S, and the Harley-Seal step;
        // returns the two weight-8 carries
        auto step8 = [&](uint32_t e, uint32_t& c0, uint32_t& c1) {
            const uint4 a0l = *reinterpret_cast<const uint4*>(pa0 + e), a0h = *reinterpret_cast<const uint4*>(pa0 + e + 4);
            const uint4 a1l = *reinterpret_cast<const uint4*>(pa1 + e), a1h = *reinterpret_cast<const uint4*>(pa1 + e + 4);
            const uint4 bq = *reinterpret_cast<const uint4*>(pb + e);
            const uint32_t a0[8] = {a0l.x, a0l.y, a0l.z, a0l.w, a0h.x, a0h.y, a0h.z, a0h.w};
            const uint32_t a1[8] = {a1l.x, a1l.y, a1l.z, a1l.w, a1h.x, a1h.y, a1h.z, a1h.w};
            const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
            uint32_t w0[8], w1[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t q = (i & 1) ? ml_b_q1(bw[i >> 1]) : ml_b_q0(bw[i >> 1]);
                const uint32_t r0 = rbase + a0[i] + q, r1 = rbase + a1[i] + q;
                uint32_t l0, h0, l1, h1;
                asm("ld.shared.u32 %0, [%1];" : "=r"(l0) : "r"(r0));
                asm("ld.shared.u32 %0, [%1];" : "=r"(h0) : "r"(r0 + kPlane));
                asm("ld.shared.u32 %0, [%1];" : "=r"(w0[i]) : "r"(r0 + 2u * kPlane));
                asm("ld.shared.u32 %0, [%1];" : "=r"(l1) : "r"(r1));
                asm("ld.shared.u32 %0, [%1];" : "=r"(h1) : "r"(r1 + kPlane));
                asm("ld.shared.u32 %0, [%1];" : "=r"(w1[i]) : "r"(r1 + 2u * kPlane));
                acc[0].S += ((uint64_t)h0 << 32) | l0;
                acc[1].S += ((uint64_t)h1 << 32) | l1;
            }
            c0 = acc[0].csa8(w0);
            c1 = acc[1].csa8(w1);
        };
        // the same on a G-sum tile: ONE table lookup per pair (W); S as min of two staged powers of two, eight per 32-bit batch;
        // the query side's column offset and G term are extracted once for the two rows.
        // ncu, one-row version (n = 6000): ALU pipe 80-86 %, FMA-heavy 29 %.  Moving the field shifts onto the FMA pipe
        // (IMAD.HI by run-time powers of two: ALU 60 %, FMA-heavy 59 %) did not shorten it (9.0 -> 9.5 ms, dispatch stalls x10).
        const uint32_t* pg0 = sga + ty * a_stride;
        const uint32_t* pg1 = sga + (ty + 16) * a_stride;
        auto step8g = [&](uint32_t e, uint32_t& c0, uint32_t& c1) {
            const uint4 a0l = *reinterpret_cast<const uint4*>(pa0 + e), a0h = *reinterpret_cast<const uint4*>(pa0 + e + 4);
            const uint4 a1l = *reinterpret_cast<const uint4*>(pa1 + e), a1h = *reinterpret_cast<const uint4*>(pa1 + e + 4);
            const uint4 g0l = *reinterpret_cast<const uint4*>(pg0 + e), g0h = *reinterpret_cast<const uint4*>(pg0 + e + 4);
            const uint4 g1l = *reinterpret_cast<const uint4*>(pg1 + e), g1h = *reinterpret_cast<const uint4*>(pg1 + e + 4);
            const uint4 bq = *reinterpret_cast<const uint4*>(pb + e);
            const uint32_t a0[8] = {a0l.x, a0l.y, a0l.z, a0l.w, a0h.x, a0h.y, a0h.z, a0h.w};
            const uint32_t a1[8] = {a1l.x, a1l.y, a1l.z, a1l.w, a1h.x, a1h.y, a1h.z, a1h.w};
            const uint32_t ga0[8] = {g0l.x, g0l.y, g0l.z, g0l.w, g0h.x, g0h.y, g0h.z, g0h.w};
            const uint32_t ga1[8] = {g1l.x, g1l.y, g1l.z, g1l.w, g1h.x, g1h.y, g1h.z, g1h.w};
            const uint32_t bw[4] = {bq.x, bq.y, bq.z, bq.w};
            uint32_t w0[8], w1[8], s0 = 0u, s1 = 0u;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t x = bw[i >> 1];
                const uint32_t q = (i & 1) ? ml_b_q1(x) : ml_b_q0(x);
                const uint32_t gb = ml_gs_term((i & 1) ? ml_b_n1(x) : ml_b_n0(x));
                const uint32_t g0 = min(ga0[i], gb), g1 = min(ga1[i], gb);
                asm("ld.shared.u32 %0, [%1];" : "=r"(w0[i]) : "r"(mad_one(q, one, a0[i])));
                asm("ld.shared.u32 %0, [%1];" : "=r"(w1[i]) : "r"(mad_one(q, one, a1[i])));
                s0 = i == 0 ? g0 : mad_one(g0, one, s0);
                s1 = i == 0 ? g1 : mad_one(g1, one, s1);
            }
            acc[0].S += (uint64_t)s0;
            acc[1].S += (uint64_t)s1;
            c0 = acc[0].csa8(w0);
            c1 = acc[1].csa8(w1);
        };
        if (gs) {   // G-sum tiles only exist for 16-byte-aligned sketches of >= 16 registers: chunk >= 16
            if (chunk >= 64u) {
#pragma unroll 1
                for (uint32_t e = 0; e < chunk; e += 64) {
                    uint32_t c0[8], c1[8];
#pragma unroll
                    for (int g = 0; g < 8; ++g) step8g(e + 8u * g, c0[g], c1[g]);
                    acc[0].fold64(c0);
                    acc[1].fold64(c1);
                }
            } else {
#pragma unroll 1
                for (uint32_t e = 0; e < chunk; e += 8) {
                    uint32_t c0, c1;
                    step8g(e, c0, c1);
                    acc[0].template ripple_all<3>(c0);
                    acc[1].template ripple_all<3>(c1);
                }
            }
        } else if (chunk >= 64u) {
#pragma unroll 1
            for (uint32_t e = 0; e < chunk; e += 64) {
                uint32_t c0[8], c1[8];
#pragma unroll
                for (int g = 0; g < 8; ++g) step8(e + 8u * g, c0[g], c1[g]);
                acc[0].fold64(c0);
                acc[1].fold64(c1);
            }
        } else {  // sketches of 8 .. 32 registers (p = 3 .. 5)
#pragma unroll 1
            for (uint32_t e = 0; e < chunk; e += 8) {
                uint32_t c0, c1;
                step8(e, c0, c1);
                acc[0].template ripple_all<3>(c0);
                acc[1].template ripple_all<3>(c1);
            }
        }
    }
    __syncthreads();  // sflag complete (set during the last staging pass at the latest)

#pragma unroll
    for (int b = 0; b < 2; ++b) {
        const uint64_t i = row0 + ty + 16 * b, j = col0 + tx;
        if (i >= dp.row_end || j >= dp.n_qry) continue;
        if (dp.triangular && j > i) continue;
        if (gs) acc[b].mmax = kMlGsMarker | k0;                             // S holds the G-sum: finish_union rebuilds S
        if (sflag[ty + 16 * b] | sflag[kMlTabTR + tx]) acc[b].mmax = 255u;  // -> exact per-pair path in finish_union
        const uint64_t o = dp.packed_tri ? (i * (i + 1) / 2 + j) : ((i - dp.out_row0) * dp.n_qry + j);
        if (SPLIT) {
            uint32_t* sc = dp.ml_scratch + (o - dp.ml_o_base);  // consecutive lanes -> consecutive cells: coalesced per word
            const uint64_t n = dp.ml_cells;
            sc[0] = (uint32_t)acc[b].S;
            sc[n] = (uint32_t)(acc[b].S >> 32);
            sc[2 * n] = acc[b].mmax;
#pragma unroll
            for (int l = 0; l < NPL; ++l) sc[(3 + l) * n] = acc[b].pl[l];
        } else {
            ml_pair_epilogue(dp, acc[b], i, j, o, cell_bytes);
        }
    }
    }  // tiles
}

// ------------------------------------------------------------------------------------------------
// K4m, small sketches: expectedCollision's 41 x 1024 loop as per-sketch term vectors and an exact-order tile product
// (estimators.cuh: hmh_ec_term).  Per pair the loop costs 2 x 41 x 1024 pow calls in the reference; here every small sketch
// pays them once (hmh_ec_fill_kernel), and a pair costs 41 984 multiply-adds summed in the reference's (i, j) order with one
// scalar accumulator per pair -- the same doubles in the same order, so the sum is the per-pair loop's bit for bit.
// ------------------------------------------------------------------------------------------------
__global__ void hmh_count_small_kernel(const double* __restrict__ card, uint64_t begin, uint64_t end, uint32_t* count) {
    uint32_t mine = 0;
    for (uint64_t i = begin + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < end; i += (uint64_t)gridDim.x * blockDim.x)
        mine += hmh_ec_is_small(card[i]) ? 1u : 0u;
    mine = __reduce_add_sync(0xffffffffu, mine);
    if ((threadIdx.x & 31u) == 0 && mine) atomicAdd(count, mine);
}
cudaError_t launch_hmh_count_small(const double* card, uint64_t begin, uint64_t end, uint32_t* count_dev, cudaStream_t st) {
    if (end <= begin) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<uint64_t>((end - begin + 255) / 256, 1024);
    hmh_count_small_kernel<<<grid, 256, 0, st>>>(card, begin, end, count_dev);
    return cudaGetLastError();
}

// one CTA: ranks of the small sketches in index order (so slots ascend with the sketch index)
__global__ void __launch_bounds__(1024) hmh_slots_kernel(const double* __restrict__ card, uint64_t begin, uint64_t end, uint32_t cap,
                                                         int32_t* __restrict__ slot, uint32_t* __restrict__ src, uint32_t* count) {
    __shared__ uint32_t s_warp[32];
    __shared__ uint32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = begin; base < end; base += 1024) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = (i < end && hmh_ec_is_small(card[i])) ? 1u : 0u;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if ((threadIdx.x & 31u) >= (uint32_t)d) x += y;
        }
        if ((threadIdx.x & 31u) == 31u) s_warp[threadIdx.x >> 5] = x;
        __syncthreads();
        if (threadIdx.x < 32) {
            const uint32_t w = s_warp[threadIdx.x];
            uint32_t xw = w;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, xw, d);
                if (threadIdx.x >= (uint32_t)d) xw += y;
            }
            s_warp[threadIdx.x] = xw - w;
        }
        __syncthreads();
        const uint32_t rank = s_carry + s_warp[threadIdx.x >> 5] + (x - v);
        if (i < end) {
            const bool take = v && rank < cap;
            slot[i - begin] = take ? (int32_t)rank : -1;
            if (take) src[rank] = (uint32_t)i;
        }
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = rank + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = min(s_carry, cap);
}
cudaError_t launch_hmh_slots(const double* card, uint64_t begin, uint64_t end, uint32_t cap, int32_t* slot, uint32_t* src,
                             uint32_t* count_dev, cudaStream_t st) {
    hmh_slots_kernel<<<1, 1024, 0, st>>>(card, begin, end, cap, slot, src, count_dev);
    return cudaGetLastError();
}

// Per sketch: 41 x 1024 terms, then kHmhEcTail doubles of which the first 41 are the rows' largest |term| (the tile product
// uses them to stop early).  One CTA per (row, sketch).
__global__ void __launch_bounds__(256) hmh_ec_fill_kernel(const double* __restrict__ card, const uint32_t* __restrict__ src,
                                                          const uint32_t* __restrict__ count, double* __restrict__ terms) {
    __shared__ double s_max[8];
    const uint32_t s = blockIdx.y;
    if (s >= *count) return;
    const double n = card[src[s]];
    double* out = terms + (size_t)s * kHmhEcTermsPerSketch;
    const int i = 1 + (int)blockIdx.x;
    double mx = 0.0;
    for (uint32_t j = threadIdx.x; j < 1024u; j += blockDim.x) {
        const double t = hmh_ec_term(i, 1 + (int)j, n);
        out[(size_t)blockIdx.x * 1024u + j] = t;
        mx = fmax(mx, fabs(t));
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
    if ((threadIdx.x & 31u) == 0) s_max[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) mx = fmax(mx, s_max[w]);
        out[kHmhEcLen + blockIdx.x] = mx;
    }
}
cudaError_t launch_hmh_ec_fill(const double* card, const uint32_t* src, const uint32_t* count_dev, uint32_t cap, double* terms,
                               cudaStream_t st) {
    if (cap == 0) return cudaSuccess;
    if (cap > 65535) return cudaErrorInvalidValue;   // grid.y limit: the caller caps the slots below it
    hmh_ec_fill_kernel<<<dim3(kHmhEcRows, cap), 256, 0, st>>>(card, src, count_dev, terms);
    return cudaGetLastError();
}

// 128 x 128 pairs per CTA, 8 x 8 per thread: 16 LDS.64 + 128 un-fused f64 operations per k (the 64 x 64 / 4 x 4 first version
// ran at 45 % of the f64 pipe).  Every term is >= 0 (both factors are <= 0), so a pair's sum only grows; after each row of
// 1024 terms the CTA checks whether, for every pair it owns, the largest product the remaining rows can hold is below half an
// ulp of the sum so far (largest |term| of the rows still to come, per sketch, from the fill kernel; estimators.cuh:
// hmh_ec_rest_is_absorbed): then every further addition of the reference's loop rounds back to the same double, and
// the loop can stop -- typically after 23-25 of the 41 rows -- with the bit-identical result.
constexpr int kEcTile = 128, kEcKC = 16, kEcThreads = 256, kEcM = 8;
__global__ void __launch_bounds__(kEcThreads, 1) hmh_ec_gemm_kernel(const double* __restrict__ tr, const uint32_t* __restrict__ src_r,
                                                                     const uint32_t* __restrict__ count_r, const double* __restrict__ tq,
                                                                     const uint32_t* __restrict__ src_q, const uint32_t* __restrict__ count_q,
                                                                     int triangular, double* __restrict__ ec, uint32_t ld) {
    // k-major tiles: sA[kk][row], sB[kk][col] (+1 pad against the transposing stores)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    double (*sA)[kEcTile + 1] = reinterpret_cast<double (*)[kEcTile + 1]>(smem_raw);
    double (*sB)[kEcTile + 1] = sA + kEcKC;
    double (*sufA)[kEcTile] = reinterpret_cast<double (*)[kEcTile]>(sB + kEcKC);   // sufA[i][row] = largest |term| of rows i.. (from 0) of that sketch
    double (*sufB)[kEcTile] = sufA + (kHmhEcRows + 1);
    const uint32_t nr = *count_r, nq = *count_q;
    const uint32_t r0 = blockIdx.y * kEcTile, q0 = blockIdx.x * kEcTile;
    if (r0 >= nr || q0 >= nq) return;
    if (triangular && src_r[min(r0 + kEcTile, nr) - 1] < src_q[q0]) return;   // every pair of the tile has j > i
    const uint32_t ty = threadIdx.x >> 4, tx = threadIdx.x & 15u;
    // suffix maxima of the row maxima (rows past the count: 0, they never hold the loop back)
    {
        const uint32_t e = threadIdx.x & (kEcTile - 1);
        const bool side_b = threadIdx.x >= (uint32_t)kEcTile;
        const uint32_t idx = (side_b ? q0 : r0) + e;
        const bool valid = idx < (side_b ? nq : nr);
        const double* rm = (side_b ? tq : tr) + (size_t)(valid ? idx : 0) * kHmhEcTermsPerSketch + kHmhEcLen;
        double (*suf)[kEcTile] = side_b ? sufB : sufA;
        double m = 0.0;
        suf[kHmhEcRows][e] = 0.0;
        for (int i = kHmhEcRows - 1; i >= 0; --i) {
            m = valid ? fmax(m, rm[i]) : 0.0;
            suf[i][e] = m;
        }
    }
    double acc[kEcM][kEcM];
#pragma unroll
    for (int i = 0; i < kEcM; ++i)
#pragma unroll
        for (int j = 0; j < kEcM; ++j) acc[i][j] = 0.0;
    // loader: thread t brings 4 consecutive k of two rows of each operand (rows past the count read row 0: never stored)
    const uint32_t lrow = threadIdx.x >> 2, lk = (threadIdx.x & 3u) * 4u;
    const double* ga0 = tr + (size_t)(r0 + lrow < nr ? r0 + lrow : 0) * kHmhEcTermsPerSketch + lk;
    const double* ga1 = tr + (size_t)(r0 + lrow + 64 < nr ? r0 + lrow + 64 : 0) * kHmhEcTermsPerSketch + lk;
    const double* gb0 = tq + (size_t)(q0 + lrow < nq ? q0 + lrow : 0) * kHmhEcTermsPerSketch + lk;
    const double* gb1 = tq + (size_t)(q0 + lrow + 64 < nq ? q0 + lrow + 64 : 0) * kHmhEcTermsPerSketch + lk;
    for (uint32_t k0 = 0; k0 < (uint32_t)kHmhEcLen; k0 += kEcKC) {
        const double4 a40 = *reinterpret_cast<const double4*>(ga0 + k0), a41 = *reinterpret_cast<const double4*>(ga1 + k0);
        const double4 b40 = *reinterpret_cast<const double4*>(gb0 + k0), b41 = *reinterpret_cast<const double4*>(gb1 + k0);
        __syncthreads();
        if ((k0 & 1023u) == 0 && k0) {
            // a row of the reference's loop is complete: can the rest still change any of this CTA's sums?
            const uint32_t row = k0 >> 10;
            bool done = true;
#pragma unroll
            for (int i = 0; i < kEcM; ++i)
#pragma unroll
                for (int j = 0; j < kEcM; ++j)
                    done = done && (hmh_ec_rest_is_absorbed(sufA[row][ty * kEcM + i], sufB[row][j * 16 + tx], acc[i][j]) ||
                                    r0 + ty * kEcM + i >= nr || q0 + j * 16 + tx >= nq);
            if (__syncthreads_and(done)) break;
        }
        sA[lk + 0][lrow] = a40.x; sA[lk + 1][lrow] = a40.y; sA[lk + 2][lrow] = a40.z; sA[lk + 3][lrow] = a40.w;
        sA[lk + 0][lrow + 64] = a41.x; sA[lk + 1][lrow + 64] = a41.y; sA[lk + 2][lrow + 64] = a41.z; sA[lk + 3][lrow + 64] = a41.w;
        sB[lk + 0][lrow] = b40.x; sB[lk + 1][lrow] = b40.y; sB[lk + 2][lrow] = b40.z; sB[lk + 3][lrow] = b40.w;
        sB[lk + 0][lrow + 64] = b41.x; sB[lk + 1][lrow + 64] = b41.y; sB[lk + 2][lrow + 64] = b41.z; sB[lk + 3][lrow + 64] = b41.w;
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < kEcKC; ++kk) {
            double a[kEcM], b[kEcM];
#pragma unroll
            for (int i = 0; i < kEcM; ++i) a[i] = sA[kk][ty * kEcM + i];
#pragma unroll
            for (int j = 0; j < kEcM; ++j) b[j] = sB[kk][j * 16 + tx];    // lanes read consecutive doubles: no bank conflicts
#pragma unroll
            for (int i = 0; i < kEcM; ++i)
#pragma unroll
                for (int j = 0; j < kEcM; ++j) acc[i][j] = acc[i][j] + a[i] * b[j];   // -fmad=false: multiply, then add, like the loop
        }
    }
#pragma unroll
    for (int i = 0; i < kEcM; ++i)
#pragma unroll
        for (int j = 0; j < kEcM; ++j) {
            const uint32_t r = r0 + ty * kEcM + i, q = q0 + j * 16 + tx;
            if (r < nr && q < nq) ec[(size_t)r * ld + q] = acc[i][j];
        }
}
cudaError_t launch_hmh_ec_gemm(const double* terms_r, const uint32_t* src_r, const uint32_t* count_r, uint32_t cap_r,
                               const double* terms_q, const uint32_t* src_q, const uint32_t* count_q, uint32_t cap_q, int triangular,
                               double* ec, uint32_t ld, cudaStream_t st) {
    if (cap_r == 0 || cap_q == 0) return cudaSuccess;
    const dim3 grid((cap_q + kEcTile - 1) / kEcTile, (cap_r + kEcTile - 1) / kEcTile);
    if (grid.y > 65535) return cudaErrorInvalidValue;
    constexpr size_t smem = 2 * sizeof(double) * ((size_t)kEcKC * (kEcTile + 1) + (size_t)(kHmhEcRows + 1) * kEcTile);
    const cudaError_t e = cudaFuncSetAttribute(hmh_ec_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    hmh_ec_gemm_kernel<<<grid, kEcThreads, smem, st>>>(terms_r, src_r, count_r, terms_q, src_q, count_q, triangular, ec, ld);
    return cudaGetLastError();
}

// ------------------------------------------------------------------------------------------------
// K3: per-sketch cardinality (utils.rs:213-219, 314-316; hyperminhash cardinality())
// One lane per sketch walks the registers in index order with the same accumulators as K4
// (the union of a sketch with itself is the sketch).
// ------------------------------------------------------------------------------------------------
constexpr int kCardWarps = 4;        // sketches per CTA
constexpr int kCardChunk = 1024;     // bytes of one sketch staged per pass

// A warp owns one sketch: all lanes stage it through shared memory (coalesced 16-byte loads), lane 0 walks the staged
// bytes in index order.  (One THREAD per sketch reading global memory directly -- the first version -- spent its time
// waiting on its own strided loads: 0.44 ms for 200 HLL p=14 sketches.)
template <class ACC, int G>
__global__ void __launch_bounds__(kCardWarps * 32) card_kernel(const unsigned char* __restrict__ regs, uint64_t n, uint32_t cell_bytes, int p,
                                                               double* __restrict__ card, uint32_t* flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables tabs;
    build_tables<ACC>(tabs, smem_raw, p);
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint64_t i = (uint64_t)blockIdx.x * kCardWarps + warp;
    if (i >= n) return;
    uint32_t* stage = reinterpret_cast<uint32_t*>(smem_raw + ((ACC::kTableBytes + 15) & ~15) + warp * kCardChunk);
    const unsigned char* g = regs + i * cell_bytes;
    const uint32_t chunk = cell_bytes < (uint32_t)kCardChunk ? cell_bytes : (uint32_t)kCardChunk;
    ACC acc;
    acc.init();
    for (uint32_t c0 = 0; c0 < cell_bytes; c0 += chunk) {
        __syncwarp();
        for (uint32_t w = lane; w < chunk / 4; w += 32) stage[w] = __ldg(reinterpret_cast<const uint32_t*>(g + c0) + w);
        __syncwarp();
        if (lane == 0) {
            for (uint32_t e = 0; e < chunk; e += G) {
                uint32_t v[G];
#pragma unroll
                for (int k = 0; k < G / 4; ++k) {
                    const uint32_t w = stage[e / 4 + k];
                    v[4 * k] = w & 0xffu;
                    v[4 * k + 1] = (w >> 8) & 0xffu;
                    v[4 * k + 2] = (w >> 16) & 0xffu;
                    v[4 * k + 3] = w >> 24;
                }
                acc_add<ACC, G>(acc, v, v, tabs, p + 1);
            }
        }
    }
    if (lane == 0) {
        bool bias;
        card[i] = finish_union(acc, p, g, g, &bias);
        if (bias && flags) atomicAdd(flags, 1u);
    }
}

__global__ void __launch_bounds__(kCardWarps * 32) card_hmh_kernel(const uint32_t* __restrict__ regs, uint64_t n, double* __restrict__ card) {
    __shared__ uint32_t s_stage[kCardWarps][kCardChunk / 4];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint64_t i = (uint64_t)blockIdx.x * kCardWarps + warp;
    if (i >= n) return;
    const uint32_t* g = regs + i * 8192u;
    // Warp-parallel when exact: the terms 2^-lz of a sketch whose leading-zero counts span at most 29 levels sum to the same
    // double in any order (no partial sum of the sequential loop rounds: 16384 * 2^28 < 2^53 units of 2^-(lo+28); K4i's
    // argument), so the lanes add their shares as integers.  Otherwise lane 0 walks the registers as the reference does.
    uint32_t mn = 63u, mx = 0u;
    for (uint32_t w = lane; w < 8192u; w += 32) {
        const uint32_t v = __ldg(g + w);
        const uint32_t l0 = (v & 0xffffu) >> 10, l1 = v >> 26;
        mn = min(mn, min(l0, l1));
        mx = max(mx, max(l0, l1));
    }
    const uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = __reduce_max_sync(0xffffffffu, mx);
    if (hi - lo <= (uint32_t)kHllIntW) {
        uint64_t s64 = 0;
        uint32_t z = 0;
        for (uint32_t w = lane; w < 8192u; w += 32) {
            const uint32_t v = __ldg(g + w);
            const uint32_t l0 = (v & 0xffffu) >> 10, l1 = v >> 26;
            s64 += (uint64_t)(1u << (kHllIntW - (l0 - lo))) + (1u << (kHllIntW - (l1 - lo)));
            z += (l0 == 0u) + (l1 == 0u);
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s64 += __shfl_xor_sync(0xffffffffu, s64, d);
            z += __shfl_xor_sync(0xffffffffu, z, d);
        }
        if (lane == 0)
            card[i] = hmh_cardinality_from((double)s64 * __hiloint2double((int)((1023u - (uint32_t)kHllIntW - lo) << 20), 0), (double)z);
        return;
    }
    uint32_t* stage = s_stage[warp];
    double sum = 0.0, ez = 0.0;
    for (uint32_t w0 = 0; w0 < 8192u; w0 += kCardChunk / 4) {
        __syncwarp();
        for (uint32_t w = lane; w < kCardChunk / 4; w += 32) stage[w] = __ldg(g + w0 + w);
        __syncwarp();
        if (lane == 0) {
            for (uint32_t w = 0; w < kCardChunk / 4; ++w) {
                const uint32_t v = stage[w];
                const uint32_t l0 = (v & 0xffffu) >> 10, l1 = v >> 26;
                if (l0 == 0) ez += 1.0;
                sum += pow2neg(l0);
                if (l1 == 0) ez += 1.0;
                sum += pow2neg(l1);
            }
        }
    }
    if (lane == 0) card[i] = hmh_cardinality_from(sum, ez);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
cudaError_t ensure_tables() {
    // constant memory is per device (per context); upload is cheap, so do it whenever asked
    static const UllConsts host = make_ull_consts();
    return cudaMemcpyToSymbol(c_ull, &host, sizeof(UllConsts));
}

static uint32_t cell_bytes_of(int algo, int p) { return algo == HMH ? 32768u : (1u << p); }
uint32_t ml_scratch_words(int p) { return 3u + (p <= 11 ? 12u : p <= 15 ? 16u : (uint32_t)kMlPlanes); }

template <class ACC, int G>
static cudaError_t launch_dist_t(const DistParams& dp, cudaStream_t st) {
    constexpr int TR = 16 * ACC::RM, TQ = 16 * ACC::QM;
    const uint32_t cb = cell_bytes_of(dp.algo, dp.p);
    const uint32_t chunk = cb < (uint32_t)kChunkBytes ? cb : (uint32_t)kChunkBytes;
    const size_t smem = ACC::kTableBytes + (size_t)(TR + TQ) * (chunk + 4);
    auto kern = dist_kernel<ACC, G>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t rows = dp.row_end - dp.row_begin;
    const uint64_t gy = (rows + TR - 1) / TR;
    uint64_t ncols = dp.n_qry;
    if (dp.triangular && dp.row_end < ncols) ncols = dp.row_end;  // nothing right of the diagonal
    const uint64_t gx = (ncols + TQ - 1) / TQ;
    // grid.y is limited to 65535: walk row bands
    for (uint64_t y0 = 0; y0 < gy; y0 += 65535) {
        DistParams q = dp;
        q.row_begin = dp.row_begin + y0 * TR;
        const uint64_t ny = (gy - y0) < 65535 ? (gy - y0) : 65535;
        dim3 grid((unsigned)gx, (unsigned)ny);
        kern<<<grid, kDistThreads, smem, st>>>(q, cb, chunk);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// smallest non-zero register byte of an array (atomicMin into *out, which the caller presets to 0xffffffff)
__global__ void regmin_kernel(const uint32_t* __restrict__ regs, uint64_t n_words, uint32_t* out) {
    uint32_t m = 0xffu;
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_words; i += stride) {
        const uint32_t w = __ldg(regs + i);
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            const uint32_t r = (w >> (8 * b)) & 0xffu;
            m = r ? min(m, r) : m;
        }
    }
    m = __reduce_min_sync(0xffffffffu, m);
    if ((threadIdx.x & 31) == 0 && m != 0xffu) atomicMin(out, m);
}
cudaError_t launch_regmin(const void* regs, uint64_t n_bytes, uint32_t* out_dev, int n_sm, cudaStream_t st) {
    const uint64_t n_words = n_bytes / 4;
    if (n_words == 0) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<uint64_t>((n_words + 255) / 256, (uint64_t)n_sm * 8);
    regmin_kernel<<<grid, 256, 0, st>>>(reinterpret_cast<const uint32_t*>(regs), n_words, out_dev);
    return cudaGetLastError();
}

// Wrapping sum of the output bit patterns (f64 -> u64, f32 -> zero-extended u32) of the cells rows [row_begin, row_end)
// define, plus their count: sums[0] += sum, sums[1] += cells.  Order-free, so it is independent of how the rows were cut
// into blocks / ranks: the checksum of checksums of a tiled or sharded run equals the one of a single call.
// Addressing as DistParams: packed_tri ? out[i*(i+1)/2 + j] : out[(i - out_row0) * n_qry + j]; triangular: only j <= i.
__global__ void __launch_bounds__(256) out_checksum_kernel(DistParams dp, unsigned long long* sums) {
    unsigned long long acc = 0ull, cnt = 0ull;
    for (uint64_t i = dp.row_begin + blockIdx.x; i < dp.row_end; i += gridDim.x) {
        const uint64_t ncol = dp.triangular ? min(dp.n_qry, i + 1) : dp.n_qry;
        const uint64_t o0 = dp.packed_tri ? i * (i + 1) / 2 : (i - dp.out_row0) * dp.n_qry;
        if (dp.fp32) {
            const uint32_t* q = reinterpret_cast<const uint32_t*>(dp.out) + o0;
            for (uint64_t j = threadIdx.x; j < ncol; j += blockDim.x) acc += q[j];
        } else {
            const unsigned long long* q = reinterpret_cast<const unsigned long long*>(dp.out) + o0;
            for (uint64_t j = threadIdx.x; j < ncol; j += blockDim.x) acc += q[j];
        }
        if (threadIdx.x == 0) cnt += ncol;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, d);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, d);
    }
    if ((threadIdx.x & 31u) == 0) {
        if (acc) atomicAdd(sums, acc);
        if (cnt) atomicAdd(sums + 1, cnt);
    }
}
cudaError_t launch_out_checksum(const DistParams& dp, unsigned long long* sums_dev, cudaStream_t st) {
    if (dp.row_end <= dp.row_begin) return cudaSuccess;
    const unsigned grid = (unsigned)std::min<uint64_t>(dp.row_end - dp.row_begin, (uint64_t)dp.n_sm * 8);
    out_checksum_kernel<<<grid, 256, 0, st>>>(dp, sums_dev);
    return cudaGetLastError();
}

static cudaError_t launch_dist_fgra_tab(const DistParams& dp, cudaStream_t st) {
    const uint32_t cb = cell_bytes_of(dp.algo, dp.p);
    const uint32_t chunk = cb < (uint32_t)kTabChunk ? cb : (uint32_t)kTabChunk;
    const size_t smem = (size_t)kTabN * kTabN * 8 + (size_t)kTabTR * (chunk + 4) * 4 + (size_t)kTabTQ * (chunk + 8) * 2;
    cudaError_t e = cudaFuncSetAttribute(dist_fgra_tab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t rows = dp.row_end - dp.row_begin;
    const uint64_t gy = (rows + kTabTR - 1) / kTabTR;
    uint64_t ncols = dp.n_qry;
    if (dp.triangular && dp.row_end < ncols) ncols = dp.row_end;
    const uint64_t gx = (ncols + kTabTQ - 1) / kTabTQ;
    if (gx * gy > 0xffffffffull) return cudaErrorInvalidValue;
    if (dp.tile_counter) {
        e = cudaMemsetAsync(dp.tile_counter, 0, 4, st);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(gx * gy, (uint64_t)dp.n_sm);
    dist_fgra_tab_kernel<<<grid, kTabThreads, smem, st>>>(dp, cb, chunk, (uint32_t)gx, (uint32_t)gy);
    return cudaGetLastError();
}

static cudaError_t launch_dist_hll_fast(const DistParams& dp, cudaStream_t st) {
    const uint32_t cb = cell_bytes_of(dp.algo, dp.p);
    const uint32_t chunk = cb < (uint32_t)kHllChunk ? cb : (uint32_t)kHllChunk;
    const size_t smem = (size_t)(kHllTR + kHllTQ) * (chunk + 4) * 4;
    cudaError_t e = cudaFuncSetAttribute(dist_hll_fast_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t rows = dp.row_end - dp.row_begin;
    const uint64_t gy = (rows + kHllTR - 1) / kHllTR;
    uint64_t ncols = dp.n_qry;
    if (dp.triangular && dp.row_end < ncols) ncols = dp.row_end;  // nothing right of the diagonal
    const uint64_t gx = (ncols + kHllTQ - 1) / kHllTQ;
    for (uint64_t y0 = 0; y0 < gy; y0 += 65535) {  // grid.y is limited to 65535: walk row bands
        DistParams q = dp;
        q.row_begin = dp.row_begin + y0 * kHllTR;
        const uint64_t ny = (gy - y0) < 65535 ? (gy - y0) : 65535;
        dist_hll_fast_kernel<<<dim3((unsigned)gx, (unsigned)ny), kHllThreads, smem, st>>>(q, cb, chunk);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

template <int RM>
static cudaError_t launch_dist_hll_int_t(const DistParams& dp, uint32_t cb, uint32_t chunk, uint64_t gx, cudaStream_t st) {
    constexpr int TR = HiShape<RM>::kTR;
    const size_t smem = (size_t)(TR + kHiTQ) * hll_int_stride(chunk) * 4;
    cudaError_t e = cudaFuncSetAttribute(dist_hll_int_kernel<RM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t gy = (dp.row_end - dp.row_begin + TR - 1) / TR;
    for (uint64_t y0 = 0; y0 < gy; y0 += 65535) {  // grid.y is limited to 65535: walk row bands
        DistParams q = dp;
        q.row_begin = dp.row_begin + y0 * TR;
        const uint64_t ny = (gy - y0) < 65535 ? (gy - y0) : 65535;
        dist_hll_int_kernel<RM><<<dim3((unsigned)gx, (unsigned)ny), kHllThreads, smem, st>>>(q, cb, chunk, 1u);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}
static cudaError_t launch_dist_hll_int(const DistParams& dp, cudaStream_t st) {
    const uint32_t cb = cell_bytes_of(dp.algo, dp.p);
    const uint32_t chunk = cb < (uint32_t)kHiChunk ? cb : (uint32_t)kHiChunk;
    const uint64_t rows = dp.row_end - dp.row_begin;
    uint64_t ncols = dp.n_qry;
    if (dp.triangular && dp.row_end < ncols) ncols = dp.row_end;  // nothing right of the diagonal
    const uint64_t gx = (ncols + kHiTQ - 1) / kHiTQ;
    // 64 x 64 tiles once they fill the GPU's 2 x n_sm resident slots about four times over (the triangle keeps half of them)
    uint64_t tiles8 = gx * ((rows + HiShape<8>::kTR - 1) / HiShape<8>::kTR);
    if (dp.triangular) tiles8 /= 2;
    static const int force = [] { const char* v = getenv("LASH_HLL_INT_RM"); return v ? atoi(v) : 0; }();   // A/B measurements
    const bool big = force ? force == 8 : tiles8 >= 8ull * (uint64_t)dp.n_sm;
    return big ? launch_dist_hll_int_t<8>(dp, cb, chunk, gx, st) : launch_dist_hll_int_t<4>(dp, cb, chunk, gx, st);
}

static cudaError_t launch_dist_hmh_fast(const DistParams& dp, cudaStream_t st) {
    const uint64_t rows = dp.row_end - dp.row_begin;
    const uint64_t gy = (rows + kHmhTR - 1) / kHmhTR;
    uint64_t ncols = dp.n_qry;
    if (dp.triangular && dp.row_end < ncols) ncols = dp.row_end;  // nothing right of the diagonal
    const uint64_t gx = (ncols + kHmhTQ - 1) / kHmhTQ;
    for (uint64_t y0 = 0; y0 < gy; y0 += 65535) {  // grid.y is limited to 65535: walk row bands
        DistParams q = dp;
        q.row_begin = dp.row_begin + y0 * kHmhTR;
        const uint64_t ny = (gy - y0) < 65535 ? (gy - y0) : 65535;
        dist_hmh_fast_kernel<<<dim3((unsigned)gx, (unsigned)ny), kHmhThreads, 0, st>>>(q);
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

static cudaError_t launch_dist_ml_tab(const DistParams& dp, cudaStream_t st) {
    const uint32_t cb = cell_bytes_of(dp.algo, dp.p);
    const uint32_t chunk = cb < (uint32_t)kMlTabChunk ? cb : (uint32_t)kMlTabChunk;
    const size_t smem = (size_t)kTabN * kTabN * 12 + 2 * (size_t)kMlTabTR * (chunk + 4) * 4 + (size_t)kMlTabTQ * (chunk + 8) * 2 +
                        (size_t)(kMlTabTR + kMlTabTQ) * 4;
    const bool split = dp.ml_scratch != nullptr;
    const int npl = dp.p <= 11 ? 12 : dp.p <= 15 ? 16 : kMlPlanes;
    using Kern = void (*)(DistParams, uint32_t, uint32_t, uint32_t, uint32_t, uint32_t);
    Kern kern = npl == 12 ? (split ? dist_ml_tab_kernel<12, true> : dist_ml_tab_kernel<12, false>)
              : npl == 16 ? (split ? dist_ml_tab_kernel<16, true> : dist_ml_tab_kernel<16, false>)
                          : (split ? dist_ml_tab_kernel<kMlPlanes, true> : dist_ml_tab_kernel<kMlPlanes, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const uint64_t rows = dp.row_end - dp.row_begin;
    const uint64_t gy = (rows + kMlTabTR - 1) / kMlTabTR;
    uint64_t ncols = dp.n_qry;
    if (dp.triangular && dp.row_end < ncols) ncols = dp.row_end;
    const uint64_t gx = (ncols + kMlTabTQ - 1) / kMlTabTQ;
    if (gx * gy > 0xffffffffull) return cudaErrorInvalidValue;
    if (dp.tile_counter) {
        e = cudaMemsetAsync(dp.tile_counter, 0, 4, st);
        if (e != cudaSuccess) return e;
    }
    const unsigned grid = (unsigned)std::min<uint64_t>(gx * gy, (uint64_t)dp.n_sm);
    kern<<<grid, kMlTabThreads, smem, st>>>(dp, cb, chunk, (uint32_t)gx, (uint32_t)gy, 1u);
    e = cudaGetLastError();
    if (e != cudaSuccess || !split) return e;
    const unsigned fgrid = (unsigned)std::min<uint64_t>((dp.ml_cells + 255) / 256, (uint64_t)dp.n_sm * 32);
    if (npl == 12) ml_finish_kernel<12><<<fgrid, 256, 0, st>>>(dp, cb);
    else if (npl == 16) ml_finish_kernel<16><<<fgrid, 256, 0, st>>>(dp, cb);
    else ml_finish_kernel<kMlPlanes><<<fgrid, 256, 0, st>>>(dp, cb);
    return cudaGetLastError();
}

cudaError_t launch_dist(const DistParams& dp, cudaStream_t st, uint32_t* n_launches) {
    if (dp.row_end <= dp.row_begin || dp.n_qry == 0) return cudaSuccess;
    if (n_launches) *n_launches += 1;
    // LASH_FGRA_KERNEL=merge selects the ALU-merge kernels K4 for ULL (A/B measurements); default: pair-table kernels K4b / K4c
    static const bool fgra_merge = [] {
        const char* v = getenv("LASH_FGRA_KERNEL");
        return v && std::string(v) == "merge";
    }();
    if (dp.algo == ULL && dp.estimator == 0 && !fgra_merge) return launch_dist_fgra_tab(dp, st);
    if (dp.algo == ULL && dp.estimator == 1 && !fgra_merge) {
        if (n_launches && dp.ml_scratch) *n_launches += 1;  // tile kernel + ml_finish_kernel
        return launch_dist_ml_tab(dp, st);
    }
    // LASH_HLL_KERNEL=table selects K4 (LDS.64 table of 2^-r + per-register zero test) for A/B measurements; default K4h
    static const bool hll_table = [] {
        const char* v = getenv("LASH_HLL_KERNEL");
        return v && std::string(v) == "table";
    }();
    // K4h stages with 16-byte loads: register arrays that are not 16-byte aligned (a caller's odd device pointer) use K4
    const bool hll_aligned = (((uintptr_t)dp.ref | (uintptr_t)dp.qry) & 15u) == 0;
    // LASH_HLL_KERNEL=float selects K4h (f64 adds in register order); default K4i (32-bit fixed point, needs the per-sketch
    // min / max the API computes before the launch; 16-register rows as the 16-byte staging loads)
    static const bool hll_float = [] {
        const char* v = getenv("LASH_HLL_KERNEL");
        return v && std::string(v) == "float";
    }();
    if (dp.algo == HLL) {
        if (hll_table || !hll_aligned) return launch_dist_t<HllAcc, 16>(dp, st);
        if (hll_float || !dp.reg_mm_ref || !dp.reg_mm_qry) return launch_dist_hll_fast(dp, st);
        return launch_dist_hll_int(dp, st);
    }
    // LASH_HMH_KERNEL=generic selects K4 for A/B measurements; K4m stages with 16-byte loads like K4h
    static const bool hmh_generic = [] {
        const char* v = getenv("LASH_HMH_KERNEL");
        return v && std::string(v) == "generic";
    }();
    if (dp.algo == HMH) return (hmh_generic || !hll_aligned) ? launch_dist_t<HmhAcc, 16>(dp, st) : launch_dist_hmh_fast(dp, st);
    const bool tiny = dp.p == 3;  // 8 registers per sketch
    if (dp.estimator == 0) return tiny ? launch_dist_t<FgraAcc, 8>(dp, st) : launch_dist_t<FgraAcc, 16>(dp, st);
    return tiny ? launch_dist_t<MlAcc, 8>(dp, st) : launch_dist_t<MlAcc, 16>(dp, st);
}

// K3 for HLL, warp-parallel: the sum of 2^-r over a sketch whose registers lie within 29 levels of its smallest one is exact in
// f64 in any order (dist_tables.cuh, K4i), so the 32 lanes sum their shares as integers and the total equals the reference's
// sequential loop bit for bit; card_kernel<HllAcc> (one lane walking 2^p registers: 67 us per sketch at p = 14, 1.8 ms for C3's
// 10 000 sketches on EVERY rank) remains for sketches outside the window (10^-3 of them), for unaligned arrays and p < 4.
__global__ void __launch_bounds__(kCardWarps * 32) card_hll_int_kernel(const unsigned char* __restrict__ regs, uint64_t n, uint32_t cell_bytes, int p,
                                                                       double* __restrict__ card, uint32_t* flags) {
    const uint32_t lane = threadIdx.x & 31u;
    const uint64_t i = (uint64_t)blockIdx.x * kCardWarps + (threadIdx.x >> 5);
    if (i >= n) return;
    const unsigned char* g = regs + i * cell_bytes;
    const uint4* src = reinterpret_cast<const uint4*>(g);
    const uint32_t n16 = cell_bytes / 16;
    uint32_t mn = 0xffffffffu, mx = 0u;
    for (uint32_t e = lane; e < n16; e += 32) {
        const uint4 v = __ldg(src + e);
        mn = __vminu4(__vminu4(mn, v.x), __vminu4(__vminu4(v.y, v.z), v.w));
        mx = __vmaxu4(__vmaxu4(mx, v.x), __vmaxu4(__vmaxu4(v.y, v.z), v.w));
    }
    mn = __vminu4(mn, mn >> 16), mn = __vminu4(mn, mn >> 8) & 0xffu;
    mx = __vmaxu4(mx, mx >> 16), mx = __vmaxu4(mx, mx >> 8) & 0xffu;
    const uint32_t lo = __reduce_min_sync(0xffffffffu, mn), hi = __reduce_max_sync(0xffffffffu, mx);
    double sum;
    uint32_t zero;
    if (hi - lo <= (uint32_t)kHllIntW) {
        uint64_t s64 = 0;
        uint32_t z = 0;
        for (uint32_t e = lane; e < n16; e += 32) {
            const uint4 v = __ldg(src + e);
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const uint4 t = hll_int_recode(w4[k], lo);
                s64 += (uint64_t)t.x + t.y + t.z + t.w;                                  // 4 * 2^28 per step: no overflow
                z += (uint32_t)__popc(__vcmpeq4(w4[k], 0u)) >> 3;                           // empty registers (0xff per zero byte)
            }
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s64 += __shfl_xor_sync(0xffffffffu, s64, d);
            z += __shfl_xor_sync(0xffffffffu, z, d);
        }
        sum = (double)s64 * __hiloint2double((int)((1023u - (uint32_t)kHllIntW - lo) << 20), 0);
        zero = z;
    } else {
        sum = 0.0;
        zero = 0;
        if (lane == 0) {
            for (uint32_t e = 0; e < cell_bytes; ++e) {   // the reference's loop
                const uint32_t r = g[e];
                zero += r == 0u;
                sum += __hiloint2double((int)(kHllOne - (r << 20)), 0);
            }
        }
    }
    if (lane == 0) {
        bool bias;
        card[i] = hll_len(sum, zero, p, &bias);
        if (bias && flags) atomicAdd(flags, 1u);
    }
}

template <class ACC, int G>
static cudaError_t launch_card_t(int algo, int p, const void* regs, uint64_t n, double* card, uint32_t* flags,
                                 cudaStream_t st) {
    const uint32_t cb = cell_bytes_of(algo, p);
    const unsigned grid = (unsigned)((n + kCardWarps - 1) / kCardWarps);
    const size_t smem = ((ACC::kTableBytes + 15) & ~15) + (size_t)kCardWarps * kCardChunk;
    card_kernel<ACC, G><<<grid, kCardWarps * 32, smem, st>>>(reinterpret_cast<const unsigned char*>(regs), n, cb, p, card, flags);
    return cudaGetLastError();
}

cudaError_t launch_cardinality(int algo, int p, int estimator, const void* regs, uint64_t n, double* card,
                               uint32_t* flags, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (algo == HMH) {
        card_hmh_kernel<<<(unsigned)((n + kCardWarps - 1) / kCardWarps), kCardWarps * 32, 0, st>>>(reinterpret_cast<const uint32_t*>(regs), n, card);
        return cudaGetLastError();
    }
    if (algo == HLL) {
        // LASH_HLL_KERNEL=float|table: the sequential kernel (A/B measurements)
        static const bool seq = [] { const char* v = getenv("LASH_HLL_KERNEL"); return v && (std::string(v) == "float" || std::string(v) == "table"); }();
        const uint32_t cb = cell_bytes_of(algo, p);
        if (!seq && ((uintptr_t)regs & 15u) == 0 && cb % 16 == 0) {
            card_hll_int_kernel<<<(unsigned)((n + kCardWarps - 1) / kCardWarps), kCardWarps * 32, 0, st>>>(reinterpret_cast<const unsigned char*>(regs), n, cb, p, card, flags);
            return cudaGetLastError();
        }
        return launch_card_t<HllAcc, 16>(algo, p, regs, n, card, flags, st);
    }
    const bool tiny = p == 3;
    if (estimator == 0)
        return tiny ? launch_card_t<FgraAcc, 8>(algo, p, regs, n, card, flags, st)
                    : launch_card_t<FgraAcc, 16>(algo, p, regs, n, card, flags, st);
    return tiny ? launch_card_t<MlAcc, 8>(algo, p, regs, n, card, flags, st)
                : launch_card_t<MlAcc, 16>(algo, p, regs, n, card, flags, st);
}

}  // namespace lash
