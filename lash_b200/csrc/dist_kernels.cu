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
#include <cmath>
#include <type_traits>

#include "estimators.cuh"
#include "kernels.h"
#include "registers.cuh"

namespace lash {


constexpr int kDistThreads = 256;
constexpr int kChunkWords = 256;  // 1 KiB of registers per sketch per stage

// ------------------------------------------------------------------------------------------------
// per-pair accumulators
// ------------------------------------------------------------------------------------------------
struct SharedTables {
    double* fgra_tab;    // [256] contribution of a merged register byte (0 outside [4p+4, 252))
    uint64_t* ml_ret;    // [256] ML alpha contribution (scaled by 2^64) of a register byte
};

struct HllAcc {
    static constexpr int RM = 2, QM = 2;
    static constexpr int kTableBytes = 0;
    double sum;
    uint32_t zero;
    __device__ __forceinline__ void init() { sum = 0.0; zero = 0; }
    __device__ __forceinline__ void add(uint32_t a, uint32_t b, const SharedTables&, int) {
        const uint32_t m = __vmaxu4(a, b);
        zero += __popc(__vcmpeq4(m, 0u)) >> 3;
        sum += pow2neg(m & 0xffu);
        sum += pow2neg((m >> 8) & 0xffu);
        sum += pow2neg((m >> 16) & 0xffu);
        sum += pow2neg(m >> 24);
    }
};

struct FgraAcc {
    static constexpr int RM = 2, QM = 2;
    static constexpr int kTableBytes = 256 * 8;
    double sum;
    uint32_t c0, c4, c8, c10, w0, w1, w2, w3;
    __device__ __forceinline__ void init() { sum = 0.0; c0 = c4 = c8 = c10 = w0 = w1 = w2 = w3 = 0; }
    __device__ __forceinline__ void classify(uint32_t r, int off) {
        const int r2 = (int)r - off;
        c0 += (r2 < -8);
        c4 += (r2 == -8);
        c8 += (r2 == -4);
        c10 += (r2 == -2);
        w0 += (r == 252u);
        w1 += (r == 253u);
        w2 += (r == 254u);
        w3 += (r == 255u);
    }
    __device__ __forceinline__ void add(uint32_t a, uint32_t b, const SharedTables& t, int p) {
        const uint32_t m = ull_merge4(a, b);
        sum += t.fgra_tab[m & 0xffu];
        sum += t.fgra_tab[(m >> 8) & 0xffu];
        sum += t.fgra_tab[(m >> 16) & 0xffu];
        sum += t.fgra_tab[m >> 24];
        // registers outside [4p+4, 252) contribute through counts, not through the table
        const uint32_t off = 4u * p + 4u;
        const uint32_t rel = __vsub4(m, off * 0x01010101u);
        if (__vcmpgeu4(rel, (252u - off) * 0x01010101u)) {
            classify(m & 0xffu, (int)off);
            classify((m >> 8) & 0xffu, (int)off);
            classify((m >> 16) & 0xffu, (int)off);
            classify(m >> 24, (int)off);
        }
    }
};

struct MlAcc {
    static constexpr int RM = 1, QM = 1;
    static constexpr int kTableBytes = 256 * 8;
    uint64_t S;
    int b[66];
    __device__ __forceinline__ void init() {
        S = 0;
#pragma unroll 1
        for (int i = 0; i < 66; ++i) b[i] = 0;
    }
    __device__ __forceinline__ void one(uint32_t r, const SharedTables& t, int off) {
        S += t.ml_ret[r];
        const int r2 = (int)r - off;
        if (r2 >= 0) {
            const int k = r2 >> 2;
            b[k] += (int)(r & 1u);
            b[k + 1] += (int)((r >> 1) & 1u);
            b[k + 2] += 1;
        } else {
            if (r2 == -2 || r2 == -8) b[0] += 1;
            if (r2 == -2 || r2 == -4) b[1] += 1;
        }
    }
    __device__ __forceinline__ void add(uint32_t a, uint32_t bb, const SharedTables& t, int p) {
        const uint32_t m = ull_merge4(a, bb);
        const int off = 4 * p + 4;
        one(m & 0xffu, t, off);
        one((m >> 8) & 0xffu, t, off);
        one((m >> 16) & 0xffu, t, off);
        one(m >> 24, t, off);
    }
};

struct HmhAcc {
    static constexpr int RM = 2, QM = 2;
    static constexpr int kTableBytes = 0;
    uint32_t C, N;
    __device__ __forceinline__ void init() { C = N = 0; }
    __device__ __forceinline__ void add(uint32_t a, uint32_t b, const SharedTables&, int) {
        const uint32_t eq = __vcmpeq2(a, b) & __vcmpne2(a, 0u);
        C += __popc(eq) >> 4;
        N += __popc(__vcmpne2(a | b, 0u)) >> 4;
    }
};

// ML alpha contribution of a register byte (hash4j contribute(), scaled by 2^64)
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

template <class ACC>
__device__ __forceinline__ void build_tables(SharedTables& t, unsigned char* smem_tab, int p) {
    t.fgra_tab = reinterpret_cast<double*>(smem_tab);
    t.ml_ret = reinterpret_cast<uint64_t*>(smem_tab);
    if (ACC::kTableBytes == 0) return;
    const int off = 4 * p + 4;
    for (int r = threadIdx.x; r < 256; r += blockDim.x) {
        if constexpr (std::is_same<ACC, FgraAcc>::value) {
            t.fgra_tab[r] = (r >= off && r < 252) ? c_ull.reg[r - off] : 0.0;
        } else {
            t.ml_ret[r] = ml_ret_of((uint32_t)r, p);
        }
    }
}

// union cardinality from a finished accumulator
__device__ __forceinline__ double finish_union(HllAcc& a, int p, uint32_t, bool* bias) { return hll_len(a.sum, a.zero, p, bias); }
__device__ __forceinline__ double finish_union(FgraAcc& a, int p, uint32_t, bool* bias) {
    *bias = false;
    uint32_t cnt[8] = {a.c0, a.c4, a.c8, a.c10, a.w0, a.w1, a.w2, a.w3};
    return ull_fgra_finalize(a.sum, cnt, p);
}
__device__ __forceinline__ double finish_union(MlAcc& a, int p, uint32_t reg0, bool* bias) {
    *bias = false;
    return ull_ml_finalize(a.S, a.b, p, reg0);
}

template <class ACC>
struct IsHmh { static constexpr bool value = false; };
template <>
struct IsHmh<HmhAcc> { static constexpr bool value = true; };

// ------------------------------------------------------------------------------------------------
// K4: distance tiles
// ------------------------------------------------------------------------------------------------
template <class ACC>
__global__ void __launch_bounds__(kDistThreads) dist_kernel(DistParams dp, uint32_t cell_words, uint32_t chunk_words) {
    constexpr int RM = ACC::RM, QM = ACC::QM;
    constexpr int TR = 16 * RM, TQ = 16 * QM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables tabs;
    build_tables<ACC>(tabs, smem_raw, dp.p);
    uint32_t* sref = reinterpret_cast<uint32_t*>(smem_raw + ACC::kTableBytes);
    const uint32_t stride = chunk_words + 1;
    uint32_t* sqry = sref + TR * stride;

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

    const uint32_t* gref = reinterpret_cast<const uint32_t*>(dp.ref);
    const uint32_t* gqry = reinterpret_cast<const uint32_t*>(dp.qry);

    for (uint32_t c0 = 0; c0 < cell_words; c0 += chunk_words) {
        __syncthreads();  // previous chunk fully consumed (also orders the table build)
        for (uint32_t e = threadIdx.x; e < (uint32_t)TR * chunk_words; e += kDistThreads) {
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gi = row0 + r;
            sref[r * stride + w] = gi < dp.row_end ? __ldg(gref + gi * cell_words + c0 + w) : 0u;
        }
        for (uint32_t e = threadIdx.x; e < (uint32_t)TQ * chunk_words; e += kDistThreads) {
            const uint32_t r = e / chunk_words, w = e % chunk_words;
            const uint64_t gj = col0 + r;
            sqry[r * stride + w] = gj < dp.n_qry ? __ldg(gqry + gj * cell_words + c0 + w) : 0u;
        }
        __syncthreads();
        const uint32_t* pr = sref + ty * stride;
        const uint32_t* pq = sqry + tx * stride;
#pragma unroll 2
        for (uint32_t w = 0; w < chunk_words; ++w) {
            uint32_t ra[RM], qb[QM];
#pragma unroll
            for (int a = 0; a < RM; ++a) ra[a] = pr[a * 16 * stride + w];
#pragma unroll
            for (int b = 0; b < QM; ++b) qb[b] = pq[b * 16 * stride + w];
#pragma unroll
            for (int a = 0; a < RM; ++a)
#pragma unroll
                for (int b = 0; b < QM; ++b) acc[a][b].add(ra[a], qb[b], tabs, dp.p);
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
                double sim = hmh_similarity_from(acc[a][b].C, acc[a][b].N, dp.card_qry[j], dp.card_ref[i]);
                s = fmax(sim, 0.0);
            } else {
                bool bias;
                // register 0 of the union is only needed by ML's S == 0 corner (all-empty vs saturated)
                uint32_t reg0 = 0;
                if constexpr (std::is_same<ACC, MlAcc>::value) {
                    uint32_t ra0 = __ldg(gref + i * cell_words) & 0xffu, rb0 = __ldg(gqry + j * cell_words) & 0xffu;
                    reg0 = ull_merge1(ra0, rb0);
                }
                const double U = finish_union(acc[a][b], dp.p, reg0, &bias);
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
// K3: per-sketch cardinality (utils.rs:213-219, 314-316; hyperminhash cardinality())
// One thread per sketch, registers walked in index order with the same accumulators as K4
// (the union of a sketch with itself is the sketch).
// ------------------------------------------------------------------------------------------------
template <class ACC>
__global__ void __launch_bounds__(128) card_kernel(const uint32_t* __restrict__ regs, uint64_t n, uint32_t cell_words, int p,
                                                   double* __restrict__ card, uint32_t* flags) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SharedTables tabs;
    build_tables<ACC>(tabs, smem_raw, p);
    __syncthreads();
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* g = regs + i * cell_words;
    ACC acc;
    acc.init();
    for (uint32_t w = 0; w < cell_words; ++w) {
        const uint32_t v = __ldg(g + w);
        acc.add(v, v, tabs, p);
    }
    bool bias;
    card[i] = finish_union(acc, p, __ldg(g) & 0xffu, &bias);
    if (bias && flags) atomicAdd(flags, 1u);
}

__global__ void __launch_bounds__(128) card_hmh_kernel(const uint32_t* __restrict__ regs, uint64_t n, double* __restrict__ card) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t* g = regs + i * 8192u;
    double sum = 0.0, ez = 0.0;
    for (uint32_t w = 0; w < 8192u; ++w) {
        const uint32_t v = __ldg(g + w);
        const uint32_t l0 = (v & 0xffffu) >> 10, l1 = v >> 26;
        if (l0 == 0) ez += 1.0;
        sum += pow2neg(l0);
        if (l1 == 0) ez += 1.0;
        sum += pow2neg(l1);
    }
    card[i] = hmh_cardinality_from(sum, ez);
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static UllConsts make_ull_consts() {
    UllConsts c;
    c.pow2tau = std::pow(2.0, kUllTau);
    c.pow2mtau = std::pow(2.0, -kUllTau);
    c.pow4mtau = std::pow(4.0, -kUllTau);
    c.etaX = kUllEta0 - kUllEta1 - kUllEta2 + kUllEta3;
    c.eta23X = (kUllEta2 - kUllEta3) / c.etaX;
    c.eta13X = (kUllEta1 - kUllEta3) / c.etaX;
    c.eta3012XX = (kUllEta3 * kUllEta0 - kUllEta1 * kUllEta2) / (c.etaX * c.etaX);
    c.phi1 = kUllEta0 / (c.pow2tau * (2.0 * c.pow2tau - 1.0));
    c.pinit = c.etaX * (c.pow4mtau / (2.0 - c.pow2mtau));
    c.minus_inv_tau = -1.0 / kUllTau;
    const double eta[4] = {kUllEta0, kUllEta1, kUllEta2, kUllEta3};
    for (int i = 0; i < 256; ++i) c.reg[i] = eta[i & 3] * std::pow(2.0, -kUllTau * (double)(3 + (i >> 2)));
    for (int p = 0; p < 27; ++p) {
        double m = (double)(1ull << p);
        c.factor[p] = m * std::pow(m, 1.0 / kUllTau) / (1.0 + kUllV * (1.0 + kUllTau) / (2.0 * m));
    }
    return c;
}

cudaError_t ensure_tables() {
    // constant memory is per device (per context); upload is cheap, so do it whenever asked
    static const UllConsts host = make_ull_consts();
    return cudaMemcpyToSymbol(c_ull, &host, sizeof(UllConsts));
}

static uint32_t cell_words_of(int algo, int p) { return algo == HMH ? 8192u : (p >= 2 ? (1u << p) / 4u : 1u); }

template <class ACC>
static cudaError_t launch_dist_t(const DistParams& dp, cudaStream_t st) {
    constexpr int TR = 16 * ACC::RM, TQ = 16 * ACC::QM;
    const uint32_t cw = cell_words_of(dp.algo, dp.p);
    const uint32_t chunk = cw < (uint32_t)kChunkWords ? cw : (uint32_t)kChunkWords;
    const size_t smem = ACC::kTableBytes + (size_t)(TR + TQ) * (chunk + 1) * 4;
    auto kern = dist_kernel<ACC>;
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
        kern<<<grid, kDistThreads, smem, st>>>(q, cw, chunk);
        e = cudaGetLastError();
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

cudaError_t launch_dist(const DistParams& dp, cudaStream_t st, uint32_t* n_launches) {
    if (dp.row_end <= dp.row_begin || dp.n_qry == 0) return cudaSuccess;
    if (n_launches) *n_launches += 1;
    if (dp.algo == HLL) return launch_dist_t<HllAcc>(dp, st);
    if (dp.algo == HMH) return launch_dist_t<HmhAcc>(dp, st);
    if (dp.estimator == 0) return launch_dist_t<FgraAcc>(dp, st);
    return launch_dist_t<MlAcc>(dp, st);
}

template <class ACC>
static cudaError_t launch_card_t(int algo, int p, const void* regs, uint64_t n, double* card, uint32_t* flags,
                                 cudaStream_t st) {
    const uint32_t cw = cell_words_of(algo, p);
    const unsigned grid = (unsigned)((n + 127) / 128);
    card_kernel<ACC><<<grid, 128, ACC::kTableBytes, st>>>(reinterpret_cast<const uint32_t*>(regs), n, cw, p, card, flags);
    return cudaGetLastError();
}

cudaError_t launch_cardinality(int algo, int p, int estimator, const void* regs, uint64_t n, double* card,
                               uint32_t* flags, cudaStream_t st) {
    if (n == 0) return cudaSuccess;
    if (algo == HMH) {
        card_hmh_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(reinterpret_cast<const uint32_t*>(regs), n, card);
        return cudaGetLastError();
    }
    if (algo == HLL) return launch_card_t<HllAcc>(algo, p, regs, n, card, flags, st);
    if (estimator == 0) return launch_card_t<FgraAcc>(algo, p, regs, n, card, flags, st);
    return launch_card_t<MlAcc>(algo, p, regs, n, card, flags, st);
}

}  // namespace lash
