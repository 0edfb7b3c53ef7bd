// Test infrastructure: the estimator epilogues of the distance kernels (lash_b200/csrc/estimators.cuh: HLL++ len(), ULL FGRA
// finalisation, Ertl's ML solver, HyperMinHash cardinality / similarity, Mash distance) compiled with g++ so that they can be
// checked against the oracle on a machine without a GPU (tests/test_device_math.py).  Nothing here is part of the product.
#include <cmath>
#include <cstdint>
#include <cstring>

#define __CUDACC__ 1
#define __device__
#define __host__
#define __constant__
#define __forceinline__ inline

static inline double __hiloint2double(int hi, int lo) {
    const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo;
    double d;
    memcpy(&d, &b, 8);
    return d;
}
static inline int __double2hiint(double d) {
    uint64_t b;
    memcpy(&b, &d, 8);
    return (int)(b >> 32);
}
static inline double __longlong_as_double(long long v) {
    double d;
    memcpy(&d, &v, 8);
    return d;
}
static inline double __ull2double_rn(unsigned long long v) { return (double)v; }  // round to nearest even, like cvt.rn.f64.u64

#include "../../lash_b200/csrc/estimators.cuh"

using namespace lash;

namespace {
struct Init {
    Init() { c_ull = make_ull_consts(); }
} init_once;
}  // namespace

extern "C" {
double dm_ull_reg(int i) { return c_ull.reg[i]; }
double dm_hll_len(double sum, uint32_t zero, int p, int* bias) {
    bool b = false;
    const double r = hll_len(sum, zero, p, &b);
    *bias = b ? 1 : 0;
    return r;
}
double dm_fgra(double sum, const uint32_t* cnt, int p) { return ull_fgra_finalize(sum, cnt, p); }
double dm_ml(uint64_t S, const int* b_in, int p, uint32_t reg0) {
    int b[66];
    for (int i = 0; i < 66; ++i) b[i] = b_in[i];
    return ull_ml_finalize(S, b, p, reg0);
}
double dm_hmh_card(double sum, double ez) { return hmh_cardinality_from(sum, ez); }
double dm_hmh_similarity(uint32_t C, uint32_t N, double card_q, double card_r) { return hmh_similarity_from(C, N, card_q, card_r); }
double dm_hmh_ec_term(int i, int j, double n) { return hmh_ec_term(i, j, n); }
double dm_hmh_ec(double n, double m) { return hmh_expected_collisions(n, m); }
int dm_hmh_ec_rows(void) { return kHmhEcRows; }
// the tile product's early end (hmh_ec_gemm_kernel): full 41-row sum in the reference's order, and the sum at the first row
// boundary where hmh_ec_rest_is_absorbed holds; returns that row (41 = never), both sums through the pointers
int dm_hmh_ec_early(double n, double m, double* full, double* early) {
    static thread_local double tn[kHmhEcLen], tm[kHmhEcLen];
    double sn[kHmhEcRows + 1], sm[kHmhEcRows + 1];
    for (int i = 0; i < kHmhEcRows; ++i)
        for (int j = 0; j < 1024; ++j) {
            tn[i * 1024 + j] = hmh_ec_term(i + 1, j + 1, n);
            tm[i * 1024 + j] = hmh_ec_term(i + 1, j + 1, m);
        }
    sn[kHmhEcRows] = sm[kHmhEcRows] = 0.0;
    for (int i = kHmhEcRows - 1; i >= 0; --i) {
        double a = 0.0, b = 0.0;
        for (int j = 0; j < 1024; ++j) {
            a = fmax(a, fabs(tn[i * 1024 + j]));
            b = fmax(b, fabs(tm[i * 1024 + j]));
        }
        sn[i] = fmax(sn[i + 1], a);
        sm[i] = fmax(sm[i + 1], b);
    }
    double x = 0.0;
    int stop = kHmhEcRows;
    *early = 0.0;
    for (int i = 0; i < kHmhEcRows; ++i) {
        if (i && stop == kHmhEcRows && hmh_ec_rest_is_absorbed(sn[i], sm[i], x)) {
            stop = i;
            *early = x;
        }
        for (int j = 0; j < 1024; ++j) x = x + tn[i * 1024 + j] * tm[i * 1024 + j];
    }
    if (stop == kHmhEcRows) *early = x;
    *full = x;
    return stop;
}
double dm_mash64(double frac, int k, int model) { return mash_distance_f64(frac, k, model); }
float dm_mash32(float frac, int k, int model) { return mash_distance_f32(frac, k, model); }
}
