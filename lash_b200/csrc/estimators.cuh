// Scalar estimator epilogues, evaluated once per pair (or per sketch) by the thread that owns it.
// All arithmetic follows the order of the crates the reference calls (see each function);
// this translation unit is compiled with -fmad=false so no multiply-add is contracted and the only
// operations that are not bit-reproducible against a host libm are pow / log / log1p themselves.
//
//   HLL   streaming_algorithms 0.3.3 HyperLogLog::len()        (reference utils.rs:315,358)
//   ULL   ultraloglog 0.1.6 get_distinct_count_estimate (FGRA) (utils.rs:215,266)
//         ultraloglog 0.1.6 MaximumLikelihoodEstimator          (utils.rs:216,267)
//   HMH   hyperminhash 0.1.4 cardinality()/similarity()         (utils.rs:164)
//   Mash  main.rs:415-423 compute_distance<F>
#pragma once
#include <cmath>
#include <cstdint>

#include "ddmath.cuh"

namespace lash {

struct UllConsts {
    double reg[256];     // REGISTER_CONTRIBUTIONS[i] = eta[i&3] * 2^(-tau*(3+(i>>2)))
    double factor[27];   // ESTIMATION_FACTORS[p] = m^(1+1/tau) / (1 + V(1+tau)/(2m))
    double pow2tau, pow2mtau, pow4mtau, etaX, eta23X, eta13X, eta3012XX, phi1, pinit, minus_inv_tau;
};

constexpr double kUllTau = 0.8194911375910897;
constexpr double kUllV = 0.6118931496978437;
constexpr double kUllEta0 = 4.663135422063788;
constexpr double kUllEta1 = 2.1378502137958524;
constexpr double kUllEta2 = 2.781144650979996;
constexpr double kUllEta3 = 0.9824082545153715;
constexpr double kUllInvSqrtFisher = 0.7608621002725182;
constexpr double kUllMlBias = 0.48147376527720065;

// host side: the table of constants uploaded into c_ull (and used by the CPU-side check of these epilogues, tests/host_shim)
inline UllConsts make_ull_consts() {
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

#ifdef __CUDACC__
__constant__ UllConsts c_ull;  // this header belongs to exactly one translation unit (dist_kernels.cu)

// ---------------------------------------------------------------- Mash distance, main.rs:415-423
// CR: log / pow correctly rounded (ddmath.cuh) -- equal to glibc's except in ~0.01-0.1 % of the arguments, where glibc itself
// is one ulp off the correctly rounded value.  The ML epilogue passes CR = false: its union estimate comes out of a secant
// iteration whose path already depends on the last bit of its starting value, so the extra 6 % of kernel time (measured at
// 100k x 100k) would buy +0.2 % bit-identical cells; everywhere else the cost is noise next to the register loop.
template <bool CR = true>
__device__ __forceinline__ double mash_distance_f64(double frac, int k, int model) {
    const double kk = (double)k;
    if (model == 2) return frac;  // LASH_MODEL_FRAC: what the reference's emit() carries (utils.rs:176,277,364)
    if (model == 1) return fmin(-(CR ? log_cr(frac) : log(frac)) / kk, 1.0);
    return 1.0 - (CR ? pow_cr(frac, 1.0 / kk) : pow(frac, 1.0 / kk));
}
__device__ __forceinline__ float mash_distance_f32(float frac, int k, int model) {
    const float kk = (float)k;
    if (model == 2) return frac;
    if (model == 1) return fminf(-logf(frac) / kk, 1.0f);
    return 1.0f - powf(frac, 1.0f / kk);
}

// ---------------------------------------------------------------- HLL++ len()
__device__ __forceinline__ double pow2neg(uint32_t r) {  // 2^-r, exact
    return __hiloint2double((int)((1023u - r) << 20), 0);
}
__device__ inline double hll_threshold(int p) {
    switch (p) {
        case 4: return 10; case 5: return 20; case 6: return 40; case 7: return 80; case 8: return 220;
        case 9: return 400; case 10: return 900; case 11: return 1800; case 12: return 3100; case 13: return 6500;
        case 14: return 11500; case 15: return 20000; case 16: return 50000; case 17: return 120000;
        default: return 350000;
    }
}
__device__ inline double hll_alpha(int p) {
    if (p == 4) return 0.673;
    if (p == 5) return 0.697;
    if (p == 6) return 0.709;
    return 0.7213 / (1.0 + 1.079 / (double)(1ull << p));
}
// returns NaN and sets *bias when the estimate lands in the bias-table regime (see lash_gpu.h)
__device__ inline double hll_len(double sum, uint32_t zero, int p, bool* bias) {
    const double m = (double)(1ull << p);
    *bias = false;
    if (zero > 0) {
        double h = m * log_cr(m / (double)zero);
        if (h <= hll_threshold(p)) return h;
    }
    double e = hll_alpha(p) * (m * m) / sum;
    if (e <= 5.0 * m) {
        *bias = true;
        return __longlong_as_double(0x7ff8000000000000LL);
    }
    return e;
}

// ---------------------------------------------------------------- ULL FGRA
__device__ __forceinline__ double ull_psi_prime(double z, double z2) {
    return (z + c_ull.eta23X) * (z2 + c_ull.eta13X) + c_ull.eta3012XX;
}
__device__ inline double ull_sigma(double z) {
    if (z <= 0.0) return kUllEta3;
    if (z >= 1.0) return __longlong_as_double(0x7ff0000000000000LL);
    double powZ = z, nextPowZ = z * z, s = 0.0, powTau = c_ull.etaX;
    for (;;) {
        double oldS = s;
        double nn = nextPowZ * nextPowZ;
        s += powTau * (powZ - nextPowZ) * ull_psi_prime(nextPowZ, nn);
        if (!(s > oldS)) return s / z;
        powZ = nextPowZ;
        nextPowZ = nn;
        powTau *= c_ull.pow2tau;
    }
}
__device__ inline double ull_phi(double z, double zSquare) {
    if (z <= 0.0) return 0.0;
    if (z >= 1.0) return c_ull.phi1;
    double previousPowZ = zSquare, powZ = z, nextPowZ = sqrt(powZ);
    double pp = c_ull.pinit / (1.0 + nextPowZ);
    double ps = ull_psi_prime(powZ, previousPowZ);
    double s = nextPowZ * (ps + ps) * pp;
    for (;;) {
        previousPowZ = powZ;
        powZ = nextPowZ;
        double oldS = s;
        nextPowZ = sqrt(powZ);
        double nextPs = ull_psi_prime(powZ, previousPowZ);
        pp *= c_ull.pow2mtau / (1.0 + nextPowZ);
        s += nextPowZ * ((nextPs + nextPs) - (powZ + nextPowZ) * ps) * pp;
        if (!(s > oldS)) return s;
        ps = nextPs;
    }
}
// cnt[0..3] = c0,c4,c8,c10 (registers below 4p+4), cnt[4..7] = registers 252..255
__device__ inline double ull_fgra_finalize(double sum, const uint32_t* cnt, int p) {
    const long long m = 1ll << p;
    const long long c0 = cnt[0], c4 = cnt[1], c8 = cnt[2], c10 = cnt[3];
    const long long w0 = cnt[4], w1 = cnt[5], w2 = cnt[6], w3 = cnt[7];
    if (c0 > 0 || c4 > 0 || c8 > 0 || c10 > 0) {
        double alpha = (double)(m + 3 * (c0 + c4 + c8 + c10));
        double beta = (double)(m - c0 - c4);
        double gamma = (double)(4 * c0 + 2 * c4 + 3 * c8 + c10);
        double q = (sqrt(beta * beta + 4.0 * alpha * gamma) - beta) / (2.0 * alpha);
        double rz = q * q;
        double z = rz * rz;
        if (c0 > 0) sum += (double)c0 * ull_sigma(z);
        if (c4 > 0) sum += (double)c4 * (c_ull.pow2mtau * c_ull.etaX) * ull_psi_prime(z, z * z);
        if (c8 > 0) sum += (double)c8 * (z * (c_ull.pow4mtau * (kUllEta0 - kUllEta1)) + c_ull.pow4mtau * kUllEta1);
        if (c10 > 0) sum += (double)c10 * (z * (c_ull.pow4mtau * (kUllEta2 - kUllEta3)) + c_ull.pow4mtau * kUllEta3);
    }
    if (w0 > 0 || w1 > 0 || w2 > 0 || w3 > 0) {
        double c = (double)(w0 + w1 + w2 + w3);
        double alpha = (double)m + 3.0 * c;
        double beta = (double)(w0 + w1 + 2 * (w2 + w3));
        double gamma = (double)(m + 2 * w0 + w2 - w3);
        double z = sqrt((sqrt(beta * beta + 4.0 * alpha * gamma) - beta) / (2.0 * alpha));
        double rz = sqrt(z);
        double s = ull_phi(rz, z) * c;
        s += z * (1.0 + rz) * ((double)w0 * kUllEta0 + (double)w1 * kUllEta1 + (double)w2 * kUllEta2 + (double)w3 * kUllEta3);
        s += rz * ((double)(w0 + w1) * (z * (c_ull.pow2mtau * (kUllEta0 - kUllEta2)) + c_ull.pow2mtau * kUllEta2) +
                   (double)(w2 + w3) * (z * (c_ull.pow2mtau * (kUllEta1 - kUllEta3)) + c_ull.pow2mtau * kUllEta3));
        sum += s * pow_cr(c_ull.pow2mtau, (double)(65 - p)) / ((1.0 + rz) * (1.0 + z));
    }
    return c_ull.factor[p] * pow_cr(sum, c_ull.minus_inv_tau);
}

// ---------------------------------------------------------------- ULL ML (Ertl 2017 Alg. 8 as in hash4j)
// 2^k for |k| <= 1022, built from bits.  x * pow2i(k) is ldexp(x, k) exactly (one correctly rounded operation either
// way); the library ldexp / ilogb are calls with denormal handling and showed up in the per-pair epilogue.
__device__ __forceinline__ double pow2i(int k) { return __hiloint2double((1023 + k) << 20, 0); }
__device__ __forceinline__ int ilogb_pos(double x) {
    const int e = (__double2hiint(x) >> 20) & 0x7ff;
    return (e == 0 || e == 0x7ff) ? ilogb(x) : e - 1023;  // normal numbers: the exponent field
}
__device__ inline double ull_solve_ml(double a, const int* b, int n, double eps) {
    if (a == 0.0) return __longlong_as_double(0x7ff0000000000000LL);
    int kMax = n;
    while (kMax >= 0 && b[kMax] == 0) --kMax;
    if (kMax < 0) return 0.0;
    int kMin = kMax;
    long long s1 = b[kMax];
    double s2 = (double)b[kMax] * pow2i(kMax);
    for (int k = kMax - 1; k >= 0; --k) {
        int t = b[k];
        if (t > 0) {
            s1 += t;
            s2 += (double)t * pow2i(k);
            kMin = k;
        }
    }
    double gPrev = 0.0, x;
    if (s2 <= 1.5 * a)
        x = (double)s1 / (0.5 * s2 + a);
    else
        x = log1p(s2 / a) * ((double)s1 / s2);
    double dx = x;
    while (dx > x * eps) {
        int kappa = ilogb_pos(x) + 2;
        int sh = (kMax > kappa ? kMax : kappa) + 1;
        double xp = x * pow2i(-sh);
        double xp2 = xp * xp;
        double h = xp - xp2 / 3.0 + (xp2 * xp2) * (1.0 / 45.0 - xp2 / 472.5);
        for (int k = kappa - 1; k >= kMax; --k) {
            double hp = 1.0 - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
        }
        double g = (double)b[kMax] * h;
        for (int k = kMax - 1; k >= kMin; --k) {
            double hp = 1.0 - h;
            h = (xp + h * hp) / (xp + hp);
            xp += xp;
            g += (double)b[k] * h;
        }
        g += x * a;
        if (gPrev < g && g <= (double)s1)
            dx *= (g - (double)s1) / (gPrev - g);
        else
            dx = 0.0;
        x += dx;
        gPrev = g;
    }
    return x;
}
// S: wrapping sum of the per-register contributions; b[0..65]: bit statistics; reg0: register 0
__device__ inline double ull_ml_finalize(uint64_t S, int* b, int p, uint32_t reg0) {
    if (S == 0) return reg0 == 0 ? 0.0 : __longlong_as_double(0x7ff0000000000000LL);
    b[63 - p] += b[64 - p];
    const double m = (double)(1ull << p);
    const double factor = m + m;
    const double a = __ull2double_rn(S) * factor * 0x1p-64;
    const double eps = 1e-3 * kUllInvSqrtFisher / sqrt(m);
    return factor * ull_solve_ml(a, b, 63 - p, eps) / (1.0 + kUllMlBias / m);
}

// ---------------------------------------------------------------- HyperMinHash
__device__ inline double hmh_beta(double ez) {
    double zl = log_cr(ez + 1.0);
    return -0.370393911 * ez + 0.070471823 * zl + 0.17393686 * pow_cr(zl, 2.0) + 0.16339839 * pow_cr(zl, 3.0) +
           -0.09237745 * pow_cr(zl, 4.0) + 0.03738027 * pow_cr(zl, 5.0) + -0.005384159 * pow_cr(zl, 6.0) +
           0.00042419 * pow_cr(zl, 7.0);
}
__device__ inline double hmh_cardinality_from(double sum, double ez) {
    const double M = 16384.0;
    const double alpha = 0.7213 / (1.0 + 1.079 / M);
    return alpha * M * (M - ez) / (hmh_beta(ez) + sum);
}
// expectedCollision(n, m) of hyperminhash for n <= 2^19 is a 64 x 1024 double loop  x += prx(i, j; n) * pry(i, j; m)  with
//     prx(i, j; n) = (1 - b2)^n - (1 - b1)^n,   b1 = (1024 + j) / 2^(24 + i),  b2 = (1025 + j) / 2^(24 + i)     (i < 64)
// (row i = 64 uses b1 = j / 2^87, b2 = (j + 1) / 2^87).  The term depends on ONE cardinality, so it is a per-sketch vector;
// and from i = 42 on b <= 2049 / 2^66 < 2^-54 makes 1 - b == 1.0 in double, the term exactly 0 and the product an exact +0
// that leaves x untouched -- rows 1..41 are all that can matter (row 41 still holds one non-zero term, j = 1024;
// tests/test_device_math.py checks the zero rows).  K4m's small-sketch path stores the 41 x 1024 vector of every small
// sketch once (hmh_ec_fill_kernel) and sums the products of two vectors in the reference's (i, j) order
// (hmh_ec_gemm_kernel); the per-pair loop below is the same arithmetic for pairs without a stored vector.
constexpr int kHmhEcRows = 41;
constexpr int kHmhEcLen = kHmhEcRows * 1024;
constexpr double kHmhEcSmall = 524288.0;   // 2^(P + 5): above it the closed form applies

__device__ __forceinline__ double hmh_ec_term(int i, int j, double n) {   // i in 1..63, j in 1..1024
    const double inv_den = __hiloint2double((1023 - (24 + i)) << 20, 0);    // 2^-(24+i), exact; x / 2^k == x * 2^-k exactly
    const double b1 = (1024.0 + j) * inv_den;
    const double b2 = (1024.0 + j + 1.0) * inv_den;
    return pow_cr(1.0 - b2, n) - pow_cr(1.0 - b1, n);
}
// Early end of the loop, exact: every term is >= 0, so the sum x only grows, and an addition x + t with t < ulp(x) / 2 rounds
// back to x.  bound_n / bound_m = the largest |term| of the rows still to come of the two sketches; their product bounds every
// remaining product before rounding (after rounding at most one part in 2^53 more), and x * 2^-54 <= ulp(x) / 4.
__device__ __forceinline__ bool hmh_ec_rest_is_absorbed(double bound_n, double bound_m, double x) { return bound_n * bound_m < x * 0x1p-54; }
// which branch expectedCollision takes: true = the double loop
__device__ __forceinline__ bool hmh_ec_is_small(double card) { return !(card > kHmhEcSmall); }

// x_pre: the loop sum computed elsewhere (both sketches small), or nullptr
__device__ inline double hmh_expected_collisions(double n, double m, const double* x_pre = nullptr) {
    if (n < m) { double t = n; n = m; m = t; }
    if (n > 0x1p74) return 18446744073709551615.0;
    if (n > kHmhEcSmall) {
        double r = (1.0 + n) / m;
        double d = (4.0 * n / m) / pow_cr(r, 2.0);
        return 0.169919487159739093975315012348 * 16.0 * d + 0.5;
    }
    double x = 0.0;
    if (x_pre) {
        x = *x_pre;
    } else {
        for (int i = 1; i <= kHmhEcRows; ++i)
            for (int j = 1; j <= 1024; ++j) x += hmh_ec_term(i, j, n) * hmh_ec_term(i, j, m);
    }
    return (x * 14.0 + 0.5) / 14.0;
}
__device__ inline double hmh_similarity_from(uint32_t C, uint32_t N, double card_q, double card_r, const double* x_pre = nullptr) {
    if (C == 0) return 0.0;
    double ec = hmh_expected_collisions(card_q, card_r, x_pre);
    if ((double)C < ec) return 0.0;
    return ((double)C - ec) / (double)N;
}
#endif

}  // namespace lash
