// Correctly rounded pow(x, y) for the FGRA epilogue, in double-double arithmetic.
//
// Why: every estimator SUM of the distance path is bit-identical to the scalar CPU order by construction; what is left
// between the GPU's distances and the reference's is libm -- CUDA's pow is faithful to ~1-2 ulp, glibc's to 0.52 ulp, and
// a 1-ulp difference in U = factor * sum^(-1/tau) is amplified by s = (a + b - U) / U as 1/s.  A pow that is
// correctly rounded (error < 2^-70, rounded once) agrees with glibc's except where the exact value lies within 0.02 ulp of
// a rounding boundary, so U, a and b become bit-identical to the oracle's in all but a fraction of a percent of the cases.
//
// pow(x, y) = exp(y * log(x)):  x = 2^e * m, m in [1/sqrt2, sqrt2);  log(m) = 2 atanh(z), z = (m-1)/(m+1), as a double-double
// series (first two terms in double-double, the tail in double);  t = y * (e ln2 + log m) in double-double;  t = k ln2 + r;
// exp(r) = exp(r/16)^16 with exp(r/16) a double-double Taylor sum (three terms in double-double, the tail in double).
// Pure arithmetic (+, *, /, fma): the same code runs on the host (tests/host_shim) and is checked there against mpmath
// (correct rounding) and glibc.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace lash {

#ifndef LASH_DD_FN
#ifdef __CUDACC__
#define LASH_DD_FN __host__ __device__ __forceinline__
#else
#define LASH_DD_FN inline
#endif
#endif

struct dd {
    double hi, lo;
};

LASH_DD_FN dd two_sum(double a, double b) {
    const double s = a + b, bb = s - a;
    return dd{s, (a - (s - bb)) + (b - bb)};
}
LASH_DD_FN dd quick_two_sum(double a, double b) {  // |a| >= |b|
    const double s = a + b;
    return dd{s, b - (s - a)};
}
LASH_DD_FN dd two_prod(double a, double b) {
    const double p = a * b;
    return dd{p, fma(a, b, -p)};
}
LASH_DD_FN dd dd_add(dd a, dd b) {
    dd s = two_sum(a.hi, b.hi);
    const dd t = two_sum(a.lo, b.lo);
    s.lo += t.hi;
    s = quick_two_sum(s.hi, s.lo);
    s.lo += t.lo;
    return quick_two_sum(s.hi, s.lo);
}
LASH_DD_FN dd dd_add_d(dd a, double b) {
    dd s = two_sum(a.hi, b);
    s.lo += a.lo;
    return quick_two_sum(s.hi, s.lo);
}
LASH_DD_FN dd dd_neg(dd a) { return dd{-a.hi, -a.lo}; }
LASH_DD_FN dd dd_mul(dd a, dd b) {
    dd p = two_prod(a.hi, b.hi);
    p.lo = fma(a.hi, b.lo, fma(a.lo, b.hi, p.lo));
    return quick_two_sum(p.hi, p.lo);
}
LASH_DD_FN dd dd_mul_d(dd a, double b) {
    dd p = two_prod(a.hi, b);
    p.lo = fma(a.lo, b, p.lo);
    return quick_two_sum(p.hi, p.lo);
}
LASH_DD_FN dd dd_div(dd a, dd b) {
    const double q1 = a.hi / b.hi;
    dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
    const double q2 = r.hi / b.hi;
    r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
    const double q3 = r.hi / b.hi;
    return dd_add_d(quick_two_sum(q1, q2), q3);
}

LASH_DD_FN uint64_t dd_bits(double x) {
    uint64_t b;
    memcpy(&b, &x, 8);
    return b;
}
LASH_DD_FN double dd_from_bits(uint64_t b) {
    double x;
    memcpy(&x, &b, 8);
    return x;
}

// constants as double-doubles (hi = nearest double, lo = nearest double of the rest)
#define LASH_DD_LN2 dd{0.6931471805599453, 2.3190468138462996e-17}
#define LASH_DD_THIRD dd{0.3333333333333333, 1.850371707708594e-17}
#define LASH_DD_FIFTH dd{0.2, -1.1102230246251566e-17}
#define LASH_DD_SIXTH dd{0.16666666666666666, 9.25185853854297e-18}

// log(x), x positive, finite, normal; absolute error < 2^-72 for |log x| < 750
LASH_DD_FN dd dd_log(double x) {
    const uint64_t b = dd_bits(x);
    int e = (int)((b >> 52) & 0x7ff) - 1023;
    double m = dd_from_bits((b & 0x000fffffffffffffULL) | 0x3ff0000000000000ULL);  // [1, 2)
    if (m > 1.4142135623730951) {
        m *= 0.5;
        e += 1;
    }
    const dd z = dd_div(dd{m - 1.0, 0.0}, two_sum(m, 1.0));   // m - 1 is exact (m in [0.707, 1.415])
    const dd p = dd_mul(z, z);
    const double ph = p.hi;
    const double poly = 1.0 / 7.0 + ph * (1.0 / 9.0 + ph * (1.0 / 11.0 + ph * (1.0 / 13.0 + ph * (1.0 / 15.0 + ph * (1.0 / 17.0 +
                        ph * (1.0 / 19.0 + ph * (1.0 / 21.0 + ph * (1.0 / 23.0 + ph * (1.0 / 25.0 + ph * (1.0 / 27.0 + ph * (1.0 / 29.0)))))))))));
    const double tail = (ph * ph * ph) * poly;
    dd S = dd_add(dd_mul(p, LASH_DD_THIRD), dd_mul(dd_mul(p, p), LASH_DD_FIFTH));
    S = dd_add_d(S, tail);
    dd lm = dd_add(z, dd_mul(z, S));   // atanh(z) = z (1 + S)
    lm.hi *= 2.0;
    lm.lo *= 2.0;
    return dd_add(dd_mul_d(LASH_DD_LN2, (double)e), lm);
}

// exp(t) rounded to nearest, |t| < 700
LASH_DD_FN double dd_exp_round(dd t) {
    const double kd = nearbyint(t.hi * 1.4426950408889634);
    const dd r = dd_add(t, dd_neg(dd_mul_d(LASH_DD_LN2, kd)));
    const dd s = dd{r.hi * 0.0625, r.lo * 0.0625};
    const double sh = s.hi;
    const double q = (sh * sh) * (sh * sh) *
                     (1.0 / 24.0 + sh * (1.0 / 120.0 + sh * (1.0 / 720.0 + sh * (1.0 / 5040.0 + sh * (1.0 / 40320.0 + sh * (1.0 / 362880.0 +
                      sh * (1.0 / 3628800.0 + sh * (1.0 / 39916800.0))))))));
    const dd s2 = dd_mul(s, s);
    dd in = dd_add(s, dd{s2.hi * 0.5, s2.lo * 0.5});
    in = dd_add(in, dd_mul(dd_mul(s2, s), LASH_DD_SIXTH));
    in = dd_add_d(in, q);
    dd E = dd_add(dd{1.0, 0.0}, in);
    E = dd_mul(E, E);
    E = dd_mul(E, E);
    E = dd_mul(E, E);
    E = dd_mul(E, E);
    // E.hi is the double nearest to E.hi + E.lo (the pair is normalised); scaling by 2^k is exact for a normal result
    const int k = (int)kd;
    return E.hi * dd_from_bits((uint64_t)(1023 + k) << 52);
}

// pow(x, y) correctly rounded for positive normal finite x and results well inside the normal range; everything else
// (zero, negative, subnormal, infinite, NaN arguments, overflowing / underflowing results) goes to the library pow
LASH_DD_FN double pow_cr(double x, double y) {
    const uint64_t bx = dd_bits(x);
    const int ex = (int)((bx >> 52) & 0x7ff);
    if ((bx >> 63) || ex == 0 || ex == 0x7ff || !(y == y) || y - y != 0.0) return pow(x, y);
    const dd t = dd_mul_d(dd_log(x), y);
    if (!(fabs(t.hi) < 700.0)) return pow(x, y);
    return dd_exp_round(t);
}

// log(x) correctly rounded for positive normal finite x (the poisson Mash distance -ln(frac)/k); the rest goes to the library
LASH_DD_FN double log_cr(double x) {
    const uint64_t bx = dd_bits(x);
    const int ex = (int)((bx >> 52) & 0x7ff);
    if ((bx >> 63) || ex == 0 || ex == 0x7ff) return log(x);
    return dd_log(x).hi;   // the pair is normalised: hi is the double nearest to hi + lo
}

// log1p(x) correctly rounded for x > -1 (the initial guess of the ML solver): 1 + x = h + l exactly, and
// log(h + l) = log(h) + w - w^2 / 2 with w = l / h (|w| < 2^-52; the next term is below 2^-158)
LASH_DD_FN double log1p_cr(double x) {
    if (!(x > -1.0) || !(x < 1e300)) return log1p(x);
    const dd u = two_sum(1.0, x);
    const uint64_t bu = dd_bits(u.hi);
    if (((bu >> 52) & 0x7ff) == 0) return log1p(x);
    const double q = u.lo / u.hi;
    const dd w = quick_two_sum(q, fma(-q, u.hi, u.lo) / u.hi);   // l / h to double-double: its rounding error matters when log(h) ~ 0
    return dd_add(dd_log(u.hi), dd_add_d(w, -0.5 * q * q)).hi;
}

}  // namespace lash
