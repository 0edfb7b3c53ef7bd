// Test infrastructure: lash_b200/csrc/ddmath.cuh (the correctly rounded pow of the FGRA epilogue) compiled with g++ so that
// it can be checked against mpmath (correct rounding) and glibc on a machine without a GPU.  Nothing here is the product.
#include "../../lash_b200/csrc/ddmath.cuh"

extern "C" {
double dm_pow_cr(double x, double y) { return lash::pow_cr(x, y); }
void dm_log_dd(double x, double* hi, double* lo) {
    const lash::dd r = lash::dd_log(x);
    *hi = r.hi;
    *lo = r.lo;
}
double dm_log_cr(double x) { return lash::log_cr(x); }
double dm_log1p_cr(double x) { return lash::log1p_cr(x); }
void dm_log_cr_many(const double* x, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = lash::log_cr(x[i]);
}
void dm_log_libm_many(const double* x, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = log(x[i]);
}
void dm_pow_cr_many(const double* x, double y, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = lash::pow_cr(x[i], y);
}
void dm_pow_libm_many(const double* x, double y, double* out, long n) {
    for (long i = 0; i < n; ++i) out[i] = pow(x[i], y);
}
}
