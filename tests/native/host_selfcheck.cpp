// Host build of the DEVICE math header (py_vollib_vectorized_b200/csrc/pvv_math.cuh) so that the
// CPU test tier can diff the product's scalar code against the oracle and the host libm without a
// GPU.  This is a test harness: it is never loaded by the product package.
//   g++ -O2 -std=c++17 -ffp-contract=off -mfma -shared -fPIC -o libpvv_hostcheck.so host_selfcheck.cpp
#include "../../py_vollib_vectorized_b200/csrc/pvv_math.cuh"

#define X extern "C" __attribute__((visibility("default")))

X double pvvh_exp(double x) { return pvv::exp_g(x); }
X double pvvh_log(double x) { return pvv::log_g(x); }
X double pvvh_erfcx(double x) { return pvv::erfcx_cody(x); }
X double pvvh_erfc(double x) { return pvv::erfc_cody(x); }
X double pvvh_norm_cdf(double x) { return pvv::norm_cdf(x); }
X double pvvh_inverse_norm_cdf(double x) { return pvv::inverse_norm_cdf(x); }
X double pvvh_pow(double x, double y) { return pvv::pow_g(x, y); }
X void pvvh_pow_array(long n, const double *x, double y, double *out) { for (long i = 0; i < n; ++i) out[i] = pvv::pow_g(x[i], y); }
X void pvvh_exp_array(long n, const double *x, double *y) { for (long i = 0; i < n; ++i) y[i] = pvv::exp_g(x[i]); }
X void pvvh_log_array(long n, const double *x, double *y) { for (long i = 0; i < n; ++i) y[i] = pvv::log_g(x[i]); }
// x / c through the zero-safe constant division of pvv_math.cuh (c = 6: Householder factor; c = Cody's B[3])
X void pvvh_div_const_array(long n, const double *x, double c, double *y) {
    const double rc = 1.0 / c;
    for (long i = 0; i < n; ++i) y[i] = pvv::div_by_const_zn(x[i], c, rc);
}
// region of the sort key against the reference-order classification (the key's region must be exact)
X long pvvh_sort_key_mismatches(long n, const double *x, const double *s) {
    long bad = 0;
    for (long i = 0; i < n; ++i) bad += pvv::key_region(pvv::nbc_sort_key(x[i], s[i])) != pvv::nbc_classify(x[i], s[i]);
    return bad;
}
X double pvvh_const(int which) {
    switch (which) {
        case 0: return PVV_SQRT_EPS;
        case 1: return PVV_FOURTH_ROOT_EPS;
        case 2: return PVV_SIXTEENTH_ROOT_EPS;
        case 3: return PVV_SMALL_T_THRESHOLD;
        case 4: return PVV_SQRT_DBL_MIN;
        case 5: return PVV_SQRT_DBL_MAX;
        case 6: return PVV_RC_MIN;
        case 7: return PVV_RC_MAX;
        case 8: return PVV_NORM_CDF_SECOND_THRESHOLD;
        case 9: return 0x1.a36e2eb1c432dp-14;
    }
    return 0;
}

template <int MODEL>
static void price_loop(long n, const signed char *flag, const double *S, const double *K, const double *t, const double *r, const double *q,
                       const double *sigma, double *out, unsigned char *status) {
    for (long i = 0; i < n; ++i) {
        unsigned st = 0;
        out[i] = pvv::price_model<MODEL>((double)flag[i], S[i], K[i], t[i], r[i], sigma[i], q[i], st);
        status[i] = (unsigned char)st;
    }
}
X void pvvh_price(int model, long n, const signed char *flag, const double *S, const double *K, const double *t, const double *r,
                  const double *q, const double *sigma, double *out, unsigned char *status) {
    if (model == 0) price_loop<0>(n, flag, S, K, t, r, q, sigma, out, status);
    else if (model == 1) price_loop<1>(n, flag, S, K, t, r, q, sigma, out, status);
    else price_loop<2>(n, flag, S, K, t, r, q, sigma, out, status);
}
X void pvvh_greeks(int model, long n, const signed char *flag, const double *S, const double *K, const double *t, const double *r,
                   const double *q, const double *sigma, double *out6, unsigned char *status) {
    for (long i = 0; i < n; ++i) {
        unsigned st = 0;
        pvv::Greeks g = (model == 2) ? pvv::fd_greeks<2>((double)flag[i], S[i], K[i], t[i], r[i], sigma[i], q[i], st)
                                     : pvv::fd_greeks<1>((double)flag[i], S[i], K[i], t[i], r[i], sigma[i], 0.0, st);
        out6[6 * i + 0] = g.price; out6[6 * i + 1] = g.delta; out6[6 * i + 2] = g.gamma;
        out6[6 * i + 3] = g.theta; out6[6 * i + 4] = g.rho; out6[6 * i + 5] = g.vega;
        status[i] = (unsigned char)st;
    }
}
X void pvvh_iv(int model, long n, const signed char *flag, const double *price, const double *S, const double *K, const double *t,
               const double *r, const double *q, int mask, double *iv, unsigned char *status, signed char *branch) {
    for (long i = 0; i < n; ++i) {
        unsigned st = 0;
        int br = -1;
        double f = (double)flag[i];
        iv[i] = model == 0 ? pvv::implied_vol<0>(f, price[i], S[i], K[i], t[i], r[i], 0.0, mask, st, &br)
              : model == 1 ? pvv::implied_vol<1>(f, price[i], S[i], K[i], t[i], r[i], 0.0, mask, st, &br)
                           : pvv::implied_vol<2>(f, price[i], S[i], K[i], t[i], r[i], q[i], mask, st, &br);
        status[i] = (unsigned char)st;
        if (branch) branch[i] = (signed char)br;
    }
}
