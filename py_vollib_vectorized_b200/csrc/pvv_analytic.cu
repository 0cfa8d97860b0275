// pvv_analytic.cu — opt-in closed-form greeks.
//
// The reference's greeks are finite differences (kernel K2, FP64-bound: eight price evaluations per
// contract).  This kernel is the alternative the hot path's roofline statement asks for: the
// closed-form Black-Scholes(-Merton) price and greeks with py_vollib's `greeks.analytical`
// conventions (theta per calendar day, vega and rho per point), one pass, 41-49 bytes in / 48 out per
// contract, arithmetic small enough (~230 DP instructions) to sit near the HBM roofline.  There is
// no numba counterpart to be bit-identical with, so this translation unit is compiled WITH FMA
// contraction and uses its own minimal-operation-count special functions:
//   N(x) via Cody's scaled complementary error function evaluated once per |d| together with the
//   density (N(-|d|) = phi(d) * erfcx(|d|/sqrt2) * sqrt(pi/2)), exp/log from the CUDA math library.
// Parity: against the closed-form float64 checker used by the tests (pvv_oracle_analytic_greeks),
// 1e-12 relative (+1e-300 absolute), and against the reference's regression table columns
// CD/CG/CT/CV/CR/PD/PG/PT/PV/PR (6 decimals) — tests/test_gpu_analytic.py.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>

#include "../../include/pvv.h"

namespace {

struct AStrides {
    int s[7];
};

// erfcx(y) for y >= 0: Cody's rational Chebyshev approximations (Math. Comp. 1969), FMA Horner form.
__device__ __forceinline__ double erfcx_pos(double y) {
    if (y <= 0.46875) {
        const double ysq = y * y;
        double xnum = .185777706184603153 * ysq, xden = ysq;
        xnum = (xnum + 3.1611237438705656) * ysq;
        xden = (xden + 23.6012909523441209) * ysq;
        xnum = (xnum + 113.864154151050156) * ysq;
        xden = (xden + 244.024637934444173) * ysq;
        xnum = (xnum + 377.485237685302021) * ysq;
        xden = (xden + 1282.61652607737228) * ysq;
        const double erf = y * (xnum + 3209.37758913846947) / (xden + 2844.23683343917062);
        return (1.0 - erf) * exp(ysq);
    }
    if (y <= 4.0) {
        double xnum = 2.15311535474403846e-8 * y, xden = y;
        xnum = (xnum + .564188496988670089) * y;
        xden = (xden + 15.7449261107098347) * y;
        xnum = (xnum + 8.88314979438837594) * y;
        xden = (xden + 117.693950891312499) * y;
        xnum = (xnum + 66.1191906371416295) * y;
        xden = (xden + 537.181101862009858) * y;
        xnum = (xnum + 298.635138197400131) * y;
        xden = (xden + 1621.38957456669019) * y;
        xnum = (xnum + 881.95222124176909) * y;
        xden = (xden + 3290.79923573345963) * y;
        xnum = (xnum + 1712.04761263407058) * y;
        xden = (xden + 4362.61909014324716) * y;
        xnum = (xnum + 2051.07837782607147) * y;
        xden = (xden + 3439.36767414372164) * y;
        return (xnum + 1230.33935479799725) / (xden + 1230.33935480374942);
    }
    const double ysq = 1.0 / (y * y);
    double xnum = .0163153871373020978 * ysq, xden = ysq;
    xnum = (xnum + .305326634961232344) * ysq;
    xden = (xden + 2.56852019228982242) * ysq;
    xnum = (xnum + .360344899949804439) * ysq;
    xden = (xden + 1.87295284992346047) * ysq;
    xnum = (xnum + .125781726111229246) * ysq;
    xden = (xden + .527905102951428412) * ysq;
    xnum = (xnum + .0160837851487422766) * ysq;
    xden = (xden + .0605183413124413191) * ysq;
    const double r = ysq * (xnum + 6.58749161529837803e-4) / (xden + .00233520497626869185);
    return (0.56418958354775628695 - r) / y;
}

// phi(d), N(d) and N(-d) with one exp: N(-|d|) = 0.5 * exp(-d^2/2) * erfcx(|d|/sqrt 2).  Both tails
// keep full relative accuracy (a put's N(-d1) must not be formed as 1 - N(d1)).
__device__ __forceinline__ void pdf_cdf(double d, double &pdf, double &cdf, double &ccdf) {
    const double ad = fabs(d);
    const double e = exp(-0.5 * d * d);
    pdf = 0.3989422804014326779399460599343818684758586311649 * e;
    const double tail = 0.5 * e * erfcx_pos(ad * 0.7071067811865475244008443621048490392848359376887);
    cdf = d < 0 ? tail : 1.0 - tail;
    ccdf = d < 0 ? 1.0 - tail : tail;
    if (d != d) cdf = ccdf = d;
}

template <int MODEL>
__global__ void __launch_bounds__(256) analytic_kernel(long long n, const int8_t *__restrict__ flag, const double *__restrict__ S,
                                                       const double *__restrict__ K, const double *__restrict__ t,
                                                       const double *__restrict__ r, const double *__restrict__ q,
                                                       const double *__restrict__ sigma, AStrides st, double *__restrict__ price,
                                                       double *__restrict__ delta, double *__restrict__ gamma, double *__restrict__ theta,
                                                       double *__restrict__ rho, double *__restrict__ vega) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const bool call = flag[i * st.s[0]] > 0;
    const double Si = S[i * st.s[1]], Ki = K[i * st.s[2]], ti = t[i * st.s[3]], ri = r[i * st.s[4]], sg = sigma[i * st.s[6]];
    const double qi = (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? q[i * st.s[5]] : 0.0;
    const double sqt = sqrt(ti);
    const double ssqt = sg * sqt;
    const double inv = 1.0 / ssqt;
    const double d1 = (log(Si / Ki) + (ri - qi + 0.5 * sg * sg) * ti) * inv;
    const double d2 = d1 - ssqt;
    double pdf1, N1, M1, pdf2, N2, M2;  // N = N(d), M = N(-d)
    pdf_cdf(d1, pdf1, N1, M1);
    pdf_cdf(d2, pdf2, N2, M2);
    const double Dr = exp(-ri * ti);
    const double Dq = (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? exp(-qi * ti) : 1.0;
    const double SDq = Si * Dq, KDr = Ki * Dr;
    const double Nd1 = call ? N1 : M1;  // N(d1) for a call, N(-d1) for a put
    const double Nd2 = call ? N2 : M2;
    const double sgn = call ? 1.0 : -1.0;
    if (price) price[i] = sgn * (SDq * Nd1 - KDr * Nd2);
    if (delta) delta[i] = sgn * Dq * Nd1;
    if (gamma) gamma[i] = Dq * pdf1 * inv / Si;
    if (vega) vega[i] = SDq * pdf1 * sqt * 0.01;
    if (theta) theta[i] = (-(SDq * pdf1 * sg) / (2 * sqt) - sgn * ri * KDr * Nd2 + sgn * qi * SDq * Nd1) * (1.0 / 365.0);
    if (rho) rho[i] = sgn * ti * KDr * Nd2 * 0.01;
}

}  // namespace

extern "C" {

int pvv_analytic_greeks_dev(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                            const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta, double *gamma,
                            double *theta, double *rho, double *vega, void *stream) {
    AStrides st;
    if (n < 0 || (model != PVV_MODEL_BLACK_SCHOLES && model != PVV_MODEL_BLACK_SCHOLES_MERTON)) return PVV_E_INVALID_ARGUMENT;
    if (model == PVV_MODEL_BLACK_SCHOLES_MERTON && !q) return PVV_E_INVALID_ARGUMENT;
    for (int i = 0; i < 7; ++i) {
        if (stride[i] != 0 && stride[i] != 1) return PVV_E_INVALID_ARGUMENT;
        st.s[i] = (int)stride[i];
    }
    if (n == 0) return 0;
    const int grid = (int)((n + 255) / 256);
    cudaStream_t s = (cudaStream_t)stream;
    if (model == PVV_MODEL_BLACK_SCHOLES)
        analytic_kernel<PVV_MODEL_BLACK_SCHOLES><<<grid, 256, 0, s>>>(n, flag, S, K, t, r, q, sigma, st, price, delta, gamma, theta, rho, vega);
    else
        analytic_kernel<PVV_MODEL_BLACK_SCHOLES_MERTON><<<grid, 256, 0, s>>>(n, flag, S, K, t, r, q, sigma, st, price, delta, gamma, theta, rho, vega);
    pvv_note_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
