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

// Cody's coefficients in __constant__ memory (a literal fp64 constant costs two UMOV per use on sm_100a:
// 29 % of this kernel's instructions in the first profile).
__constant__ double CA[5] = {3.1611237438705656, 113.864154151050156, 377.485237685302021, 3209.37758913846947, .185777706184603153};
__constant__ double CB[4] = {23.6012909523441209, 244.024637934444173, 1282.61652607737228, 2844.23683343917062};
__constant__ double CC[9] = {.564188496988670089, 8.88314979438837594, 66.1191906371416295, 298.635138197400131, 881.95222124176909,
                             1712.04761263407058, 2051.07837782607147, 1230.33935479799725, 2.15311535474403846e-8};
__constant__ double CD[8] = {15.7449261107098347, 117.693950891312499, 537.181101862009858, 1621.38957456669019, 3290.79923573345963,
                             4362.61909014324716, 3439.36767414372164, 1230.33935480374942};
__constant__ double CP[6] = {.305326634961232344, .360344899949804439, .125781726111229246, .0160837851487422766, 6.58749161529837803e-4,
                             .0163153871373020978};
__constant__ double CQ[5] = {2.56852019228982242, 1.87295284992346047, .527905102951428412, .0605183413124413191, .00233520497626869185};

// erfcx(y) for y >= 0: Cody's rational Chebyshev approximations (Math. Comp. 1969), FMA Horner form.
// `e` = exp(-y^2) is supplied by the caller (it has exp(-d^2/2) at hand); only the small-|y| interval needs it.
__device__ __forceinline__ double half_erfc_pos(double y, double e) {  // 0.5 * erfc(y) for y >= 0
    if (y <= 0.46875) {
        const double ysq = y * y;
        double xnum = CA[4] * ysq, xden = ysq;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xnum = (xnum + CA[i]) * ysq;
            xden = (xden + CB[i]) * ysq;
        }
        return 0.5 - 0.5 * (y * (xnum + CA[3]) / (xden + CB[3]));
    }
    if (y <= 4.0) {
        double xnum = CC[8] * y, xden = y;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            xnum = (xnum + CC[i]) * y;
            xden = (xden + CD[i]) * y;
        }
        return 0.5 * e * ((xnum + CC[7]) / (xden + CD[7]));
    }
    const double ysq = 1.0 / (y * y);
    double xnum = CP[5] * ysq, xden = ysq;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        xnum = (xnum + CP[i]) * ysq;
        xden = (xden + CQ[i]) * ysq;
    }
    const double r = ysq * (xnum + CP[4]) / (xden + CQ[4]);
    return 0.5 * e * ((0.56418958354775628695 - r) / y);
}

// N(d) and N(-d) from e = exp(-d^2/2): N(-|d|) = 0.5 erfc(|d|/sqrt 2).  Both tails keep full relative
// accuracy (a put's N(-d1) must not be formed as 1 - N(d1)).
__device__ __forceinline__ void cdf_pair(double d, double e, double &cdf, double &ccdf) {
    const double tail = half_erfc_pos(fabs(d) * 0.7071067811865475244008443621048490392848359376887, e);
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
    const double Dr = exp(-ri * ti);
    const double Dq = (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? exp(-qi * ti) : 1.0;
    const double SDq = Si * Dq, KDr = Ki * Dr;
    // exp(-d2^2/2) = exp(-d1^2/2) * S Dq / (K Dr)  (the identity S Dq phi(d1) = K Dr phi(d2)): one exp, not two
    const double e1 = exp(-0.5 * d1 * d1);
    const double e2 = e1 * (SDq / KDr);
    const double pdf1 = 0.3989422804014326779399460599343818684758586311649 * e1;
    double N1, M1, N2, M2;  // N = N(d), M = N(-d)
    cdf_pair(d1, e1, N1, M1);
    cdf_pair(d2, e2, N2, M2);
    const double Nd1 = call ? N1 : M1;  // N(d1) for a call, N(-d1) for a put
    const double Nd2 = call ? N2 : M2;
    const double sgn = call ? 1.0 : -1.0;
    if (price) price[i] = sgn * (SDq * Nd1 - KDr * Nd2);
    if (delta) delta[i] = sgn * Dq * Nd1;
    if (gamma) gamma[i] = Dq * pdf1 * inv / Si;
    if (vega) vega[i] = SDq * pdf1 * sqt * 0.01;
    if (theta) theta[i] = (-(SDq * pdf1 * sg) * (0.5 * sg * inv) - sgn * ri * KDr * Nd2 + sgn * qi * SDq * Nd1) * (1.0 / 365.0);
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
