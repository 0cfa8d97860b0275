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
#include <initializer_list>

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
// scalar constants with a non-zero low word (two UMOV each as literals): 1/sqrt(2 pi), 1/sqrt 2, 1/sqrt(pi), 1/365, 0.01
__constant__ double CK[5] = {0.3989422804014326779399460599343818684758586311649, 0.7071067811865475244008443621048490392848359376887,
                             0.56418958354775628695, 1.0 / 365.0, 0.01};

// 1 / b to within an ulp: MUFU.RCP64H seed, one cubic and one quadratic Newton step — the fast path of
// nvcc's own divider without its operand guards and slow-path call (7 of the 16 instructions of every
// IEEE division).  This translation unit has no bit-exactness contract (tolerance 1e-12 relative), and a
// quotient a / b becomes a * rcp(b); zero, infinite and NaN divisors return the seed (inf, 0, NaN).
__device__ __forceinline__ double rcp_nr(double b) {
    double r0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r0) : "d"(b));
    double e = __fma_rn(-b, r0, 1.0);
    const bool special = e != e;
    e = __fma_rn(e, e, e);
    double r = __fma_rn(r0, e, r0);
    e = __fma_rn(-b, r, 1.0);
    r = __fma_rn(r, e, r);
    return special ? r0 : r;
}

// sqrt(t) to within an ulp from the MUFU.RSQ64H seed: two Newton steps on y ~ 1/sqrt(t), then one correction of
// s = t y.  Zero, negative, infinite and NaN arguments take the library sqrt.
__device__ __forceinline__ double sqrt_nr(double t) {
    if (!(t >= 0x1p-1000 && t <= 0x1p1000)) return sqrt(t);
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(t));
    const double ht = 0.5 * t;
    y = __fma_rn(y, __fma_rn(-ht * y, y, 0.5), y);
    y = __fma_rn(y, __fma_rn(-ht * y, y, 0.5), y);
    const double s = t * y;
    return __fma_rn(__fma_rn(-s, s, t), 0.5 * y, s);
}

// 0.5 * erfc(y) for y >= 0 as A + B * (N / D): Cody's rational Chebyshev approximations (Math. Comp. 1969),
// FMA Horner form.  The division is left to the caller, which shares ONE reciprocal between the two calls
// of a contract.  `e` = exp(-y^2) is supplied by the caller (it has exp(-d^2/2) at hand).
struct ErfcParts {
    double A, B, N, D;
};
__device__ __forceinline__ ErfcParts half_erfc_parts(double y, double e) {
    ErfcParts p;
    if (y <= 0.46875) {
        const double ysq = y * y;
        double xnum = CA[4] * ysq, xden = ysq;
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xnum = (xnum + CA[i]) * ysq;
            xden = (xden + CB[i]) * ysq;
        }
        p.A = 0.5;
        p.B = -0.5 * y;
        p.N = xnum + CA[3];
        p.D = xden + CB[3];
    } else if (y <= 4.0) {
        double xnum = CC[8] * y, xden = y;
#pragma unroll
        for (int i = 0; i < 7; ++i) {
            xnum = (xnum + CC[i]) * y;
            xden = (xden + CD[i]) * y;
        }
        p.A = 0.0;
        p.B = 0.5 * e;
        p.N = xnum + CC[7];
        p.D = xden + CD[7];
    } else {
        const double ry = rcp_nr(y);
        const double ysq = ry * ry;
        double xnum = CP[5] * ysq, xden = ysq;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            xnum = (xnum + CP[i]) * ysq;
            xden = (xden + CQ[i]) * ysq;
        }
        const double her = 0.5 * e * ry;
        p.A = her * CK[2];
        p.B = -her * ysq;
        p.N = xnum + CP[4];
        p.D = xden + CQ[4];
    }
    return p;
}

// N(d) and N(-d) from the tail 0.5 erfc(|d|/sqrt 2).  Both tails keep full relative accuracy (a put's
// N(-d1) must not be formed as 1 - N(d1)).
__device__ __forceinline__ void cdf_pair(double d, double tail, double &cdf, double &ccdf) {
    cdf = d < 0 ? tail : 1.0 - tail;
    ccdf = d < 0 ? 1.0 - tail : tail;
    if (d != d) cdf = ccdf = d;
}

struct AnalyticOut {
    double price, delta, gamma, theta, rho, vega;
};

// One contract.  py_vollib's greeks.analytical conventions: theta per calendar day, vega and rho per point.
template <int MODEL>
__device__ __forceinline__ AnalyticOut analytic_one(bool call, double Si, double Ki, double ti, double ri, double qi, double sg) {
    const double sqt = sqrt(ti);
    const double ssqt = sg * sqt;
    const double rSs = rcp_nr(ssqt * Si);  // 1 / (S sigma sqrt t): gamma's denominator, and 1 / (sigma sqrt t) with it
    const double inv = rSs * Si;
    const double rK = rcp_nr(Ki);
    double ratio = Si * rK;
    ratio = __fma_rn(__fma_rn(-Ki, ratio, Si), rK, ratio);  // S / K correctly rounded: its log feeds d1 / (sigma sqrt t)
    const double d1 = (log(ratio) + (ri - qi + 0.5 * sg * sg) * ti) * inv;
    const double d2 = d1 - ssqt;
    const double Dr = exp(-ri * ti);
    const double Dq = (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? exp(-qi * ti) : 1.0;
    const double SDq = Si * Dq, KDr = Ki * Dr;
    // exp(-d2^2/2) = exp(-d1^2/2) * S Dq / (K Dr)  (the identity S Dq phi(d1) = K Dr phi(d2)): one exp, not two
    const double e1 = exp(-0.5 * d1 * d1);
    const double e2 = e1 * (SDq * rcp_nr(KDr));
    const double pdf1 = CK[0] * e1;
    const ErfcParts p1 = half_erfc_parts(fabs(d1) * CK[1], e1), p2 = half_erfc_parts(fabs(d2) * CK[1], e2);
    const double rD = rcp_nr(p1.D * p2.D);  // both denominators are positive polynomials between 2e-3 and 1e6
    const double tail1 = __fma_rn(p1.B, p1.N * p2.D * rD, p1.A), tail2 = __fma_rn(p2.B, p2.N * p1.D * rD, p2.A);
    double N1, M1, N2, M2;  // N = N(d), M = N(-d)
    cdf_pair(d1, tail1, N1, M1);
    cdf_pair(d2, tail2, N2, M2);
    const double Nd1 = call ? N1 : M1;  // N(d1) for a call, N(-d1) for a put
    const double Nd2 = call ? N2 : M2;
    const double sgn = call ? 1.0 : -1.0;
    AnalyticOut o;
    o.price = sgn * (SDq * Nd1 - KDr * Nd2);
    o.delta = sgn * Dq * Nd1;
    o.gamma = Dq * pdf1 * rSs;
    o.vega = SDq * pdf1 * sqt * CK[4];
    o.theta = (-(SDq * pdf1 * sg) * (0.5 * sg * inv) - sgn * ri * KDr * Nd2 + sgn * qi * SDq * Nd1) * CK[3];
    o.rho = sgn * ti * KDr * Nd2 * CK[4];
    return o;
}

// General form: one contract per thread, dense (stride 1) or broadcast (stride 0) columns, any alignment.
template <int MODEL>
__global__ void __launch_bounds__(256, 5) analytic_kernel(long long n, const int8_t *__restrict__ flag, const double *__restrict__ S,
                                                          const double *__restrict__ K, const double *__restrict__ t,
                                                          const double *__restrict__ r, const double *__restrict__ q,
                                                          const double *__restrict__ sigma, AStrides st, double *__restrict__ price,
                                                          double *__restrict__ delta, double *__restrict__ gamma, double *__restrict__ theta,
                                                          double *__restrict__ rho, double *__restrict__ vega) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n) return;
    const double qi = (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? q[i * st.s[5]] : 0.0;
    const AnalyticOut o = analytic_one<MODEL>(flag[i * st.s[0]] > 0, S[i * st.s[1]], K[i * st.s[2]], t[i * st.s[3]], r[i * st.s[4]], qi,
                                              sigma[i * st.s[6]]);
    if (price) price[i] = o.price;
    if (delta) delta[i] = o.delta;
    if (gamma) gamma[i] = o.gamma;
    if (vega) vega[i] = o.vega;
    if (theta) theta[i] = o.theta;
    if (rho) rho[i] = o.rho;
}

// Streaming form: all columns dense and 16-byte aligned, all six outputs wanted.  Two consecutive contracts per
// thread: every column moves as one 128-bit access per thread (half the memory instructions and index
// arithmetic, two independent dependency chains per thread).  n2 = number of contract PAIRS.
template <int MODEL>
__global__ void __launch_bounds__(256, 4) analytic_kernel_x2(long long n2, const char2 *__restrict__ flag, const double2 *__restrict__ S,
                                                             const double2 *__restrict__ K, const double2 *__restrict__ t,
                                                             const double2 *__restrict__ r, const double2 *__restrict__ q,
                                                             const double2 *__restrict__ sigma, double2 *__restrict__ price,
                                                             double2 *__restrict__ delta, double2 *__restrict__ gamma,
                                                             double2 *__restrict__ theta, double2 *__restrict__ rho,
                                                             double2 *__restrict__ vega) {
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= n2) return;
    const char2 f = flag[i];
    const double2 Si = S[i], Ki = K[i], ti = t[i], ri = r[i], sg = sigma[i];
    double2 qi = make_double2(0.0, 0.0);
    if (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) qi = q[i];
    const AnalyticOut a = analytic_one<MODEL>(f.x > 0, Si.x, Ki.x, ti.x, ri.x, qi.x, sg.x);
    const AnalyticOut b = analytic_one<MODEL>(f.y > 0, Si.y, Ki.y, ti.y, ri.y, qi.y, sg.y);
    price[i] = make_double2(a.price, b.price);
    delta[i] = make_double2(a.delta, b.delta);
    gamma[i] = make_double2(a.gamma, b.gamma);
    theta[i] = make_double2(a.theta, b.theta);
    rho[i] = make_double2(a.rho, b.rho);
    vega[i] = make_double2(a.vega, b.vega);
}

inline bool aligned16(const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

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
    cudaStream_t s = (cudaStream_t)stream;
    const bool bsm = model == PVV_MODEL_BLACK_SCHOLES_MERTON;
    // streaming form for the bulk of a dense, aligned, all-outputs call; the general form for everything else
    bool fast = price && delta && gamma && theta && rho && vega && n >= 2 && (reinterpret_cast<uintptr_t>(flag) & 1u) == 0;
    for (int i = 0; i < 7 && fast; ++i) fast = st.s[i] == 1 || (i == 5 && !bsm);
    for (const void *p : {(const void *)S, (const void *)K, (const void *)t, (const void *)r, (const void *)sigma, (const void *)price,
                          (const void *)delta, (const void *)gamma, (const void *)theta, (const void *)rho, (const void *)vega})
        fast = fast && aligned16(p);
    if (bsm) fast = fast && aligned16(q);
    long long done = 0;
    if (fast) {
        const long long n2 = n / 2;
        const int grid2 = (int)((n2 + 255) / 256);
#define PVV_X2(M)                                                                                                                      \
    analytic_kernel_x2<M><<<grid2, 256, 0, s>>>(n2, (const char2 *)flag, (const double2 *)S, (const double2 *)K, (const double2 *)t,   \
                                                (const double2 *)r, (const double2 *)q, (const double2 *)sigma, (double2 *)price,      \
                                                (double2 *)delta, (double2 *)gamma, (double2 *)theta, (double2 *)rho, (double2 *)vega)
        if (bsm) PVV_X2(PVV_MODEL_BLACK_SCHOLES_MERTON);
        else PVV_X2(PVV_MODEL_BLACK_SCHOLES);
#undef PVV_X2
        done = 2 * n2;
    }
    if (done < n) {  // the odd last contract of the streaming form, or the whole call
        const long long m = n - done;
        const int grid = (int)((m + 255) / 256);
        const double *qq = q ? q + done * st.s[5] : q;
#define PVV_X1(M)                                                                                                                          \
    analytic_kernel<M><<<grid, 256, 0, s>>>(m, flag + done * st.s[0], S + done * st.s[1], K + done * st.s[2], t + done * st.s[3],           \
                                            r + done * st.s[4], qq, sigma + done * st.s[6], st, price ? price + done : price,               \
                                            delta ? delta + done : delta, gamma ? gamma + done : gamma, theta ? theta + done : theta,       \
                                            rho ? rho + done : rho, vega ? vega + done : vega)
        if (bsm) PVV_X1(PVV_MODEL_BLACK_SCHOLES_MERTON);
        else PVV_X1(PVV_MODEL_BLACK_SCHOLES);
#undef PVV_X1
    }
    pvv_note_launch();
    return (int)cudaGetLastError();
}

}  // extern "C"
