// pvv_analytic.cu — opt-in closed-form greeks.
//
// The reference's greeks are finite differences (kernel K2, FP64-bound: eight price evaluations per
// contract).  This kernel is the alternative the hot path's roofline statement asks for: the
// closed-form Black-Scholes(-Merton) price and greeks with py_vollib's `greeks.analytical`
// conventions (theta per calendar day, vega and rho per point), one pass, 41-49 bytes in / 48 out per
// contract, arithmetic small enough (~230 DP instructions) to sit near the HBM roofline.  There is
// no numba counterpart to be bit-identical with, so this translation unit is compiled WITH FMA
// contraction and uses its own minimal-operation-count special functions:
//   N(-|d|) = exp(-d^2/2) erfcx(|d|/sqrt 2) / 2 with erfcx from a 100-interval table of degree-6 polynomials in
//   t = 400 / (4 + y) (csrc/erfcx_table.cuh, generated with mpmath by tools/gen_erfcx_table.py: 1.6e-16 relative) held
//   in shared memory — seven FMAs and no interval branch, where Cody's rational functions cost 32 dependent operations on
//   one of three divergent paths (round 1: 223 issued vs 174 useful DP instructions per contract); exp/log from the CUDA
//   math library.
// Parity: against the closed-form float64 checker used by the tests (pvv_oracle_analytic_greeks),
// 1e-12 relative (+1e-300 absolute), and against the reference's regression table columns
// CD/CG/CT/CV/CR/PD/PG/PT/PV/PR (6 decimals) — tests/test_gpu_analytic.py.
#include <cuda_runtime.h>

#include <cfloat>
#include <cstdint>
#include <initializer_list>

#include "../../include/pvv.h"
#include "erfcx_table.cuh"

#ifndef PVV_ANALYTIC_STAGES
#define PVV_ANALYTIC_STAGES 2
#endif
#ifndef PVV_ANALYTIC_RESIDENT
#define PVV_ANALYTIC_RESIDENT 2    /* resident CTAs per SM: 128 registers per thread, four contracts side by side, no spills */
#endif
#ifndef PVV_ANALYTIC_CTAS_PER_SM
#define PVV_ANALYTIC_CTAS_PER_SM (2 * PVV_ANALYTIC_RESIDENT)  /* grid: two waves of the resident CTAs */
#endif

namespace {

struct AStrides {
    int s[7];
};

// erfcx table, coefficient-major in shared memory (tab[k * kTabStride + i]: the 32 lanes of a warp read the k-th coefficient of
// 32 different intervals — consecutive intervals fall into consecutive bank pairs)
constexpr int kTabRows = 100, kTabStride = 104;
__device__ const double ERFCX_TAB[kTabRows * 8] = {PVV_ERFCX_TAB_VALUES};
__device__ __forceinline__ void load_erfcx_table(double *tab) {
    for (int j = threadIdx.x; j < kTabRows * 7; j += blockDim.x) {
        const int i = j / 7, k = j - 7 * i;
        tab[k * kTabStride + i] = ERFCX_TAB[i * 8 + k];
    }
    __syncthreads();
}

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

// erfcx(y) for y >= 0 given rinv = 1 / (4 + y): t = 400 rinv in (0, 100], interval floor(t), u in [-1, 1], degree-6 Horner.
// y beyond 396 (interval 0) or NaN evaluates a finite polynomial that the caller multiplies by exp(-y^2) = 0 (or NaN).
__device__ __forceinline__ double erfcx_tab(double rinv, const double *__restrict__ tab) {
    const double t = 400.0 * rinv;
    int i = (int)t;
    i = i < 0 ? 0 : (i > kTabRows - 1 ? kTabRows - 1 : i);  // t == 100 (y == 0) belongs to the last interval, u = 1
    const double u = __fma_rn(2.0, t, -(double)(2 * i + 1));
    const double *c = tab + i;
    double p = c[6 * kTabStride];
#pragma unroll
    for (int k = 5; k >= 0; --k) p = __fma_rn(p, u, c[k * kTabStride]);
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
__device__ __forceinline__ AnalyticOut analytic_one(bool call, double Si, double Ki, double ti, double ri, double qi, double sg,
                                                    const double *__restrict__ tab) {
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
    // both tails from one reciprocal: 1 / (4 + y1) = (4 + y2) / ((4 + y1)(4 + y2))
    const double w1 = __fma_rn(fabs(d1), CK[1], 4.0), w2 = __fma_rn(fabs(d2), CK[1], 4.0);
    const double rw = rcp_nr(w1 * w2);
    const double tail1 = 0.5 * e1 * erfcx_tab(rw * w2, tab), tail2 = 0.5 * e2 * erfcx_tab(rw * w1, tab);
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
// A grid-stride loop over a few CTAs per SM: the erfcx table is copied to shared memory once per CTA.
template <int MODEL>
__global__ void __launch_bounds__(256, PVV_ANALYTIC_RESIDENT) analytic_kernel(long long n, const int8_t *__restrict__ flag, const double *__restrict__ S,
                                                          const double *__restrict__ K, const double *__restrict__ t,
                                                          const double *__restrict__ r, const double *__restrict__ q,
                                                          const double *__restrict__ sigma, AStrides st, double *__restrict__ price,
                                                          double *__restrict__ delta, double *__restrict__ gamma, double *__restrict__ theta,
                                                          double *__restrict__ rho, double *__restrict__ vega) {
    __shared__ double tab[7 * kTabStride];
    load_erfcx_table(tab);
    for (long long i = (long long)blockIdx.x * 256 + threadIdx.x; i < n; i += (long long)gridDim.x * 256) {
        const double qi = (MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? q[i * st.s[5]] : 0.0;
        const AnalyticOut o = analytic_one<MODEL>(flag[i * st.s[0]] > 0, S[i * st.s[1]], K[i * st.s[2]], t[i * st.s[3]], r[i * st.s[4]], qi,
                                                  sigma[i * st.s[6]], tab);
        if (price) price[i] = o.price;
        if (delta) delta[i] = o.delta;
        if (gamma) gamma[i] = o.gamma;
        if (vega) vega[i] = o.vega;
        if (theta) theta[i] = o.theta;
        if (rho) rho[i] = o.rho;
    }
}

// ---- streaming form with bulk-asynchronous loads (TMA, cp.async.bulk) ----------------------------------------------
// A persistent CTA of 256 threads walks chunks of 1024 contracts (four per thread: sixteen resident warps with four independent
// chains each issue more than twenty-four warps with two — measured 61.7 vs 59.7 G/s).  One elected thread issues the chunk's columns as 1-D bulk
// copies global -> shared (one per column: 8 KB, the flag column 1 KB) that complete on an mbarrier; PVV_ANALYTIC_STAGES
// stages, so the copies of the next chunk(s) are in flight while chunk k is computed — the loads cost no registers and no issue slots of the
// 255 other threads, and their latency is off the critical path (the register-loading form above stalls 3 warps per
// issue on its long scoreboard).  Outputs go straight from registers to global memory as 128-bit stores.
#ifndef PVV_ANALYTIC_PAIRS
#define PVV_ANALYTIC_PAIRS 2
#endif
constexpr int kPairs = PVV_ANALYTIC_PAIRS;   // double2 (two contracts) per thread and chunk
constexpr int kChunk = 512 * kPairs;  // contracts per chunk
template <int MODEL>
struct alignas(128) AnalyticStage {
    double2 col[(MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON) ? 6 : 5][kChunk / 2];  // S, K, t, r, sigma[, q]
    char2 flag[kChunk / 2];
};
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

template <int MODEL>
__global__ void __launch_bounds__(256, PVV_ANALYTIC_RESIDENT) analytic_kernel_tma(long long nchunks, const int8_t *__restrict__ flag,
                                                                                  const double *__restrict__ S, const double *__restrict__ K,
                                                                                  const double *__restrict__ t, const double *__restrict__ r,
                                                                                  const double *__restrict__ q, const double *__restrict__ sigma,
                                                                                  double2 *__restrict__ price, double2 *__restrict__ delta,
                                                                                  double2 *__restrict__ gamma, double2 *__restrict__ theta,
                                                                                  double2 *__restrict__ rho, double2 *__restrict__ vega) {
    constexpr bool BSM = MODEL == PVV_MODEL_BLACK_SCHOLES_MERTON;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    AnalyticStage<MODEL> *stage = reinterpret_cast<AnalyticStage<MODEL> *>(smem_raw);
    constexpr int NS = PVV_ANALYTIC_STAGES;
    double *tab = reinterpret_cast<double *>(smem_raw + NS * sizeof(AnalyticStage<MODEL>));
    uint64_t *full = reinterpret_cast<uint64_t *>(tab + 7 * kTabStride);
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int k = 0; k < NS; ++k) mbar_init(&full[k], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    load_erfcx_table(tab);  // ends with a barrier: the mbarriers are initialised for everyone
    const auto issue = [&](int s, long long c) {  // thread 0: the columns of chunk c into stage s
        constexpr uint32_t kBytes = (BSM ? 6u : 5u) * kChunk * 8u + kChunk;
        mbar_expect_tx(&full[s], kBytes);
        const long long o = c * kChunk;
        bulk_g2s(stage[s].col[0], S + o, kChunk * 8, &full[s]);
        bulk_g2s(stage[s].col[1], K + o, kChunk * 8, &full[s]);
        bulk_g2s(stage[s].col[2], t + o, kChunk * 8, &full[s]);
        bulk_g2s(stage[s].col[3], r + o, kChunk * 8, &full[s]);
        bulk_g2s(stage[s].col[4], sigma + o, kChunk * 8, &full[s]);
        if (BSM) bulk_g2s(stage[s].col[BSM ? 5 : 4], q + o, kChunk * 8, &full[s]);
        bulk_g2s(stage[s].flag, flag + o, kChunk, &full[s]);
    };
    const long long step = gridDim.x;
    long long c = blockIdx.x;
    if (tid == 0) {
        for (int k = 0; k < NS; ++k)
            if (c + k * step < nchunks) issue(k, c + k * step);
    }
    int s = 0;
    uint32_t parity = 0;
    for (; c < nchunks; c += step) {
        mbar_wait(&full[s], parity);
        double2 Si[kPairs], Ki[kPairs], ti[kPairs], ri[kPairs], sg[kPairs], qi[kPairs];
        char2 f[kPairs];
#pragma unroll
        for (int h = 0; h < kPairs; ++h) {
            const int j = h * 256 + tid;
            Si[h] = stage[s].col[0][j];
            Ki[h] = stage[s].col[1][j];
            ti[h] = stage[s].col[2][j];
            ri[h] = stage[s].col[3][j];
            sg[h] = stage[s].col[4][j];
            qi[h] = BSM ? stage[s].col[BSM ? 5 : 4][j] : make_double2(0.0, 0.0);
            f[h] = stage[s].flag[j];
        }
        __syncthreads();  // every thread has read the stage: it can be refilled
        if (tid == 0 && c + NS * step < nchunks) issue(s, c + NS * step);
#pragma unroll
        for (int h = 0; h < kPairs; ++h) {
            const AnalyticOut a = analytic_one<MODEL>(f[h].x > 0, Si[h].x, Ki[h].x, ti[h].x, ri[h].x, qi[h].x, sg[h].x, tab);
            const AnalyticOut b = analytic_one<MODEL>(f[h].y > 0, Si[h].y, Ki[h].y, ti[h].y, ri[h].y, qi[h].y, sg[h].y, tab);
            const long long i = c * (kChunk / 2) + h * 256 + tid;
            price[i] = make_double2(a.price, b.price);
            delta[i] = make_double2(a.delta, b.delta);
            gamma[i] = make_double2(a.gamma, b.gamma);
            theta[i] = make_double2(a.theta, b.theta);
            rho[i] = make_double2(a.rho, b.rho);
            vega[i] = make_double2(a.vega, b.vega);
        }
        if (++s == NS) {
            s = 0;
            parity ^= 1u;
        }
    }
}
template <int MODEL>
constexpr size_t analytic_tma_smem() {
    return PVV_ANALYTIC_STAGES * sizeof(AnalyticStage<MODEL>) + 7 * kTabStride * sizeof(double) + PVV_ANALYTIC_STAGES * sizeof(uint64_t);
}

// grid of a grid-stride launch: every SM gets `per_sm` CTAs (or fewer CTAs than that when the column is short)
inline int stride_grid(long long work_items, int per_sm) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    const long long need = (work_items + 255) / 256, cap = (long long)sms * per_sm;
    return (int)(need < cap ? need : cap);
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
    bool fast = price && delta && gamma && theta && rho && vega;
    for (int i = 0; i < 7 && fast; ++i) fast = st.s[i] == 1 || (i == 5 && !bsm);
    for (const void *p : {(const void *)S, (const void *)K, (const void *)t, (const void *)r, (const void *)sigma, (const void *)price,
                          (const void *)delta, (const void *)gamma, (const void *)theta, (const void *)rho, (const void *)vega})
        fast = fast && aligned16(p);
    if (bsm) fast = fast && aligned16(q);
    long long done = 0;
    if (fast && n >= kChunk && aligned16(flag)) {
        const long long nchunks = n / kChunk;
        const int grid = stride_grid(nchunks * 256, PVV_ANALYTIC_RESIDENT);
#define PVV_TMA(M)                                                                                                                     \
    do {                                                                                                                               \
        cudaFuncSetAttribute(analytic_kernel_tma<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)analytic_tma_smem<M>());         \
        analytic_kernel_tma<M><<<grid, 256, analytic_tma_smem<M>(), s>>>(nchunks, flag, S, K, t, r, q, sigma, (double2 *)price,        \
                                                                         (double2 *)delta, (double2 *)gamma, (double2 *)theta,        \
                                                                         (double2 *)rho, (double2 *)vega);                             \
    } while (0)
        if (bsm) PVV_TMA(PVV_MODEL_BLACK_SCHOLES_MERTON);
        else PVV_TMA(PVV_MODEL_BLACK_SCHOLES);
#undef PVV_TMA
        done = nchunks * kChunk;
    }
    if (done < n) {  // the last partial chunk of the streaming form, or the whole call
        const long long m = n - done;
        const int grid = stride_grid(m, PVV_ANALYTIC_CTAS_PER_SM);
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
