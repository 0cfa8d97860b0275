// pvv_kernels.cu — sm_100a kernels and the C ABI (include/pvv.h) of the B200 engine.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false
//        (-fmad=false is part of the numerical contract, see pvv_math.cuh)
//
// Data layout in HBM: structure of arrays — one int8 flag column and one float64 column per input,
// each either dense (stride 1) or a single broadcast scalar (stride 0); one float64 column per
// output (+ one uint8 status column for the IV kernels).  Consecutive threads read consecutive
// elements: every warp-level access is a fully coalesced 256-byte line pair.  The kernels are
// FP64-pipe bound (hundreds to thousands of DP instructions per 50-100 bytes of traffic), so the
// launch shapes are chosen for uniform, balanced warps on the FP64 pipe, not for bandwidth: a CTA
// stages a tile of work in shared memory, counting-sorts it by code path (pvv_tile.cuh) and
// evaluates it in sorted order, several items per thread.  DESIGN.md section 5 has the measurements.
#include <cuda_runtime.h>

#include <atomic>
#include <thread>
#include <vector>
#include <climits>
#include <cstdlib>
#include <cstdio>
#include <new>

#include "../../include/pvv.h"
#include "pvv_math.cuh"
#include "pvv_tile.cuh"

namespace {


std::atomic<long long> g_launches{0};

struct Strides {
    int s[7];
};


__device__ __forceinline__ void note_zde(long long *first_zde, long long i) {
    if (first_zde) atomicMin(first_zde, i);
}

// Sets ST_ZERO_DIVISION in a status byte another kernel wrote (the greeks pass of the IV -> greeks route): an atomic OR on
// the aligned word that holds the byte, because several threads may report the same contract and its neighbours.
__device__ __noinline__ void or_status_zde(uint8_t *status, long long i) {
    const unsigned long long a = (unsigned long long)(status + i);
    atomicOr((unsigned *)(a & ~3ull), (unsigned)pvv::ST_ZERO_DIVISION << (8 * (unsigned)(a & 3ull)));
}

// ---- K1: price --------------------------------------------------------------------------------
// Tile = 1024 contracts per CTA of 256 threads (4 per thread, strided so that global accesses stay
// coalesced).  P1 stages each contract's (x, s, sqrt(F)sqrt(K), intrinsic, deflater) in shared
// memory, P2 evaluates b(x, s) in region-sorted order (pvv_tile.cuh), P3 stores the prices.
#ifndef PVV_TILE_THREADS
#define PVV_TILE_THREADS 256  // measured: 128 / 192 / 256 / 512 threads per CTA (tools/ab_perf.sh), DESIGN.md section 5
#endif
#ifndef PVV_TILE_IPT
#define PVV_TILE_IPT 4
#endif
constexpr int kTileThreads = PVV_TILE_THREADS;
constexpr int kTileCtasPerSm = 1024 / kTileThreads;
constexpr int kItemsPerThread = PVV_TILE_IPT;
constexpr int kTileItems = kTileThreads * kItemsPerThread;
constexpr int kGreekContracts = kTileItems / 8;  // contracts per CTA of the greeks kernel: 8 stencil points each

// shared memory of K1 / K2 (43 KB; dynamic like the larger tiles of K3 / K4)
struct PriceTileShared {
    pvv::PriceItems<kTileItems> items;
    pvv::TileSort<kTileThreads, kItemsPerThread> ts;
};
extern __shared__ __align__(16) unsigned char dyn_smem[];

// P1 of K1 for one contract: _model_calls.py:56-67 up to the call of the undiscounted Black price
template <int MODEL, bool BOUNDED>
__device__ __forceinline__ void stage_price(pvv::PriceItems<kTileItems> &items, int item, double f, double Si, double Ki, double ti,
                                            double ri, double qi, double sg, unsigned &status) {
    double F = Si, D = 1.0;
    if (MODEL != pvv::MODEL_BLACK) {
        D = pvv::exp_g(-ri * ti);
        if (MODEL == pvv::MODEL_BLACK_SCHOLES) {
            if (D == 0.0) status |= pvv::ST_ZERO_DIVISION;
            F = pvv::div_sel<BOUNDED>(Si, D);
        } else {
            F = Si * pvv::exp_g((ri - qi) * ti);
        }
    }
    pvv::stage_price_item<BOUNDED>(items, item, f, F, Ki, pvv::sqrt_sel<BOUNDED>(Ki), sg * pvv::sqrt_sel<BOUNDED>(ti), D, status);
}

template <int MODEL>
__global__ void __launch_bounds__(kTileThreads, kTileCtasPerSm) price_kernel(long long n, const int8_t *__restrict__ flag, const double *__restrict__ S,
                                                             const double *__restrict__ K, const double *__restrict__ t,
                                                             const double *__restrict__ r, const double *__restrict__ q,
                                                             const double *__restrict__ sigma, Strides st, double *__restrict__ price,
                                                             long long *first_zde, long long base) {
    PriceTileShared &psh = *reinterpret_cast<PriceTileShared *>(dyn_smem);
    pvv::PriceItems<kTileItems> &items = psh.items;
    pvv::TileSort<kTileThreads, kItemsPerThread> &ts = psh.ts;
    pvv::tile_sort_init(ts);
    const int tid = threadIdx.x;
    const long long tile_base = (long long)blockIdx.x * kTileItems;
#pragma unroll 1
    for (int j = 0; j < kItemsPerThread; ++j) {
        const int item = j * kTileThreads + tid;
        const long long i = tile_base + item;
        if (i >= n) {
            pvv::stage_null_item(items, item);
            continue;
        }
        const double f = (double)flag[i * st.s[0]];
        const double Si = S[i * st.s[1]], Ki = K[i * st.s[2]], ti = t[i * st.s[3]], sg = sigma[i * st.s[6]];
        const double ri = (MODEL != pvv::MODEL_BLACK) ? r[i * st.s[4]] : 0.0;
        const double qi = (MODEL == pvv::MODEL_BLACK_SCHOLES_MERTON) ? q[i * st.s[5]] : ri;
        unsigned status = 0;
        // ordinary magnitudes (warp-uniform in practice): the guard-free divisions and square roots
        if (pvv::staging_is_bounded(Si, Ki, ti, ri, ri - qi)) stage_price<MODEL, true>(items, item, f, Si, Ki, ti, ri, qi, sg, status);
        else stage_price<MODEL, false>(items, item, f, Si, Ki, ti, ri, qi, sg, status);
        if (status & pvv::ST_ZERO_DIVISION) note_zde(first_zde, base + i);
    }
    pvv::eval_price_items<kTileThreads, kItemsPerThread, MODEL != pvv::MODEL_BLACK>(
        items, ts, [&](int item) { note_zde(first_zde, base + tile_base + item); });
#pragma unroll
    for (int j = 0; j < kItemsPerThread; ++j) {
        const int item = j * kTileThreads + tid;
        const long long i = tile_base + item;
        if (i < n) price[i] = items.x[item];
    }
}

// ---- K2: finite-difference greeks ---------------------------------------------------------------
// `want` bits: 1 price, 2 delta, 4 gamma, 8 theta, 16 rho, 32 vega (stencil points nobody needs are
// staged as null items and cost nothing).
enum : unsigned { W_PRICE = 1, W_DELTA = 2, W_GAMMA = 4, W_THETA = 8, W_RHO = 16, W_VEGA = 32 };

// Tiled K2: 128 contracts per CTA of 256 threads; the 8 stencil points of a contract are 8 work
// items (item = e * 128 + c).  Two threads stage one contract (points 0,1,2,6,7 / points 3,4,5).
// P1 shares every bit-neutral sub-expression between stencil points
// (SURVEY.md 8(a) G7): exp(-r t) once for E0,E1,E2,E6,E7; F, ln(F/K), sqrt(F) once for E0,E6,E7;
// sqrt(K) once; sqrt(t) twice.  P2 = region-sorted evaluation; P3 = the difference quotients.
__device__ __forceinline__ unsigned stencil_need(unsigned want) {
    // stencil points needed by the requested outputs: 0 base, 1/2 S+-dS, 3 t - 1/365, 4/5 r+-0.01, 6/7 sigma+-0.01
    unsigned need = 0;
    if (want & (W_PRICE | W_GAMMA | W_THETA)) need |= 1u;
    if (want & (W_DELTA | W_GAMMA)) need |= 6u;
    if (want & W_THETA) need |= 8u;
    if (want & W_RHO) need |= 0x30u;
    if (want & W_VEGA) need |= 0xc0u;
    return need;
}

// P1 for contract slot c of the tile, done by thread `half` (0: points 0,1,2,6,7; 1: points 3,4,5).
template <bool BSM, int T, bool BOUNDED>
__device__ __forceinline__ void stage_greeks_body(pvv::PriceItems<T> &items, int c, int half, double f, double Si, double Ki, double ti,
                                                  double ri, double sg, double qi, unsigned need, unsigned &status);

template <bool BSM, int T>
__device__ __forceinline__ void stage_greeks(pvv::PriceItems<T> &items, int c, int half, bool valid, double f, double Si, double Ki,
                                             double ti, double ri, double sg, double qi, unsigned need, unsigned &status) {
    constexpr int kGreekContracts = T / 8;  // contracts of this tile (shadows the K2 constant: K4 uses larger sub-tiles)
    if (!valid) {
#pragma unroll
        for (int k = 0; k < 3; ++k) pvv::stage_null_item(items, (half * 3 + k) * kGreekContracts + c);
        if (half == 0) {
            pvv::stage_null_item(items, 6 * kGreekContracts + c);
            pvv::stage_null_item(items, 7 * kGreekContracts + c);
        }
        return;
    }
    // contracts of ordinary magnitude (all of them, in practice) take the copy of the staging code whose divisions
    // and square roots are the dividers' fast paths without guards; the other copy is never fetched
    if (pvv::staging_is_bounded(Si, Ki, ti, ri, BSM ? ri - qi : 0.0)) stage_greeks_body<BSM, T, true>(items, c, half, f, Si, Ki, ti, ri, sg, qi, need, status);
    else stage_greeks_body<BSM, T, false>(items, c, half, f, Si, Ki, ti, ri, sg, qi, need, status);
}

template <bool BSM, int T, bool BOUNDED>
__device__ __forceinline__ void stage_greeks_body(pvv::PriceItems<T> &items, int c, int half, double f, double Si, double Ki, double ti,
                                                  double ri, double sg, double qi, unsigned need, unsigned &status) {
    constexpr int kGreekContracts = T / 8;
    const double dS = .01;
    const double b = BSM ? ri - qi : ri;   // api.py:241: b = r - q on the host side of the reference
    const double qe = BSM ? ri - b : 0.0;  // _numerical_greeks.py: black_scholes_merton(..., r - b)
    const double sqrtK = pvv::sqrt_sel<BOUNDED>(Ki), sqrtT = pvv::sqrt_sel<BOUNDED>(ti);
    double D0 = 0.0, g0 = 0.0;
#pragma unroll 1
    for (int k = 0; k < 3; ++k) {
        const int e = half * 3 + k;
        const int item = e * kGreekContracts + c;
        const bool wanted = ((need >> e) & 1u) || (e == 0 && (need & 0xc6u));
        if (!wanted) {
            pvv::stage_null_item(items, item);
            if (e == 0) {
                pvv::stage_null_item(items, 6 * kGreekContracts + c);
                pvv::stage_null_item(items, 7 * kGreekContracts + c);
            }
            continue;
        }
        const double Se = (e == 1) ? Si + dS : ((e == 2) ? Si - dS : Si);
        const double te = (e == 3) ? ((ti <= 1. / 365.) ? 0.00001 : ti - 1. / 365.) : ti;
        const double re = (e == 4) ? ri + 0.01 : ((e == 5) ? ri - 0.01 : ri);
        const double qv = (e == 4) ? ri - b + 0.01 : ((e == 5) ? ri - b - 0.01 : qe);
        double D, F;
        if (e == 1 || e == 2) {  // same deflater (and carry factor) as the base point
            D = D0;
            F = BSM ? Se * g0 : pvv::div_sel<BOUNDED>(Se, D0);
        } else {
            D = pvv::exp_g(-re * te);
            if (BSM) {
                const double g = pvv::exp_g((re - qv) * te);
                F = Se * g;
                if (e == 0) g0 = g;
            } else {
                if (D == 0.0) status |= pvv::ST_ZERO_DIVISION;
                F = pvv::div_sel<BOUNDED>(Se, D);
            }
            if (e == 0) D0 = D;
        }
        const double sq = (e == 3) ? pvv::sqrt_sel<BOUNDED>(te) : sqrtT;
        pvv::stage_price_item<BOUNDED>(items, item, f, F, Ki, sqrtK, sg * sq, D, status);
        if (e == 0) {  // sigma +- 0.01: same forward, only s changes
#pragma unroll
            for (int v = 6; v < 8; ++v) {
                const int iv_ = v * kGreekContracts + c;
                if ((need >> v) & 1u) {
                    items.x[iv_] = items.x[item];
                    items.s[iv_] = ((v == 6) ? sg + 0.01 : sg - 0.01) * sq;
                    items.scale[iv_] = items.scale[item];
                    items.intr[iv_] = items.intr[item];
                    items.defl[iv_] = D;
                } else {
                    pvv::stage_null_item(items, iv_);
                }
            }
            if (!(need & 1u)) pvv::stage_null_item(items, item);  // base point only served as the template
        }
    }
}

// P3: the difference quotients of _numerical_greeks.py from the eight finished prices of slot c.
template <int T>
__device__ __forceinline__ void finish_greeks(const pvv::PriceItems<T> &items, int c, long long i, double f, double Si, double Ki,
                                              double ti, double *price, double *delta, double *gamma, double *theta, double *rho,
                                              double *vega) {
    constexpr int kGreekContracts = T / 8;
    const double dS = .01;
    double p[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) p[e] = items.x[e * kGreekContracts + c];
    if (price) price[i] = p[0];
    if (ti == 0.0) {
        if (delta) delta[i] = (Si == Ki) ? (f > 0 ? 0.5 : -0.5) : (Si > Ki ? (f > 0 ? 1.0 : 0.0) : (f > 0 ? 0.0 : -1.0));
        if (gamma) gamma[i] = (Si == Ki) ? pvv::u2d(0x7ff0000000000000ull) : 0.0;
    } else {
        if (delta) delta[i] = pvv::div_by_const(p[1] - p[2], 2 * dS, 1.0 / (2 * dS));
        if (gamma) gamma[i] = pvv::div_by_const(p[1] - 2. * p[0] + p[2], 0x1.a36e2eb1c432dp-14, 1.0 / 0x1.a36e2eb1c432dp-14);  // dS ** 2.
    }
    if (theta) theta[i] = p[3] - p[0];
    if (rho) rho[i] = (p[4] - p[5]) / 2.;
    if (vega) vega[i] = (p[6] - p[7]) / 2.;
}

// FULL: every column dense and all six outputs wanted (get_all_greeks / price_dataframe): the stride multiplications, the
// want / need bit tests of the staging loop and the null checks of the stores are compiled out.
// ORSTAT: the greeks pass of the IV -> greeks route: zero divisions are ORed into the status bytes K3 wrote (`status_or`).
// A separate instantiation, because even this cold path costs the plain kernel 1.5 % (4.98 -> 4.90 G/s, A/B on one box).
template <int MODEL, bool FULL, bool ORSTAT = false>
__global__ void __launch_bounds__(kTileThreads, kTileCtasPerSm) greeks_kernel(long long n, const int8_t *__restrict__ flag, const double *__restrict__ S,
                                                                 const double *__restrict__ K, const double *__restrict__ t,
                                                                 const double *__restrict__ r, const double *__restrict__ q,
                                                                 const double *__restrict__ sigma, Strides st, double *__restrict__ price,
                                                                 double *__restrict__ delta, double *__restrict__ gamma,
                                                                 double *__restrict__ theta, double *__restrict__ rho,
                                                                 double *__restrict__ vega, long long *first_zde, long long base,
                                                                 uint8_t *status_or) {
    constexpr bool BSM = (MODEL == pvv::MODEL_BLACK_SCHOLES_MERTON);
    PriceTileShared &psh = *reinterpret_cast<PriceTileShared *>(dyn_smem);
    pvv::PriceItems<kTileItems> &items = psh.items;
    pvv::TileSort<kTileThreads, kItemsPerThread> &ts = psh.ts;
    pvv::tile_sort_init(ts);
    const int tid = threadIdx.x;
    const int c = tid % kGreekContracts;     // contract within the tile
    const int half = tid / kGreekContracts;  // warp-uniform split of the staging work
    const long long i = (long long)blockIdx.x * kGreekContracts + c;
    const unsigned want = FULL ? 63u
                               : (price ? W_PRICE : 0u) | (delta ? W_DELTA : 0u) | (gamma ? W_GAMMA : 0u) | (theta ? W_THETA : 0u) |
                                     (rho ? W_RHO : 0u) | (vega ? W_VEGA : 0u);
    const bool valid = i < n;
    double Si = 0, Ki = 0, ti = 0, f = 1, ri = 0, sg = 0, qi = 0;
    if (valid) {
        if (FULL) {
            f = (double)flag[i];
            Si = S[i];
            Ki = K[i];
            ti = t[i];
            ri = r[i];
            sg = sigma[i];
            if (BSM) qi = q[i];
        } else {
            f = (double)flag[i * st.s[0]];
            Si = S[i * st.s[1]];
            Ki = K[i * st.s[2]];
            ti = t[i * st.s[3]];
            ri = r[i * st.s[4]];
            sg = sigma[i * st.s[6]];
            if (BSM) qi = q[i * st.s[5]];
        }
    }
    unsigned status = 0;
    stage_greeks<BSM, kTileItems>(items, c, half, valid, f, Si, Ki, ti, ri, sg, qi, FULL ? 0xffu : stencil_need(want), status);
    if (status & pvv::ST_ZERO_DIVISION) {
        note_zde(first_zde, base + i);
        if (ORSTAT) or_status_zde(status_or, i);
    }
    __syncthreads();  // the sort classifies items staged by other threads (two threads per contract)
    pvv::eval_price_items<kTileThreads, kItemsPerThread, true>(items, ts, [&](int item) {
        const long long iz = (long long)blockIdx.x * kGreekContracts + (item % kGreekContracts);
        note_zde(first_zde, base + iz);
        if (ORSTAT) or_status_zde(status_or, iz);
    });
    if (!valid || half != 0) return;
    finish_greeks(items, c, i, f, Si, Ki, ti, price, delta, gamma, theta, rho, vega);
}

// ---- K3 / K4: implied volatility ("Let's be rational"), staged over a CTA tile ------------------
// A CTA of NT threads owns C = NT * IPT contracts.  The solver is run as a sequence of stages for the
// whole tile; the per-contract state (8 doubles) lives in SHARED memory between the stages, and
// wherever the reference's control flow would diverge the work is regrouped so that warps stay uniform:
//   B1  b_c = b(x, s_c)              key-sorted batch over all C contracts (pvv_tile.cuh)
//   B2  b_2 = b(x, s_l or s_h)       key-sorted batch
//   G   initial guess                contracts sorted by solver branch (lowest / lower-middle /
//                                    upper-middle / highest); that order is kept to the end
//   B3, B4  b(x, s) of the two Householder(3) passes, key-sorted batches
// Per-contract arithmetic is exactly pvv::normalised_iv (pvv_math.cuh), so results are unchanged.
// Why IPT > 1: the size of the sorted population decides how uniform the warps are, and with one item
// per thread a warp that drew an expensive region cannot be balanced.  Measured (10 M BSM contracts,
// G solves/s): registers-resident state, 512 x 1: 4.30, 1024 x 1: 4.45; shared-memory state (72 B per
// contract), 512 x 2: 4.36, 512 x 3: 5.33, 1024 x 3: 5.49, 768 x 4: 5.19, 384 x 4: 4.98, 512 x 6 (one CTA
// per SM): 4.42; compact state (56 B), 1024 x 3: 5.60, 1024 x 4: 5.66.
#ifndef PVV_IV_REUSE_ORDER
#define PVV_IV_REUSE_ORDER 1
#endif
template <int NT, int IPT>
struct IvTile {
    static constexpr int C = NT * IPT;
    double x[C], beta[C];
    double q[5][C];             // until the guess: b_c, v_c, s_2, b_2, v_2;  afterwards: s, s_left, s_right, b, b'
    unsigned char run[C];       // the contract takes part in the current batch
    unsigned char mode[C];      // iteration objective (MODE_*)
    unsigned char zde[C];       // a batch evaluation divided by zero
    unsigned char stat[C];      // status bits of the front
    pvv::TileSort<NT, IPT> ts;  // order of the b(x, s) batches
    pvv::TileSort<NT, IPT> tg;  // order by solver branch: fixed from the initial guess to the end
};
// 64 bytes per contract: three contracts per thread fit the 227 KB of a 1024-thread CTA.  The centre node
// s_c = sqrt(|2x|) and b_max = exp(x/2) are recomputed where they are used (b_max only on the upper
// branches) instead of being stored: same expressions, same bits.  Every batch returns the normalised vega b'(x, s)
// next to b(x, s): in the two common regions it shares the evaluator's exponential (nbc_eval_core), which saves the
// solver four divisions and four exponentials per contract.

// b(x[item], s_of(item)) -> out[item] for every running contract of the tile, in key-sorted order.
// Entered with all writes to x / s / run visible (barrier by the caller); ends with a barrier.
// RESORT = false reuses the order of the previous batch (the second Householder pass: s has moved by one converging
// step, nearly every item still groups with the same neighbours) and takes each item's exact region from
// nbc_classify instead of the sort key: one counting sort, its two barriers and the key refinement less per tile.
template <bool RESORT = true, int NT, int IPT, class SOf>
__device__ __forceinline__ void iv_tile_batch(IvTile<NT, IPT> &sh, SOf s_of, double *out, double *out_vega) {
    const int tid = threadIdx.x;
    if (RESORT) pvv::tile_sort(sh.ts, [&](int c) { return sh.run[c] ? pvv::nbc_sort_key(sh.x[c], s_of(c)) : (int)pvv::K_ZERO; });
#pragma unroll 1
    for (int j = 0; j < IPT; ++j) {
        const unsigned e = sh.ts.order[j * NT + tid];
        const int item = e & 0xfffu;
        const double xi = sh.x[item], si = s_of(item);
        const int region = RESORT ? pvv::key_region(e >> 12) : (sh.run[item] ? pvv::nbc_classify(xi, si) : (int)pvv::R_ZERO);
        if (pvv::nbc_zde(region, xi, si)) sh.zde[item] = 1;
        const pvv::D2 bv = pvv::nbc_eval_bv(region, xi, si);
        out[item] = bv.a;
        out_vega[item] = bv.b;
    }
    __syncthreads();
}

// Everything between the front and sigma = s / sqrt(t).  Entered with x, beta, run, zde = 0 filled in
// natural order by the owner threads (contract c = j * NT + tid); leaves s in q[0] (0 where the
// reference returns 0 without solving), zde flags in zde[]; ends with a barrier.
template <int NT, int IPT>
__device__ __forceinline__ void iv_tile_solve2(IvTile<NT, IPT> &sh) {
    const int tid = threadIdx.x;
    double *b_c = sh.q[0], *v_c = sh.q[1], *s_2 = sh.q[2], *b_2 = sh.q[3], *v_2 = sh.q[4];
    const auto s_c = [&](int c) { return sqrt(fabs(2 * sh.x[c])); };
    // the centre node is computed once per contract into the (still unused) s_2 slot: the sort key, the evaluation and
    // the second node read it from there; the guess, which runs after s_2 has replaced it, computes it again
#pragma unroll
    for (int j = 0; j < IPT; ++j) s_2[j * NT + tid] = s_c(j * NT + tid);
    iv_tile_batch(sh, [&](int c) { return s_2[c]; }, b_c, v_c);  // own slots only so far: no barrier needed before the sort
#pragma unroll 1
    for (int j = 0; j < IPT; ++j) {
        const int c = j * NT + tid;
        if (!sh.run[c]) continue;
        const double x = sh.x[c], sc = s_2[c], beta = sh.beta[c], bc = b_c[c];
        const double b_max = (beta < bc) ? 0.0 : pvv::exp_g(0.5 * x);  // the lower side does not look at it
        s_2[c] = pvv::iv_second_node(beta, b_max, sc, bc, v_c[c]);
    }
    iv_tile_batch(sh, [&](int c) { return s_2[c]; }, b_2, v_2);
    // initial guess, contracts regrouped by branch.  From here on thread (j, tid) works on contract
    // tg.order[j * NT + tid]: the tile stays grouped by solver branch through the Householder passes.
    pvv::tile_sort(sh.tg, [&](int c) { return sh.run[c] ? pvv::iv_branch(sh.beta[c], b_c[c], b_2[c]) : (int)pvv::IV_DONE; });
    double *s = sh.q[0], *s_left = sh.q[1], *s_right = sh.q[2], *b = sh.q[3], *bp = sh.q[4];
#pragma unroll 1
    for (int j = 0; j < IPT; ++j) {
        const unsigned e = sh.tg.order[j * NT + tid];
        const int item = e & 0xfffu, branch = e >> 12;
        pvv::IvState o;
        o.s = 0.0;
        o.s_left = 0.0;
        o.s_right = 0.0;
        o.mode = pvv::MODE_MIDDLE;
        bool running = branch != pvv::IV_DONE;
        if (running) {
            const double x = sh.x[item];
            const double b_max = (branch == pvv::IV_HIGHEST) ? pvv::exp_g(0.5 * x) : 0.0;  // only the highest segment uses it
            o = pvv::iv_initial_guess(branch, x, sh.beta[item], b_max, s_c(item), b_c[item], v_c[item], s_2[item], b_2[item], v_2[item]);
            running = pvv::iv_iteration_head(0, -DBL_MAX, o);
        }
        s[item] = o.s;  // the item's inputs are consumed: its slots become the iteration state
        s_left[item] = o.s_left;
        s_right[item] = o.s_right;
        sh.mode[item] = (unsigned char)o.mode;
        sh.run[item] = running;
    }
    __syncthreads();
#pragma unroll 1
    for (int it = 0; it < 2; ++it) {
        if (it == 0 || !PVV_IV_REUSE_ORDER) iv_tile_batch<true>(sh, [&](int c) { return s[c]; }, b, bp);
        else iv_tile_batch<false>(sh, [&](int c) { return s[c]; }, b, bp);
#pragma unroll 1
        for (int j = 0; j < IPT; ++j) {
            const int item = sh.tg.order[j * NT + tid] & 0xfffu;
            if (!sh.run[item]) continue;
            pvv::IvState o;
            o.s = s[item];
            o.s_left = s_left[item];
            o.s_right = s_right[item];
            o.mode = sh.mode[item];
            const double x = sh.x[item];
            const double b_max = (o.mode == pvv::MODE_HIGHEST) ? pvv::exp_g(0.5 * x) : 0.0;  // only that objective uses it
            const double ds = pvv::iv_iteration_tail(x, sh.beta[item], b_max, b[item], bp[item], o);
            bool running = false;
            if (it == 0) running = pvv::iv_iteration_head(1, ds, o);
            s[item] = o.s;
            s_left[item] = o.s_left;
            s_right[item] = o.s_right;
            sh.run[item] = running;
        }
        __syncthreads();
    }
}

// Front of the tile in natural order (contract c = j * NT + tid, coalesced loads): implied_volatility.py:48-86
// and the first lines of the numba loop -> x, beta, the run flag and the status bits of every contract.
template <int MODEL, int NT, int IPT>
__device__ __forceinline__ void iv_tile_front(IvTile<NT, IPT> &sh, long long tile0, long long n, const int8_t *__restrict__ flag,
                                              const double *__restrict__ price, const double *__restrict__ S, const double *__restrict__ K,
                                              const double *__restrict__ t, const double *__restrict__ r, const double *__restrict__ q,
                                              const Strides &st) {
    const int tid = threadIdx.x;
    if (tid < pvv::kSortKeys) {
        sh.ts.cnt[tid] = 0;
        sh.tg.cnt[tid] = 0;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < IPT; ++j) {
        const int c = j * NT + tid;
        const long long i = tile0 + c;
        unsigned status = 0;
        pvv::IvFront fr;
        fr.beta = 0.0;
        fr.x = 0.0;
        if (i < n) {
            const double qq = (MODEL == pvv::MODEL_BLACK_SCHOLES_MERTON) ? q[i * st.s[6]] : 0.0;
            fr = pvv::iv_front<MODEL>((double)flag[i * st.s[0]], price[i * st.s[1]], S[i * st.s[2]], K[i * st.s[3]], t[i * st.s[4]],
                                      r[i * st.s[5]], qq, status);
        }
        sh.x[c] = fr.x;
        sh.beta[c] = fr.beta;
        sh.run[c] = (i < n) && !(fr.beta <= 0);  // beta <= 0: the reference returns 0 without solving
        sh.zde[c] = 0;
        sh.stat[c] = (unsigned char)status;
    }
}

// sigma = s / sqrt(t) with the on_error mask, and the final status byte of contract c
template <int NT, int IPT>
__device__ __forceinline__ double iv_tile_vol(const IvTile<NT, IPT> &sh, int c, double ti, int mask_errors, unsigned &status) {
    status = sh.stat[c];
    if (sh.zde[c]) status |= pvv::ST_ZERO_DIVISION;
    const double vol = pvv::div_zn(sh.q[0][c], sqrt(ti));  // s == 0 where the reference returns 0 without solving
    return (mask_errors && (status & (pvv::ST_BELOW_INTRINSIC | pvv::ST_ABOVE_MAXIMUM))) ? pvv::u2d(0x7ff8000000000000ull) : vol;
}

// K3 comes in two shapes: 1024 threads x 3 contracts (one CTA per SM, up to 3072 contracts, 189 KB of shared
// memory), and 512 threads x 3 contracts (two CTAs per SM, up to 1536 contracts each) for calls of a few thousand
// contracts.  Tiles are filled evenly (iv_tile_plan), so a partial last wave no longer decides between them; measured
// with tools/probe_iv_sizes.py the big shape wins from 64 Ki contracts on (128 Ki: 43 vs 48 us, 512 Ki: 107 vs 117 us,
// 1 M: 184 vs 194 us).
#ifndef PVV_IV_NT
#define PVV_IV_NT 1024
#define PVV_IV_IPT 3
#endif
constexpr long long kIvBigTileMin = 65536;  // with even tiles the big shape wins from 64 Ki contracts on (tools/probe_iv_sizes.py)

template <int MODEL, int NT, int IPT>
__global__ void __launch_bounds__(NT, (NT > 512 ? 1 : 1024 / NT)) iv_kernel(long long n, const int8_t *__restrict__ flag, const double *__restrict__ price,
                                                           const double *__restrict__ S, const double *__restrict__ K,
                                                           const double *__restrict__ t, const double *__restrict__ r,
                                                           const double *__restrict__ q, Strides st, int mask_errors,
                                                           double *__restrict__ iv, uint8_t *__restrict__ status_out, long long *first_zde,
                                                           long long base, int cpt) {
    using Tile = IvTile<NT, IPT>;
    Tile &sh = *reinterpret_cast<Tile *>(dyn_smem);
    // the CTA owns contracts [tile0, tile0 + cpt), cpt <= Tile::C chosen by the launch so that the tiles fill whole
    // waves evenly (iv_tile_plan); slots beyond cpt are treated like the ragged tail of the column
    const long long tile0 = (long long)blockIdx.x * cpt;
    if (tile0 + cpt < n) n = tile0 + cpt;
    iv_tile_front<MODEL>(sh, tile0, n, flag, price, S, K, t, r, q, st);
    iv_tile_solve2(sh);
#pragma unroll 1
    for (int j = 0; j < IPT; ++j) {
        const int c = j * NT + threadIdx.x;
        const long long i = tile0 + c;
        if (i >= n) continue;
        unsigned status;
        iv[i] = iv_tile_vol(sh, c, t[i * st.s[4]], mask_errors, status);
        if (status_out) status_out[i] = (uint8_t)status;
        if (status & pvv::ST_ZERO_DIVISION) note_zde(first_zde, base + i);
    }
}

// ---- K4: implied volatility, then the greeks at the solved volatility (api.py:142-178) ------------
// One CTA = 512 threads, 1536 contracts.  Phase 1 is K3's staged solve (sigma stays on chip); phase 2
// runs the K2 tile machinery six times, 256 contracts (2048 stencil items) at a time, over the same
// shared memory: all 512 threads stage (two per contract) and each evaluates 4 sorted items.
// The contract inputs are re-read from global memory for phase 2 (L2-resident by then on the chunked host pipeline
// and for device-resident callers; the host entry point keeps zero-copy to short calls for that reason).
constexpr int kIvgThreads = 512, kIvgIpt = 3;
constexpr int kIvgTile = kIvgThreads * kIvgIpt;  // contracts per CTA
constexpr int kIvgItems = 2048;                  // stencil items per sub-tile (4 per thread)
constexpr int kIvgContracts = kIvgItems / 8;     // 256 contracts per sub-tile: every thread stages one half of a contract
static_assert(kIvgTile % kIvgContracts == 0, "sub-tiles must cover the tile");
using IvTile4 = IvTile<kIvgThreads, kIvgIpt>;
struct IvGreeksShared {
    pvv::PriceItems<kIvgItems> items;
    pvv::TileSort<kIvgThreads, kIvgItems / kIvgThreads> ts;
    unsigned char zf[3][kIvgContracts];  // zero-division flags of a sub-tile: staging half 0 / half 1 / evaluation
    double vol[kIvgTile];                // phase 1 -> phase 2
    unsigned char stat[kIvgTile];
};
union IvGreeksUnion {
    IvTile4 iv;
    IvGreeksShared g;
};

template <int MODEL>
__global__ void __launch_bounds__(kIvgThreads, 2) iv_greeks_kernel(long long n, const int8_t *__restrict__ flag,
                                                                   const double *__restrict__ price, const double *__restrict__ S,
                                                                   const double *__restrict__ K, const double *__restrict__ t,
                                                                   const double *__restrict__ r, const double *__restrict__ q, Strides st,
                                                                   int mask_errors, double *__restrict__ iv, uint8_t *__restrict__ status_out,
                                                                   double *__restrict__ delta, double *__restrict__ gamma,
                                                                   double *__restrict__ theta, double *__restrict__ rho,
                                                                   double *__restrict__ vega, long long *first_zde, long long base) {
    constexpr bool BSM = (MODEL == pvv::MODEL_BLACK_SCHOLES_MERTON);
    IvTile4 &sh = reinterpret_cast<IvGreeksUnion *>(dyn_smem)->iv;
    IvGreeksShared &gs = reinterpret_cast<IvGreeksUnion *>(dyn_smem)->g;
    const int tid = threadIdx.x;
    const long long tile0 = (long long)blockIdx.x * kIvgTile;
    iv_tile_front<MODEL>(sh, tile0, n, flag, price, S, K, t, r, q, st);
    iv_tile_solve2(sh);
    // sigma leaves the solver's arrays through registers: phase 2's arrays overlay them
    double vol[kIvgIpt];
    unsigned status[kIvgIpt];
#pragma unroll
    for (int j = 0; j < kIvgIpt; ++j) {
        const int c = j * kIvgThreads + tid;
        const long long i = tile0 + c;
        vol[j] = iv_tile_vol(sh, c, i < n ? t[i * st.s[4]] : 1.0, mask_errors, status[j]);
        if (i < n) iv[i] = vol[j];
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kIvgIpt; ++j) {
        gs.vol[j * kIvgThreads + tid] = vol[j];
        gs.stat[j * kIvgThreads + tid] = (unsigned char)status[j];
    }
    pvv::tile_sort_init(gs.ts);
    const unsigned want = (delta ? W_DELTA : 0u) | (gamma ? W_GAMMA : 0u) | (theta ? W_THETA : 0u) | (rho ? W_RHO : 0u) | (vega ? W_VEGA : 0u);
    const unsigned need = stencil_need(want);
    const int c = tid % kIvgContracts, half = tid / kIvgContracts;
#pragma unroll 1
    for (int sub = 0; sub < kIvgTile / kIvgContracts; ++sub) {
        __syncthreads();
        const long long gi = tile0 + sub * kIvgContracts + c;
        const bool gvalid = gi < n;
        double Si = 0, Ki = 0, ti = 0, f = 1, ri = 0, qi = 0, sg = 0;
        if (gvalid) {
            f = (double)flag[gi * st.s[0]];
            Si = S[gi * st.s[2]];
            Ki = K[gi * st.s[3]];
            ti = t[gi * st.s[4]];
            ri = r[gi * st.s[5]];
            if (BSM) qi = q[gi * st.s[6]];
            sg = gs.vol[sub * kIvgContracts + c];
        }
        unsigned sst = 0;
        stage_greeks<BSM, kIvgItems>(gs.items, c, half, gvalid, f, Si, Ki, ti, ri, sg, qi, need, sst);
        gs.zf[half][c] = (sst & pvv::ST_ZERO_DIVISION) ? 1 : 0;
        if (half == 0) gs.zf[2][c] = 0;
        __syncthreads();  // the sort classifies items staged by other threads
        pvv::eval_price_items<kIvgThreads, kIvgItems / kIvgThreads, true>(gs.items, gs.ts, [&](int item) { gs.zf[2][item % kIvgContracts] = 1; });
        if (gvalid && half == 0) finish_greeks(gs.items, c, gi, f, Si, Ki, ti, nullptr, delta, gamma, theta, rho, vega);
        __syncthreads();  // zf[2] written during the evaluation
        if (half == 0 && (gs.zf[0][c] | gs.zf[1][c] | gs.zf[2][c])) gs.stat[sub * kIvgContracts + c] |= (unsigned char)pvv::ST_ZERO_DIVISION;
    }
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < kIvgIpt; ++j) {
        const int cc = j * kIvgThreads + tid;
        const long long i = tile0 + cc;
        if (i >= n) continue;
        const unsigned stt = gs.stat[cc];
        if (status_out) status_out[i] = (uint8_t)stt;
        if (stt & pvv::ST_ZERO_DIVISION) note_zde(first_zde, base + i);
    }
}

// ---- diagnostics ---------------------------------------------------------------------------------
__global__ void unary_kernel(int which, long long n, const double *__restrict__ x, double *__restrict__ y) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double v = x[i];
    double o;
    switch (which) {
        case 0: o = pvv::exp_g(v); break;
        case 1: o = pvv::log_g(v); break;
        case 2: o = pvv::erfcx_cody(v); break;
        case 3: o = pvv::erfc_cody(v); break;
        case 4: o = pvv::norm_cdf(v); break;
        case 6: o = pvv::pow_g(v, 1. / 3.); break;
        case 7: o = pvv::sqrt_bounded(v); break;
        case 8: o = pvv::div_bounded(v, x[i ^ 1]); break;  // every element divided by its pair partner (n even)
        default: o = pvv::inverse_norm_cdf(v); break;
    }
    y[i] = o;
}

// Dependent chains of DP instructions, 8 independent chains per thread: measures the FP64 issue
// rate the roofline of the FP64-bound kernels is quoted against (MEASURED_PEAKS.json has none).
template <bool FMA>
__global__ void fp64_probe_kernel(int iters, double *out) {
    double a[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    const double m = 1.0000000001, c = 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (FMA) {
                a[j] = __fma_rn(a[j], m, c);
                a[j] = __fma_rn(a[j], m, c);
            } else {
                a[j] = __dmul_rn(a[j], m);
                a[j] = __dadd_rn(a[j], c);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
}

// The same with CH independent chains per thread and an occupancy set by the caller (dynamic shared memory as
// ballast): how the FP64 issue rate depends on instruction-level parallelism and on resident warps (tools/probe_ilp.py;
// the numbers behind DESIGN.md section 5's "what does not limit the kernels").
template <int CH>
__global__ void fp64_ilp_probe_kernel(int iters, double *out) {
    double a[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) a[j] = 1.0 + 1e-9 * (threadIdx.x + j);
    const double m = 1.0000000001, c = 1e-12;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 8 / CH; ++rep) {
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                a[j] = __dmul_rn(a[j], m);
                a[j] = __dadd_rn(a[j], c);
            }
        }
    }
    double s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += a[j];
    if (s == 123.456) out[0] = s;
}

inline int grid_tiles(long long n, int per_cta) { return (int)((n + per_cta - 1) / per_cta); }

// dynamic shared memory above the 48 KB default needs the opt-in once per kernel (and per device: cheap, so every launch)
template <class Kern>
inline void allow_smem(Kern k, size_t bytes) {
    if (bytes > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// SMs of the current device (cached per device: a process may drive several)
inline int sm_count() {
    static int cache[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (!cache[dev]) {
        int v = 0;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = v > 0 ? v : 148;
    }
    return cache[dev];
}

// Contracts per tile for a K3 launch.  A tile is a long sequential program (50-70 us), so a last wave that fills a
// fraction of the SMs costs a whole tile time, and a short column occupies only n / capacity SMs.  Instead of full
// tiles plus a remainder, the column is cut into waves x slots equal tiles (slots = SMs x resident CTAs), each
// a multiple of 32 contracts so that every tile starts on a 256-byte boundary of the columns.  PVV_IV_EVEN_TILES=0
// restores full tiles (A/B runs).
inline int iv_tile_plan_for(long long n, int capacity, long long slots) {
    const long long waves = (n + slots * capacity - 1) / (slots * capacity);
    long long cpt = (n + waves * slots - 1) / (waves * slots);
    cpt = (cpt + 31) / 32 * 32;
    return (int)(cpt > capacity ? capacity : cpt);
}
inline int iv_tile_plan(long long n, int capacity, int ctas_per_sm) {
    static const bool even = [] { const char *e = getenv("PVV_IV_EVEN_TILES"); return !(e && e[0] == '0'); }();
    if (!even) return capacity;
    return iv_tile_plan_for(n, capacity, (long long)sm_count() * ctas_per_sm);
}
// host_path: launches of the host entry points (chunks of the copy pipeline, zero-copy calls).  There the kernel shares
// the GPU with the DMA traffic of its neighbours, and the two-CTA shape is the better citizen up to four full waves of
// big tiles: 10 M BSM contracts end to end 0.960 vs 0.938 G IV/s, A/B on the same box.
inline long long iv_big_tile_min(bool host_path) {
    static const long long v = [] { const char *e = getenv("PVV_IV_BIG_MIN"); return e ? atoll(e) : -1ll; }();
    if (v >= 0) return v;
    return host_path ? 4ll * 148 * PVV_IV_NT * PVV_IV_IPT : kIvBigTileMin;
}

template <int MODEL>
inline void launch_iv_shape(long long n, const int8_t *flag, const double *price, const double *S, const double *K, const double *t,
                            const double *r, const double *q, const Strides &st, int mask_errors, double *iv, uint8_t *status, long long *fz,
                            long long base, cudaStream_t s, bool host_path) {
    if (n >= iv_big_tile_min(host_path)) {
        using Tile = IvTile<PVV_IV_NT, PVV_IV_IPT>;
        const int cpt = iv_tile_plan(n, Tile::C, 1);
        allow_smem(iv_kernel<MODEL, PVV_IV_NT, PVV_IV_IPT>, sizeof(Tile));
        iv_kernel<MODEL, PVV_IV_NT, PVV_IV_IPT><<<grid_tiles(n, cpt), PVV_IV_NT, sizeof(Tile), s>>>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, fz, base, cpt);
    } else {
        using Tile = IvTile<512, 3>;
        const int cpt = iv_tile_plan(n, Tile::C, 2);
        allow_smem(iv_kernel<MODEL, 512, 3>, sizeof(Tile));
        iv_kernel<MODEL, 512, 3><<<grid_tiles(n, cpt), 512, sizeof(Tile), s>>>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, fz, base, cpt);
    }
}

inline bool fill_strides(const int64_t *stride, Strides &out) {
    for (int i = 0; i < 7; ++i) {
        if (stride[i] != 0 && stride[i] != 1) return false;
        out.s[i] = (int)stride[i];
    }
    return true;
}

inline int launch_status() {
    cudaError_t e = cudaGetLastError();
    return (int)e;
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

int pvv_version(void) { return 100; }

const char *pvv_error_string(int code) {
    if (code == 0) return "success";
    if (code == PVV_E_INVALID_ARGUMENT) return "pvv: invalid argument";
    if (code == PVV_E_NO_DEVICE) return "pvv: no CUDA device";
    if (code == PVV_E_OUT_OF_MEMORY) return "pvv: out of memory";
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "pvv: unknown error";
}

int64_t pvv_launch_count(void) { return g_launches.load(); }
void pvv_note_launch(void) { g_launches++; }

static int launch_price(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                        const double *q, const double *sigma, const int64_t stride[7], double *price, int64_t *first_zde, long long base,
                        void *stream) {
    Strides st;
    if (n < 0 || model < 0 || model > 2 || !fill_strides(stride, st)) return PVV_E_INVALID_ARGUMENT;
    if (model == 2 && !q) return PVV_E_INVALID_ARGUMENT;
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    long long *fz = (long long *)first_zde;
    if (model == 0) { allow_smem(price_kernel<0>, sizeof(PriceTileShared)); price_kernel<0><<<grid_tiles(n, kTileItems), kTileThreads, sizeof(PriceTileShared), s>>>(n, flag, S, K, t, r, q, sigma, st, price, fz, base); }
    else if (model == 1) { allow_smem(price_kernel<1>, sizeof(PriceTileShared)); price_kernel<1><<<grid_tiles(n, kTileItems), kTileThreads, sizeof(PriceTileShared), s>>>(n, flag, S, K, t, r, q, sigma, st, price, fz, base); }
    else { allow_smem(price_kernel<2>, sizeof(PriceTileShared)); price_kernel<2><<<grid_tiles(n, kTileItems), kTileThreads, sizeof(PriceTileShared), s>>>(n, flag, S, K, t, r, q, sigma, st, price, fz, base); }
    g_launches++;
    return launch_status();
}

static int launch_greeks(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                         const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta, double *gamma,
                         double *theta, double *rho, double *vega, int64_t *first_zde, long long base, void *stream,
                         uint8_t *status_or = nullptr) {
    Strides st;
    if (n < 0 || model < 0 || model > 2 || !fill_strides(stride, st)) return PVV_E_INVALID_ARGUMENT;
    if (model == 2 && !q) return PVV_E_INVALID_ARGUMENT;
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    long long *fz = (long long *)first_zde;
    bool full = price && delta && gamma && theta && rho && vega;
    for (int k = 0; k < 7 && full; ++k) full = st.s[k] == 1 || (k == 5 && model != 2);
    const int grid = grid_tiles(n, kGreekContracts);
#define PVV_GREEKS(M, F, O)                                                                                                         \
    do {                                                                                                                            \
        allow_smem(greeks_kernel<M, F, O>, sizeof(PriceTileShared));                                                                \
        greeks_kernel<M, F, O><<<grid, kTileThreads, sizeof(PriceTileShared), s>>>(n, flag, S, K, t, r, q, sigma, st, price, delta, \
                                                                                   gamma, theta, rho, vega, fz, base, status_or);   \
    } while (0)
    if (status_or) {
        if (model == 2) PVV_GREEKS(2, false, true);
        else PVV_GREEKS(1, false, true);
    } else if (model == 2) {
        if (full) PVV_GREEKS(2, true, false);
        else PVV_GREEKS(2, false, false);
    } else {
        if (full) PVV_GREEKS(1, true, false);
        else PVV_GREEKS(1, false, false);
    }
#undef PVV_GREEKS
    g_launches++;
    return launch_status();
}

static int launch_iv(int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K, const double *t,
                     const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv, uint8_t *status,
                     int64_t *first_zde, long long base, void *stream, bool host_path = false) {
    Strides st;
    if (n < 0 || model < 0 || model > 2 || !fill_strides(stride, st)) return PVV_E_INVALID_ARGUMENT;
    if (model == 2 && !q) return PVV_E_INVALID_ARGUMENT;
    if (n == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    long long *fz = (long long *)first_zde;
    if (model == 0) launch_iv_shape<0>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, fz, base, s, host_path);
    else if (model == 1) launch_iv_shape<1>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, fz, base, s, host_path);
    else launch_iv_shape<2>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, fz, base, s, host_path);
    g_launches++;
    return launch_status();
}

// PVV_IVG_FUSED=1 routes device-resident and chunked IV -> greeks calls through the single fused kernel again (A/B runs;
// read at every call, so a process can time both).
static bool ivg_fused_env() {
    const char *e = getenv("PVV_IVG_FUSED");
    return e && e[0] == '1';
}

static int launch_iv_greeks(int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K,
                            const double *t, const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv,
                            uint8_t *status, double *delta, double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde,
                            long long base, void *stream, bool fused, bool host_path = false) {
    Strides st;
    if (n < 0 || model < 0 || model > 2 || !iv || !fill_strides(stride, st)) return PVV_E_INVALID_ARGUMENT;
    if (model == 2 && !q) return PVV_E_INVALID_ARGUMENT;
    if (n == 0) return 0;
    if (!fused) {
        // Two launches on one stream: K3 writes sigma (the iv column, masked as the caller asked) and the status bytes, K2
        // evaluates the greeks at it and ORs its own zero divisions into them.  Every size measured prefers this to the
        // fused kernel (12.5 M contracts: 2.88 vs 2.67 G/s; 128 Ki: 1.51 vs 1.06; tools/probe_k4_split.py): each pass
        // keeps its own tile shape (3072 contracts per CTA for the solve, four 128-contract CTAs per SM for the stencil),
        // and the 16 B per contract that sigma spends in HBM are invisible next to 3000 DP instructions.
        int rc = launch_iv(model, n, flag, price, S, K, t, r, q, stride, mask_errors, iv, status, first_zde, base, stream, host_path);
        if (rc != 0 || !(delta || gamma || theta || rho || vega)) return rc;
        const int64_t gst[7] = {stride[0], stride[2], stride[3], stride[4], stride[5], stride[6], 1};
        return launch_greeks(model == 2 ? 2 : 1, n, flag, S, K, t, r, q, iv, gst, nullptr, delta, gamma, theta, rho, vega, first_zde,
                             base, stream, status);
    }
    cudaStream_t s = (cudaStream_t)stream;
    long long *fz = (long long *)first_zde;
    if (model == 0)
        { allow_smem(iv_greeks_kernel<0>, sizeof(IvGreeksUnion)); iv_greeks_kernel<0><<<grid_tiles(n, kIvgTile), kIvgThreads, sizeof(IvGreeksUnion), s>>>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, delta, gamma, theta, rho, vega, fz, base); }
    else if (model == 1)
        { allow_smem(iv_greeks_kernel<1>, sizeof(IvGreeksUnion)); iv_greeks_kernel<1><<<grid_tiles(n, kIvgTile), kIvgThreads, sizeof(IvGreeksUnion), s>>>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, delta, gamma, theta, rho, vega, fz, base); }
    else
        { allow_smem(iv_greeks_kernel<2>, sizeof(IvGreeksUnion)); iv_greeks_kernel<2><<<grid_tiles(n, kIvgTile), kIvgThreads, sizeof(IvGreeksUnion), s>>>(n, flag, price, S, K, t, r, q, st, mask_errors, iv, status, delta, gamma, theta, rho, vega, fz, base); }
    g_launches++;
    return launch_status();
}

int pvv_price_dev(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                  const double *q, const double *sigma, const int64_t stride[7], double *price, int64_t *first_zde, void *stream) {
    return launch_price(model, n, flag, S, K, t, r, q, sigma, stride, price, first_zde, 0, stream);
}
int pvv_greeks_dev(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                   const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta, double *gamma,
                   double *theta, double *rho, double *vega, int64_t *first_zde, void *stream) {
    return launch_greeks(model, n, flag, S, K, t, r, q, sigma, stride, price, delta, gamma, theta, rho, vega, first_zde, 0, stream);
}
int pvv_iv_dev(int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K, const double *t,
               const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv, uint8_t *status,
               int64_t *first_zde, void *stream) {
    return launch_iv(model, n, flag, price, S, K, t, r, q, stride, mask_errors, iv, status, first_zde, 0, stream);
}
int pvv_iv_greeks_dev(int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K, const double *t,
                      const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv, uint8_t *status,
                      double *delta, double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde, void *stream) {
    return launch_iv_greeks(model, n, flag, price, S, K, t, r, q, stride, mask_errors, iv, status, delta, gamma, theta, rho, vega,
                            first_zde, 0, stream, ivg_fused_env());
}

// host logic only (no device involved): contracts per K3 tile for a column of n contracts, tiles of at most `capacity`
// contracts and `slots` = SMs x resident CTAs (tests/test_host_logic.py checks the plan's invariants)
int pvv_iv_tile_plan(int64_t n, int capacity, int slots) {
    if (n <= 0 || capacity < 32 || capacity % 32 || slots <= 0) return PVV_E_INVALID_ARGUMENT;
    return iv_tile_plan_for(n, capacity, slots);
}

int pvv_unary_dev(int which, int64_t n, const double *x, double *y, void *stream) {
    if (n <= 0) return n == 0 ? 0 : PVV_E_INVALID_ARGUMENT;
    unary_kernel<<<(int)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(which, n, x, y);
    g_launches++;
    return launch_status();
}

int pvv_fp64_probe(int64_t threads, int iters, int use_fma, float *ms, void *stream) {
    cudaStream_t s = (cudaStream_t)stream;
    double *out = nullptr;
    cudaError_t e = cudaMalloc(&out, 8);
    if (e != cudaSuccess) return (int)e;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    const int block = 256;
    const int grid = (int)((threads + block - 1) / block);
    for (int rep = 0; rep < 2; ++rep) {  // first pass warms up
        cudaEventRecord(a, s);
        if (use_fma) fp64_probe_kernel<true><<<grid, block, 0, s>>>(iters, out);
        else fp64_probe_kernel<false><<<grid, block, 0, s>>>(iters, out);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
    }
    g_launches += 2;
    cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    return launch_status();
}

// chains: 1, 2, 4 or 8 dependent DMUL+DADD chains per thread (16 DP instructions per thread and iteration in every
// case); warps_per_sm: resident warps (multiple of 4, at most 64), enforced with shared-memory ballast.
int pvv_fp64_ilp_probe(int chains, int warps_per_sm, int iters, float *ms, void *stream) {
    if ((chains != 1 && chains != 2 && chains != 4 && chains != 8) || warps_per_sm < 4 || warps_per_sm > 64 || warps_per_sm % 4) return PVV_E_INVALID_ARGUMENT;
    cudaStream_t s = (cudaStream_t)stream;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *out = nullptr;
    cudaError_t e = cudaMalloc(&out, 8);
    if (e != cudaSuccess) return (int)e;
    const int block = 128;  // 4 warps: one per scheduler
    const int ctas_per_sm = warps_per_sm / 4;
    const size_t ballast = (size_t)(227 * 1024) / ctas_per_sm - 1024 - 256;  // leaves no room for one more CTA
    void (*k)(int, double *) = chains == 1 ? fp64_ilp_probe_kernel<1> : chains == 2 ? fp64_ilp_probe_kernel<2> : chains == 4 ? fp64_ilp_probe_kernel<4> : fp64_ilp_probe_kernel<8>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ballast);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(a, s);
        k<<<sms * ctas_per_sm, block, ballast, s>>>(iters, out);
        cudaEventRecord(b, s);
        cudaEventSynchronize(b);
    }
    g_launches += 2;
    cudaEventElapsedTime(ms, a, b);
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(out);
    return launch_status();
}

// -------------------------------------------------------------------------------------------------
// Host-buffer pipeline: chunks of `chunk` contracts round-robin over `nstreams` streams; each stream
// owns one device slot (inputs + outputs).  H2D(k+1), kernel(k) and D2H(k-1) overlap across streams
// on the two copy engines.  Dense columns are copied chunk by chunk, broadcast scalars once.
// With pinned (cudaHostAlloc / cudaHostRegister) user buffers every copy is a true async DMA; with
// pageable buffers cudaMemcpyAsync falls back to the driver's staged copy (correct, slower).
// -------------------------------------------------------------------------------------------------
}  // extern "C"

struct pvv_ctx {
    int device;
    int nstreams;
    long long chunk;
    cudaStream_t stream[8];
    char *slot[8];        // device memory of one chunk: 8 double columns in, 6 double columns out, flag, status
    double *scalars;      // device copies of broadcast scalars: 8 doubles + 8 bytes for the flag
    long long *first_zde; // device int64
};

namespace {
constexpr int kInCols = 7;   // double input columns per slot (price,S,K,t,r,q,sigma — a superset of every entry point)
constexpr int kOutCols = 6;  // double output columns per slot

inline size_t slot_bytes(long long chunk) { return (size_t)chunk * (8 * (kInCols + kOutCols) + 2); }

struct SlotView {
    double *in[kInCols];
    double *out[kOutCols];
    int8_t *flag;
    uint8_t *status;
    bool zero_copy = false;  // the pointers are mappings of the caller's pinned host buffers, not a device slot
};
inline SlotView view(const pvv_ctx *c, int i) {
    SlotView v;
    char *p = c->slot[i];
    for (int k = 0; k < kInCols; ++k) { v.in[k] = (double *)p; p += 8 * c->chunk; }
    for (int k = 0; k < kOutCols; ++k) { v.out[k] = (double *)p; p += 8 * c->chunk; }
    v.flag = (int8_t *)p; p += c->chunk;
    v.status = (uint8_t *)p;
    return v;
}

#define PVV_CK(call)                      \
    do {                                  \
        cudaError_t e_ = (call);          \
        if (e_ != cudaSuccess) return (int)e_; \
    } while (0)

// Error exit of the pipeline: copies and kernels already enqueued keep writing into the caller's arrays and the ctx
// slots, so every stream is drained before the code is handed back (secondary errors ignored, sticky error cleared).
inline int pipeline_fail(pvv_ctx *c, int code) {
    for (int i = 0; i < c->nstreams; ++i)
        if (c->stream[i]) cudaStreamSynchronize(c->stream[i]);
    cudaGetLastError();
    return code;
}
#define PVV_CKP(call)                                            \
    do {                                                         \
        cudaError_t e_ = (call);                                 \
        if (e_ != cudaSuccess) return pipeline_fail(c, (int)e_); \
    } while (0)

// out[i] = vals[j] where in[i] == keys[j], else 0; returns the number of unmatched cells.  The keys live in
// registers and the key loop is unrolled (NK is a template parameter): 1.2 ns per cell on one core — the first
// version, which read keys/vals through the lambda's by-reference captures, took 12 ns.  Threaded from 2 Mi
// cells on (spawning and joining the threads costs more than the pass below that).
template <class T, int NK>
int64_t match_range(const T *__restrict__ in, int8_t *__restrict__ out, int64_t lo, int64_t hi, const T *keys, const int8_t *vals) {
    T k[NK];
    int8_t v[NK];
    for (int j = 0; j < NK; ++j) {
        k[j] = keys[j];
        v[j] = vals[j];
    }
    int64_t miss = 0;
    for (int64_t i = lo; i < hi; ++i) {
        const T x = in[i];
        int8_t o = 0;
#pragma GCC unroll 8
        for (int j = 0; j < NK; ++j) o = (x == k[j]) ? v[j] : o;
        out[i] = o;
        miss += (o == 0);
    }
    return miss;
}

inline int host_threads(int64_t n, int threads) {
    if (threads <= 0) threads = (int)std::thread::hardware_concurrency();
    const int64_t by_size = n >> 20;  // one thread per Mi cells
    if (threads > by_size) threads = (int)by_size;
    if (threads > 16) threads = 16;
    return threads < 2 ? 1 : threads;
}

template <class F>
int64_t run_ranges(int64_t n, int threads, F range) {  // sum of range(lo, hi) over `threads` contiguous slices
    threads = host_threads(n, threads);
    if (threads == 1) return range(0, n);
    std::vector<int64_t> part((size_t)threads, 0);
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([&, t] { part[(size_t)t] = range(n * t / threads, n * (t + 1) / threads); });
    for (auto &th : pool) th.join();
    int64_t total = 0;
    for (int64_t m : part) total += m;
    return total;
}

template <class T>
int64_t match_cells(int64_t n, const T *in, int nkeys, const T *keys, const int8_t *vals, int8_t *out, int threads) {
    if (n <= 0) return 0;
    if (nkeys < 1 || nkeys > 8) return -1;
    return run_ranges(n, threads, [=](int64_t lo, int64_t hi) -> int64_t {
        switch (nkeys) {
            case 1: return match_range<T, 1>(in, out, lo, hi, keys, vals);
            case 2: return match_range<T, 2>(in, out, lo, hi, keys, vals);
            case 3: return match_range<T, 3>(in, out, lo, hi, keys, vals);
            case 4: return match_range<T, 4>(in, out, lo, hi, keys, vals);
            case 5: return match_range<T, 5>(in, out, lo, hi, keys, vals);
            case 6: return match_range<T, 6>(in, out, lo, hi, keys, vals);
            case 7: return match_range<T, 7>(in, out, lo, hi, keys, vals);
            default: return match_range<T, 8>(in, out, lo, hi, keys, vals);
        }
    });
}

// Device-accessible alias of a pinned (cudaHostAlloc / cudaHostRegister) host pointer, nullptr for pageable memory.
inline void *mapped_pointer(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return nullptr;
    }
    return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
}
inline long long zero_copy_max() {
    static const long long v = [] {
        const char *e = getenv("PVV_ZERO_COPY_MAX");
        return e ? atoll(e) : (3ll << 19);  // 1.5 Mi contracts: crossover measured with tools/probe_zc_threshold.py
    }();
    return v;
}
// K4 reads its inputs three times (the solve, then both staging halves of the greeks phase): from device memory the
// repeats hit the L2, over a zero-copy mapping each one is another PCIe read.  Zero-copy therefore only serves its
// latency-bound short calls; PVV_ZERO_COPY_MAX_IVG overrides (tools/probe_zc_threshold.py measures the crossover).
inline long long zero_copy_max_iv_greeks() {
    static const long long v = [] {
        const char *e = getenv("PVV_ZERO_COPY_MAX_IVG");
        return e ? atoll(e) : (1ll << 17);
    }();
    return v < zero_copy_max() ? v : zero_copy_max();
}

// Shared driver of the four host entry points.
//   in[k]/in_stride[k]: host double columns (k < n_in), flag/flag_stride: host int8 column
//   out[k]: host double output columns (k < n_out, NULL = not wanted), status: host uint8 (may be NULL)
//   launch(slotview, dev_in[], dev_flag, strides, n_chunk, stream): enqueues the kernel
template <class Launch>
int run_pipeline(pvv_ctx *c, long long n, const int8_t *flag, long long flag_stride, int n_in, const double *const *in,
                 const long long *in_stride, int n_out, double *const *out, uint8_t *status, int64_t *first_zde, Launch launch,
                 long long zc_max = -1) {
    PVV_CK(cudaSetDevice(c->device));
    const long long big = LLONG_MAX;
    PVV_CKP(cudaMemcpyAsync(c->first_zde, &big, 8, cudaMemcpyHostToDevice, c->stream[0]));
    // broadcast scalars: one upload
    double hs[8] = {0};
    for (int k = 0; k < n_in; ++k)
        if (in_stride[k] == 0 && in[k]) hs[k] = in[k][0];
    PVV_CKP(cudaMemcpyAsync(c->scalars, hs, sizeof(hs), cudaMemcpyHostToDevice, c->stream[0]));
    int8_t hf[8] = {0};
    if (flag_stride == 0) hf[0] = flag[0];
    PVV_CKP(cudaMemcpyAsync((char *)c->scalars + 64, hf, 8, cudaMemcpyHostToDevice, c->stream[0]));
    PVV_CKP(cudaStreamSynchronize(c->stream[0]));
    // Short calls on pinned buffers: ONE launch straight on the host memory (UVA zero-copy).  Reads and
    // writes stream over PCIe in both directions while the SMs compute, with no per-chunk copy latency:
    // measured 1.22 ms vs 1.44 ms for 1 M price+greeks contracts; the chunked DMA pipeline wins again
    // from a few million contracts on (tools/probe_zerocopy.py).
    if (n <= (zc_max >= 0 ? zc_max : zero_copy_max())) {
        bool ok = true;
        SlotView z{};
        z.zero_copy = true;
        const double *zin[kInCols] = {nullptr};
        for (int j = 0; j < n_in && ok; ++j) {
            if (!in[j]) continue;
            if (in_stride[j] == 0) zin[j] = c->scalars + j;
            else ok = (zin[j] = (const double *)mapped_pointer(in[j])) != nullptr;
        }
        const int8_t *zflag = (flag_stride == 0) ? (const int8_t *)((char *)c->scalars + 64) : (const int8_t *)mapped_pointer(flag);
        ok = ok && zflag;
        for (int j = 0; j < n_out && ok; ++j)
            if (out[j]) ok = (z.out[j] = (double *)mapped_pointer(out[j])) != nullptr;
        if (status && ok) ok = (z.status = (uint8_t *)mapped_pointer(status)) != nullptr;
        if (ok) {
            int rc = launch(z, zin, zflag, n, 0, c->stream[0]);
            if (rc != 0) return pipeline_fail(c, rc);
            long long fz0 = big;
            PVV_CKP(cudaMemcpyAsync(&fz0, c->first_zde, 8, cudaMemcpyDeviceToHost, c->stream[0]));
            PVV_CKP(cudaStreamSynchronize(c->stream[0]));
            if (first_zde) *first_zde = (fz0 == big) ? -1 : fz0;
            return 0;
        }
    }
    // effective chunk: the ctx's chunk for long columns; for short ones about a quarter of the call so
    // that H2D, kernel and D2H of neighbouring chunks still overlap (measured: tools/tune_e2e.py)
    long long chunk = c->chunk;
    if (n < 4 * chunk) {
        chunk = ((n + 3) / 4 + 1023) / 1024 * 1024;
        if (chunk < 65536) chunk = 65536;
        if (chunk > c->chunk) chunk = c->chunk;
    }
    long long nchunks = (n + chunk - 1) / chunk;
    for (long long k = 0; k < nchunks; ++k) {
        const int si = (int)(k % c->nstreams);
        cudaStream_t s = c->stream[si];
        const long long off = k * chunk;
        const long long m = (n - off < chunk) ? (n - off) : chunk;
        SlotView v = view(c, si);
        const double *dev_in[kInCols];
        for (int j = 0; j < n_in; ++j) {
            if (!in[j]) { dev_in[j] = nullptr; continue; }
            if (in_stride[j] == 0) dev_in[j] = c->scalars + j;
            else {
                PVV_CKP(cudaMemcpyAsync(v.in[j], in[j] + off, 8 * m, cudaMemcpyHostToDevice, s));
                dev_in[j] = v.in[j];
            }
        }
        const int8_t *dev_flag;
        if (flag_stride == 0) dev_flag = (const int8_t *)((char *)c->scalars + 64);
        else {
            PVV_CKP(cudaMemcpyAsync(v.flag, flag + off, m, cudaMemcpyHostToDevice, s));
            dev_flag = v.flag;
        }
        int rc = launch(v, dev_in, dev_flag, m, off, s);
        if (rc != 0) return pipeline_fail(c, rc);
        for (int j = 0; j < n_out; ++j)
            if (out[j]) PVV_CKP(cudaMemcpyAsync(out[j] + off, v.out[j], 8 * m, cudaMemcpyDeviceToHost, s));
        if (status) PVV_CKP(cudaMemcpyAsync(status + off, v.status, m, cudaMemcpyDeviceToHost, s));
    }
    for (int i = 0; i < c->nstreams; ++i) PVV_CKP(cudaStreamSynchronize(c->stream[i]));
    long long fz = big;
    PVV_CKP(cudaMemcpy(&fz, c->first_zde, 8, cudaMemcpyDeviceToHost));
    if (first_zde) *first_zde = (fz == big) ? -1 : fz;
    return 0;
}

}  // namespace

extern "C" {

int pvv_ctx_create(int device, int64_t chunk_contracts, int n_streams, pvv_ctx **out) {
    if (!out) return PVV_E_INVALID_ARGUMENT;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PVV_E_NO_DEVICE;
    if (device < 0 || device >= ndev) return PVV_E_INVALID_ARGUMENT;
    pvv_ctx *c = new (std::nothrow) pvv_ctx();
    if (!c) return PVV_E_OUT_OF_MEMORY;
    c->device = device;
    c->chunk = chunk_contracts > 0 ? chunk_contracts : (512ll << 10);
    c->nstreams = n_streams > 0 ? (n_streams > 8 ? 8 : n_streams) : 3;
    cudaError_t e = cudaSetDevice(device);
    for (int i = 0; i < c->nstreams && e == cudaSuccess; ++i) {
        e = cudaStreamCreateWithFlags(&c->stream[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMalloc(&c->slot[i], slot_bytes(c->chunk));
    }
    if (e == cudaSuccess) e = cudaMalloc(&c->scalars, 128);
    if (e == cudaSuccess) e = cudaMalloc(&c->first_zde, 8);
    if (e != cudaSuccess) {
        pvv_ctx_destroy(c);
        return (int)e;
    }
    *out = c;
    return 0;
}

void pvv_ctx_destroy(pvv_ctx *c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (int i = 0; i < c->nstreams; ++i) {
        if (c->stream[i]) cudaStreamSynchronize(c->stream[i]);
        if (c->slot[i]) cudaFree(c->slot[i]);
        if (c->stream[i]) cudaStreamDestroy(c->stream[i]);
    }
    if (c->scalars) cudaFree(c->scalars);
    if (c->first_zde) cudaFree(c->first_zde);
    delete c;
}

int pvv_price_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t,
                   const double *r, const double *q, const double *sigma, const int64_t stride[7], double *price, int64_t *first_zde) {
    if (!ctx || n < 0 || model < 0 || model > 2 || !price || (model == 2 && !q)) return PVV_E_INVALID_ARGUMENT;
    if (first_zde) *first_zde = -1;
    if (n == 0) return 0;
    const double *in[7] = {nullptr, S, K, t, (model == 0) ? nullptr : r, (model == 2) ? q : nullptr, sigma};
    const long long ist[7] = {0, stride[1], stride[2], stride[3], stride[4], stride[5], stride[6]};
    double *out[1] = {price};
    return run_pipeline(ctx, n, flag, stride[0], 7, in, ist, 1, out, nullptr, first_zde,
                        [&](const SlotView &v, const double *const *d, const int8_t *df, long long m, long long off, cudaStream_t s) {
                            return launch_price(model, m, df, d[1], d[2], d[3], d[4], d[5], d[6], stride, v.out[0],
                                                (int64_t *)ctx->first_zde, off, s);
                        });
}

int pvv_greeks_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t,
                    const double *r, const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta,
                    double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde) {
    if (!ctx || n < 0 || model < 0 || model > 2 || (model == 2 && !q)) return PVV_E_INVALID_ARGUMENT;
    if (first_zde) *first_zde = -1;
    if (n == 0) return 0;
    const double *in[7] = {nullptr, S, K, t, r, (model == 2) ? q : nullptr, sigma};
    const long long ist[7] = {0, stride[1], stride[2], stride[3], stride[4], stride[5], stride[6]};
    double *out[6] = {price, delta, gamma, theta, rho, vega};
    return run_pipeline(ctx, n, flag, stride[0], 7, in, ist, 6, out, nullptr, first_zde,
                        [&](const SlotView &v, const double *const *d, const int8_t *df, long long m, long long off, cudaStream_t s) {
                            return launch_greeks(model, m, df, d[1], d[2], d[3], d[4], d[5], d[6], stride, price ? v.out[0] : nullptr,
                                                 delta ? v.out[1] : nullptr, gamma ? v.out[2] : nullptr, theta ? v.out[3] : nullptr,
                                                 rho ? v.out[4] : nullptr, vega ? v.out[5] : nullptr, (int64_t *)ctx->first_zde, off, s);
                        });
}

int pvv_iv_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K,
                const double *t, const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv,
                uint8_t *status, int64_t *first_zde) {
    if (!ctx || n < 0 || model < 0 || model > 2 || !iv || (model == 2 && !q)) return PVV_E_INVALID_ARGUMENT;
    if (first_zde) *first_zde = -1;
    if (n == 0) return 0;
    const double *in[7] = {nullptr, price, S, K, t, r, (model == 2) ? q : nullptr};
    const long long ist[7] = {0, stride[1], stride[2], stride[3], stride[4], stride[5], stride[6]};
    double *out[1] = {iv};
    return run_pipeline(ctx, n, flag, stride[0], 7, in, ist, 1, out, status, first_zde,
                        [&](const SlotView &v, const double *const *d, const int8_t *df, long long m, long long off, cudaStream_t s) {
                            return launch_iv(model, m, df, d[1], d[2], d[3], d[4], d[5], d[6], stride, mask_errors, v.out[0], v.status,
                                             (int64_t *)ctx->first_zde, off, s, true);
                        });
}

int pvv_iv_greeks_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K,
                       const double *t, const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv,
                       uint8_t *status, double *delta, double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde) {
    if (!ctx || n < 0 || model < 0 || model > 2 || !iv || (model == 2 && !q)) return PVV_E_INVALID_ARGUMENT;
    if (first_zde) *first_zde = -1;
    if (n == 0) return 0;
    const double *in[7] = {nullptr, price, S, K, t, r, (model == 2) ? q : nullptr};
    const long long ist[7] = {0, stride[1], stride[2], stride[3], stride[4], stride[5], stride[6]};
    double *out[6] = {iv, delta, gamma, theta, rho, vega};
    return run_pipeline(ctx, n, flag, stride[0], 7, in, ist, 6, out, status, first_zde,
                        [&](const SlotView &v, const double *const *d, const int8_t *df, long long m, long long off, cudaStream_t s) {
                            return launch_iv_greeks(model, m, df, d[1], d[2], d[3], d[4], d[5], d[6], stride, mask_errors, v.out[0],
                                                    v.status, delta ? v.out[1] : nullptr, gamma ? v.out[2] : nullptr,
                                                    theta ? v.out[3] : nullptr, rho ? v.out[4] : nullptr, vega ? v.out[5] : nullptr,
                                                    (int64_t *)ctx->first_zde, off, s, v.zero_copy || ivg_fused_env(), true);
                        },
                        zero_copy_max_iv_greeks());
}

int pvv_analytic_greeks_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t,
                             const double *r, const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta,
                             double *gamma, double *theta, double *rho, double *vega) {
    if (!ctx || n < 0 || (model != 1 && model != 2) || (model == 2 && !q)) return PVV_E_INVALID_ARGUMENT;
    if (n == 0) return 0;
    const double *in[7] = {nullptr, S, K, t, r, (model == 2) ? q : nullptr, sigma};
    const long long ist[7] = {0, stride[1], stride[2], stride[3], stride[4], stride[5], stride[6]};
    double *out[6] = {price, delta, gamma, theta, rho, vega};
    int64_t unused = -1;
    return run_pipeline(ctx, n, flag, stride[0], 7, in, ist, 6, out, nullptr, &unused,
                        [&](const SlotView &v, const double *const *d, const int8_t *df, long long m, long long, cudaStream_t s) {
                            return pvv_analytic_greeks_dev(model, m, df, d[1], d[2], d[3], d[4], d[5], d[6], stride, price ? v.out[0] : nullptr,
                                                           delta ? v.out[1] : nullptr, gamma ? v.out[2] : nullptr, theta ? v.out[3] : nullptr,
                                                           rho ? v.out[4] : nullptr, vega ? v.out[5] : nullptr, s);
                        },
                        zero_copy_max_iv_greeks());  // K5 streams its inputs with TMA bulk copies from two CTAs per SM: over a zero-copy
                                                     // mapping that is too few reads in flight (0.66 vs 0.71 G contracts/s at 1 M)
}

// Host-side helper of the flag ingest (util/data_format.py): out[i] = vals[j] where in[i] == keys[j],
// 0 where nothing matches; returns the number of unmatched elements.  The Python side hands over the
// PyObject* values of an object column and the addresses of the handful of singleton objects a flag
// cell can be ('c', 'p', 1, -1, ...): one multi-threaded pass over 8 bytes per element replaces the
// reference's per-element dict lookup (util/data_format.py:10-11, 51 % of its wall time at 1 M rows).
int64_t pvv_match_u64(int64_t n, const uint64_t *in, int nkeys, const uint64_t *keys, const int8_t *vals, int8_t *out, int threads) {
    return match_cells(n, in, nkeys, keys, vals, out, threads);
}

// The same over 4-byte cells: numpy '<U1' flag arrays (one UCS4 code point per cell), int32 / float32 flags.
int64_t pvv_match_u32(int64_t n, const uint32_t *in, int nkeys, const uint32_t *keys, const int8_t *vals, int8_t *out, int threads) {
    return match_cells(n, in, nkeys, keys, vals, out, threads);
}

// 1 if offsets[i + 1] - offsets[i] == 1 for all i < n (every cell of an Arrow string column is one byte long: its
// data buffer then IS the flag column), else 0.  width = 4 (string) or 8 (large_string) bytes per offset.
int pvv_unit_offsets(int64_t n, const void *offsets, int width, int threads) {
    if (n <= 0) return 1;
    if (width != 4 && width != 8) return 0;
    return run_ranges(n, threads, [=](int64_t lo, int64_t hi) -> int64_t {
        int64_t bad = 0;
        if (width == 8) {
            const int64_t *o = (const int64_t *)offsets;
            for (int64_t i = lo; i < hi; ++i) bad += (o[i + 1] - o[i] != 1);
        } else {
            const int32_t *o = (const int32_t *)offsets;
            for (int64_t i = lo; i < hi; ++i) bad += (o[i + 1] - o[i] != 1);
        }
        return bad;
    }) == 0;
}

// out[i] = lut[in[i]] over bytes (flag columns that arrive as one byte per cell: Arrow string data of
// one-character cells); returns the number of zeros written (= unknown flags).  Threaded like pvv_match_u64.
int64_t pvv_lut_u8(int64_t n, const uint8_t *in, const int8_t *lut, int8_t *out, int threads) {
    if (n <= 0) return 0;
    return run_ranges(n, threads, [=](int64_t lo, int64_t hi) -> int64_t {
        int8_t tab[256];
        for (int j = 0; j < 256; ++j) tab[j] = lut[j];
        const uint8_t *__restrict__ src = in;
        int8_t *__restrict__ dst = out;
        int64_t miss = 0;
        for (int64_t i = lo; i < hi; ++i) {
            const int8_t o = tab[src[i]];
            dst[i] = o;
            miss += (o == 0);
        }
        return miss;
    });
}

}  // extern "C"
