// pvv_tile.cuh — CTA-wide regrouping of divergent work through shared memory.
//
// Why: one thread per contract leaves a warp with a mix of the four normalised-Black regions (on a
// random chain 52 % small-t, 47.5 % erfcx): ncu on the straightforward kernels showed 15.4 of 32
// lanes active per instruction and "no instruction" (instruction-cache) as the first stall reason
// (profiles/r01_v1_ncu_full_summary.txt).  Here a CTA stages a tile of work items in shared memory,
// counting-sorts them by region (a handful of compares per item, warp-aggregated shared atomics),
// and then walks the items in sorted order: every warp executes ONE region's code, and all warps of
// the CTA run the same code at about the same time.  Results go back to the items' home slots, so
// the arithmetic per item — and therefore every bit of the output — is unchanged.
#pragma once

#include "pvv_math.cuh"

namespace pvv {

// order[] entries: low 12 bits item id (tile size <= 4096), high 4 bits the sort key
constexpr int kSortKeys = 16;
template <int NT, int IPT>
struct TileSort {
    static constexpr int T = NT * IPT;
    static_assert(T <= 4096, "item id must fit 12 bits");
    unsigned short order[T];
    int cnt[kSortKeys];
};

// Counting sort of the tile's T items by key_of(item) in [0, 16).  All NT threads must call it.
// Contract: ts.cnt[] is zero on entry (tile_sort_init once per kernel, then every sort re-zeroes it
// on exit), and at least one __syncthreads() separates two sorts' counting phases — the callers'
// own end-of-stage barrier provides it.  key_of(item) is called for the items j*NT + tid: if those
// were written by OTHER threads the caller must put a barrier before the sort.  Two barriers per sort.
template <int NT, int IPT>
__device__ __forceinline__ void tile_sort_init(TileSort<NT, IPT> &ts) {
    if (threadIdx.x < kSortKeys) ts.cnt[threadIdx.x] = 0;
    __syncthreads();
}

template <int NT, int IPT, class KeyOf>
__device__ __forceinline__ void tile_sort(TileSort<NT, IPT> &ts, KeyOf key_of) {
    const int tid = threadIdx.x, lane = tid & 31;
    int key[IPT], pos[IPT];
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        key[j] = key_of(j * NT + tid);
        const unsigned peers = __match_any_sync(0xffffffffu, key[j]);
        const int leader = __ffs(peers) - 1;
        int base = 0;
        if (lane == leader) base = atomicAdd(&ts.cnt[key[j]], __popc(peers));
        base = __shfl_sync(0xffffffffu, base, leader);
        pos[j] = base + __popc(peers & ((1u << lane) - 1u));
    }
    __syncthreads();
    // exclusive prefix of the 16 counters, redundantly in every warp: lane k holds the start of key k
    const int c = (lane < kSortKeys) ? ts.cnt[lane] : 0;
    int incl = c;
#pragma unroll
    for (int d = 1; d < kSortKeys; d <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += v;
    }
    const int excl = incl - c;
#pragma unroll
    for (int j = 0; j < IPT; ++j) {
        const int start = __shfl_sync(0xffffffffu, excl, key[j]);
        ts.order[start + pos[j]] = (unsigned short)((j * NT + tid) | (key[j] << 12));
    }
    __syncthreads();
    if (tid < kSortKeys) ts.cnt[tid] = 0;  // every thread has read cnt[] before the barrier above
}

// One work item of the pricing tiles: everything needed to turn b(x, s) into a discounted price.
//   price = (itm ? intrinsic + max(0, scale*b) : max(intrinsic, scale*b)) * deflater
// `intr` carries the in-the-money flag in its sign (itm => intrinsic > 0 strictly).
template <int T>
struct PriceItems {
    double x[T];      // in: x (<= 0 after the put-call flip); out: the finished price
    double s[T];      // sigma * sqrt(t)
    double scale[T];  // sqrt(F) * sqrt(K)
    double intr[T];   // +intrinsic (out of the money) or -intrinsic (in the money)
    double defl[T];   // exp(-r t); unused for model 0
};

// _model_calls.py:70-78 up to the call of normalised_black: intrinsic, put-call parity flip,
// x = ln(F/K) oriented for the out-of-the-money twin, sqrt(F) sqrt(K).
template <bool BOUNDED, int T>
__device__ __forceinline__ void stage_price_item(PriceItems<T> &it, int item, double flag, double F, double K, double sqrtK, double s,
                                                 double deflater, unsigned &st) {
    const double payoff = flag < 0 ? K - F : F - K;
    const double intrinsic = abs_max0(payoff);
    const bool itm = flag * (F - K) > 0;
    const double qq = itm ? -flag : flag;
    if (K == 0.0) st |= ST_ZERO_DIVISION;
    const double x = log_g(div_sel<BOUNDED>(F, K));
    it.x[item] = qq < 0 ? -x : x;
    it.s[item] = s;
    it.scale[item] = sqrt_sel<BOUNDED>(F) * sqrtK;
    it.intr[item] = itm ? -intrinsic : intrinsic;
    it.defl[item] = deflater;
}

template <int T>
__device__ __forceinline__ void stage_null_item(PriceItems<T> &it, int item) {
    it.x[item] = 0.0;
    it.s[item] = 0.0;  // classifies as R_ZERO: costs nothing
    it.scale[item] = 0.0;
    it.intr[item] = 0.0;
    it.defl[item] = 0.0;
}

// Evaluate every item of the tile in region-sorted order and finish it to a price (in place, x[]).
// `on_zde(item)` is called for every item whose evaluation divided by zero (numba would raise).
template <int NT, int IPT, bool DEFLATE, class OnZde>
__device__ __forceinline__ void eval_price_items(PriceItems<NT * IPT> &it, TileSort<NT, IPT> &ts, OnZde on_zde) {
    tile_sort(ts, [&](int item) { return nbc_sort_key(it.x[item], it.s[item]); });
    const int tid = threadIdx.x;
#pragma unroll 1
    for (int j = 0; j < IPT; ++j) {
        const unsigned e = ts.order[j * NT + tid];
        const int item = e & 0xfffu, region = key_region(e >> 12);
        const double xi = it.x[item], si = it.s[item];
        const bool zde = nbc_zde(region, xi, si);  // before the call: one predicate, not two doubles, lives across it
        const double nb = nbc_eval(region, xi, si);
        const double tv = it.scale[item] * nb;
        const double is = it.intr[item];
        double v = (is < 0) ? (-is) + np_maximum(0.0, tv) : np_maximum(is, tv);
        if (DEFLATE) v = v * it.defl[item];
        it.x[item] = v;
        if (zde) on_zde(item);
    }
    __syncthreads();
}

}  // namespace pvv
