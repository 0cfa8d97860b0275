/*
 * pvv.h — C ABI of the B200 engine behind the py_vollib_vectorized hot path (libpvv.so).
 *
 * The reference has no FFI; its seam is the set of module-level numba entry points that its public
 * Python functions import by name.  Each function below replaces those loops one for one
 * (file:line under /root/reference/py_vollib_vectorized):
 *
 *   pvv_price_*      _black_vectorized_call / _black_scholes_vectorized_call /
 *                    _black_scholes_merton_vectorized_call            _model_calls.py:81-102
 *   pvv_greeks_*     numerical_{delta,gamma,theta,rho,vega}_black_scholes[_merton]
 *                                                                      _numerical_greeks.py:87-253
 *   pvv_iv_*         implied_volatility_from_a_transformed_rational_guess _iv_models.py:18-44,
 *                    fused with the numpy pre/post-processing of implied_volatility.py:48-86 and
 *                    the masks of util/data_format.py:42-72
 *   pvv_iv_greeks_*  the price_col branch of price_dataframe            api.py:142-178
 *
 * Conventions
 *   - plain C: pointers + sizes, no torch / numpy types.  `_dev` entry points take DEVICE pointers
 *     and enqueue asynchronously on `stream` (a cudaStream_t passed as void*; NULL = default stream);
 *     `_host` entry points take HOST pointers (pinned or pageable) and run a chunked, multi-stream
 *     H2D -> kernel -> D2H pipeline inside the library, returning when the outputs are complete.
 *   - SoA columns: int8 flags (+1 call, -1 put), float64 everything else.  Every input column has an
 *     element stride in `stride[]`: 1 = dense array of n elements, 0 = one broadcast scalar.
 *     Column order of stride[]: price/greeks {flag,S,K,t,r,q,sigma}; iv {flag,price,S,K,t,r,q}.
 *   - model: 0 black (Black-76 on the forward passed as S; price is undiscounted, r unused — models.py:38),
 *            1 black_scholes, 2 black_scholes_merton (q required; otherwise q may be NULL).
 *     pvv_greeks / pvv_iv_greeks with model 0 use the Black-Scholes stencil, as api.py:217-225 does.
 *   - return value: 0 on success, > 0 a cudaError_t, < 0 a PVV_E_* code.  Nothing throws, the
 *     library keeps no global mutable state besides what lives in a pvv_ctx; one ctx per device,
 *     calls on one ctx must be serialised by the caller.
 *   - numba raises ZeroDivisionError for the WHOLE call when a contract divides by zero
 *     (K == 0, F == 0, t == 0 in the IV path, exp(-r t) == 0 in Black-Scholes pricing).  The kernels
 *     keep going and report the smallest offending contract index through `first_zde`
 *     (-1 if none; device pointer to int64 for _dev, which must be pre-set to INT64_MAX) and,
 *     for the IV entry points, bit 4 of the per-contract status byte.
 */
#ifndef PVV_H
#define PVV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PVV_MODEL_BLACK 0
#define PVV_MODEL_BLACK_SCHOLES 1
#define PVV_MODEL_BLACK_SCHOLES_MERTON 2

/* per-contract status byte of the IV entry points */
#define PVV_ST_BELOW_INTRINSIC 1 /* undiscounted price < intrinsic          data_format.py:55-62  */
#define PVV_ST_ABOVE_MAXIMUM 2   /* undiscounted price >= F (call) / K (put) data_format.py:63-70  */
#define PVV_ST_ZERO_DIVISION 4   /* numba would raise ZeroDivisionError                             */
#define PVV_ST_DEFLATER_ZERO 8   /* exp(-r t) == 0                          implied_volatility.py:50 */

#define PVV_E_INVALID_ARGUMENT (-1)
#define PVV_E_NO_DEVICE (-2)
#define PVV_E_OUT_OF_MEMORY (-3)

#if defined(__GNUC__)
#define PVV_API __attribute__((visibility("default")))
#else
#define PVV_API
#endif

typedef struct pvv_ctx pvv_ctx;

PVV_API int pvv_version(void);
PVV_API const char *pvv_error_string(int code);

/* ---- device-resident entry points (async on `stream`) ---------------------------------------- */
PVV_API int pvv_price_dev(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                  const double *q, const double *sigma, const int64_t stride[7], double *price, int64_t *first_zde, void *stream);

/* any of price/delta/gamma/theta/rho/vega may be NULL (that column is not written; stencil points
 * nobody needs are not evaluated) */
PVV_API int pvv_greeks_dev(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t, const double *r,
                   const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta, double *gamma,
                   double *theta, double *rho, double *vega, int64_t *first_zde, void *stream);

/* mask_errors: 1 for on_error "warn"/"raise" (rows below intrinsic / above maximum get NaN),
 * 0 for "ignore" (the solver's raw value is kept, implied_volatility.py:83-86 writes nothing) */
PVV_API int pvv_iv_dev(int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K, const double *t,
               const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv, uint8_t *status,
               int64_t *first_zde, void *stream);

/* The price_col branch of price_dataframe (api.py:142-178) in one call: the solve, then the finite-difference greeks at
 * the solved (and, with mask_errors, masked) volatility.  `iv` is required (it carries sigma from the first pass to
 * the second); status bit 4 and first_zde cover zero divisions of BOTH passes.  Device-resident and chunked calls are
 * two launches on `stream` (K3, then K2: measured faster than one fused kernel at every size);
 * short zero-copy host calls keep the single fused kernel. */
PVV_API int pvv_iv_greeks_dev(int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K, const double *t,
                      const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv, uint8_t *status,
                      double *delta, double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde, void *stream);

/* Opt-in closed-form greeks (py_vollib `greeks.analytical` conventions: theta per calendar day, vega
 * and rho per point).  model 1 or 2.  Any output may be NULL.  No reference
 * counterpart in py_vollib_vectorized (its greeks are finite differences): tolerance 1e-12 against the
 * closed-form oracle, 6 decimals against the reference's regression table. */
PVV_API int pvv_analytic_greeks_dev(int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t,
                                    const double *r, const double *q, const double *sigma, const int64_t stride[7], double *price,
                                    double *delta, double *gamma, double *theta, double *rho, double *vega, void *stream);
PVV_API int pvv_analytic_greeks_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *S, const double *K,
                                     const double *t, const double *r, const double *q, const double *sigma, const int64_t stride[7],
                                     double *price, double *delta, double *gamma, double *theta, double *rho, double *vega);
/* internal: bumps the launch counter from other translation units of the library */
PVV_API void pvv_note_launch(void);

/* ---- host-buffer entry points (chunked multi-stream pipeline inside the library) ------------- */
/* chunk_contracts <= 0 and n_streams <= 0 pick defaults (512 Ki contracts, 3 streams: the measured
 * optimum on PCIe Gen5, tools/tune_e2e.py). */
PVV_API int pvv_ctx_create(int device, int64_t chunk_contracts, int n_streams, pvv_ctx **out);
PVV_API void pvv_ctx_destroy(pvv_ctx *ctx);

PVV_API int pvv_price_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t,
                   const double *r, const double *q, const double *sigma, const int64_t stride[7], double *price, int64_t *first_zde);

PVV_API int pvv_greeks_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *S, const double *K, const double *t,
                    const double *r, const double *q, const double *sigma, const int64_t stride[7], double *price, double *delta,
                    double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde);

PVV_API int pvv_iv_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K,
                const double *t, const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv,
                uint8_t *status, int64_t *first_zde);

PVV_API int pvv_iv_greeks_host(pvv_ctx *ctx, int model, int64_t n, const int8_t *flag, const double *price, const double *S, const double *K,
                       const double *t, const double *r, const double *q, const int64_t stride[7], int mask_errors, double *iv,
                       uint8_t *status, double *delta, double *gamma, double *theta, double *rho, double *vega, int64_t *first_zde);

/* number of kernel launches issued by this library since load (bench.py's gpu_launches claim) */
PVV_API int64_t pvv_launch_count(void);

/* ---- host ingest helper (no GPU involved) ------------------------------------------------------- */
/* out[i] = vals[j] where in[i] == keys[j], else 0; returns the number of unmatched elements.  Used by
 * the Python host to encode an object-dtype flag column ('c'/'p' singletons) in one threaded pass;
 * replaces the per-element dict lookup of util/data_format.py:10-11. */
PVV_API int64_t pvv_match_u64(int64_t n, const uint64_t *in, int nkeys, const uint64_t *keys, const int8_t *vals, int8_t *out, int threads);
/* the same over 4-byte cells: numpy '<U1' flag arrays (one UCS4 code point per cell), int32 / float32 flags */
PVV_API int64_t pvv_match_u32(int64_t n, const uint32_t *in, int nkeys, const uint32_t *keys, const int8_t *vals, int8_t *out, int threads);

/* out[i] = lut[in[i]] (256-entry table) for flag columns that arrive as one byte per cell; returns the
 * number of zeros written. */
PVV_API int64_t pvv_lut_u8(int64_t n, const uint8_t *in, const int8_t *lut, int8_t *out, int threads);

/* 1 if every cell of an Arrow string column is exactly one byte long (offsets[i + 1] - offsets[i] == 1 for all
 * i < n; width = 4 or 8 bytes per offset): its data buffer then is the flag column for pvv_lut_u8. */
PVV_API int pvv_unit_offsets(int64_t n, const void *offsets, int width, int threads);

/* ---- diagnostics used by the parity tests ----------------------------------------------------- */
/* y[i] = f(x[i]) on the device; which: 0 exp, 1 log, 2 erfcx, 3 erfc, 4 norm_cdf, 5 inverse_norm_cdf,
 * 6 pow(x, 1/3), 7 sqrt_bounded(x), 8 div_bounded(x[i], x[i ^ 1]) (the guard-free fast paths of csrc/pvv_math.cuh; n even) */
PVV_API int pvv_unary_dev(int which, int64_t n, const double *x, double *y, void *stream);
/* measured FP64 issue rate: runs a dependent DADD/DMUL (fmad off) chain, returns elapsed ms via *ms
 * for `iters` x 2 instructions per thread over `threads` threads */
PVV_API int pvv_fp64_probe(int64_t threads, int iters, int use_fma, float *ms, void *stream);

/* the same with `chains` (1, 2, 4, 8) independent chains per thread at `warps_per_sm` resident warps (multiple of 4):
 * elapsed ms for iters x 16 DP instructions per thread over warps_per_sm x 32 threads on every SM */
PVV_API int pvv_fp64_ilp_probe(int chains, int warps_per_sm, int iters, float *ms, void *stream);

/* host logic, no device: contracts per K3 tile when a column of n contracts is cut into waves x slots equal tiles of at
 * most `capacity` (multiple of 32) contracts; slots = SMs x resident CTAs.  Returns the tile size (multiple of 32,
 * <= capacity) or PVV_E_INVALID_ARGUMENT. */
PVV_API int pvv_iv_tile_plan(int64_t n, int capacity, int slots);

#ifdef __cplusplus
}
#endif
#endif /* PVV_H */
