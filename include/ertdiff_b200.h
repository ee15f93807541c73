/*
 * ertdiff_b200.h -- C ABI of the B200-native ensemble posterior-sampling path.
 *
 * The reference (pnnl/ERT-Conditional-Diffusion-Model) is a Python script with no FFI of
 * its own; its call surface for this path is
 *     model(x, t, condition)                      ERT_Conditional_Diffusion.py:155-164
 *     sample_model(model, condition, T, ...)      ERT_Conditional_Diffusion.py:102-119
 *     model.state_dict() / load_state_dict()      ERT_Conditional_Diffusion.py:133-153, 371
 *     np.mean/std/var/percentile, gaussian_kde    ERT_Conditional_Diffusion.py:747-762, 867-872
 * Each entry point below names the reference lines it replaces.  INTEGRATION.md shows the
 * ctypes binding a maintainer of the reference would add.
 *
 * Conventions
 *   - every pointer named d_* is a DEVICE pointer on the model's (or the current) device;
 *     h_* is a HOST pointer; sizes are element counts unless said otherwise;
 *   - `stream` is a cudaStream_t passed as void* (NULL = the legacy default stream); all work
 *     is enqueued on it and the call returns without synchronising unless stated;
 *   - every function returns 0 on success or a negative ertdiff_status; the message of the
 *     last failure on the calling thread is ertdiff_last_error();
 *   - a model handle is bound to one device, is not thread-safe, distinct handles are;
 *   - the handle-less statistics / metric functions (ensemble_*, minmax, misfit_metrics,
 *     wasserstein_distance) share one scratch buffer per device that grows on demand: on one device
 *     issue them from one host thread and on one stream at a time (work already enqueued is never
 *     invalidated: a growing buffer synchronises the device before it is replaced);
 *   - there is no CPU fallback: without a CUDA device every compute call fails with
 *     ERTDIFF_ERR_CUDA.
 */
#ifndef ERTDIFF_B200_H
#define ERTDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ERTDIFF_ABI_VERSION 1

typedef enum ertdiff_status {
    ERTDIFF_OK = 0,
    ERTDIFF_ERR_ARG = -1,         /* bad argument (null pointer, size, unsupported shape)  */
    ERTDIFF_ERR_CUDA = -2,        /* a CUDA runtime call or kernel launch failed           */
    ERTDIFF_ERR_UNSUPPORTED = -3, /* shape outside what the kernels are built for          */
    ERTDIFF_ERR_STATE = -4        /* handle used before weights were loaded, etc.          */
} ertdiff_status;

typedef enum ertdiff_dtype { ERTDIFF_F32 = 0, ERTDIFF_F64 = 1 } ertdiff_dtype;

/* how the reverse loop is executed (north_star item 6) */
typedef enum ertdiff_loop_mode {
    ERTDIFF_LOOP_PERSISTENT = 0,  /* one launch; x lives in registers/smem for all steps   */
    ERTDIFF_LOOP_GRAPH = 1,       /* one step kernel per timestep, captured in a CUDA graph */
    ERTDIFF_LOOP_STREAM = 2       /* same step kernels launched one by one (for reference)  */
} ertdiff_loop_mode;

/* arithmetic of the denoiser contractions */
typedef enum ertdiff_precision {
    ERTDIFF_PREC_FP32 = 0,        /* fp32 FFMA everywhere (BASELINE config 2)               */
    ERTDIFF_PREC_BF16 = 1,        /* bf16 operands, fp32 accumulate on tcgen05 (configs 3/4); */
                                  /* hidden_dim = 128 or 256, param_dim <= 29                 */
    ERTDIFF_PREC_BF16X3 = 2       /* split precision on tcgen05: every operand as two bf16    */
                                  /* terms (hi + residual), three accumulating products per   */
                                  /* projection -- fp32-class fields at tensor-core speed;    */
                                  /* hidden_dim = 128, param_dim <= 29; the condition encoder */
                                  /* stays in fp32                                            */
} ertdiff_precision;

typedef struct ertdiff_model ertdiff_model;

/* ---- library ------------------------------------------------------------------------- */
int ertdiff_abi_version(void);
const char* ertdiff_last_error(void);
/* number of kernels this library has launched since load / since the last reset (bench.py's
 * gpu_launches claim is read from here) */
int64_t ertdiff_launch_count(void);
void ertdiff_launch_count_reset(void);

/* ---- model handle: ConditionalDiffusionModel.__init__ / state_dict (ECD.py:123-153) ---- */
/* param_dim <= 32, hidden_dim in {32,64,128,256,512}; in_channels is 14 as hard-coded at
 * ECD.py:134. */
int ertdiff_model_create(ertdiff_model** out, int device, int param_dim, int hidden_dim);
int ertdiff_model_destroy(ertdiff_model* m);
/* measurement aid: when enabled, the persistent chain kernel is bracketed by CUDA events on
 * the launching stream; ertdiff_model_last_chain_ms synchronises on the second event and
 * returns that kernel's duration in milliseconds. */
int ertdiff_model_profile(ertdiff_model* m, int enable);
int ertdiff_model_last_chain_ms(ertdiff_model* m, float* h_ms);
/* diagnostics of the tensor-core chain: *h_status != 0 if a tile's MMA never signalled completion
 * (that tile's output was filled with NaN).  Synchronises the device. */
int ertdiff_model_umma_status(ertdiff_model* m, int* h_status);
/* development aid for the tensor-core chain: reads (if h_out16 != NULL) the 16 int64 cycle sums
 * CTA 0 recorded during the last chain launch -- epilogue thread 0: [0] loop top incl. the noise-ring
 * rendezvous, [1] wait D, [2] epilogue 1, [4] wait E, [5] epilogue 2 + operand publish, [6]/[7] the proxy
 * fence / arrive inside epilogue 1; MMA warp: [8] wait X, [9] GEMM1 issue, [10] waits on H + GEMM2
 * issue; noise warp 0: [11] wait for a free slot, [12] generate (summed over ITS items); [15] steps --
 * then enables or
 * disables the recording for the following launches.  Synchronises the device. */
int ertdiff_debug_umma_timing(ertdiff_model* m, int enable, int64_t* h_out16);
/* The 12 tensors in the reference's state_dict order:
 *  0 condition_encoder.0.weight (32,14,3)   1 condition_encoder.0.bias (32)
 *  2 condition_encoder.2.weight (64,32,3)   3 condition_encoder.2.bias (64)
 *  4 condition_encoder.6.weight (H,64)      5 condition_encoder.6.bias (H)
 *  6 time_embed.0.weight (H,H)              7 time_embed.0.bias (H)
 *  8 mlp.0.weight (H,P+2H)                  9 mlp.0.bias (H)
 * 10 mlp.2.weight (P,H)                    11 mlp.2.bias (P)
 * Contiguous fp32.  `on_device` says whether the 12 pointers are device or host pointers.
 * The data is copied (the caller may free its tensors) and re-packed for the kernels.
 * `h_freq` (H/2 floats, host) is the frequency table exp(-i*ln(1e4)/(H/2-1)) of
 * ECD.py:82-83, evaluated by the caller exactly as the reference does (torch fp32 exp), so
 * that sin/cos arguments are bit-identical.  Synchronises `stream`. */
int ertdiff_model_load(ertdiff_model* m, const float* const* tensors12, int on_device,
                       const float* h_freq, void* stream);
/* copy the 12 tensors back out (state_dict()); same order/layout; device or host dst. */
int ertdiff_model_export(ertdiff_model* m, float* const* tensors12, int on_device, void* stream);

/* ---- denoiser forward as written: model(x, t, condition) (ECD.py:155-164) -------------- */
/* x (B,P) f32, t (B) int64 (per-row values), condition (B,14,L) f32 NCL with member stride
 * `cond_member_stride` elements (0 = one condition shared by all rows); out (B,P) f32. */
int ertdiff_forward(ertdiff_model* m, const float* d_x, const int64_t* d_t,
                    const float* d_condition, int64_t B, int64_t L, int64_t cond_member_stride,
                    float* d_out, void* stream);

/* ---- condition encoder, hoisted out of the loop (ECD.py:133-142, 161) ------------------- */
/* condition (n_cond,14,L) -> d_cond_emb (n_cond,H) [= condition_encoder(condition), may be
 * NULL] and d_cond_bias (n_cond,H) [= mlp.0.weight[:,P+H:] @ cond_emb + mlp.0.bias, may be
 * NULL]: the per-member constant of the first MLP layer. */
int ertdiff_encode_condition(ertdiff_model* m, const float* d_condition, int64_t n_cond,
                             int64_t L, int64_t cond_member_stride, float* d_cond_emb,
                             float* d_cond_bias, void* stream);
/* the same with the arithmetic selectable: ERTDIFF_PREC_FP32 = the call above (CUDA-core FFMA);
 * ERTDIFF_PREC_BF16 = both convolutions as implicit GEMMs on the tensor cores (tcgen05, bf16
 * operands, fp32 accumulation, input fed by TMA), which is also what ertdiff_sample_model uses when
 * args->precision is ERTDIFF_PREC_BF16. */
int ertdiff_encode_condition_prec(ertdiff_model* m, const float* d_condition, int64_t n_cond,
                                  int64_t L, int64_t cond_member_stride, float* d_cond_emb,
                                  float* d_cond_bias, int32_t precision, void* stream);

/* ---- reverse chain: sample_model (ECD.py:102-119) --------------------------------------- */
typedef struct ertdiff_chain_args {
    int64_t B;               /* ensemble members                                            */
    int64_t n_cond;          /* rows of d_cond_bias; member i uses row (i % n_cond)         */
    int32_t T;               /* length of the schedule arrays                               */
    int32_t num_steps;       /* timesteps run: t = num_steps-1 .. 0 (truncation, ECD.py:108) */
    double temperature;      /* ECD.py:118 (a python float: sigma = sqrt(beta)*temperature   */
                             /* is formed in double before rounding to fp32)                */
    const float* d_betas;    /* (T) caller's schedule tensors (ECD.py:90-94); consumed as    */
    const float* d_alphas;   /* given, never recomputed                                     */
    const float* d_alpha_bar;
    const float* d_cond_bias;/* (n_cond,H) from ertdiff_encode_condition                    */
    const float* d_x_T;      /* (B,P) initial noise, or NULL = draw it (Philox)             */
    const float* d_noise;    /* (num_steps-1,B,P) injected noise in draw order (row k is    */
                             /* used at t = num_steps-1-k), or NULL = draw it (Philox)      */
    uint64_t seed;           /* Philox key / offset when drawing on the device              */
    uint64_t offset;
    int64_t member_offset;   /* global index of member 0 (multi-GPU shards keep RNG streams  */
                             /* and noise rows aligned with the unsharded ensemble)         */
    int64_t noise_member_stride_B; /* row length (in members) of d_noise / d_x_T, >= B; lets */
                             /* a shard read its slice of the full (.., B_total, P) tensor   */
    int32_t loop_mode;       /* ertdiff_loop_mode                                           */
    int32_t precision;       /* ertdiff_precision                                           */
    float* d_x_out;          /* (B,P) x_0                                                   */
    float* d_eps_trace;      /* optional (num_steps,B,P): predicted noise per step, or NULL */
} ertdiff_chain_args;

int ertdiff_sample_chain(ertdiff_model* m, const ertdiff_chain_args* args, void* stream);

/* encode + chain in one call: what sample_model(model, condition, ...) does.  condition is
 * (n_cond,14,L); d_cond_bias in `args` is ignored (computed internally). */
int ertdiff_sample_model(ertdiff_model* m, const float* d_condition, int64_t L,
                         int64_t cond_member_stride, const ertdiff_chain_args* args,
                         void* stream);

/* per-step scalars of ECD.py:111-118 with the reference's rounding, as a (num_steps,4) table
 * [coef, 1/sqrt(alpha), sqrt(beta)*temperature, 0] (4 floats per row) indexed by t (exposed for tests) */
int ertdiff_step_coefficients(const float* d_betas, const float* d_alphas,
                              const float* d_alpha_bar, int32_t num_steps, double temperature,
                              float* d_table, void* stream);

/* the device noise source that replaces torch.randn (ECD.py:107,116) when no noise is injected:
 * Philox4x32-10 + Box-Muller with the chain's stream layout.  Fills d_out (draws,B,P): row 0 is
 * what the chain uses as x_T, row k the k-th in-loop draw (exposed so tests can feed the very
 * same draws to the oracle). */
int ertdiff_philox_normal(uint64_t seed, uint64_t offset, int64_t member_offset, int64_t B,
                          int32_t P, int32_t draws, float* d_out, void* stream);

/* ---- posterior update alone (ECD.py:111-118), bit-exact, n = B*P elements --------------- */
/* out = c1*(x - coef*eps) [+ sigma*z if d_z != NULL]; each op rounded to fp32, no FMA.       */
int ertdiff_posterior_update(const float* d_x, const float* d_eps, const float* d_z,
                             float coef, float c1, float sigma, int64_t n, float* d_out,
                             void* stream);

/* ---- ensemble statistics over members (axis 0) of a (N,Q) array ------------------------- */
/* np.mean / np.std / np.var (ddof=0), ECD.py:867-869.  Outputs (Q) in the array's dtype; any
 * may be NULL.  Bit-exact with numpy for Q > 1 (left-to-right accumulation per column). */
int ertdiff_ensemble_moments(const void* d_a, int dtype, int64_t N, int64_t Q, void* d_mean,
                             void* d_std, void* d_var, void* stream);
/* np.percentile(a, q, axis=0), method 'linear', ECD.py:870-872, 612, 1126-1127, 1199-1200.
 * h_q: nq percentiles in [0,100] (host).  index_dtype: the dtype numpy does the index
 * arithmetic in (F32 when a is f32 and q was a python scalar, else F64).  Output (nq,Q) in
 * result_type(dtype, index_dtype).  Bit-exact with numpy. */
int ertdiff_ensemble_percentiles(const void* d_a, int dtype, int64_t N, int64_t Q,
                                 const double* h_q, int32_t nq, int index_dtype, void* d_out,
                                 void* stream);
/* UQ calibration, the coverage step (ECD.py:1121-1132, 1195-1206): d_low / d_upp are
 * (n_intervals, Q) float64 lower / upper bounds (rows of ertdiff_ensemble_percentiles), d_truth (Q)
 * float64, column c belongs to parameter c % P.  d_counts (n_intervals, P + 1) int32:
 * [k][0] = #{c : low < truth <= upp}, [k][1 + j] = the same over the columns of parameter j. */
int ertdiff_interval_coverage(const double* d_low, const double* d_upp, const double* d_truth,
                              int32_t n_intervals, int64_t Q, int32_t P, int32_t* d_counts, void* stream);

/* Per-member data misfit of N simulated maps d_sims (N, L, C) against the observed map d_obs (L, C),
 * both `dtype` (F32 / F64), C fastest:
 *   d_wsse (N, C)      np.average((pred-obs)**2 / (A*|obs|+B)**2) per survey column, ECD.py:764-783
 *   d_wsse_total (N)   WSSE_sim.sum(axis=1), ECD.py:785
 *   d_mse (N)          mean_squared_error(obs.flatten(), sim.flatten()), ECD.py:927-930, 939-940
 * Outputs are `dtype`; each follows numpy's pairwise-summation tree and is bit-identical to numpy's.
 * Either the two WSSE outputs or d_mse may be NULL.  C <= 128. */
int ertdiff_misfit_metrics(const void* d_sims, const void* d_obs, int dtype, int64_t N, int64_t L, int64_t C,
                           double A, double B, void* d_wsse, void* d_wsse_total, void* d_mse, void* stream);

/* scipy.stats.wasserstein_distance(u_row.flatten(), v.flatten()) for each of the N rows of d_u (N, n)
 * against d_v (m), ECD.py:860, 898-899 (simulated / mean / mode map against the observed map).  Inputs
 * `dtype`, arithmetic in float64 as scipy's; d_out (N) float64.  scipy's final dot product runs through
 * BLAS, so parity is to a relative 1e-12, not bitwise. */
int ertdiff_wasserstein_distance(const void* d_u, const void* d_v, int dtype, int64_t N, int64_t n, int64_t m,
                                 double* d_out, void* stream);

/* global min and max of n elements (the KDE grid's end points, ECD.py:749-750) -> d_out[2]
 * as float64. */
int ertdiff_minmax(const void* d_a, int dtype, int64_t n, double* d_out2, void* stream);
/* Gaussian-KDE mode, ECD.py:751-762: grid = linspace(lo, hi, n_grid) read from d_lohi[2];
 * per column argmax of the Scott-bandwidth KDE.  d_mode (Q) float64 grid value,
 * d_index (Q) int64 grid index (either may be NULL). */
int ertdiff_ensemble_kde_mode(const void* d_a, int dtype, int64_t N, int64_t Q,
                              const double* d_lohi, int32_t n_grid, double* d_mode,
                              int64_t* d_index, void* stream);
/* the same with the grid range taken from the data: d_lohi (2 float64, output) = global min / max of
 * the whole (N, Q) array (ECD.py:749-750).  Small ensembles (N < 1024, N*Q <= 65,536 values) run as ONE fused
 * launch (range + column constants + scan + selection); larger ones as ertdiff_minmax followed by
 * ertdiff_ensemble_kde_mode. */
int ertdiff_ensemble_kde_mode_auto(const void* d_a, int dtype, int64_t N, int64_t Q, int32_t n_grid,
                                   double* d_lohi, double* d_mode, int64_t* d_index, void* stream);

/* ---- after the chain: un-transform + bounds, ECD.py:42-53, 402-406, 183-218
 * s    = a + (b-a)*sigmoid(u)                       fp32   (inverse_transform, ECD.py:48-50)
 * phys = fp32( fp32( (double)s - scaler_min[p] ) / scaler_scale[p] )
 *                                                   sklearn MinMaxScaler.inverse_transform on an
 *                                                   fp32 array with float64 min_/scale_ (ECD.py:405)
 * valid[i] = all_p( lim_lo[p] <= phys[i,p] <= lim_hi[p] )       check_param_bounds, ECD.py:183-218
 * u (B,P) f32; d_scaler_min / d_scaler_scale / d_lim_lo / d_lim_hi (P) f64 (scaler and limit
 * pointers may be NULL to skip that stage); outputs d_phys (B,P) f32, d_valid (B) uint8 and
 * d_first_bad (B) int32 = index of the first out-of-bounds parameter or -1; any may be NULL. */
int ertdiff_untransform_bounds(const float* d_u, int64_t B, int32_t P, float a, float b,
                               const double* d_scaler_min, const double* d_scaler_scale,
                               const double* d_lim_lo, const double* d_lim_hi, float* d_phys,
                               uint8_t* d_valid, int32_t* d_first_bad, void* stream);

/* Everything the path reports about an ensemble (ECD.py:747-762, 867-872), for columns [col0, col0 + ncols) of the
 * row-major (N, Q) array d_a, in one call: mean / std / var, the nq percentiles h_q (float64 index arithmetic), and the
 * KDE mode on the common grid of n_grid points spanning the min / max of the WHOLE array -- packed as float64 records
 * d_out[c * ld + r], r = 0 mean, 1 std, 2 var, 3..2+nq percentiles, 3+nq mode, 4+nq mode grid index (integral), one record
 * of ld >= 5 + nq doubles per column.  d_lohi_out (optional, 2 doubles) receives the grid range.  Same kernels and bits as
 * the separate entry points; this is what each rank of the column-sharded statistics calls once per step. */
int ertdiff_ensemble_summary(const void* d_a, int dtype, int64_t N, int64_t Q, int64_t col0, int64_t ncols,
                             const double* h_q, int32_t nq, int32_t n_grid, double* d_lohi_out, double* d_out, int64_t ld,
                             void* stream);

/* Pack n_rows (<= 64) device row vectors of ncols values each -- dtype per row: ERTDIFF_F32, ERTDIFF_F64 or 2 =
 * int64 -- into one float64 block d_out[c * ld + r] (one contiguous record of ld >= n_rows doubles per column).
 * h_rows / h_dtypes are HOST arrays.  Used by the column-sharded statistics: one launch, then one all-gather. */
int ertdiff_pack_rows_f64(const void* const* h_rows, const int32_t* h_dtypes, int32_t n_rows, int64_t ncols,
                          int64_t ld, double* d_out, void* stream);

/* check_param_bounds alone (ECD.py:183-218) for values (B, P) of dtype ERTDIFF_F32 / F64, P <= 32: d_valid[row] = 1
 * unless some parameter is `< lim_lo or > lim_hi` (compared in float64, as numpy promotes; NaN never drops a row, as
 * in the reference); d_first_bad[row] = the first offending parameter (the one the reference prints) or -1.
 * Either output may be NULL. */
int ertdiff_check_bounds(const void* d_v, int dtype, int64_t B, int32_t P, const double* d_lim_lo,
                         const double* d_lim_hi, uint8_t* d_valid, int32_t* d_first_bad, void* stream);

/* Stable ascending argsort of a device vector (n values, dtype ERTDIFF_F32 / F64), NaN last: the ranking
 * `np.argsort(WSSE_sim_total)` of ECD.py:786.  d_order[k] = index of the k-th smallest value. */
int ertdiff_argsort_stable(const void* d_v, int dtype, int64_t n, int64_t* d_order, void* stream);

/* ---- all-gather over NVLink peer memory (one process per GPU on one node) ----------------------------
 * The path's collectives -- the final fields, (B/G, 29) fp32 per rank, and the packed statistics records -- are
 * latency-bound.  A peer group replaces the NCCL call by ONE kernel: it stores this rank's slice straight into every
 * peer's buffer (P2P stores over NVLink), publishes an epoch flag to every peer and waits for theirs.
 *   create : allocates this rank's buffer (room for `bytes` = all ranks' slices, double-buffered) and returns its
 *            64-byte CUDA IPC handle in h_ipc_handle64; the caller exchanges the handles (e.g. one
 *            torch.distributed all-gather at set-up) and passes all `world` of them, in rank order, to connect.
 *   all_gather : d_src (nbytes <= slot_bytes) lands at offset rank * slot_bytes of every rank's buffer;
 *            *d_gathered = this rank's buffer for this call, valid until the next-but-one call.  Every rank must call
 *            in the same order.  A peer that never arrives sets the status word after ~2 s instead of hanging.
 * world <= 16.  Not thread-safe; one stream at a time. */
typedef struct ertdiff_peer ertdiff_peer;
int ertdiff_peer_create(ertdiff_peer** out, int device, int rank, int world, size_t bytes, void* h_ipc_handle64);
int ertdiff_peer_connect(ertdiff_peer* p, const void* h_all_handles);
int ertdiff_peer_all_gather(ertdiff_peer* p, const void* d_src, size_t nbytes, size_t slot_bytes, void** d_gathered,
                            void* stream);
int ertdiff_peer_status(ertdiff_peer* p, int* h_status);
int ertdiff_peer_destroy(ertdiff_peer* p);

/* ---- self-test of the tcgen05 building blocks -------------------------------------------------
 * D (128,N) = A (128,K) @ B (N,K)^T on the tensor cores (bf16 operands, fp32 accumulate in TMEM),
 * one CTA; (N,K) in {(128,32), (32,128), (128,128), (64,96), (256,32), (32,256)}.  Row-major fp32 in and out.
 * Synchronises the stream; fails with ERTDIFF_ERR_CUDA if the MMA never signals completion. */
int ertdiff_debug_umma_gemm(const float* d_A, const float* d_B, int32_t N, int32_t K, float* d_D,
                            void* stream);

/* ---- measurement aids -----------------------------------------------------------------------------
 * Latency floor of the fp32 chain kernel (SURVEY.md §8 d): while enabled, fp32 persistent chains of
 * this handle launch the SAME kernel with the two matrix-vector products removed -- every barrier,
 * shared-memory hand-off, shuffle, staging copy, RNG refill and the posterior update stay -- so that
 * its duration (ertdiff_model_last_chain_ms) is the cost of num_steps dependent steps of this
 * structure.  The fields it writes are meaningless.  hidden_dim 128 / 256, device RNG, no trace. */
int ertdiff_debug_chain_floor(ertdiff_model* m, int enable);
/* Graph-mode bookkeeping: h_out2[0] = graphs instantiated, h_out2[1] = in-place updates of the
 * instantiated graph (cudaGraphExecUpdate) since the handle was created. */
int ertdiff_debug_graph_stats(ertdiff_model* m, int64_t* h_out2);

#ifdef __cplusplus
}
#endif
#endif /* ERTDIFF_B200_H */
