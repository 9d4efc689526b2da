/*
 * hyperbo_b200 -- C ABI of the B200-native GP pre-training / inference engine.
 *
 * This is the drop-in boundary.  google-research/hyperbo has no
 * FFI of its own: its boundary is the function-level Python API, so every entry
 * point below names the reference function (file:line under
 * /root/reference/hyperbo/) whose arithmetic it replaces.  INTEGRATION.md shows
 * the ctypes / pybind11 stub a maintainer of the reference would add.
 *
 * Conventions
 *  - plain C: pointers + sizes, no torch / C++ types.
 *  - every `const void*` / `void*` data pointer is a DEVICE pointer owned by the
 *    caller, of the handle's scalar type (double for HB_F64, float for HB_F32);
 *    `offs` arrays are HOST pointers.
 *  - `stream` is the caller's cudaStream_t (passed as void*); all work is
 *    enqueued asynchronously on it, nothing here synchronises the device.
 *    (The first call with a new task shape set allocates workspace with
 *    cudaMalloc; later calls with the same shapes do not, so the call sequence
 *    is CUDA-graph capturable after one warm-up call.)
 *  - return value: HB_OK or an HB_ERR_* code for API misuse only.  Numerical
 *    breakdown (non-PD matrix) never raises: info[t] = failing column + 1 and
 *    NaN propagates into nll / outputs, exactly like the reference, whose
 *    Python loop then stops on the non-finite loss (gp_utils/gp.py:135-142).
 *  - parameter vector layout (P = 3 + d scalars, handle dtype):
 *        raw[0] = constant        (mean.constant,            mean.py:60-64)
 *        raw[1] = signal_variance (kernel.py:78-81)
 *        raw[2] = noise_variance  (linalg.py:64-68)
 *        raw[3+k] = lengthscale[k], k < d   (ARD; kernel.py:80)
 *    `warp_mask` bit p set  <=>  theta_p = softplus(raw_p) + 1e-10
 *    (utils.DEFAULT_WARP_FUNC, gp_utils/utils.py:73-81); clear <=> identity
 *    (warp_func=None, params_utils.py:97-111).
 *  - ragged task batches: tasks are concatenated row-wise; task t owns rows
 *    offs[t] .. offs[t+1]-1 of X (row-major, d columns) and y.  n_t = 0 is
 *    allowed and skipped (objectives.py:184).
 */
#ifndef HYPERBO_B200_H_
#define HYPERBO_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct hb_handle_s* hb_handle_t;

enum hb_status {
  HB_OK = 0,
  HB_ERR_BAD_ARG = 1,      /* null pointer, bad id, negative size */
  HB_ERR_UNSUPPORTED = 2,  /* d > HB_MAX_DIM, unsupported dtype/kernel combo */
  HB_ERR_CUDA = 3,         /* a CUDA runtime call failed (see hb_last_error) */
  HB_ERR_NO_DEVICE = 4
};

/* HB_F64: fp64 DMMA tile products (1e-10 parity with the fp64 oracle).
 * HB_F32: fp32 storage, 3xTF32 tensor-core tile products (the reference's JAX
 * default precision). */
enum hb_dtype { HB_F64 = 0, HB_F32 = 1 };

/* gp_utils/kernel.py:63-123 */
enum hb_kernel { HB_KERNEL_SE = 0, HB_KERNEL_MATERN32 = 1, HB_KERNEL_MATERN52 = 2 };
/* gp_utils/mean.py:54-64 */
enum hb_mean { HB_MEAN_ZERO = 0, HB_MEAN_CONSTANT = 1 };
/* bo_utils/acfun.py:96-142 */
enum hb_acq { HB_ACQ_NONE = 0, HB_ACQ_EI = 1, HB_ACQ_PI = 2, HB_ACQ_UCB = 3 };

#define HB_MAX_DIM 32
#define HB_TILE 64 /* edge of the packed lower-triangular tiles */

/* ---- lifetime ----------------------------------------------------------- */
/* One handle per device per host thread; owns the workspace arena.          */
int hb_create(hb_handle_t* out, int device, int dtype);
int hb_destroy(hb_handle_t h);
const char* hb_last_error(hb_handle_t h);
const char* hb_version(void);
/* Number of kernel launches this handle has enqueued (bench `gpu_launches`). */
int64_t hb_launch_count(hb_handle_t h);
/* Bytes of device workspace currently held. */
int64_t hb_workspace_bytes(hb_handle_t h);
/* Generation counter of the handle's device state: bumped whenever a workspace
 * buffer is re-allocated or a cached plan is evicted.  A caller that captured
 * engine calls into a CUDA graph must re-capture when it changes (the graph
 * holds raw workspace pointers). */
int64_t hb_generation(hb_handle_t h);
/* Tuning / test knobs (defaults come from the HB_* environment variables at
 * hb_create): "fused" 0 = launch-per-column path, 1 = automatic (persistent
 * kernel except for few tasks per GPU), 2 = always persistent; "fastpath",
 * "groups", "skew", "grid" of the persistent kernel's scheduler. */
int hb_set_option(hb_handle_t h, const char* name, double value);
/* Debug / tests: the persistent kernel's work-item order for a batch shape:
 * out receives (task, kind, a, b) per item (kind 0 DIAG, 1 PANEL, 2 TRTRI,
 * 3 LAUUM, 4 ALPHA); variant 0 factor, 1 + inverse, 2 + gradient.  Returns the
 * number of items.  The order must be a topological order of the tile DAG. */
int64_t hb_debug_items(hb_handle_t h, int T, const int64_t* offs_host, int d, int variant,
                       int32_t* out, int64_t max_items);
/* Debug: synchronises the device and returns 1 if a work item of the persistent
 * kernel ever gave up waiting for a dependency (a scheduling bug), else 0. */
int hb_debug_fused_timeout(hb_handle_t h);
/* Debug: synchronises and copies an internal workspace buffer to the host;
 * which: 0 L tiles, 1 M tiles, 2 W tiles, 3 z, 4 alpha, 5 apart|rpart, 6 gpart,
 * 7 gtask, 8 logdet, 9 nll_task, 10 sync words.  Returns the buffer's capacity
 * in bytes (or -1). */
int64_t hb_debug_read(hb_handle_t h, int which, void* host_out, int64_t max_bytes);
/* Per-kernel timing for bench.py's roofline: when enabled, CUDA events are
 * recorded on the caller's stream around each section of
 * hb_nll_grad_batched / hb_factorize_batched:
 *   0 = factorisation launches (k_prep + k_step x (nblk+1))
 *   1 = k_alpha   2 = k_lauum_grad   3 = reductions
 * hb_profile_read synchronises on the recorded events and returns the
 * accumulated milliseconds and launch-group counts since the last enable. */
#define HB_PROFILE_SECTIONS 4
int hb_profile_enable(hb_handle_t h, int enable);
int hb_profile_read(hb_handle_t h, double* ms_out, int64_t* count_out);

/* ---- a4-a7: Gram / cross-Gram ------------------------------------------- */
/* kernel.covariance_matrix.matrix_map (gp_utils/kernel.py:33-58) for
 * squared_exponential / matern32 / matern52 (kernel.py:63-123), optionally
 * followed by  + I*(noise_variance + jitter)  (linalg.compute_delta_y_and_cov,
 * basics/linalg.py:64-68) when add_noise != 0 and X2 == NULL.
 * out: (n1, n2) row-major; or (n1,) when diag_only != 0 and X2 == NULL
 * (kernel.py:54-56 -- diag is ignored when X2 is given). */
int hb_kernel_matrix(hb_handle_t h, int kernel_id, const void* X1, int64_t n1,
                     const void* X2_or_null, int64_t n2, int d,
                     const void* raw, uint64_t warp_mask, int diag_only,
                     int add_noise, double jitter, void* out, void* stream);

/* ---- a7-a10: batched factorisation + per-task NLL ----------------------- */
/* linalg.solve_gp_linear_system (basics/linalg.py:72-110) for every task, plus
 * objectives.neg_log_marginal_likelihood's per-task value
 * (gp_utils/objectives.py:144-156):
 *   K~_t = K(X_t,X_t) + (noise_variance + 1e-6) I,  L_t = chol(K~_t),
 *   alpha_t = K~_t^{-1} (y_t - m),  nll_t = .5 r'alpha + sum log L_ii + .5 n log 2pi.
 * chol_out_or_null: concatenated row-major (n_t, n_t) lower factors with the
 *   strict upper triangle zeroed (== jspla.cholesky(lower=True)), task t at
 *   element offset sum_{s<t} n_s^2.
 * alpha_out_or_null: (sum n,)  nll_out_or_null: (T,)  info_out_or_null: (T,) int32. */
int hb_factorize_batched(hb_handle_t h, int kernel_id, int mean_id, int T,
                         const int64_t* offs_host, int d, const void* X,
                         const void* y, const void* raw, uint64_t warp_mask,
                         void* chol_out_or_null, void* alpha_out_or_null,
                         void* nll_out_or_null, int32_t* info_out_or_null,
                         void* stream);

/* ---- a10: batched NLL + gradient w.r.t. the raw parameters -------------- */
/* What jax.value_and_grad(loss_func) computes at gp_utils/gp.py:134 for
 * objective = neg_log_marginal_likelihood, restated in closed form
 * (G_t = .5 (K~^{-1} - alpha alpha')), as SUMS over the tasks of this call so
 * that task shards on several GPUs combine with one all-reduce(sum):
 *   sums_out[0]        = sum_t nll_t
 *   sums_out[1 + p]    = sum_t d nll_t / d raw_p     (p < P = 3 + d)
 *   sums_out[1 + P]    = number of non-empty tasks
 * sums_out has P + 2 scalars.  nll_task_out_or_null: (T,) per-task nll
 * (return_key2nll, objectives.py:208-209). */
int hb_nll_grad_batched(hb_handle_t h, int kernel_id, int mean_id, int T,
                        const int64_t* offs_host, int d, const void* X,
                        const void* y, const void* raw, uint64_t warp_mask,
                        void* sums_out, void* nll_task_out_or_null,
                        int32_t* info_out_or_null, void* stream);

/* Same launch sequence with two generalisations, which turn it into the
 * building block of the reference's OTHER training objective, the empirical KL
 * divergence on aligned data (objectives.multivariate_normal_divergence,
 * gp_utils/objectives.py:29-101, with utils.kl_multivariate_normal /
 * partial_kl_mvn, gp_utils/utils.py:84-141):
 *   task_weight_or_null : (T,) device scalars w_t; sums_out[0] = sum w_t nll_t,
 *                         sums_out[1+p] = sum w_t d nll_t / d raw_p
 *                         (sums_out[1+P] stays the unweighted task count);
 *   jitter              : what is added to the diagonal next to the noise
 *                         variance (1e-6 in linalg.py:42; `eps`, default 0, in
 *                         utils.kl_multivariate_normal).
 * With B = [Yc / sqrt(m) | mu_model - mu_data] (n x (m+1)) the partial KL is
 *   tr(K1^-1 S0) + d'K1^-1 d + logdet K1 = 2 sum_q nll(y = B_q) - 2 m nll(y = 0)
 *                                          - n log 2pi,
 * i.e. ONE weighted call over m+2 tasks that share X (hyperbo_b200/gp_utils/
 * objectives.py builds them), gradient included. */
int hb_nll_grad_weighted(hb_handle_t h, int kernel_id, int mean_id, int T,
                         const int64_t* offs_host, int d, const void* X,
                         const void* y, const void* raw, uint64_t warp_mask,
                         const void* task_weight_or_null, double jitter,
                         void* sums_out, void* nll_task_out_or_null,
                         int32_t* info_out_or_null, void* stream);

/* Several right-hand sides on ONE factorisation per task: the same objective
 * family as hb_nll_grad_weighted without refactorising X once per column.
 *   B          : task t owns the (R, n_t) block at element offset offs[t] * R,
 *                column q contiguous (column-major per task);
 *   col_weight : (T, R) device scalars c_tq;
 *   col_mean_or_null : (R) device int32, 1 = subtract the model mean m(x)
 *                (mean_id) from column q, 0 = use the column as it is;
 *   task_weight_or_null : (T,) device scalars w_t (1 if null).
 *   sums_out[0]   = sum_t { w_t (sum_i log L_ii + .5 n_t log 2pi)
 *                           + sum_q c_tq .5 r_tq' K~_t^-1 r_tq },
 *   sums_out[1+p] = d sums_out[0] / d raw_p,   sums_out[1+P] = task count.
 * The partial KL of utils.kl_multivariate_normal (gp_utils/utils.py:84-106) on
 * an aligned sub-dataset with m columns is the call with R = m + 1,
 * B = [Yc / sqrt(m) | mu_data], col_mean = [0,...,0,1], w_t = c_tq = 2 (x the
 * term's weight), minus n log 2pi: one Cholesky / inverse per sub-dataset
 * instead of the m + 2 of the hb_nll_grad_weighted decomposition. */
int hb_nll_grad_mrhs(hb_handle_t h, int kernel_id, int mean_id, int T,
                     const int64_t* offs_host, int d, const void* X, int R,
                     const void* B, const void* col_weight,
                     const int32_t* col_mean_or_null, const void* raw,
                     uint64_t warp_mask, const void* task_weight_or_null,
                     double jitter, void* sums_out, int32_t* info_out_or_null,
                     void* stream);

/* The reference's other regulariser on aligned data: the Euclidean distance of
 * utils.euclidean_multivariate_normal (gp_utils/utils.py:151-173) inside
 * objectives.multivariate_normal_divergence (gp_utils/objectives.py:29-101),
 * value AND gradient (objectives.nll_regeuc*, objectives.py:218,238):
 *   sums_out[0]   = sum_t w_t ( mean_weight ||mu0_t - m(x_t)||_2
 *                             + cov_weight  ||Yc_t Yc_t' - (K_t + nv I)||_F ),
 *   sums_out[1+p] = d sums_out[0] / d raw_p,   sums_out[1+P] = task count.
 * Yc: per task the (R, n_t) block of centred columns ALREADY scaled by
 * 1/sqrt(m) (so Yc Yc' = cov(y, bias=True)), same layout as B of
 * hb_nll_grad_mrhs; mu0: (sum n) row means.  No factorisation: one pass over
 * the lower-triangular 64x64 tiles evaluates K and dK from X. */
int hb_euclid_grad(hb_handle_t h, int kernel_id, int mean_id, int T,
                   const int64_t* offs_host, int d, const void* X, int R,
                   const void* Yc, const void* mu0, const void* raw,
                   uint64_t warp_mask, double mean_weight, double cov_weight,
                   const void* task_weight_or_null, void* sums_out, void* stream);

/* ---- a11: one optax.adam update (gp_utils/gp.py:124,143-144) ------------ */
/* state (device, handle dtype): raw[P], m[P], v[P], accepted[P].
 * scalars_io (device, 4 scalars): [0] loss of this step (written),
 *   [1] step counter t (read, incremented), [2] stopped flag (0/1),
 *   [3] number of accepted steps.
 * Semantics of gp.py:135-146: loss = sums[0]/sums[1+P]; if finite:
 * accepted <- raw, then raw <- raw - lr*mhat/(sqrt(vhat)+eps); else stopped=1
 * and nothing changes any more (the host loop `break`s).
 * tie_lengthscale != 0: the model has ONE scalar lengthscale broadcast over the
 * d inputs (kernel.py:80): its gradient is the sum over k and all d entries of
 * raw move together. */
int hb_adam_step(hb_handle_t h, int P, void* raw, void* m, void* v,
                 void* accepted, const void* sums, void* scalars_io, double lr,
                 double b1, double b2, double eps, int tie_lengthscale,
                 void* stream);

/* ---- 8(e): the one exchange step of task-sharded training ---------------- */
/* The loss is a mean over tasks (gp_utils/objectives.py:178-195), so task shards
 * on several GPUs (one process per GPU) combine with ONE all-reduce(sum) of the
 * P+2 partial sums per optimiser step.  These entries do it over NVLink peer
 * memory inside one small kernel per rank -- no NCCL call, capturable in a CUDA
 * graph -- and reduce in rank order, so every rank holds bit-identical sums.
 *
 *   hb_comm_export : allocates this rank's exchange buffer and writes its CUDA
 *                    IPC handle (HB_IPC_HANDLE_BYTES bytes) to ipc_handle_out;
 *   (the caller all-gathers the handles with whatever transport it has,)
 *   hb_comm_import : maps the peers' buffers; ipc_handles = world consecutive
 *                    handles in rank order.  Every rank must return from it
 *                    (barrier) before any rank enqueues an all-reduce.
 *   hb_allreduce   : buf[0..count) <- sum over ranks, count <= 64, in place;
 *   hb_allreduce_adam_step : the same on sums[0..P+2) followed by hb_adam_step's
 *                    update in the same kernel (what gp.infer_parameters needs
 *                    per step, gp_utils/gp.py:134-144).
 * Collective: every rank must enqueue the same sequence of these calls. */
#define HB_IPC_HANDLE_BYTES 64
int hb_comm_export(hb_handle_t h, void* ipc_handle_out);
int hb_comm_import(hb_handle_t h, int rank, int world, const void* ipc_handles);
int hb_allreduce(hb_handle_t h, void* buf, int count, void* stream);
int hb_allreduce_adam_step(hb_handle_t h, int P, void* raw, void* m, void* v,
                           void* accepted, void* sums, void* scalars_io, double lr,
                           double b1, double b2, double eps, int tie_lengthscale,
                           void* stream);

/* ---- a12/a13: predictor cache, predict, acquisition --------------------- */
/* Bytes of the opaque predictor cache for n observations (packed L^{-1} tiles,
 * alpha, padded).  */
int64_t hb_predictor_bytes(hb_handle_t h, int64_t n);
/* GP.setup_predictor (gp_utils/gp.py:540-560): factorise one task and fill the
 * caller-owned cache; optionally also emit the reference-visible GPCache fields
 * chol (n,n) row-major lower and kinvy (n,). */
int hb_build_predictor(hb_handle_t h, int kernel_id, int mean_id, int64_t n,
                       int d, const void* X, const void* y, const void* raw,
                       uint64_t warp_mask, void* cache, void* chol_out_or_null,
                       void* kinvy_out_or_null, void* nll_out_or_null,
                       int32_t* info_out_or_null, void* stream);
/* gp.predict (gp_utils/gp.py:242-305, full_cov=False) + GP.predict's noise /
 * N/(N-1) handling (gp.py:607-619) + acfun_sub (bo_utils/acfun.py:96-142):
 *   mu_q  = K*' alpha + m
 *   var_q = (k(xq,xq) - |L^{-1} K*_q|^2 + noise_add) * var_scale
 *   acq_q = acfun_sub(mu_q, sqrt(var_q), acq_param)        (if acq_id != NONE)
 * Outputs (nq,) each; any of mu/var/acq may be NULL. */
int hb_predict(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int d,
               const void* X, const void* cache, const void* raw,
               uint64_t warp_mask, int64_t nq, const void* Xq,
               double noise_add_flag, double var_scale, int acq_id,
               double acq_param, void* mu_out, void* var_out, void* acq_out,
               void* stream);
/* gp.predict(..., full_cov=True) (gp_utils/gp.py:295-300) with GP.predict's noise
 * and N/(N-1) handling (gp.py:607-619):
 *   cov = (k(Xq, Xq) - (L^{-1} K*)' (L^{-1} K*) + noise_add I) * var_scale
 * cov_out: (nq, nq) row-major, both triangles; mu_out (nq,) may be NULL.
 * nq <= 16384 (one pass; HB_ERR_UNSUPPORTED beyond).  n = 0: the prior. */
int hb_predict_cov(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int d,
                   const void* X, const void* cache, const void* raw,
                   uint64_t warp_mask, int64_t nq, const void* Xq,
                   double noise_add_flag, double var_scale, void* mu_out_or_null,
                   void* cov_out, void* stream);

/* acfun_sub alone on given mu / var vectors (bo_utils/acfun.py:96-142). */
int hb_acquisition(hb_handle_t h, int acq_id, double acq_param, int64_t nq,
                   const void* mu, const void* var, void* out, void* stream);

/* ---- 8(f) rank 4: S hyper-parameter sets x the same data ---------------- */
/* The reference evaluates the same dataset under many parameter vectors by
 * vmapping the whole pipeline (bo_utils/acfun_test.py:74-118, S = 100) or by
 * looping over HGP samples (gp_utils/gp.py:623-682).  Here the parameter set is
 * a second batch axis of ONE launch sequence: virtual task (s, t) reads the
 * data rows of task t and raw_sets[s] ((S, 3+d) row-major).
 *   hb_nll_grad_multi         : sums_out (S, P+2) as hb_nll_grad_batched per
 *                               set; nll_task_out_or_null (S, T)
 *                               (speculative line-search points, HGP stats);
 *   hb_build_predictors_multi : S predictor caches of one task (cache s at
 *                               caches + s * cache_stride_bytes, each as
 *                               hb_build_predictor's), nll_out_or_null (S,). */
int hb_nll_grad_multi(hb_handle_t h, int kernel_id, int mean_id, int S, int T,
                      const int64_t* offs_host, int d, const void* X, const void* y,
                      const void* raw_sets, uint64_t warp_mask, void* sums_out,
                      void* nll_task_out_or_null, void* stream);
int hb_build_predictors_multi(hb_handle_t h, int kernel_id, int mean_id, int S, int64_t n,
                              int d, const void* X, const void* y, const void* raw_sets,
                              uint64_t warp_mask, void* caches,
                              int64_t cache_stride_bytes, void* nll_out_or_null,
                              int32_t* info_out_or_null, void* stream);

/* ---- a11: per-step sub-sampling on the device --------------------------- */
/* data_utils.sub_sample_dataset_iterator (basics/data_utils.py:72-100): tasks
 * with n >= batch_size are replaced, every step, by batch_size of their points
 * drawn uniformly without replacement; others pass unchanged.  hb_subsample
 * gathers such a sample of a packed source batch into a packed destination
 * batch of FIXED shape (n'_t = offs_dst[t+1] - offs_dst[t] = min(n_t,
 * batch_size) or n_t), keyed by (seed, step, task_ids[t]) with a counter-based
 * generator, so the result does not depend on the task sharding.  The step is
 * read from step_scalars_dev[1] (the step counter hb_adam_step maintains) when
 * that pointer is given -- the call can then sit in the step's CUDA graph --
 * else from `step`.  offs_* / task_ids: int64 DEVICE arrays (T+1 / T). */
int hb_subsample(hb_handle_t h, int T, int d, const void* offs_src_dev,
                 const void* offs_dst_dev, const void* task_ids_dev, int64_t max_rows,
                 const void* Xs, const void* ys, void* Xd, void* yd, uint64_t seed,
                 const void* step_scalars_dev_or_null, int64_t step, void* stream);
/* The permutation hb_subsample applies: source row of destination row i. */
uint32_t hb_subsample_perm(uint32_t i, uint32_t n, uint64_t seed, uint64_t step,
                           int64_t task_id);

/* ---- 8(f) rank 1: device-resident simulated Bayesian optimisation -------- */
/* bo_utils/bayesopt.py:137-193 per iteration: evaluate the acquisition on ALL
 * candidates, take the arg-max, append the chosen (x, y) to the queried task
 * (gp.py:446-452) and re-condition the GP.  The reference refactorises from
 * scratch every iteration (gp.py:284: "one can potentially support rank-1
 * updates"); hb_bo_step keeps observations, candidates and the factor on the
 * device and appends one row to the packed L^{-1} tiles and alpha in O(n^2),
 * with no host synchronisation inside the loop.
 *   cache      : hb_bo_cache_bytes(n_cap) bytes, filled by hb_bo_init from the
 *                first n0 observations (n0 = 0 allowed: prior prediction, EI/PI
 *                target 0.0 as acfun.py:145-148);
 *   X, y       : (n_cap, d) / (n_cap,) device buffers; rows < n are the
 *                observations, hb_bo_step writes row n;
 *   acquisition: acq_id as hb_predict; its parameter is acq_param (UCB beta), or
 *                max(y_observed) + acq_param when target_is_ymax != 0 (EI: +0,
 *                PI: + zeta), GP.predict's with_noise / N/(N-1) conventions via
 *                noise_add_flag / var_scale;
 *   sel_out    : device int32, receives the index of the chosen candidate.
 * Call with n = n0, n0+1, ... (n + 1 <= n_cap). */
int64_t hb_bo_cache_bytes(hb_handle_t h, int64_t n_cap);
int hb_bo_init(hb_handle_t h, int kernel_id, int mean_id, int64_t n0, int64_t n_cap,
               int d, const void* X, const void* y, const void* raw,
               uint64_t warp_mask, void* cache, void* stream);
int hb_bo_step(hb_handle_t h, int kernel_id, int mean_id, int64_t n, int64_t n_cap,
               int d, void* X, void* y, const void* raw, uint64_t warp_mask,
               void* cache, int64_t nq, const void* Xq, const void* yq,
               double noise_add_flag, double var_scale, int acq_id, double acq_param,
               int target_is_ymax, int32_t* sel_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HYPERBO_B200_H_ */
