/* b200lap.h -- C ABI of libb200lap.so, the B200 (sm_100a) implementation of the warm-start LAP
 * hot path of egbariajad/GNN-Accelerated-LAP-Warm-Start-Pipeline.
 *
 * Plain pointers and sizes only (no torch types).  Every entry point names the reference
 * interface it replaces (paths relative to the reference repository root).  Unless stated
 * otherwise a function returns 0 on success, a negative reference return code where the
 * reference defines one (-1 allocation, -2 n <= 0, -3 infeasible, -4 not square), or
 * B200LAP_ERR_CUDA (-100) / B200LAP_ERR_UNSUPPORTED (-101) / B200LAP_ERR_ARG (-102);
 * b200lap_last_error() then describes the failure.  There is no CPU fallback: without a
 * usable CUDA device every compute entry point fails with B200LAP_ERR_CUDA.
 */
#ifndef B200LAP_H
#define B200LAP_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200LAP_ERR_CUDA (-100)
#define B200LAP_ERR_UNSUPPORTED (-101)
#define B200LAP_ERR_ARG (-102)

#define B200LAP_ROW_FEAT_DIM 21   /* gnn/features.py:223-241 */
#define B200LAP_TRACE_WORDS 48
#define B200LAP_TOPK_MAX 32

typedef struct b200lap_ctx b200lap_ctx;       /* one per (device, stream): workspaces + stream */
typedef struct b200lap_model b200lap_model;   /* OneGNN weights packed on a device              */
typedef struct b200lap_job b200lap_job;       /* one batch submitted to the asynchronous host-buffer pipeline */

/* ------------------------------------------------------------------------------------------
 * 1. Drop-in: the reference's only extern "C" symbol.
 *    Replaces LAP/lap/lapjv_seeded.h:8-13 (definition LAP/_lapjv_cpp/lapjv_seeded.cpp:19-23),
 *    bound by LAP/lap/_seeded_jv.pyx:7-12.  HOST pointers, borrowed for the call; C row-major
 *    n_rows x n_cols binary64; x[n_rows], y[n_cols] written only on success.
 * ------------------------------------------------------------------------------------------ */
int lapjv_seeded(const double* C, int n_rows, int n_cols, long long* x, long long* y,
                 const double* u_seed, const double* v_seed, double eps);

/* Cold solve = lapjv_internal (LAP/_lapjv_cpp/lapjv.cpp:323-346, C++ linkage in the reference,
 * reached through LAP/_lapjv_cpp/_lapjv.pyx:38-129).  HOST pointers, square n x n. */
int b200lap_lapjv(const double* C, int n, int* x, int* y);

/* Batched host entry: `batch` independent instances, C [batch][n][n], seeds [batch][n].
 * rc[batch] receives each instance's lapjv_seeded return code; trace (nullable)
 * [batch][B200LAP_TRACE_WORDS] the phase counters (proj_triggers, tight_edges, greedy_matched,
 * took_fallback, micro_bumps, free_after_cr, arr_iters, aug_paths, collect_calls, relax_cols, rc, then SM-clock
 * cycles spent in relax steps, collect steps, ARR scans, ARR serial updates, the whole solve; collect records,
 * cycles of the serial collect / relax replays, relax hits). */
int b200lap_lapjv_seeded_batch(const double* C, int batch, int n, long long* x, long long* y,
                               const double* u_seed, const double* v_seed, double eps, int* rc,
                               long long* trace);

/* ------------------------------------------------------------------------------------------
 * 2. Contexts (device-pointer API).  A context owns a stream and cached workspaces on one
 *    device; calls on one context are serialised by the caller.  `stream` may be 0 (the
 *    context then creates its own non-blocking stream) or an existing cudaStream_t.
 * ------------------------------------------------------------------------------------------ */
int b200lap_ctx_create(int device, void* stream, b200lap_ctx** out);
void b200lap_ctx_destroy(b200lap_ctx* ctx);
void* b200lap_ctx_stream(b200lap_ctx* ctx);          /* the cudaStream_t work is enqueued on */
int b200lap_ctx_sync(b200lap_ctx* ctx);               /* waits for every lane */
/* A context has up to eight LANES (stream + workspace each).  With the option overlap_steps = L (2..8), consecutive
 * b200lap_dev_pipeline calls rotate over L of them, so L independent batches are in flight (a 64-instance solve
 * occupies 64 of the 148 SMs; overlap_steps < 2 = one batch at a time).  Results of such calls are complete after
 * b200lap_ctx_sync, or -- for work enqueued on lane 0's stream -- after b200lap_ctx_join, which makes lane 0's stream
 * wait for the other lanes on the device.  A lane's first call allocates its workspace: warm every lane before timing. */
void* b200lap_ctx_lane_stream(b200lap_ctx* ctx, int lane);
int b200lap_ctx_last_lane(b200lap_ctx* ctx);            /* lane the most recent b200lap_dev_pipeline call was enqueued on */
int b200lap_ctx_join(b200lap_ctx* ctx);
/* Tuning / test options (0 = automatic unless stated): solver_threads, solver_cluster (CTAs per instance: 1 = single
 * CTA, 2/4/8 = thread-block cluster; auto = 8 from solver_cluster_min_n = 8192 on), force_global_state,
 * solver_smem_budget, front_rows_per_cta, mlp_impl (1 = FFMA instead of tcgen05), feat_impl (1 = register-resident
 * row-feature kernel), feat_threads, feat_nbuf, feat_nsamp, feat_ctas, feat_ept, feat_group (warps per row of the group
 * kernel), feat_torch_mode (1 = the definitions of compute_row_features_torch, gnn/features.py:246-351, binary32 input).  Unknown keys return
 * B200LAP_ERR_ARG.  No option changes an assignment or a selected order statistic; the feat_* launch shapes change
 * floating-point summation orders (features stay inside the stated 1e-4 tolerance). */
int b200lap_ctx_set_option(b200lap_ctx* ctx, const char* key, long long value);
long long b200lap_ctx_launch_count(b200lap_ctx* ctx); /* kernels launched so far on this context */
/* rows of the LAST row-feature call on this context that the fast kernel handed to the exact fall-back kernel
 * (diagnostic; synchronises the stream; valid until the next call that resets the context workspace) */
long long b200lap_ctx_feature_redo_rows(b200lap_ctx* ctx);
const char* b200lap_last_error(void);
int b200lap_device_count(void);

/* ------------------------------------------------------------------------------------------
 * 3. Dense half on DEVICE buffers.  `C` points to `batch` row-major n x n matrices, instance
 *    stride n*n elements, stored as binary32 (is_f64 = 0; the caller guarantees every entry is
 *    exactly representable) or binary64 (is_f64 = 1).  All outputs are device buffers.
 * ------------------------------------------------------------------------------------------ */
/* binary64 -> binary32 copy; *exact (device int, set to 1 by the caller) is cleared when any entry
 * does not survive the round trip. */
int b200lap_dev_narrow(b200lap_ctx* ctx, const double* src, long long count, float* dst, int* exact);
/* column minima + first row attaining them (gnn/features.py:218 col_min; lapjv.cpp:21-32). colmin has C's type. */
int b200lap_dev_col_argmin(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, void* colmin, int* colarg);
/* gnn/features.py:161-243 compute_row_features -> feat [batch][n][21] f32; topv [batch][n][topk] the
 * topk smallest entries of every row, ascending, rounded to binary32 (feeds gnn/one_gnn.py:143-147). */
int b200lap_dev_row_features(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, int topk,
                             const void* colmin, float* feat, float* topv);
/* gnn/one_gnn.py:89-120 OneGNN.forward (inference, all-true mask) from features + top-k values
 * -> u [batch][n] f32 (mean-centred) and raw [batch][n] f32 (head output before centring, nullable). */
int b200lap_dev_onegnn_forward(b200lap_ctx* ctx, const b200lap_model* model, const float* feat, const float* topv,
                               int has_cost, int batch, int n, float* u, float* raw);
/* scripts/gnn_benchmark.py:262 min-trick: v_j = min_i (C_ij - (double)u_i), binary64, exact. */
int b200lap_dev_min_trick(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const float* u, double* v);
/* GNNPredictor.predict (scripts/gnn_benchmark.py:213-289): features -> OneGNN -> min-trick.
 * u64/v64 [batch][n] binary64 (u64 holds the binary32 values widened); feat/topv nullable. */
int b200lap_dev_predict_duals(b200lap_ctx* ctx, const b200lap_model* model, const void* C, int is_f64, int batch,
                              int n, double* u64, double* v64, float* u32, float* feat);

/* ------------------------------------------------------------------------------------------
 * 4. Solver half on DEVICE buffers (same matrix conventions).  x,y int32 [batch][n] (written
 *    for instances with rc == 0), rc int32 [batch], trace int64 [batch][B200LAP_TRACE_WORDS] nullable,
 *    v_out binary64 [batch][n] nullable (final column potentials).
 * ------------------------------------------------------------------------------------------ */
int b200lap_dev_solve_seeded(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const double* u_seed,
                             const double* v_seed, double eps, int* x, int* y, int* rc, long long* trace,
                             double* v_out);
int b200lap_dev_solve_cold(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, int* x, int* y, int* rc,
                           long long* trace, double* v_out);
/* Front-end sweep alone (solvers/advanced_dual.py:14-63 project/reduce/check stated on the solver's
 * own sweeps): u_tight [batch][n], tight_cnt [batch][n], flags [batch][4] int32 =
 * {any_violation, infeasible, total_tight lo, total_tight hi}. */
int b200lap_dev_front_end(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const double* u_seed,
                          const double* v_seed, double eps, double* u_tight, int* tight_cnt, int* flags);
/* solvers/advanced_dual.py:14-36 project_feasible on device buffers of one instance: u, v [n] binary64 are tightened in
 * place (u = min(u, min_j(C - v)); v = min(v, min_i(C - u)); stop when no (c - u_i) - v_j < -tol, at most max_rounds
 * rounds, at least one); *rounds (host, nullable) receives the rounds used. */
int b200lap_dev_project_feasible(b200lap_ctx* ctx, const void* C, int is_f64, int n, double* u, double* v, int max_rounds,
                                 double tol, int* rounds);
/* solvers/advanced_dual.py:39-63: min_host[b] (HOST) = min_ij ((c_ij - u_i) - v_j) per instance -- the quantity
 * check_dual_feasible tests and reduce_costs shifts by; out (device, nullable) [batch][n][n] receives the unshifted
 * reduced-cost matrices (C - u 1^T) - 1 v^T. */
int b200lap_dev_reduced_costs(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const double* u, const double* v,
                              double* out, double* min_host);

/* Oracle duals by difference constraints -- solvers/dual_computation.py:13-47 (dual_from_matching_diff_constraints):
 * Bellman-Ford relaxation from v = 0 over the edges p -> j of weight C[r_p, j] - C[r_p, p] built from an optimal
 * matching x (row -> column, int32 [batch][n], device), every edge evaluated as the reference evaluates it, all n^2
 * edges of a round at once; returns the column potentials v (double [batch][n], device) and the rounds taken.
 * B200LAP_ERR_ARG when the relaxation still moves after n - 1 rounds (the reference's "Negative cycle" error). */
int b200lap_dev_bf_duals(b200lap_ctx* ctx, const void* C, int is_f64, int batch, int n, const int* x, double* v, int* rounds);
/* features -> OneGNN -> min-trick -> seeded solve without leaving the device. */
int b200lap_dev_pipeline(b200lap_ctx* ctx, const b200lap_model* model, const void* C, int is_f64, int batch, int n,
                         double eps, int* x, int* y, int* rc, double* u64, double* v64, long long* trace);

/* ------------------------------------------------------------------------------------------
 * 5. Model weights.  `params` is one HOST binary32 blob holding the reference state_dict tensors
 *    in the order of gnn/one_gnn.py's module definition:
 *    input_proj.0.{weight[H,F],bias[H]}, input_proj.2.{weight,bias}[H],
 *    per block: fc1.{weight[H,H],bias}, fc2.{weight,bias}, norm.{weight,bias};
 *    pre_out.{weight[1,H],bias[1]}, row_out.0.{weight[H/2,H],bias}, row_out.3.{weight[1,H/2],bias[1]},
 *    edge_mlp.0.{weight[H,1],bias[H]}, edge_mlp.2.{weight[H,H],bias[H]}, message_norm.{weight,bias}[H].
 * ------------------------------------------------------------------------------------------ */
int b200lap_model_create(b200lap_ctx* ctx, const float* params, long long n_params, int in_dim, int hidden,
                         int layers, int topk, b200lap_model** out);
void b200lap_model_destroy(b200lap_model* model);

/* ------------------------------------------------------------------------------------------
 * 6. Host-buffer conveniences used by the Python mirrors (upload, run, download).
 * ------------------------------------------------------------------------------------------ */
/* gnn.compute_row_features(C) -> feat[n][21] (gnn/features.py:161-243). */
int b200lap_compute_row_features(const double* C, int n, float* feat);
/* predict + solve for a batch of host matrices [batch][n][n]; u,v nullable outputs [batch][n]. */
int b200lap_pipeline_batch(const b200lap_model* model, const double* C, int batch, int n, double eps,
                           long long* x, long long* y, int* rc, double* u, double* v, long long* trace);
/* solvers/advanced_dual.py:14-36 project_feasible(C, u, v, max_rounds, tol): u, v [n] updated in place. */
int b200lap_project_feasible(const double* C, int n, double* u, double* v, int max_rounds, double tol, int* rounds);
/* solvers/advanced_dual.py:39-53 reduce_costs(C, u, v, shift_nonneg) -> out [n][n] (nullable: then only *min_out, the
 * minimum of the unshifted reduced costs, is produced -- what check_dual_feasible, :56-63, compares with -tol). */
int b200lap_reduce_costs(const double* C, int n, const double* u, const double* v, int shift_nonneg, double* out,
                         double* min_out);
/* Asynchronous form of b200lap_pipeline_batch for throughput: submit() enqueues upload, pipeline and download on one of
 * the default context's lanes (option overlap_steps = 2..8 through b200lap_ctx_set_option, at least two here) and returns once the host-side
 * marshalling of the upload is done (pass pinned host memory for C, or the upload blocks); wait() blocks until that
 * batch is complete, writes x, y [batch][n] (rows of instances with rc != 0 read -1) and rc, and frees the job.  One
 * batch may be outstanding per lane: with L lanes, submit k+L needs wait k first. */
int b200lap_pipeline_batch_submit(const b200lap_model* model, const double* C, int batch, int n, double eps, long long* x,
                                  long long* y, int* rc, b200lap_job** job);
int b200lap_pipeline_batch_wait(b200lap_job* job);
/* How submit() moves a batch to the device: the first `head` = round(batch * percent / 100) instances are narrowed to
 * binary32 by `threads` host threads into pinned staging and uploaded at 4 bytes per entry, while the remaining
 * instances go up as binary64 by DMA and are narrowed on the device (csrc/host_narrow.cpp; the reference hands over
 * binary64 matrices, scripts/gnn_benchmark.py:226).  B200LAP_HOST_NARROW_THREADS (0 = everything as binary64) and
 * B200LAP_HOST_NARROW_PERCENT set them; a head that is not binary32-representable is uploaded as binary64 too.
 * Returns the bytes submit() copies host -> device for a representable (batch, n) batch. */
long long b200lap_host_narrow_config(int batch, int n, int* threads, int* percent);
b200lap_ctx* b200lap_default_ctx(void);

#ifdef __cplusplus
}
#endif
#endif /* B200LAP_H */
