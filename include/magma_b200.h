/*
 * magma_b200.h -- C ABI of the B200-native MagmaClustR hot path.
 *
 * Every entry point is `extern "C"`, takes plain pointers/sizes to CALLER-OWNED HOST
 * memory (unless the name ends in _dev), returns an int status and never throws:
 *     0  ok
 *    <0  argument / CUDA error   (mb_last_error() gives the text)
 *    >0  numerical failure       (e.g. jitter retries exhausted; per-task `info` arrays say where)
 *
 * It replaces, for the batched-over-tasks GP numerics, what the reference reaches through
 * its own FFI (`.Call` symbols registered in /root/reference/src/init.c:17-30) and the R
 * functions above it.  Each declaration cites the reference interface it stands for.
 * INTEGRATION.md shows the R-side `.Call` shim a maintainer would add.
 *
 * Conventions
 *   - matrices are column-major doubles with explicit leading dimension where one is given;
 *     symmetric results are returned full (both triangles), like chol2inv().
 *   - ragged tasks are CSR: task t owns rows offs[t] .. offs[t+1]-1 of X (row-major rows x D),
 *     y, and grid_index (0-based position of each row in the sorted union grid,
 *     = match(db$Reference, all_input) - 1 computed by the caller).
 *   - hyper-parameters are log-scale doubles laid out per mb_kernel (below); a table of
 *     per-task HPs is row-major n_tasks x n_hp.
 *   - one caller thread per context; contexts are independent.
 */
#ifndef MAGMA_B200_H
#define MAGMA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MB_MAX_TERMS 4
#define MB_MAX_HP 12
#define MB_MAX_CLUSTERS 16

/* base kernels, /root/reference/R/kernels.R:164-398; MB_KERN_CONV = the multi-output convolution kernel with one latent
 * process, /root/reference/R/kernels.R:16-140 */
enum { MB_KERN_SE = 0, MB_KERN_PERIO = 1, MB_KERN_RQ = 2, MB_KERN_LIN = 3, MB_KERN_CONV = 4 };

/*
 * A kernel string of the reference's grammar "(k1 * k2 * ...) + k + k ..."
 * (/root/reference/R/kernel-to-matrices.R:163-340): `types[0..n_prod-1]` are multiplied, the
 * remaining terms are added.  HP layout, in term order: SE (variance, lengthscale);
 * PERIO (variance, lengthscale, period); RQ (variance, lengthscale, scale); LIN (slope, offset);
 * then `noise` last when has_noise != 0.  n_hp is the total (filled by mb_kernel_parse).
 *
 * Multi-output: kern = "CONV<O>" (e.g. "CONV2"; what `kern = convolution_kernel` is in R) is ONE term of type MB_KERN_CONV
 * with n_prod = O outputs.  Inputs then carry the 0-based output index of every point as their LAST column (the
 * reference's Output_ID, part of the Reference key "o<id>;<coords>", kernel-to-matrices.R:94-96) and the HP vector is
 * flatten_hp_mo()'s (R/hyperparameters.R:348-396): l_t_1..l_t_O, S_t_1..S_t_O, l_u_t, then noise_1..noise_O when
 * has_noise != 0 (per-output noise exp(noise_o) on the diagonal, kernel-to-matrices.R:146-157).  O <= 3 with noise.
 */
typedef struct mb_kernel {
  int32_t n_terms;
  int32_t n_prod;
  int32_t types[MB_MAX_TERMS];
  int32_t has_noise;
  int32_t n_hp;
} mb_kernel;

typedef struct mb_context mb_context;

/* ---- context -------------------------------------------------------------------------- */
int mb_create(mb_context** ctx, int device);
int mb_destroy(mb_context* ctx);
const char* mb_last_error(void);
/* parse "SE * RQ + PERIO" (kernel-to-matrices.R:163-204); has_noise appends the noise HP */
int mb_kernel_parse(const char* kern, int has_noise, mb_kernel* out);
/* name of HP `p` of a kernel ("se_variance", ..., "noise"); NULL when out of range */
const char* mb_kernel_hp_name(const mb_kernel* k, int p);
/* counters: kernels launched by this context since creation (bench.py `gpu_launches`) */
int64_t mb_launch_count(mb_context* ctx);
/* tuning / diagnostic switches of one context: "fused_below" (MB_FUSED_BELOW), "reserve_sms" (MB_RESERVE_SMS),
 * "dense_split" (0 = MB_DENSE_NOSPLIT), "logdet_chol" (0 = the reference's LU log-determinant, R/likelihoods.R:37-41,
 * like MB_LOGDET=lu; 1 = read off the Cholesky diagonal), "monitor_recompute" (MB_MONITOR_RECOMPUTE: 1 = mb_em_step* /
 * mb_train_magma evaluate logL_monitoring from scratch; 0, the default = the per-task and mean-process values
 * logL_GP_mod(new HPs) are taken from the M step's optimisers, which ended on exactly those values -- same numbers, one
 * all-task launch less per EM iteration), "pair_tables" (MB_PAIR_TABLES: 1, the default = the SE kernel of a resident
 * task is evaluated once per DISTINCT squared distance of the task and looked up per matrix element, see
 * mb_pair_table_counts; 0 = element by element; same bits either way).  The environment variables are read once at
 * mb_create. */
int mb_set_option(mb_context* ctx, const char* name, int64_t value);
/* Pair tables of the resident dataset (built by mb_upload_dataset for tasks of at most 96 points): counts[t] = number of
 * distinct values of sum_q (x_rq - x_cq)^2 over the pairs of task t, or -1 when the task keeps the element-by-element
 * evaluation (more than 1024 distinct values, or NaN inputs).  The reference evaluates every element (src/code.cpp:6-131
 * under R/kernel-to-matrices.R:71-503); tasks observed on a common grid repeat the same distances many times.
 * Returns -1 when no tables exist (no dataset, larger tasks, or the option switched off at upload). */
int mb_pair_table_counts(mb_context* ctx, int32_t* counts);

/* optional sum-all-reduce over ranks of `count` doubles in DEVICE memory, called on the
 * context's stream after it was synchronised (multi-GPU sharding).  The hook
 * must return 0 and leave the reduced values in place. */
typedef int (*mb_allreduce_fn)(void* user, void* dev_ptr, int64_t count);
int mb_set_allreduce(mb_context* ctx, mb_allreduce_fn fn, void* user, int rank, int world);
/* NCCL owned by the library (the E step's union-grid sums, the shared-HP objective and the monitoring
 * scalar are summed over the ranks).  Rank 0 calls mb_nccl_unique_id (128 bytes) and the host program distributes the id
 * by whatever channel it has (MPI, a file, torch.distributed ...); every rank then calls mb_nccl_init with its own
 * context.  The all-reduces are enqueued on the library's streams (ncclAllReduce, ncclDouble, ncclSum) -- no host round
 * trip.  libnccl.so.2 is resolved with dlopen, so single-GPU users need no NCCL at all. */
int mb_nccl_unique_id(char* id_out, int bytes);
int mb_nccl_init(mb_context* ctx, const char* id, int bytes, int rank, int world);
/* Sharding that keeps the results independent of the number of GPUs.  The reference adds the tasks one after the other
 * (R/em-magma.R:39-46, R/likelihoods.R:253-312); the library cuts the GLOBAL task list (ID order, training.R:699-701) into
 * mb_estep_leaves(M, U) contiguous LEAVES -- a power of two fixed by the task count M and the union-grid size U alone -- sums
 * the tasks of a leaf in task order and the leaves in a fixed binary tree.  A rank must hold whole leaves:
 * mb_shard_range gives rank `rank` of `world` its task range [first, last), and mb_set_shard tells a context which range
 * of the global list its resident dataset is (call it before the first E step; a single-GPU run needs neither).  With
 * that, post_mean / post_cov / log-likelihoods / EM step counts are bit-identical on 1, 2, 4, 8 GPUs. */
int mb_estep_leaves(int64_t n_tasks_global, int32_t n_grid);
int mb_shard_range(int64_t n_tasks_global, int32_t n_grid, int rank, int world, int64_t* first, int64_t* last);
int mb_set_shard(mb_context* ctx, int64_t first_task_global, int64_t n_tasks_global);

/* ---- the five legacy builders: /root/reference/src/code.cpp, src/init.c:17-24 ----------
 * m1 is n1 x d, m2 is n2 x d, out is n1 x n2, all column-major (R numeric matrices). */
int mb_cpp_dist(mb_context*, const double* m1, int n1, const double* m2, int n2, int d, double* out);
int mb_cpp_prod(mb_context*, const double* m1, int n1, const double* m2, int n2, int d, double* out);
int mb_cpp_perio(mb_context*, const double* m1, int n1, const double* m2, int n2, int d, double period, double* out);
int mb_cpp_perio_deriv(mb_context*, const double* m1, int n1, const double* m2, int n2, int d, double period, double* out);
int mb_cpp_noise(mb_context*, const double* m1, int n1, const double* m2, int n2, int d, double noise, double* out);

/* ---- kern_to_cov / kern_to_inv / chol_inv_jitter --------------------------------------- */
/* kern_to_cov(input, kern, hp, deriv, input_2) (kernel-to-matrices.R:71-503).  x1: n1 x d,
 * x2: n2 x d or NULL (= x1), ROW-major points.  deriv = -1 for the covariance, else the HP
 * index whose derivative matrix is wanted.  out: n1 x n2 column-major. */
int mb_kern_to_cov(mb_context*, const mb_kernel* k, const double* hp, const double* x1, int n1,
                   const double* x2, int n2, int d, int deriv, double* out);
/* list_kern_to_cov / list_kern_to_inv (kernel-to-matrices.R:586-652) over CSR tasks.
 * hp: n_tasks x n_hp (hp_shared = 0) or one row (hp_shared = 1).  out: blocks of n_t x n_t
 * column-major, block t at out + out_offs[t] (out_offs has n_tasks entries, caller-computed).
 * inverse != 0 applies chol_inv_jitter(pen_diag); retries[t] (may be NULL) receives the
 * number of jitter escalations of task t. */
int mb_list_kern_to_mat(mb_context*, const mb_kernel* k, int64_t n_tasks, const int64_t* offs,
                        int d, const double* X, const double* hp, int hp_shared, int deriv,
                        int inverse, double pen_diag, const int64_t* out_offs, double* out,
                        int32_t* retries);
/* chol_inv_jitter(mat, pen_diag) (kernel-to-matrices.R:670-680): inverse of mat + cumulative
 * jitter; reads the upper triangle only, like chol().  n_retry receives the retry count and
 * jitter_total the diagonal shift actually applied (both may be NULL). */
int mb_chol_inv_jitter(mb_context*, const double* mat, int n, int ld, double pen_diag,
                       double* inv, int32_t* n_retry, double* jitter_total);

/* ---- resident dataset (what `db` is to e_step/m_step) ---------------------------------- */
/* Upload the prepared training table: tasks in the caller's (string-sorted) order, rows in
 * Reference order, union grid (U points, row-major U x d) in Reference order. */
int mb_upload_dataset(mb_context*, int64_t n_tasks, const int64_t* offs, int d, const double* X,
                      const double* y, const int32_t* grid_index, int32_t n_grid,
                      const double* grid_X);

/* ---- Magma: e_step / m_step / logL_monitoring / train_magma ---------------------------- */
/* e_step (em-magma.R:25-82).  hp_0: kern_0 HPs (no noise); hp_i: n_tasks x n_hp (or one row
 * when hp_i_shared); m_0: U.  post_mean: U, post_cov: U x U (column-major, full).
 * info (may be NULL, 3 ints): jitter retries of inv_0, of post_cov, and max over tasks. */
int mb_e_step(mb_context*, const mb_kernel* kern_0, const double* hp_0, const mb_kernel* kern_i,
              const double* hp_i, int hp_i_shared, const double* m_0, double pen_diag,
              double* post_mean, double* post_cov, int32_t* info);

/* logL_GP_mod + gr_GP_mod (likelihoods.R:155-174, gradients-likelihoods.R:71-97) for every
 * task of the resident dataset in one launch, with mean = post_mean[t_i], post_cov[t_i,t_i]
 * gathered from the U-grid arrays.  f: n_tasks, g: n_tasks x n_hp (NULL = values only). */
int mb_logL_GP_mod_tasks(mb_context*, const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                         const double* post_mean, const double* post_cov, double pen_diag,
                         double* f, double* g, int32_t* retries);
/* the same pair for ONE matrix-sized problem given explicitly (mean process, cluster
 * processes): x n x d row-major, y, mean (n), post_cov n x n. */
int mb_logL_GP_mod(mb_context*, const mb_kernel* k, const double* hp, const double* x, int n, int d,
                   const double* y, const double* mean, const double* post_cov, double pen_diag,
                   double* f, double* g);

/* m_step (em-magma.R:119-235): L-BFGS-B (factr 1e13, maxit 25, lmm 5) for hp_0 and for hp_i
 * (one shared optimisation when shared_hp, else n_tasks independent problems advanced in
 * lock-step on the device).  hp_0 / hp_i are updated in place.  stats (may be NULL, 4 ints):
 * objective evaluations of hp_0, lock-step rounds for the tasks, total task evaluations,
 * tasks that ended with a non-zero optim `convergence` code. */
int mb_m_step(mb_context*, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
              double* hp_i, int shared_hp, const double* m_0, const double* post_mean,
              const double* post_cov, double pen_diag, int64_t* stats);

/* Per-task outcome of the last per-task M step: optim's `convergence` code (0 ok, 1 maxit,
 * 52 abnormal line search, 99 non-finite objective -- R raises "L-BFGS-B needs finite values of
 * 'fn'" there) and the number of objective evaluations.  Either pointer may be NULL. */
int mb_last_task_status(mb_context*, int32_t* convergence, int32_t* n_eval);

/* logL_monitoring (likelihoods.R:253-312). */
int mb_logL_monitoring(mb_context*, const mb_kernel* kern_0, const double* hp_0,
                       const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                       const double* m_0, const double* post_mean, const double* post_cov,
                       double pen_diag, double* logL);

/* One EM iteration of train_magma's loop body (training.R:905-1030) on the resident dataset,
 * everything kept on the device between the three phases: e_step, m_step (HPs updated in
 * place), logL_monitoring with the new HPs.  post_mean/post_cov may be NULL (not copied
 * back).  stats as in mb_m_step. */
int mb_em_step(mb_context*, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
               double* hp_i, int shared_hp, const double* m_0, double pen_diag, double* logL,
               double* post_mean, double* post_cov, int64_t* stats);

/* The whole EM loop with the reference's convergence rule (training.R:978-1029).
 * seq_logL: n_iter_max doubles; n_iter: iterations done; converged: 0/1. */
int mb_train_magma(mb_context*, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                   double* hp_i, int shared_hp, const double* m_0, double pen_diag,
                   int n_iter_max, double cv_threshold, double* seq_logL, int32_t* n_iter,
                   int32_t* converged, double* post_mean, double* post_cov);

/* hyperposterior (prediction.R:670-707): e_step on a caller-given grid (data U grid_inputs),
 * with per-row positions in THAT grid. */
int mb_hyperposterior(mb_context*, const mb_kernel* kern_0, const double* hp_0,
                      const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                      int32_t n_grid, const double* grid_X, const int32_t* grid_index,
                      const double* m_0, double pen_diag, double* post_mean, double* post_cov);

/* pred_magma (prediction.R:1146-1197) for one new task: obs n_obs x d, pred n_pred x d (both
 * row-major, Reference order), mean_obs/mean_pred and the three post_cov blocks gathered by
 * the caller from the hyper-posterior (column-major, ld = rows).  hp includes noise when
 * k->has_noise.  pred_cov may be NULL (only Mean and Var = diag + exp(noise) are formed). */
int mb_pred_magma(mb_context*, const mb_kernel* k, const double* hp, const double* x_obs, int n_obs,
                  const double* y_obs, const double* x_pred, int n_pred, int d,
                  const double* mean_obs, const double* mean_pred, const double* cov_obs,
                  const double* cov_pred, const double* cov_crossed, double pen_diag,
                  double* pred_mean, double* pred_var, double* pred_cov);

/* hyperposterior_clust (prediction.R:1691-1747): ve_step without the mixture update on a caller-given grid.
 * hp_k: K x n_hp(kern_k); m_k: K x n_grid; tau: n_tasks x K (row-major). post_mean: K x n_grid;
 * post_cov: K blocks of n_grid x n_grid. */
int mb_hyperposterior_clust(mb_context*, int n_clusters, const mb_kernel* kern_k, const double* hp_k,
                            const mb_kernel* kern_i, const double* hp_i, int hp_i_shared, int32_t n_grid,
                            const double* grid_X, const int32_t* grid_index, const double* m_k,
                            const double* tau, double pen_diag, double* post_mean, double* post_cov);

/* pred_magmaclust (prediction.R:2298-2413) for one new task with mixture probabilities proba (K): the
 * arguments of mb_pred_magma with a leading cluster dimension (mean_obs: K x n_obs, cov_obs: K blocks of
 * n_obs x n_obs column-major, ...).  pred_mean / pred_var: K x n_pred (the `pred` list), pred_cov: K blocks or
 * NULL; mix_mean / mix_var: n_pred (`mixture_pred`). */
int mb_pred_magmaclust(mb_context*, int n_clusters, const mb_kernel* k, const double* hp, const double* x_obs,
                       int n_obs, const double* y_obs, const double* x_pred, int n_pred, int d,
                       const double* mean_obs, const double* mean_pred, const double* cov_obs,
                       const double* cov_pred, const double* cov_crossed, const double* proba,
                       double pen_diag, double* pred_mean, double* pred_var, double* pred_cov,
                       double* mix_mean, double* mix_var);

/* ---- MagmaClust ------------------------------------------------------------------------ */
/* ve_step (vem-magmaclust.R:32-150): K hyper-posteriors weighted by tau (n_tasks x K,
 * row-major).  hp_k: K x n_hp(kern_k); m_k: K x U.  post_mean: K x U; post_cov: K blocks
 * of U x U.  When update != 0 the mixture is recomputed by update_mixture with prop_mixture
 * (K) and written to tau_out (n_tasks x K), else tau is copied. */
int mb_ve_step(mb_context*, int n_clusters, const mb_kernel* kern_k, const double* hp_k,
               const mb_kernel* kern_i, const double* hp_i, int hp_i_shared, const double* m_k,
               const double* tau, const double* prop_mixture, int update, double pen_diag,
               double* post_mean, double* post_cov, double* tau_out);
/* update_mixture (vem-magmaclust.R:330-396), rounded to 5 decimals like the reference. */
int mb_update_mixture(mb_context*, int n_clusters, const mb_kernel* kern_i, const double* hp_i,
                      int hp_i_shared, const double* post_mean, const double* post_cov,
                      const double* prop_mixture, double pen_diag, double* tau_out);
/* elbo_clust_multi_GP + gr_clust_multi_GP for every task (elbos.R:18-55,
 * gradients-elbos.R:17-70). */
int mb_elbo_clust_tasks(mb_context*, int n_clusters, const mb_kernel* kern_i, const double* hp_i,
                        int hp_i_shared, const double* post_mean, const double* post_cov,
                        const double* tau, double pen_diag, double* f, double* g);
/* vm_step (vem-magmaclust.R:191-303): m_k, hp_k, hp_i updated in place; prop_mixture out. */
int mb_vm_step(mb_context*, int n_clusters, const mb_kernel* kern_k, double* hp_k,
               const mb_kernel* kern_i, double* hp_i, int shared_hp_k, int shared_hp_i,
               double* m_k, const double* post_mean, const double* post_cov, const double* tau,
               double pen_diag, double* prop_mixture, int64_t* stats);
/* elbo_monitoring_VEM (elbos.R:200-272). */
int mb_elbo_monitoring(mb_context*, int n_clusters, const mb_kernel* kern_k, const double* hp_k,
                       const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                       const double* m_k, const double* post_mean, const double* post_cov,
                       const double* tau, const double* prop_mixture, double pen_diag,
                       double* elbo);
/* train_magmaclust's VEM loop (R/training.R:1586-1758) with the state resident on the device.  hp_k [K x P_k], hp_i [tasks x
 * P_i] and m_k [K x U] are in/out; ini_mixture [tasks x K]; fast_approx != 0 stops after the first VE step with the ELBO of
 * the initial hyper-parameters (training.R:1633-1647).  seq_elbo holds n_iter_max doubles; mixture_out [tasks x K],
 * prop_mixture_out [K], post_mean [K x U], post_cov [K x U x U] may be NULL.  With sharded tasks (mb_set_shard) every rank
 * passes its own rows of hp_i / ini_mixture and gets identical hp_k, m_k, ELBO sequence and step count. */
int mb_train_magmaclust(mb_context* ctx, int n_clusters, const mb_kernel* kern_k, double* hp_k, const mb_kernel* kern_i,
                        double* hp_i, int shared_hp_k, int shared_hp_i, double* m_k, const double* ini_mixture,
                        double pen_diag, int n_iter_max, double cv_threshold, int fast_approx, double* seq_elbo,
                        int32_t* n_iter, int32_t* converged, double* mixture_out, double* prop_mixture_out,
                        double* post_mean, double* post_cov);

/* ---- device-resident EM state and timing (measurement support; no reference counterpart:
 * the reference keeps hp_0 / hp_i in R tibbles between iterations, training.R:896-1030) ------ */
/* Store hp_0, the hp_i table (n_tasks x n_hp) and m_0 (U) on the device as the pristine
 * starting state of mb_em_step_resident. */
int mb_set_state(mb_context*, const mb_kernel* kern_0, const double* hp_0, const mb_kernel* kern_i,
                 const double* hp_i, const double* m_0);
/* One EM iteration exactly like mb_em_step but with the HPs taken from / left in device
 * memory; restart != 0 first restores the pristine state.  Only logL and stats come back. */
int mb_em_step_resident(mb_context*, const mb_kernel* kern_0, const mb_kernel* kern_i, int shared_hp,
                        double pen_diag, int restart, double* logL, int64_t* stats);
int mb_get_state(mb_context*, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                 double* hp_i);
/* CUDA-event timing on the library's own stream: which = 0 records the start event, 1 the stop
 * event; mb_timer_ms synchronises on the stop event and returns the elapsed device time. */
int mb_timer(mb_context*, int which);
int mb_timer_ms(mb_context*, double* ms);
/* Per-kernel event timing of the dominant kernel (k_task_objective).  enable != 0 brackets every
 * launch with events.  mb_profile_read returns out[0] = summed device ms, out[1] = launches,
 * out[2] = task evaluations (active CTAs) since the last read, and resets the counters. */
int mb_profile(mb_context*, int enable);
int mb_profile_read(mb_context*, double* out3);
/* Call BEFORE mb_profile_read: out2[0] = stream milliseconds between consecutive profiled launches inside the
 * lock-step M step (optimiser advance kernel + host round trip), out2[1] = the longer gaps (E step, monitoring). */
int mb_profile_read_gaps(mb_context*, double* out2);
/* Call AFTER mb_profile_read: out12[0..2] = host-clock milliseconds of the E step, the M step and the monitoring of the
 * profiled mb_em_step* calls (each phase ends with a synchronisation), [3] = device milliseconds of the persistent
 * M-step kernel among the profiled launches, [4..8] = stream milliseconds of the E step's segments (memsets + task
 * kernel, wait for inv_0, leaf reduction incl. the all-reduce, inverse of post_inv + mean, retry statistics), [9..11]
 * reserved.  Diagnostics only (bench.py). */
int mb_profile_read_phases(mb_context*, double* out12);
/* SM cycles the persistent per-task M-step kernel (R/em-magma.R:200-230: one optim() per task) of the last M step spent
 * in objective evaluations (out2[0]) and in L-BFGS-B steps (out2[1]), summed over its CTAs.  Environment knobs read at
 * mb_create: MB_FUSED_BELOW = number of still-active tasks below which the lock-step rounds hand over to the persistent
 * kernel (default 3 x the resident CTAs; 0 = lock-step only; a huge value = persistent kernel from the first evaluation);
 * MB_RESERVE_SMS = SMs kept free for the union-grid kernels of the concurrent hp_0 / cluster problem (default 4). */
int mb_mstep_cycles(mb_context*, double* out2);
/* Write `bytes` of device memory (L2 flush between timed iterations). */
int mb_flush_l2(mb_context*, int64_t bytes);

/* C = op(A) op(B) on host column-major matrices: R's `%*%`, crossprod(), tcrossprod() -> BLAS dgemm in the
 * reference (e.g. /root/reference/R/prediction.R:1165-1190, R/gradients-likelihoods.R:84-94).  transX != 0
 * means op(X) = t(X).  A is (transa ? k x m : m x k) with leading dimension lda, etc. */
int mb_matmul(mb_context*, int transa, int transb, int m, int n, int k, const double* A, int lda,
              const double* B, int ldb, double* C, int ldc);
/* Device-only timing of the large-matrix kernels (batched-Cholesky FP64 TFLOP/s of BASELINE.json's metric)
 * on an n-point SE + noise covariance: what = 0 product (2 n^3 flops), 1 chol() (n^3 / 3), 2 chol2inv(chol())
 * (n^3), 3 the single-CTA shared-memory chol2inv(chol()) of the union grid (n <= 224; ms_out[1..6] then receive the
 * cycles of its phases: load, Cholesky, log-det, triangular inverse, X X', store).  ms_out (8 doubles): ms_out[0] = mean
 * milliseconds per repetition, CUDA events on the library's stream. */
int mb_bench_dense(mb_context*, int what, int n, int reps, double* ms_out);
/* Same for the batched per-task path on the resident dataset (after mb_set_state): every task's
 * chol2inv(chol(K_i + jitter)) in one persistent launch, nothing written.  flops_out[0] = sum_i N_i^3. */
int mb_bench_task_inverse(mb_context*, const mb_kernel* kern_i, double pen_diag, int reps, double* ms_out,
                          double* flops_out);

/* ---- new tasks against a resident hyper-posterior --------------------------------------------------------------
 * pred_magma / pred_magmaclust process ONE new task per call in the reference and carry G x G matrices through R
 * (/root/reference/R/prediction.R:822-1250, 1883-2476).  Here the hyper-posterior evaluated on the union grid
 * (training inputs U prediction grid U new-task inputs) stays on the device, the new tasks are uploaded as the
 * resident dataset (mb_upload_dataset, grid_index = position of every observation in that union grid) and all of
 * them are trained / predicted in one call.
 *
 * mb_hyperpost_keep(ctx, 1): the next mb_hyperposterior / mb_hyperposterior_clust leaves its result resident
 * (its host outputs may then be NULL); keep = 0 stops that, keep < 0 also frees the resident copy.
 * mb_hyperpost_set: upload one (n_clusters x n_grid means, n_clusters x n_grid x n_grid covariances; post_cov may
 * be NULL = zero covariance, which turns mb_logL_GP_tasks into the terms of sum_logL_GP, R/likelihoods.R:113-129). */
int mb_hyperpost_keep(mb_context*, int keep);
int mb_hyperpost_set(mb_context*, int n_clusters, int32_t n_grid, const double* post_mean, const double* post_cov);
/* logL_GP + gr_GP (/root/reference/R/likelihoods.R:73-86, R/gradients-likelihoods.R:22-49) of every resident task
 * against cluster `cluster` of the resident hyper-posterior: cov = K(x; hp) + post_cov[t, t] inverted with
 * chol_inv_jitter, f[t] = -log N(y; mean[t], cov), g[t, p] = -1/2 tr((aa' - inv) dK/dhp_p).  g, retries may be NULL. */
int mb_logL_GP_tasks(mb_context*, const mb_kernel* kern, const double* hp, int hp_shared, int cluster,
                     double pen_diag, double* f, double* g, int32_t* retries);
/* train_gp (/root/reference/R/training.R:78-287) for every resident task at once: optim(L-BFGS-B, factr = 1e13,
 * maxit = 25) of logL_GP / gr_GP from hp[t, ] (n_tasks x n_hp, overwritten; a task whose result is not finite keeps
 * its initial values, :277-284).  stats as mb_m_step; mb_last_task_status gives convergence codes / counts. */
int mb_train_gp_tasks(mb_context*, const mb_kernel* kern, double* hp, int cluster, double pen_diag, int64_t* stats);
/* train_shared_gp (/root/reference/R/training.R:352-446), the default HP initialisation through precondition_hp
 * (R/preconditioning.R:24-39): one HP vector shared by all resident tasks, L-BFGS-B (factr = 1e13, maxit, 100 in the
 * reference) on sum_logL_GP (R/likelihoods.R:113-129) with optim's central-difference gradient (ndeps = 1e-3).
 * hp: n_hp values in/out; mean: the scalar prior mean.  stats[0] = optimiser evaluations (each 1 + 2 n_hp launches over
 * all tasks), stats[1] = optim's convergence code, stats[2] = task objective evaluations, stats[3] = 1 if the result
 * was not finite (hp is then left at its initial values).  Sharded datasets: the sums are all-reduced. */
int mb_train_shared_gp(mb_context*, const mb_kernel* kern, double* hp, double mean, double pen_diag, int maxit,
                       int64_t* stats);
/* train_gp_clust (/root/reference/R/training.R:1923-2217) for every resident task at once against the K resident cluster
 * hyper-posteriors: per task the reference's EM loop (M step optim(L-BFGS-B) of sum_logL_GP_clust / gr_sum_logL_GP_clust
 * (R/likelihoods.R:349-398, R/gradients-likelihoods.R:192-227), E step update_mixture, monitoring with prop_mixture),
 * each task stopping at its own iteration.  hp: n_tasks x n_hp in/out; prop_mixture: K; mixture_out: n_tasks x K;
 * n_iter_out / converged_out (n_tasks) and stats may be NULL. */
int mb_train_gp_clust_tasks(mb_context*, const mb_kernel* kern, double* hp, const double* prop_mixture, double pen_diag,
                            int n_iter_max, double cv_threshold, double* mixture_out, int32_t* n_iter_out,
                            int32_t* converged_out, int64_t* stats);
/* pred_magma (K = 1) / pred_magmaclust (K > 1, tau = n_tasks x K mixture probabilities of the new tasks) on n_pred
 * points for every resident task (/root/reference/R/prediction.R:1146-1197, 2319-2404): pred_mean, pred_var are
 * n_tasks x n_pred row-major (Var includes exp(noise) like the reference's `Var` column); the optional
 * pred_mean_k / pred_var_k (n_tasks x K x n_pred) receive the cluster-specific predictions.  pred_index[g] is the
 * position of prediction point g in the union grid.  retries: n_tasks x K (may be NULL).  Only the diagonal of the
 * predictive covariance is formed; the full matrix of one task is mb_pred_magma's business. */
int mb_pred_magma_tasks(mb_context*, const mb_kernel* kern, const double* hp, const double* tau, int32_t n_pred,
                        const double* x_pred, const int32_t* pred_index, double pen_diag, double* pred_mean,
                        double* pred_var, double* pred_mean_k, double* pred_var_k, int32_t* retries);

/* Device-only timing of the HBM-bound kernels of the path on the resident dataset (after mb_set_state):
 * what = 0 list_kern_to_cov (/root/reference/R/kernel-to-matrices.R:586-652), every task's covariance matrix
 * written to HBM; what = 1 the E step's deterministic reduction of the per-CTA partial precision sums
 * (/root/reference/R/em-magma.R:39-46).  bytes_out[0] = algorithmic bytes per repetition. */
int mb_bench_hbm(mb_context*, int what, const mb_kernel* kern_i, int reps, double* ms_out, double* bytes_out);

/* ---- host data path (no GPU involved): what train_magma / train_magmaclust / pred_* do to `data` before any arithmetic
 * (/root/reference/R/training.R:690-720, R/prediction.R:534-567,937-976, R/em-magma.R:27-30), as an R-free routine.
 * mb_r_signif = signif(x, digits) (R nmath/fprec.c); mb_r_as_character = as.character(<double>) (15 significant digits,
 * fixed unless scientific is shorter); mb_reference_key = the `Reference` string of one input row
 * (paste(signif(row, digits), collapse = ":")).
 * mb_prepare_rows: X is n_rows x d row-major (Input, covariates), ids the ID strings, y the outputs (may be NULL);
 * grid_extra (n_extra x d, may be NULL) are additional grid inputs (prediction grid / new-task inputs) to include in the
 * union grid.  Outputs (caller-allocated): perm[n_rows] = original row of every output row (rows ordered by ID string then
 * Reference string, byte-wise like dplyr::arrange); X_out / y_out the rounded, permuted table; offs_out[n_tasks + 1]
 * (capacity n_rows + 1); grid_index_out[n_rows]; grid_X_out (capacity (n_rows + n_extra) x d) the union grid in
 * Reference order.  Returns 3 when a task has two rows on the same grid point (the reference's stop(), training.R:703-714). */
double mb_r_signif(double x, int digits);
int mb_r_as_character(double x, char* out, int cap);
int mb_reference_key(const double* row, int d, int digits, char* out, int cap);
int mb_prepare_rows(int64_t n_rows, int d, const double* X, const double* y, const char* const* ids, int digits,
                    const double* grid_extra, int64_t n_extra, int64_t* perm, double* X_out, double* y_out,
                    int64_t* n_tasks_out, int64_t* offs_out, int32_t* grid_index_out, int32_t* n_grid_out,
                    double* grid_X_out);

/* ---- optimiser, exposed for tests: optim(method="L-BFGS-B") on a built-in test function
 * (0 = Rosenbrock n=2), run by the same state machine the M step uses (host build). */
int mb_lbfgsb_selftest(int which, double* x, int n, double factr, int maxit, double* fval,
                       int32_t* n_eval, int32_t* fail);

#ifdef __cplusplus
}
#endif
#endif /* MAGMA_B200_H */
