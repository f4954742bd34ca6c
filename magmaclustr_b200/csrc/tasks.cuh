// tasks.cuh -- argument block of the per-task fused kernels (one CTA per task, the task's
// matrices resident in shared memory).
#pragma once
#include "common.cuh"
#include "lbfgsb.h"

#define TK_MAXN 96        /* largest task handled by the shared-memory path */
#define TK_MAX_RETRY 40   /* jitter escalations before giving up (R has no cap) */

enum {
  TK_OBJ_MOD = 0,    // logL_GP_mod / gr_GP_mod            likelihoods.R:155-174
  TK_OBJ_ELBO = 1,   // elbo_clust_multi_GP / gr_clust_..  elbos.R:18-55
  TK_OBJ_MARG = 2,   // logL_GP / gr_GP (K + S inverted)   likelihoods.R:73-86
  TK_OBJ_MIX = 3,    // update_mixture: K values of -logL_GP_mod per task
  TK_OBJ_MARG_MIX = 4  // sum_logL_GP_clust / gr_sum_logL_GP_clust: sum_k tau_k logL_GP(.; mean_k, cov_k)   likelihoods.R:349-398
};

struct TaskData {          // resident dataset (device pointers)
  const int64_t* offs;     // n_tasks + 1
  const double* X;         // rows x D, row-major
  const double* y;         // rows
  const int32_t* gidx;     // rows: position in the union grid
  int D;
  int64_t n_tasks;
  int nmax;                // largest task
  // Pair tables (pairs3.cu, built once per upload; NULL = not built).  A stationary kernel sees a pair of points only through
  // sum_q (x_rq - x_cq)^2, and tasks observed on a common grid have few DISTINCT values of it (C2: ~385 for 2080 pairs), so
  // the SE fast path evaluates exp() once per distinct value and looks every matrix element up -- the same expression on
  // the same operand, hence the same bits as the element-by-element evaluation.
  const double* pair_d2 = nullptr;         // n_tasks x MB_PAIR_CAP: the task's distinct squared distances
  const unsigned short* pair_idx = nullptr;// n_tasks x pair_stride: per element of the packed upper tiles (tile, lane, 2: the layout of
                                 // the DMMA accumulator fragments) its index into pair_d2; bit 15 = the reference's
                                 // coincidence test sum_q |x_rq - x_cq| < 1e-6 (noise term, kernel-to-matrices.R:481-487)
  const int32_t* pair_n = nullptr;         // n_tasks: number of distinct values; -1 = more than MB_PAIR_CAP or NaN inputs (direct path)
  int pair_stride = 0;               // 16-bit entries between tasks = 64 x packed upper tiles of the 8- or 12-tile layout
};
#define MB_PAIR_CAP 1024

struct HyperPost {         // hyper-posterior on the U grid (K = 1 for Magma)
  int K;
  int U;
  const double* mean;      // K x U
  const double* cov;       // K x U x U (symmetric, so storage order is irrelevant)
  const double* tau;       // n_tasks x K row-major (NULL = weight 1)
};

struct TaskObjArgs {
  TaskData db;
  HyperPost hpst;
  DevKern kd;
  const double* hp;          // HP table, or NULL when lb != NULL
  int64_t hp_stride;         // doubles between tasks (0 = shared)
  const LbfgsbState* lb;     // device optimiser states: HPs are lb[t].x
  const int32_t* active;     // task ids to evaluate (NULL = all), n_active entries
  int mode;
  int want_grad;
  int logdet_chol;           // 0: LU of the explicit inverse (reference), 1: Cholesky diagonal
  double pen;
  double* f;                 // n_tasks (TK_OBJ_MIX: n_tasks x K)
  double* g;                 // n_tasks x n_hp
  int32_t* retries;          // n_tasks (may be NULL)
  int ld;                    // shared-memory leading dimension (odd, >= nmax)
  // optional scratch (n_tasks x kcache_stride doubles): the SE kernel values computed while K is built are kept
  // here (L2 / HBM) and re-read by the gradient contraction instead of evaluating exp() a second time
  double* kcache;
  int64_t kcache_stride;
  double* fk;                // TK_OBJ_MARG_MIX: optional n_tasks x K table of the per-cluster values logL_GP(.; k)
};

struct TaskInvArgs {       // E step / list_kern_to_inv
  TaskData db;
  DevKern kd;
  const double* hp;
  int64_t hp_stride;
  double pen;
  int K;                     // clusters (weights tau), 1 for Magma
  const double* tau;         // n_tasks x K or NULL
  int U;
  // E step accumulators, one per LEAF (contiguous task range): leaf l adds into part_inv + l * part_stride (K x U x U
  // precision sums) and part_w + l * part_stride (K x U weighted sums), in task order.  NULL = no accumulation.
  double* part_inv;
  double* part_w;
  int64_t part_stride;       // doubles between the accumulators of consecutive leaves
  const int64_t* leaf_offs;  // device, n_leaves + 1: first task of every leaf
  int n_leaves;
  int leaf_max;              // tasks in the longest leaf
  int* queue;                // device, zero-initialised work counter: item g = (leaf g % n_leaves, position g / n_leaves)
  int* tickets;              // device, n_leaves, zero-initialised: tasks of the leaf already added
  double* inv_out;           // optional: all inverses, block t at inv_offs[t], column-major
  const int64_t* inv_offs;
  int32_t* retries;
  int ld;
  int diag_skip;             // diagnostics only (MB_ESTEP_DIAG): 1 = no ticket ordering, 2 = no accumulation (timing decomposition)
  int sm_limit;              // > 0 with a work queue: CTAs that land on an SM with %smid >= sm_limit leave at once (SMs kept
                             // free for the union-grid kernels of inv_0 running on the other stream); 0 = use every SM
};

struct TaskPredArgs {      // pred_magma(clust) batched over the resident (new) tasks
  TaskData db;
  HyperPost hpst;            // resident hyper-posterior on the union grid (K clusters); tau unused here
  DevKern kd;
  const double* hp;          // n_tasks x n_hp
  int64_t hp_stride;
  const double* tau;         // n_tasks x K mixture probabilities of the new tasks (K > 1)
  double pen;
  int64_t t0;                // first task of this launch (the scratch table is indexed from t0)
  int n_pred;
  const double* x_pred;      // n_pred x D
  const int32_t* pred_index; // n_pred: position of every prediction point in the union grid
  double* ginv;              // scratch: launch tasks x K x (nmax^2 + nmax)
  double* out_mean;          // n_tasks x n_pred
  double* out_var;
  double* out_mean_k;        // optional n_tasks x K x n_pred
  double* out_var_k;
  int32_t* retries;          // n_tasks x K
  int ld;
};
