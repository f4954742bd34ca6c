// api.cu -- context, device workspaces and the extern "C" entry points of include/magma_b200.h.
// Host code here only orchestrates: the matrices, likelihoods and gradients are produced by the kernels of
// tasks*.cu, dense*.cu and lbfgsb_dev.cu; what runs on the host is O(M K) scalar glue in the reference's own
// order (softmax + round(5) of K numbers per task, colMeans, the mixture combination of K prediction
// vectors) and the L-BFGS-B state machines of the single mean / cluster problems.
// There is no CPU fallback: without a CUDA device every call fails.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <chrono>
#include <vector>

#include <dlfcn.h>
#include <nccl.h>   // types only: the library is resolved with dlopen at mb_nccl_init (no link-time dependency)

#include "common.cuh"
#include "lbfgsb.h"
#include "tasks.cuh"
#include "dense_big.cuh"

// ---- launchers defined in the other translation units ------------------------------------
cudaError_t launch_task_objective(TaskObjArgs a, int n_launch, cudaStream_t st);
cudaError_t launch_task_cov(TaskData db, DevKern kd, const double* hp, int64_t hp_stride,
                            int deriv, double* out, const int64_t* out_offs, cudaStream_t st);
cudaError_t launch_task_cov_sym(TaskData db, DevKern kd, const double* hp, int64_t hp_stride, int deriv, double* out,
                                const int64_t* out_offs, int sm_count, cudaStream_t st);
cudaError_t launch_task_pred(TaskPredArgs a, int n_launch_tasks, cudaStream_t st);
cudaError_t launch_task_objective3(TaskObjArgs a, int n_launch, cudaStream_t st);
cudaError_t launch_task_optimise3(TaskObjArgs a, int* queue, int sm_count, int reserve_sms, unsigned long long* cycles,
                                  cudaStream_t st);
cudaError_t launch_task_inverse3(TaskInvArgs a, int n_ctas, cudaStream_t st);
cudaError_t launch_task_pairs3(TaskData db, double* pair_d2, unsigned short* pair_idx, int32_t* pair_n, int pair_stride,
                               cudaStream_t st);
int task3_max_ctas_per_sm(int nmax, int D);

struct BigTaskPool;
void big_pool_destroy(BigTaskPool* p);
int big_tasks_objective(BigTaskPool** pp, const TaskObjArgs& a, const int64_t* h_offs, cudaStream_t st, int64_t* launches);
int big_tasks_inverse(BigTaskPool** pp, const TaskInvArgs& a, const int64_t* h_offs, const int64_t* h_inv_offs,
                      const int64_t* h_leaf_offs, cudaStream_t st, int64_t* launches);

// marginal-likelihood modes (new tasks, small N_obs) run the scalar shared-memory kernel of tasks.cu; everything else the
// DMMA kernels of tasks3.cu
static cudaError_t launch_task_objective_any(TaskObjArgs a, int n_launch, cudaStream_t st) {
  if (a.mode == TK_OBJ_MARG || a.mode == TK_OBJ_MARG_MIX) return launch_task_objective(a, n_launch, st);
  return launch_task_objective3(a, n_launch, st);
}
static cudaError_t launch_task_inverse_any(TaskInvArgs a, int n_ctas, cudaStream_t st) { return launch_task_inverse3(a, n_ctas, st); }
static int task_ctas_per_sm_any(int nmax, int D) { return task3_max_ctas_per_sm(nmax, D); }
cudaError_t launch_dense_cov(const DevKern& kd, const double* hp_host, const double* x1, int n1,
                             const double* x2, int n2, int D, int deriv, int with_noise,
                             double* out, int64_t rs, int64_t cs, cudaStream_t st);
cudaError_t launch_legacy_builder(int which, const double* m1, int n1, const double* m2, int n2,
                                  int d, double par, double* out, cudaStream_t st);
bool dense_sm_fits(int n);
cudaError_t launch_dense_cholinv_sm(const double* src, int ld_src, double* inv, int n, double pen, int mode,
                                    int* info, double* out, double* xt, cudaStream_t st);
cudaError_t launch_gemm_dmma(int m, int n, int k, const double* A, int64_t ars, int64_t acs, const double* B,
                             int64_t brs, int64_t bcs, double* C, int64_t ldc, double alpha, const double* Cin,
                             double beta, cudaStream_t st);
cudaError_t launch_resid_gemv(int n, const double* A, const double* y, const double* mean, double* r, double* al,
                              cudaStream_t st);
// chol_inv_jitter of a device matrix: shared-memory-resident kernel when it fits (U <= 224), else the
// multi-CTA blocked path of dense_big.cu (MB_DENSE_KERNEL=big forces the latter: used by the tests to cross-check)
static int dense_env() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("MB_DENSE_KERNEL"); v = (e && e[0] == 'b') ? 2 : 1; }
  return v;
}
static bool dense_use_sm(int n) { return dense_env() == 1 && dense_sm_fits(n); }
static bool dense_use_big(int n) { return !dense_use_sm(n); }
cudaError_t launch_dense_lu_logdet(const double* src, double* W, int n, int* done, double* out,
                                   cudaStream_t st);
cudaError_t launch_gemv(int m, int n, const double* A, int64_t ld, const double* x,
                        const double* y0, double* y, cudaStream_t st);
int dense_obj_nparts(int n);
cudaError_t launch_dense_obj(const DevKern& kd, const double* hp_host, const double* x, int n,
                             int D, const double* A, const double* S, const double* T,
                             const double* al, const double* r, const double* logdet,
                             int from_chol, int want_grad, double* part, double* res,
                             cudaStream_t st);
cudaError_t launch_mirror_upper(double* m, int U, int nmat, int64_t stride, cudaStream_t st);
cudaError_t launch_reduce_tree(int64_t count, int nparts, const double* base, const double* part, int64_t part_stride,
                               double* out, cudaStream_t st);
cudaError_t launch_leaf_sums(const double* v, int P, const int64_t* leaf_offs, int n_leaves, double* out, cudaStream_t st);
cudaError_t launch_retry_stats(const int32_t* retries, int64_t n, int* out2, cudaStream_t st);
cudaError_t launch_sum_sequential(const double* v, int64_t n, int64_t stride, double* out,
                                  cudaStream_t st);
cudaError_t launch_colsum_sequential(const double* v, int64_t n, int P, double* out,
                                     cudaStream_t st);
cudaError_t launch_axpby(int64_t n, double a, const double* x, double b, const double* y,
                         double* out, cudaStream_t st);
cudaError_t launch_fill(int64_t n, double v, double* out, cudaStream_t st);
cudaError_t launch_diag(int n, const double* M, int64_t ld, double add, double* out,
                        cudaStream_t st);
cudaError_t launch_lbfgsb_init(LbfgsbState* st, int64_t m, int n, const double* x0,
                               int64_t x0_stride, double factr, double pgtol, int maxit,
                               cudaStream_t s);
cudaError_t launch_lbfgsb_advance(LbfgsbState* st, int64_t m, int n, const double* f,
                                  const double* g, int* n_active, cudaStream_t s);
cudaError_t launch_lbfgsb_mask(LbfgsbState* st, int64_t m, const int32_t* skip, cudaStream_t s);
cudaError_t launch_lbfgsb_export(const LbfgsbState* st, int64_t m, int n, double* x,
                                 int64_t x_stride, int* fail, int* nfgv, double* fval, cudaStream_t s);

// ---- errors -------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
void mb_set_error(const char* file, int line, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s:%d: %s", file, line, msg);
}
extern "C" const char* mb_last_error(void) { return g_err; }
#define MB_FAIL(code, msg) do { mb_set_error(__FILE__, __LINE__, msg); return code; } while (0)
#define MB_TRY(call) do { int r_ = (call); if (r_ != 0) return r_; } while (0)

// ---- device buffers -------------------------------------------------------------------------
struct DBuf {
  void* p = nullptr;
  size_t cap = 0;
  int ensure(size_t bytes) {
    if (bytes <= cap) return 0;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = bytes + bytes / 8 + 256;
    if (cudaMalloc(&p, want) != cudaSuccess) { cudaGetLastError(); mb_set_error(__FILE__, __LINE__, "cudaMalloc failed"); return -3; }
    cap = want;
    return 0;
  }
  template <class T> T* as() { return (T*)p; }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DenseWs {   // workspace of one dense n x n problem
  DBuf Kmat, A, B, T, S, Ap, Bp, vecs, part, res, info, done;
  int n = 0;
  double* r() { return vecs.as<double>(); }
  double* al() { return vecs.as<double>() + n; }
  double* tmp() { return vecs.as<double>() + 2 * (size_t)n; }
  int ensure(int nn) {
    n = nn;
    size_t m = (size_t)nn * nn * sizeof(double);
    if (Kmat.ensure(m) || A.ensure(m) || B.ensure(m) || T.ensure(m) || S.ensure(m)) return -3;
    const size_t npad = (size_t)((nn + 7) / 8) * 8;
    if (Ap.ensure(npad * npad * sizeof(double)) || Bp.ensure(npad * npad * sizeof(double))) return -3;
    if (vecs.ensure(4 * (size_t)nn * sizeof(double))) return -3;
    if (part.ensure((size_t)dense_obj_nparts(nn) * (1 + MB_MAX_HP) * sizeof(double))) return -3;
    if (res.ensure(32 * sizeof(double)) || info.ensure(8 * sizeof(int))) return -3;
    if (done.ensure((size_t)nn * sizeof(int))) return -3;
    return 0;
  }
  void release() { Kmat.release(); A.release(); B.release(); T.release(); S.release(); Ap.release(); Bp.release(); vecs.release();
                   part.release(); res.release(); info.release(); done.release(); }
};

struct mb_context {
  int device = 0;
  int sm_count = 148;
  cudaStream_t st = nullptr, st2 = nullptr, st3 = nullptr;   // st3: device-to-host copies of the round counters
  int64_t launches = 0;
  int logdet_chol = 0;
  // per-task M step: lock-step rounds (one objective launch + one optimiser launch over all active tasks) while more than
  // fused_below problems are still active, then the persistent kernel of mstep3.cu finishes every remaining trajectory.
  // fused_below: -1 = 3 x the resident CTAs (default), 0 = lock-step only, a huge value = persistent kernel from the start
  int64_t fused_below = -1;
  int reserve_sms = 4;
  bool single_blocking = false, split_lauum = true;
  unsigned long long h_cycles[2] = {0, 0};
  // resident dataset
  bool have_db = false;
  TaskData db;
  int64_t n_rows = 0;
  int U = 0;
  DBuf d_offs, d_X, d_y, d_gidx, d_gridX;
  // pair tables of the resident tasks (pairs3.cu), built by mb_upload_dataset for tasks of at most TK_MAXN points;
  // option "pair_tables" / MB_PAIR_TABLES=0 keeps the per-task kernels on the element-by-element evaluation (same bits)
  DBuf d_pair_d2, d_pair_idx, d_pair_n;
  bool pair_tables = true, pairs_built = false;
  int pair_stride = 0;
  std::vector<int64_t> h_offs;
  // workspaces
  DenseWs ws0, ws1;                 // ws0: mean/cluster process on st2 ; ws1: generic on st
  DBuf hp_i, hp_k, f_t, g_t, retries, lb, counters, sums, pm, pc, m0, tau,
      inv0, post_inv, fails, nfgv, scratch, kcache;
  double* h_pin = nullptr;          // pinned host scratch (256 doubles)
  int* h_pin_i = nullptr;           // pinned host ints (1024)
  // resident EM state (mb_set_state): pristine + working HP tables
  DBuf st_hp_i0, st_hp_i;
  double st_hp_00[MB_MAX_HP], st_hp_0[MB_MAX_HP];
  bool have_state = false;
  // resident hyper-posterior on a union grid (mb_hyperpost_set / mb_hyperpost_keep): K x U means, K x U x U covariances
  DBuf hpr_mean, hpr_cov;
  int hpr_K = 0, hpr_U = 0;
  bool hpr_keep = false, hpr_has_cov = false;
  // timing / profiling
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_single = nullptr;
  cudaEvent_t ev_ring[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  cudaEvent_t ev_cp[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // after a per-task M step of train_magma: f_t holds logL_GP_mod of every task at its new HPs and m_step_f0 the mean
  // process's value at the new hp_0 -- exactly what logL_monitoring evaluates next (same functions, same hyper-posterior)
  bool mstep_values_valid = false;
  double m_step_f0 = 0.0;
  // same after a VM step of train_magmaclust: the task ELBO terms (per task in f_t, or their all-rank sum for a shared
  // HP vector) and the cluster terms, as the optimisers left them
  bool vm_valid = false, vm_tasks_in_f_t = false;
  double vm_sum_ll_i = 0.0, vm_sum_ll_k = 0.0;
  bool monitor_recompute = false;   // MB_MONITOR_RECOMPUTE / "monitor_recompute": logL_monitoring re-evaluates everything
  bool prof_on = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_ev;
  size_t prof_used = 0;
  double prof_ms = 0.0, prof_gap_ms = 0.0, prof_other_ms = 0.0;
  // wall-clock phases of dev_em_step while profiling: [0] E step, [1] M step, [2] monitoring, [3] persistent M-step kernel (device)
  double prof_phase[12] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  cudaEvent_t ev_mark[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};   // E-step segment marks on st (mb_profile)
  std::vector<char> prof_kind;     // per profiled launch: 0 = lock-step objective launch, 1 = persistent M-step kernel
  int64_t prof_launches = 0, prof_evals = 0;
  // tasks with more than TK_MAXN points: multi-stream pipelines of the large-matrix kernels (bigtasks.cu)
  BigTaskPool* big = nullptr;
  // multi-GPU: NCCL communicator owned by the context (mb_nccl_init), or a caller-supplied hook
  ncclComm_t nccl = nullptr;
  mb_allreduce_fn ar = nullptr;
  void* ar_user = nullptr;
  int rank = 0, world = 1;
  // this rank's slice of the GLOBAL task list (mb_set_shard; the whole list when never set) and the leaves it owns
  int64_t g_first = 0, g_total = 0;
  bool shard_set = false;
  int lv_NL = 0, lv_l0 = 0, lv_n = 0, lv_U = -1;
  int64_t lv_first = -1, lv_total = -1, lv_ntasks = -1;
  std::vector<int64_t> lv_offs;
  DBuf d_lv_offs, tickets, leafbuf, slotbuf, leafsum;
};

#define CU(call) MB_CUDA_OK(call)
#define LAUNCH(ctx, call) do { CU(call); (ctx)->launches++; } while (0)

// launch of the dominant kernel, optionally bracketed by events (mb_profile)
static int launch_obj(mb_context* c, const TaskObjArgs& a, int n_launch, cudaStream_t st,
                      bool profiled = false) {
  if (a.db.nmax > TK_MAXN) return big_tasks_objective(&c->big, a, c->h_offs.data(), st, &c->launches);
  TaskObjArgs a2 = a;
  if (a.want_grad && a.mode != TK_OBJ_MARG && a.mode != TK_OBJ_MARG_MIX) {
    const int n8 = (a.db.nmax + 7) / 8;
    a2.kcache_stride = (int64_t)(a.db.nmax <= 64 ? 36 : 78) * 64;   // packed upper tiles of the 8 / 12-tile layouts
    (void)n8;
    MB_TRY(c->kcache.ensure((size_t)a.db.n_tasks * a2.kcache_stride * 8));
    a2.kcache = c->kcache.as<double>();
  }
  if (!c->prof_on || !profiled) { LAUNCH(c, launch_task_objective_any(a2, n_launch, st)); return 0; }
  if (c->prof_used == c->prof_ev.size()) {
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0));
    CU(cudaEventCreate(&e1));
    c->prof_ev.push_back({e0, e1});
  }
  if (c->prof_kind.size() <= c->prof_used) c->prof_kind.resize(c->prof_used + 1);
  c->prof_kind[c->prof_used] = 0;
  auto& ev = c->prof_ev[c->prof_used++];
  CU(cudaEventRecord(ev.first, st));
  LAUNCH(c, launch_task_objective_any(a2, n_launch, st));
  CU(cudaEventRecord(ev.second, st));
  return 0;
}
static int prof_collect(mb_context* c) {
  for (size_t i = 0; i < c->prof_used; ++i) {
    float ms = 0;
    CU(cudaEventSynchronize(c->prof_ev[i].second));
    CU(cudaEventElapsedTime(&ms, c->prof_ev[i].first, c->prof_ev[i].second));
    c->prof_ms += ms;
    c->prof_launches++;
    if (i < c->prof_kind.size() && c->prof_kind[i] == 1) c->prof_phase[3] += ms;
    if (i + 1 < c->prof_used) {   // stream time between two profiled launches: advance kernel + host round trip
      float gap = 0;
      CU(cudaEventSynchronize(c->prof_ev[i + 1].first));
      CU(cudaEventElapsedTime(&gap, c->prof_ev[i].second, c->prof_ev[i + 1].first));
      if (gap < 5.0f) c->prof_gap_ms += gap;   // longer gaps are step boundaries (E step, monitoring)
      else c->prof_other_ms += gap;
    }
  }
  c->prof_used = 0;
  return 0;
}

// NCCL entry points, resolved once from libnccl.so.2 (the copy already mapped into the process when the host program
// uses one, e.g. torch's; the system library otherwise).
struct NcclApi {
  void* h = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
      if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce || !api.CommDestroy) api.h = nullptr;
    }
  }
  return api.h ? &api : nullptr;
}

extern "C" int mb_nccl_unique_id(char* id_out, int bytes) {
  if (!id_out || bytes < (int)sizeof(ncclUniqueId)) MB_FAIL(-1, "id buffer must hold 128 bytes");
  NcclApi* n = nccl_api();
  if (!n) MB_FAIL(-5, "libnccl.so.2 not found");
  ncclUniqueId id;
  if (n->GetUniqueId(&id) != ncclSuccess) MB_FAIL(-5, "ncclGetUniqueId failed");
  memcpy(id_out, &id, sizeof(id));
  return 0;
}

extern "C" int mb_nccl_init(mb_context* c, const char* id, int bytes, int rank, int world) {
  if (!c || !id || bytes < (int)sizeof(ncclUniqueId) || world < 1 || rank < 0 || rank >= world) MB_FAIL(-1, "bad argument");
  NcclApi* n = nccl_api();
  if (!n) MB_FAIL(-5, "libnccl.so.2 not found");
  CU(cudaSetDevice(c->device));
  if (c->nccl) { n->CommDestroy(c->nccl); c->nccl = nullptr; }
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  ncclResult_t r = n->CommInitRank(&c->nccl, world, uid, rank);
  if (r != ncclSuccess) { c->nccl = nullptr; MB_FAIL(-5, n->GetErrorString ? n->GetErrorString(r) : "ncclCommInitRank failed"); }
  c->rank = rank; c->world = world;
  return 0;
}

// In-place sum over the ranks of `count` doubles at `dev`: on the context's own communicator the all-reduce is
// enqueued on `s` like any kernel (no host synchronisation); the caller-supplied hook needs the data complete first.
static int ctx_allreduce(mb_context* c, void* dev, int64_t count, cudaStream_t s) {
  if (c->world <= 1) return 0;
  if (c->nccl) {
    ncclResult_t r = nccl_api()->AllReduce(dev, dev, (size_t)count, ncclDouble, ncclSum, c->nccl, s);
    if (r != ncclSuccess) MB_FAIL(-5, "ncclAllReduce failed");
    return 0;
  }
  if (!c->ar) MB_FAIL(-5, "world > 1 but neither mb_nccl_init nor an mb_set_allreduce hook is configured");
  CU(cudaStreamSynchronize(s));
  if (c->ar(c->ar_user, dev, count) != 0) MB_FAIL(-5, "allreduce hook failed");
  return 0;
}

// ================================================================================================
// Leaves: the fixed partition of the GLOBAL task list behind every sum over tasks.
// The reference adds the tasks sequentially (em-magma.R:39-46, likelihoods.R:253-312).  Here a sum over tasks is
//   (1) a sequential (E step) or compensated (scalars) sum over the tasks of each LEAF, in task order,
//   (2) a fixed binary tree over the NL leaves (dense.cu:k_reduce_tree).
// NL and the leaf boundaries depend on the global task count and the grid size only, and every rank owns whole leaves, so
// the result has the same bits on 1, 2, 4, 8 ... GPUs (and for any SM count / launch geometry).  Across ranks the leaf (or
// subtree) sums are exchanged with an all-reduce of a zero-padded slot buffer -- x + 0 is exact, so this is a gather whatever
// order the collective adds in -- and every rank finishes the same tree.
// ================================================================================================
static int leaf_count(int64_t g_total, int U) {
  const int64_t per = ((int64_t)U * U + U) * 8;
  int64_t cap = per > 0 ? ((int64_t)1 << 31) / per : 128;
  if (cap < 1) cap = 1;
  int64_t m = 128;
  if (g_total < m) m = g_total;
  if (cap < m) m = cap;
  int nl = 1;
  while ((int64_t)nl * 2 <= m) nl *= 2;
  return nl;
}
static int64_t leaf_start(int64_t l, int nl, int64_t g_total) { return (l * g_total + nl - 1) / nl; }

extern "C" int mb_estep_leaves(int64_t n_tasks_global, int32_t n_grid) {
  if (n_tasks_global < 1 || n_grid < 0) return -1;
  return leaf_count(n_tasks_global, n_grid);
}

extern "C" int mb_shard_range(int64_t n_tasks_global, int32_t n_grid, int rank, int world, int64_t* first, int64_t* last) {
  if (n_tasks_global < 1 || world < 1 || rank < 0 || rank >= world || !first || !last) MB_FAIL(-1, "bad argument");
  const int nl = leaf_count(n_tasks_global, n_grid);
  if (nl < world) MB_FAIL(-1, "more ranks than E-step leaves: use fewer GPUs for this dataset");
  *first = leaf_start((int64_t)rank * nl / world, nl, n_tasks_global);
  *last = leaf_start((int64_t)(rank + 1) * nl / world, nl, n_tasks_global);
  return 0;
}

extern "C" int mb_set_shard(mb_context* c, int64_t first_task_global, int64_t n_tasks_global) {
  if (!c) MB_FAIL(-1, "null ctx");
  if (first_task_global < 0 || n_tasks_global < 1 || first_task_global >= n_tasks_global) MB_FAIL(-1, "bad shard");
  c->g_first = first_task_global; c->g_total = n_tasks_global; c->shard_set = true;
  c->lv_U = -1;
  return 0;
}

// leaves of the resident tasks for a grid of U points (cached)
static int ctx_leaves(mb_context* c, int U) {
  const int64_t M = c->db.n_tasks;
  const int64_t first = c->shard_set ? c->g_first : 0, total = c->shard_set ? c->g_total : M;
  if (c->world > 1 && !c->shard_set) MB_FAIL(-1, "world > 1: call mb_set_shard (global position of this rank's tasks) first");
  if (first + M > total) MB_FAIL(-1, "mb_set_shard: the resident tasks run past the global task count");
  if (c->world == 1 && M != total) MB_FAIL(-1, "mb_set_shard: a partial shard needs world > 1 (mb_nccl_init / mb_set_allreduce)");
  if (c->lv_U == U && c->lv_first == first && c->lv_total == total && c->lv_ntasks == M) return 0;
  const int nl = leaf_count(total, U);
  int l0 = -1, l1 = -1;
  for (int l = 0; l <= nl; ++l) {
    const int64_t s = leaf_start(l, nl, total);
    if (s == first && l0 < 0) l0 = l;
    if (s == first + M) l1 = l;
  }
  if (l0 < 0 || l1 <= l0) MB_FAIL(-1, "shard boundaries must coincide with E-step leaf boundaries (mb_shard_range)");
  if (c->world > 1 && (c->world & (c->world - 1)) == 0 && nl % c->world == 0 &&
      (l0 != c->rank * (nl / c->world) || l1 - l0 != nl / c->world))
    MB_FAIL(-1, "with a power-of-two world every rank must hold the shard mb_shard_range gives it");
  c->lv_offs.resize(l1 - l0 + 1);
  for (int l = l0; l <= l1; ++l) c->lv_offs[l - l0] = leaf_start(l, nl, total) - first;
  MB_TRY(c->d_lv_offs.ensure(c->lv_offs.size() * 8));
  CU(cudaMemcpyAsync(c->d_lv_offs.p, c->lv_offs.data(), c->lv_offs.size() * 8, cudaMemcpyHostToDevice, c->st));
  CU(cudaStreamSynchronize(c->st));
  c->lv_NL = nl; c->lv_l0 = l0; c->lv_n = l1 - l0; c->lv_U = U;
  c->lv_first = first; c->lv_total = total; c->lv_ntasks = M;
  return 0;
}

// out[0..count) = base + (fixed tree over ALL leaves of the global task list) of the per-leaf vectors vals[l * stride ..],
// l < lv_n local leaves.  Needs ctx_leaves().  Everything is enqueued on s (the hook path synchronises inside).
static int leaf_reduce(mb_context* c, const double* vals, int64_t stride, int64_t count, const double* base, double* out,
                       cudaStream_t s) {
  if (c->world <= 1) {
    LAUNCH(c, launch_reduce_tree(count, c->lv_n, base, vals, stride, out, s));
    return 0;
  }
  const bool subtree = (c->world & (c->world - 1)) == 0 && c->lv_NL % c->world == 0;
  const int slots = subtree ? c->world : c->lv_NL;
  MB_TRY(c->slotbuf.ensure((size_t)slots * count * 8));
  double* sb = c->slotbuf.as<double>();
  CU(cudaMemsetAsync(sb, 0, (size_t)slots * count * 8, s));
  if (subtree) {
    LAUNCH(c, launch_reduce_tree(count, c->lv_n, nullptr, vals, stride, sb + (size_t)c->rank * count, s));
  } else {
    CU(cudaMemcpy2DAsync(sb + (size_t)c->lv_l0 * count, count * 8, vals, stride * 8, count * 8, c->lv_n,
                         cudaMemcpyDeviceToDevice, s));
  }
  MB_TRY(ctx_allreduce(c, sb, (int64_t)slots * count, s));
  LAUNCH(c, launch_reduce_tree(count, slots, base, sb, count, out, s));
  return 0;
}

// out[0..P) = sum over ALL tasks (every rank) of the columns of the row-major table v (local tasks x P), device pointers
static int leaf_sum_columns(mb_context* c, const double* v, int P, double* out, cudaStream_t s) {
  MB_TRY(ctx_leaves(c, c->U));
  MB_TRY(c->leafsum.ensure((size_t)c->lv_n * P * 8));
  LAUNCH(c, launch_leaf_sums(v, P, c->d_lv_offs.as<int64_t>(), c->lv_n, c->leafsum.as<double>(), s));
  return leaf_reduce(c, c->leafsum.as<double>(), P, P, nullptr, out, s);
}

// same for a host table: out_host[0..P) = column sums over the tasks of ALL ranks (vem-magmaclust.R:265-266, elbos.R:253)
static int leaf_sum_columns_host(mb_context* c, const double* tab, int P, double* out_host) {
  const size_t bytes = (size_t)c->db.n_tasks * P * 8;
  MB_TRY(c->scratch.ensure(bytes + 64 * 8));
  double* d_tab = c->scratch.as<double>();
  double* d_out = d_tab + (size_t)c->db.n_tasks * P;
  CU(cudaMemcpyAsync(d_tab, tab, bytes, cudaMemcpyHostToDevice, c->st));
  MB_TRY(leaf_sum_columns(c, d_tab, P, d_out, c->st));
  CU(cudaMemcpyAsync(out_host, d_out, (size_t)P * 8, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// Collective agreement on a status: every rank passes its own return code, all get a non-zero code when any rank failed,
// so nobody is left waiting in the next all-reduce.
static int ctx_agree(mb_context* c, int local_rc, cudaStream_t s) {
  if (c->world <= 1) return local_rc;
  if (c->sums.ensure(64 * 8)) return -3;
  const double flag = local_rc ? 1.0 : 0.0;
  if (cudaMemcpyAsync(c->sums.as<double>() + 32, &flag, 8, cudaMemcpyHostToDevice, s) != cudaSuccess) return -2;
  if (cudaStreamSynchronize(s) != cudaSuccess) return -2;
  int rc = ctx_allreduce(c, c->sums.as<double>() + 32, 1, s);
  if (rc) return rc;
  double tot = 0.0;
  if (cudaMemcpyAsync(&tot, c->sums.as<double>() + 32, 8, cudaMemcpyDeviceToHost, s) != cudaSuccess) return -2;
  if (cudaStreamSynchronize(s) != cudaSuccess) return -2;
  if (local_rc) return local_rc;
  if (tot != 0.0) MB_FAIL(3, "another rank reported a failure in this step");
  return 0;
}

// ---- context ----------------------------------------------------------------------------------
extern "C" int mb_create(mb_context** out, int device) {
  if (!out) MB_FAIL(-1, "null ctx pointer");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    MB_FAIL(-2, "no CUDA device: the magma_b200 library has no CPU fallback");
  }
  if (device < 0 || device >= ndev) MB_FAIL(-1, "bad device index");
  // one process per GPU: kernel attributes (shared-memory opt-ins) and a few scratch buffers are set up once
  // per process for the device of the first context; further contexts must use the same device
  static int process_device = -1;
  if (process_device >= 0 && process_device != device)
    MB_FAIL(-1, "this process already drives another CUDA device: use one process per GPU");
  process_device = device;
  CU(cudaSetDevice(device));
  mb_context* c = new mb_context();
  c->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  int lo = 0, hi = 0;
  cudaDeviceGetStreamPriorityRange(&lo, &hi);
  CU(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, lo));
  CU(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, hi));
  CU(cudaMallocHost((void**)&c->h_pin, 256 * sizeof(double)));
  CU(cudaMallocHost((void**)&c->h_pin_i, 1024 * sizeof(int)));
  CU(cudaEventCreate(&c->ev_t0));
  CU(cudaEventCreate(&c->ev_t1));
  // log|det(inv)|: read off the Cholesky diagonal by default; MB_LOGDET=lu reproduces the
  // reference's determinant() (LU of the explicit inverse, likelihoods.R:37-41) -- see DESIGN.md
  const char* env = getenv("MB_LOGDET");
  c->logdet_chol = (env && env[0] == 'l' && env[1] == 'u') ? 0 : 1;
  if (const char* e = getenv("MB_FUSED_BELOW")) c->fused_below = atoll(e);
  if (getenv("MB_MONITOR_RECOMPUTE")) c->monitor_recompute = true;
  if (const char* e = getenv("MB_RESERVE_SMS")) { c->reserve_sms = atoi(e); if (c->reserve_sms < 0) c->reserve_sms = 0; if (c->reserve_sms > c->sm_count / 2) c->reserve_sms = c->sm_count / 2; }
  c->single_blocking = getenv("MB_SINGLE_BLOCKING") != nullptr;
  c->split_lauum = getenv("MB_DENSE_NOSPLIT") == nullptr;
  if (const char* e = getenv("MB_PAIR_TABLES")) c->pair_tables = atoi(e) != 0;
  *out = c;
  return 0;
}

extern "C" int mb_destroy(mb_context* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  DBuf* bufs[] = {&c->d_offs, &c->d_X, &c->d_y, &c->d_gidx, &c->d_gridX, &c->d_pair_d2, &c->d_pair_idx, &c->d_pair_n, &c->hp_i, &c->hp_k, &c->f_t,
                  &c->g_t, &c->retries, &c->lb, &c->counters, &c->sums,
                  &c->pm, &c->pc, &c->m0, &c->tau, &c->inv0, &c->post_inv, &c->fails,
                  &c->nfgv, &c->scratch, &c->kcache, &c->d_lv_offs, &c->tickets, &c->leafbuf, &c->slotbuf, &c->leafsum};
  for (DBuf* b : bufs) b->release();
  c->st_hp_i0.release();
  c->st_hp_i.release();
  c->hpr_mean.release();
  c->hpr_cov.release();
  if (c->nccl && nccl_api()) nccl_api()->CommDestroy(c->nccl);
  for (auto& ev : c->prof_ev) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (auto& ev : c->ev_mark) if (ev) cudaEventDestroy(ev);
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->ev_single) cudaEventDestroy(c->ev_single);
  for (int i = 0; i < 8; ++i) if (c->ev_ring[i]) cudaEventDestroy(c->ev_ring[i]);
  for (int i = 0; i < 8; ++i) if (c->ev_cp[i]) cudaEventDestroy(c->ev_cp[i]);
  c->ws0.release();
  c->ws1.release();
  big_pool_destroy(c->big);
  if (c->h_pin) cudaFreeHost(c->h_pin);
  if (c->h_pin_i) cudaFreeHost(c->h_pin_i);
  if (c->st) cudaStreamDestroy(c->st);
  if (c->st2) cudaStreamDestroy(c->st2);
  if (c->st3) cudaStreamDestroy(c->st3);
  delete c;
  return 0;
}

extern "C" int64_t mb_launch_count(mb_context* c) { return c ? c->launches : 0; }

// Tuning / diagnostic switches of a context (the environment variables of the same meaning are read once at mb_create)
// hand the pair tables to the per-task kernels (or not): every launch copies c->db into its argument block
static void apply_pair_tables(mb_context* c) {
  const bool on = c->pair_tables && c->pairs_built;
  c->db.pair_d2 = on ? c->d_pair_d2.as<double>() : nullptr;
  c->db.pair_idx = on ? c->d_pair_idx.as<unsigned short>() : nullptr;
  c->db.pair_n = on ? c->d_pair_n.as<int32_t>() : nullptr;
  c->db.pair_stride = on ? c->pair_stride : 0;
}

extern "C" int mb_set_option(mb_context* c, const char* name, int64_t value) {
  if (!c || !name) MB_FAIL(-1, "null argument");
  const std::string n(name);
  if (n == "fused_below") c->fused_below = value;                 // MB_FUSED_BELOW
  else if (n == "reserve_sms") c->reserve_sms = (int)(value < 0 ? 0 : (value > c->sm_count / 2 ? c->sm_count / 2 : value));   // MB_RESERVE_SMS
  else if (n == "dense_split") c->split_lauum = value != 0;       // MB_DENSE_NOSPLIT (inverted)
  else if (n == "logdet_chol") c->logdet_chol = value != 0;       // MB_LOGDET=lu <-> 0
  else if (n == "monitor_recompute") c->monitor_recompute = value != 0;   // MB_MONITOR_RECOMPUTE
  else if (n == "pair_tables") { c->pair_tables = value != 0; apply_pair_tables(c); }   // MB_PAIR_TABLES
  else MB_FAIL(-1, "unknown option");
  return 0;
}

extern "C" int mb_set_allreduce(mb_context* c, mb_allreduce_fn fn, void* user, int rank, int world) {
  if (!c) MB_FAIL(-1, "null ctx");
  c->ar = fn; c->ar_user = user; c->rank = rank; c->world = world;
  return 0;
}

// ---- kernel strings ---------------------------------------------------------------------------
extern "C" int mb_kernel_parse(const char* kern, int has_noise, mb_kernel* out) {
  if (!kern || !out) MB_FAIL(-1, "null argument");
  memset(out, 0, sizeof(*out));
  std::vector<std::string> tok;
  std::string cur;
  for (const char* p = kern;; ++p) {
    if (*p == ' ' || *p == 0) { if (!cur.empty()) tok.push_back(cur); cur.clear(); if (!*p) break; }
    else cur.push_back(*p);
  }
  if (tok.empty() || tok[0] == "+" || tok[0] == "*")
    MB_FAIL(-1, "The string in 'kern' should start with a kernel not an operator.");
  if (tok.size() == 1 && tok[0].compare(0, 4, "CONV") == 0) {   // "CONV<O>": the multi-output convolution kernel
    const int O = atoi(tok[0].c_str() + 4);
    if (O < 1) MB_FAIL(-1, "multi-output kernel: write CONV<number of outputs>, e.g. CONV2");
    out->n_terms = 1; out->n_prod = O; out->types[0] = MB_KERN_CONV; out->has_noise = has_noise ? 1 : 0;
    out->n_hp = 2 * O + 1 + (has_noise ? O : 0);
    if (out->n_hp > MB_MAX_HP) MB_FAIL(-1, "too many hyper-parameters: the convolution kernel supports up to 3 outputs with noise");
    return 0;
  }
  bool seen_plus = false, expect_kernel = true;
  int nhp = 0;
  for (size_t i = 0; i < tok.size(); ++i) {
    const std::string& s = tok[i];
    if (expect_kernel) {
      int type;
      if (s == "SE") type = MB_KERN_SE; else if (s == "PERIO") type = MB_KERN_PERIO;
      else if (s == "RQ") type = MB_KERN_RQ; else if (s == "LIN") type = MB_KERN_LIN;
      else MB_FAIL(-1, "Incorrect character string specified in the 'kern' argument.");
      if (out->n_terms >= MB_MAX_TERMS) MB_FAIL(-1, "too many kernel terms");
      for (int q = 0; q < out->n_terms; ++q)
        if (out->types[q] == type) MB_FAIL(-1, "a kernel type may appear once (HP names are unique)");
      out->types[out->n_terms++] = type;
      if (!seen_plus) out->n_prod = out->n_terms;
      nhp += mb_term_nhp(type);
      expect_kernel = false;
    } else {
      if (s == "+") seen_plus = true;
      else if (s == "*") {
        if (seen_plus) MB_FAIL(-1, "Incorrect character string specified in 'kern'. The '*' operators should come before '+' operators.");
      } else MB_FAIL(-1, "Incorrect character string specified in the 'kern' argument.");
      expect_kernel = true;
    }
  }
  if (expect_kernel) MB_FAIL(-1, "Incorrect character string specified in 'kern'.");
  out->has_noise = has_noise ? 1 : 0;
  out->n_hp = nhp + out->has_noise;
  if (out->n_hp > MB_MAX_HP) MB_FAIL(-1, "too many hyper-parameters");
  return 0;
}

extern "C" const char* mb_kernel_hp_name(const mb_kernel* k, int p) {
  static const char* names[4][3] = {{"se_variance", "se_lengthscale", nullptr},
                                    {"perio_variance", "perio_lengthscale", "period"},
                                    {"rq_variance", "rq_lengthscale", "rq_scale"},
                                    {"lin_slope", "lin_offset", nullptr}};
  if (!k || p < 0 || p >= k->n_hp) return nullptr;
  if (k->n_terms == 1 && k->types[0] == MB_KERN_CONV) {   // flatten_hp_mo() names with output labels 1..O
    static thread_local char buf[32];
    const int O = k->n_prod;
    if (p < O) snprintf(buf, sizeof(buf), "l_t_%d", p + 1);
    else if (p < 2 * O) snprintf(buf, sizeof(buf), "S_t_%d", p - O + 1);
    else if (p == 2 * O) snprintf(buf, sizeof(buf), "l_u_t");
    else snprintf(buf, sizeof(buf), "noise_%d", p - 2 * O);
    return buf;
  }
  int off = 0;
  for (int t = 0; t < k->n_terms; ++t) {
    int np = mb_term_nhp(k->types[t]);
    if (p < off + np) return names[k->types[t]][p - off];
    off += np;
  }
  return "noise";
}

// ---- helpers ------------------------------------------------------------------------------------
static int h2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, s));
  return 0;
}
static int d2h(void* dst, const void* src, size_t bytes, cudaStream_t s) {
  CU(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, s));
  return 0;
}
static int check_ctx(mb_context* c) {
  if (!c) MB_FAIL(-1, "null context");
  CU(cudaSetDevice(c->device));
  return 0;
}

// ---- legacy builders --------------------------------------------------------------------------
static int legacy(mb_context* c, int which, const double* m1, int n1, const double* m2, int n2,
                  int d, double par, double* out) {
  MB_TRY(check_ctx(c));
  if (!m1 || !m2 || !out || n1 < 0 || n2 < 0 || d < 0) MB_FAIL(-1, "bad argument");
  if (n1 == 0 || n2 == 0) return 0;
  size_t b1 = (size_t)n1 * d * 8, b2 = (size_t)n2 * d * 8, bo = (size_t)n1 * n2 * 8;
  MB_TRY(c->scratch.ensure(b1 + b2 + bo + 64));
  double* d1 = c->scratch.as<double>();
  double* d2 = d1 + (size_t)n1 * d;
  double* dout = d2 + (size_t)n2 * d;
  MB_TRY(h2d(d1, m1, b1, c->st));
  MB_TRY(h2d(d2, m2, b2, c->st));
  LAUNCH(c, launch_legacy_builder(which, d1, n1, d2, n2, d, par, dout, c->st));
  MB_TRY(d2h(out, dout, bo, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}
extern "C" int mb_cpp_dist(mb_context* c, const double* m1, int n1, const double* m2, int n2, int d, double* out) { return legacy(c, 0, m1, n1, m2, n2, d, 0.0, out); }
extern "C" int mb_cpp_perio(mb_context* c, const double* m1, int n1, const double* m2, int n2, int d, double period, double* out) { return legacy(c, 1, m1, n1, m2, n2, d, period, out); }
extern "C" int mb_cpp_perio_deriv(mb_context* c, const double* m1, int n1, const double* m2, int n2, int d, double period, double* out) { return legacy(c, 2, m1, n1, m2, n2, d, period, out); }
extern "C" int mb_cpp_prod(mb_context* c, const double* m1, int n1, const double* m2, int n2, int d, double* out) { return legacy(c, 3, m1, n1, m2, n2, d, 0.0, out); }
extern "C" int mb_cpp_noise(mb_context* c, const double* m1, int n1, const double* m2, int n2, int d, double noise, double* out) { return legacy(c, 4, m1, n1, m2, n2, d, noise, out); }

// ---- kern_to_cov ------------------------------------------------------------------------------
extern "C" int mb_kern_to_cov(mb_context* c, const mb_kernel* k, const double* hp, const double* x1,
                              int n1, const double* x2, int n2, int d, int deriv, double* out) {
  MB_TRY(check_ctx(c));
  if (!k || !hp || !x1 || !out || n1 < 0 || d < 1) MB_FAIL(-1, "bad argument");
  if (deriv >= k->n_hp) MB_FAIL(-1, "deriv out of range");
  // the multi-output kernel carries its per-output noise only on the square matrix of a set of points against itself
  // (kernel-to-matrices.R:121-157); the single-output kernels add exp(noise) on every coincident pair (:481-487)
  const bool cross = x2 != nullptr && x2 != x1;
  if (!x2) { x2 = x1; n2 = n1; }
  if (n1 == 0 || n2 == 0) return 0;
  DevKern kd = make_devkern(k);
  size_t b1 = (size_t)n1 * d * 8, b2 = (size_t)n2 * d * 8, bo = (size_t)n1 * n2 * 8;
  MB_TRY(c->scratch.ensure(b1 + b2 + bo + 64));
  double* d1 = c->scratch.as<double>();
  double* d2 = d1 + (size_t)n1 * d;
  double* dout = d2 + (size_t)n2 * d;
  MB_TRY(h2d(d1, x1, b1, c->st));
  MB_TRY(h2d(d2, x2, b2, c->st));
  // column-major n1 x n2 output: element (i,j) at i + j*n1
  DevKern kd_eff = kd;
  if (cross && kd.types[0] == MB_KERN_CONV) {
    if (deriv >= 2 * kd.n_prod + 1) {       // derivative w.r.t. a noise term of a cross-covariance: zero matrix
      std::vector<double> z((size_t)n1 * n2, 0.0);
      memcpy(out, z.data(), z.size() * 8);
      return 0;
    }
    kd_eff.has_noise = 0;
  }
  LAUNCH(c, launch_dense_cov(kd_eff, hp, d1, n1, d2, n2, d, deriv, 1, dout, 1, n1, c->st));
  MB_TRY(d2h(out, dout, bo, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// ---- dense chol_inv_jitter on device data -------------------------------------------------------
// src (n x n, upper read) -> ws.A = inverse.  Synchronises `s`; returns >0 on numerical failure.
// one chol_inv_jitter (mode 0) / plain Cholesky log-diagonal (mode 1) on the multi-CTA path, host loop over
// the jitter escalation (kernel-to-matrices.R:670-680).  Synchronises `s`.  info/out: as k_dense_cholinv_sm.
static int big_cholinv_sync(mb_context* c, DenseWs& ws, const double* src, int n, double pen, int mode, double* inv,
                            int* info, double* out, cudaStream_t s) {
  const int64_t ldw = ((int64_t)n + 7) / 8 * 8;
  int retries = 0;
  for (;;) {
    CU(launch_big_cholinv_attempt(src, n, ws.Ap.as<double>(), ws.Bp.as<double>(), ldw, inv, n, pen, retries, mode,
                                  ws.info.as<int>() + 4, info, out, s, &c->launches));
    if (mode == 1) return 0;
    int* slot = c->h_pin_i + 512 + (s == c->st2 ? 1 : 0);
    MB_TRY(d2h(slot, info, sizeof(int), s));
    CU(cudaStreamSynchronize(s));
    if (*slot >= 0 || ++retries > TK_MAX_RETRY) return 0;
  }
}

static int dev_cholinv(mb_context* c, DenseWs& ws, const double* src, int n, double pen,
                       cudaStream_t s, int* retries, double* jitter, double* sumlogdiag) {
  if (dense_use_sm(n))
    LAUNCH(c, launch_dense_cholinv_sm(src, n, ws.A.as<double>(), n, pen, 0, ws.info.as<int>(), ws.res.as<double>() + 16,
                                      c->split_lauum ? ws.Bp.as<double>() : nullptr, s));
  else
    MB_TRY(big_cholinv_sync(c, ws, src, n, pen, 0, ws.A.as<double>(), ws.info.as<int>(), ws.res.as<double>() + 16, s));
  MB_TRY(d2h(c->h_pin_i, ws.info.p, sizeof(int), s));
  MB_TRY(d2h(c->h_pin + 200, ws.res.as<double>() + 16, 2 * sizeof(double), s));
  CU(cudaStreamSynchronize(s));
  if (retries) *retries = c->h_pin_i[0];
  if (jitter) *jitter = c->h_pin[200];
  if (sumlogdiag) *sumlogdiag = c->h_pin[201];
  if (c->h_pin_i[0] < 0) MB_FAIL(1, "chol_inv_jitter: no positive-definite jitter found");
  return 0;
}

// C = alpha op(A) op(B) + beta Cin: the 128 x 128-tile DMMA kernel for large products, the 32 x 32 one otherwise
static cudaError_t gemm_any(int m, int n, int k, const double* A, int64_t ars, int64_t acs, const double* B,
                            int64_t brs, int64_t bcs, double* C, int64_t ldc, double alpha, const double* Cin,
                            double beta, cudaStream_t st) {
  const bool unit = (ars == 1 || acs == 1) && (brs == 1 || bcs == 1);
  if (unit && (int64_t)m * n >= 512 * 512 && k >= 64)
    return launch_big_gemm_plain(m, n, k, A, ars, acs, B, brs, bcs, C, ldc, alpha, Cin, beta, st);
  return launch_gemm_dmma(m, n, k, A, ars, acs, B, brs, bcs, C, ldc, alpha, Cin, beta, st);
}

extern "C" int mb_chol_inv_jitter(mb_context* c, const double* mat, int n, int ld, double pen_diag,
                                  double* inv, int32_t* n_retry, double* jitter_total) {
  MB_TRY(check_ctx(c));
  if (!mat || !inv || n < 1 || ld < n) MB_FAIL(-1, "bad argument");
  MB_TRY(c->ws1.ensure(n));
  // column-major (ld) host matrix -> packed; symmetric so row/column-major agree, but the
  // reference reads the UPPER triangle only: element (i,j), i <= j, sits at mat[i + j*ld].
  std::vector<double> packed((size_t)n * n);
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) packed[(size_t)i * n + j] = (i <= j) ? mat[i + (size_t)j * ld] : mat[j + (size_t)i * ld];
  MB_TRY(h2d(c->ws1.Kmat.p, packed.data(), packed.size() * 8, c->st));
  int retries = 0;
  double jit = 0;
  int rc = dev_cholinv(c, c->ws1, c->ws1.Kmat.as<double>(), n, pen_diag, c->st, &retries, &jit, nullptr);
  if (n_retry) *n_retry = retries;
  if (jitter_total) *jitter_total = jit;
  if (rc != 0) return rc;
  MB_TRY(d2h(inv, c->ws1.A.p, (size_t)n * n * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int mb_pair_table_counts(mb_context* c, int32_t* counts) {
  MB_TRY(check_ctx(c));
  if (!counts) MB_FAIL(-1, "null argument");
  if (!c->have_db || !c->pairs_built) MB_FAIL(-1, "no pair tables for the resident dataset");
  return d2h(counts, c->d_pair_n.p, (size_t)c->db.n_tasks * 4, c->st);
}

// grid_index of the resident tasks: in range, and distinct within a task.  The E-step scatter adds one value per (task,
// grid position) and relies on the positions of one task being distinct; the reference stops on such data
// (R/training.R:703-714, "At least one individual have several Outputs on the same grid point").
static int check_grid_index(const int64_t* offs, int64_t n_tasks, const int32_t* grid_index, int32_t n_grid) {
  std::vector<int64_t> seen;
  for (int64_t t = 0; t < n_tasks; ++t) {
    // rows are Reference-sorted like the grid (R/training.R:700), so the positions of a task normally increase strictly:
    // one predictable comparison per row; only a task that breaks the order gets the general distinctness test
    int32_t prev = -1, lo = 0, hi = n_grid - 1;
    bool ordered = true;
    for (int64_t r = offs[t]; r < offs[t + 1]; ++r) {
      const int32_t g = grid_index[r];
      lo = g < lo ? g : lo;
      hi = g > hi ? g : hi;
      ordered &= g > prev;
      prev = g;
    }
    if (lo < 0 || hi >= n_grid) MB_FAIL(-1, "grid_index out of range");
    if (ordered) continue;
    if (seen.empty()) seen.assign((size_t)n_grid, -1);
    for (int64_t r = offs[t]; r < offs[t + 1]; ++r) {
      const int32_t g = grid_index[r];
      if (seen[g] == t)
        MB_FAIL(-1, "At least one individual have several Outputs on the same grid point (duplicate grid_index within a task)");
      seen[g] = t;
    }
  }
  return 0;
}

// ---- dataset -------------------------------------------------------------------------------------
extern "C" int mb_upload_dataset(mb_context* c, int64_t n_tasks, const int64_t* offs, int d,
                                 const double* X, const double* y, const int32_t* grid_index,
                                 int32_t n_grid, const double* grid_X) {
  MB_TRY(check_ctx(c));
  if (n_tasks < 1 || !offs || !X || !y || d < 1) MB_FAIL(-1, "bad argument");
  int64_t rows = offs[n_tasks];
  int nmax = 0;
  for (int64_t t = 0; t < n_tasks; ++t) {
    int64_t n = offs[t + 1] - offs[t];
    if (n < 1) MB_FAIL(-1, "empty task");
    if (n > nmax) nmax = (int)n;
  }
  c->have_db = false;   // the resident dataset is being replaced: an error below leaves none
  MB_TRY(c->d_offs.ensure((n_tasks + 1) * 8));
  MB_TRY(c->d_X.ensure((size_t)rows * d * 8));
  MB_TRY(c->d_y.ensure((size_t)rows * 8));
  MB_TRY(h2d(c->d_offs.p, offs, (n_tasks + 1) * 8, c->st));
  MB_TRY(h2d(c->d_X.p, X, (size_t)rows * d * 8, c->st));
  c->db.offs = c->d_offs.as<int64_t>();
  c->db.X = c->d_X.as<double>();
  c->db.D = d;
  c->db.n_tasks = n_tasks;
  c->db.nmax = nmax;
  // pair tables: the distinct squared distances of every task and the index of each matrix element into them (pairs3.cu).
  // They need the row offsets and the inputs only, so the kernel runs on the second stream behind those two copies while
  // the outputs and grid indices are still on their way
  c->pairs_built = false;
  apply_pair_tables(c);
  bool pairs_launched = false;
  if (c->pair_tables && nmax <= TK_MAXN) {
    c->pair_stride = (nmax <= 64 ? 36 : 78) * 64;
    MB_TRY(c->d_pair_d2.ensure((size_t)n_tasks * MB_PAIR_CAP * 8));
    MB_TRY(c->d_pair_idx.ensure((size_t)n_tasks * c->pair_stride * 2));
    MB_TRY(c->d_pair_n.ensure((size_t)n_tasks * 4));
    if (!c->ev_single) CU(cudaEventCreateWithFlags(&c->ev_single, cudaEventDisableTiming));
    CU(cudaEventRecord(c->ev_single, c->st));
    CU(cudaStreamWaitEvent(c->st2, c->ev_single, 0));
    LAUNCH(c, launch_task_pairs3(c->db, c->d_pair_d2.as<double>(), c->d_pair_idx.as<unsigned short>(),
                                 c->d_pair_n.as<int32_t>(), c->pair_stride, c->st2));
    pairs_launched = true;
  }
  MB_TRY(h2d(c->d_y.p, y, (size_t)rows * 8, c->st));
  if (grid_index) {
    MB_TRY(c->d_gidx.ensure((size_t)rows * 4));
    MB_TRY(h2d(c->d_gidx.p, grid_index, (size_t)rows * 4, c->st));
  }
  if (grid_X && n_grid > 0) {
    MB_TRY(c->d_gridX.ensure((size_t)n_grid * d * 8));
    MB_TRY(h2d(c->d_gridX.p, grid_X, (size_t)n_grid * d * 8, c->st));
  }
  // the host-side validation of grid_index runs while the copies are in flight (pinned source buffers: the copies above
  // only enqueue); on failure the streams are drained and the context is left without a dataset
  if (grid_index) {
    const int rc = check_grid_index(offs, n_tasks, grid_index, n_grid);
    if (rc) { cudaStreamSynchronize(c->st); cudaStreamSynchronize(c->st2); return rc; }
  }
  c->db.y = c->d_y.as<double>();
  c->db.gidx = grid_index ? c->d_gidx.as<int32_t>() : nullptr;
  if (pairs_launched) {
    CU(cudaStreamSynchronize(c->st2));
    c->pairs_built = true;
    apply_pair_tables(c);
  }
  CU(cudaStreamSynchronize(c->st));
  c->h_offs.assign(offs, offs + n_tasks + 1);
  c->n_rows = rows;
  c->U = n_grid;
  c->have_db = true;
  return 0;
}

static int need_db(mb_context* c, bool need_grid) {
  MB_TRY(check_ctx(c));
  if (!c->have_db) MB_FAIL(-1, "no dataset uploaded (mb_upload_dataset)");
  if (need_grid && (!c->db.gidx || c->U < 1)) MB_FAIL(-1, "dataset was uploaded without a union grid");
  return 0;
}

// upload an HP table (n_tasks x P, or one row when shared) -> device; returns stride
static int put_hp_i(mb_context* c, const mb_kernel* k, const double* hp, int shared, int64_t* stride) {
  size_t rows = shared ? 1 : (size_t)c->db.n_tasks;
  MB_TRY(c->hp_i.ensure(rows * k->n_hp * 8 + 64));
  MB_TRY(h2d(c->hp_i.p, hp, rows * k->n_hp * 8, c->st));
  *stride = shared ? 0 : k->n_hp;
  return 0;
}

// ---- list_kern_to_cov / list_kern_to_inv -------------------------------------------------------
extern "C" int mb_list_kern_to_mat(mb_context* c, const mb_kernel* k, int64_t n_tasks,
                                   const int64_t* offs, int d, const double* X, const double* hp,
                                   int hp_shared, int deriv, int inverse, double pen_diag,
                                   const int64_t* out_offs, double* out, int32_t* retries) {
  MB_TRY(check_ctx(c));
  if (!k || !offs || !X || !hp || !out_offs || !out || n_tasks < 1) MB_FAIL(-1, "bad argument");
  if (deriv >= k->n_hp) MB_FAIL(-1, "deriv out of range");
  int64_t rows = offs[n_tasks];
  int nmax = 0;
  size_t total = 0;
  for (int64_t t = 0; t < n_tasks; ++t) {
    int64_t n = offs[t + 1] - offs[t];
    if (n > nmax) nmax = (int)n;
    size_t end = (size_t)out_offs[t] + (size_t)n * n;
    if (end > total) total = end;
  }
  DevKern kd = make_devkern(k);
  DBuf b_offs, b_X, b_hp, b_oo, b_out, b_ret;
  size_t hp_rows = hp_shared ? 1 : (size_t)n_tasks;
  int rc = 0;
  if (b_offs.ensure((n_tasks + 1) * 8) || b_X.ensure((size_t)rows * d * 8) ||
      b_hp.ensure(hp_rows * k->n_hp * 8) || b_oo.ensure(n_tasks * 8) || b_out.ensure(total * 8) ||
      b_ret.ensure(n_tasks * 4)) rc = -3;
  auto cleanup = [&]() { b_offs.release(); b_X.release(); b_hp.release(); b_oo.release(); b_out.release(); b_ret.release(); };
  if (rc) { cleanup(); return rc; }
  cudaMemcpyAsync(b_offs.p, offs, (n_tasks + 1) * 8, cudaMemcpyHostToDevice, c->st);
  cudaMemcpyAsync(b_X.p, X, (size_t)rows * d * 8, cudaMemcpyHostToDevice, c->st);
  cudaMemcpyAsync(b_hp.p, hp, hp_rows * k->n_hp * 8, cudaMemcpyHostToDevice, c->st);
  cudaMemcpyAsync(b_oo.p, out_offs, n_tasks * 8, cudaMemcpyHostToDevice, c->st);
  TaskData db;
  db.offs = b_offs.as<int64_t>(); db.X = b_X.as<double>(); db.y = b_X.as<double>(); db.gidx = nullptr;
  db.D = d; db.n_tasks = n_tasks; db.nmax = nmax;
  cudaError_t e;
  if (!inverse) {
    e = launch_task_cov_sym(db, kd, b_hp.as<double>(), hp_shared ? 0 : k->n_hp, deriv, b_out.as<double>(),
                            b_oo.as<int64_t>(), c->sm_count, c->st);
    if (e == cudaErrorInvalidValue) {   // task rows larger than the staging buffer: element-per-thread kernel
      cudaGetLastError();
      e = launch_task_cov(db, kd, b_hp.as<double>(), hp_shared ? 0 : k->n_hp, deriv, b_out.as<double>(),
                          b_oo.as<int64_t>(), c->st);
    }
  } else {
    if (deriv >= 0) { cleanup(); MB_FAIL(-1, "inverse of a derivative matrix is not part of the hot path"); }
    TaskInvArgs a;
    memset(&a, 0, sizeof(a));
    a.db = db; a.kd = kd; a.hp = b_hp.as<double>(); a.hp_stride = hp_shared ? 0 : k->n_hp;
    a.pen = pen_diag; a.K = 1; a.inv_out = b_out.as<double>(); a.inv_offs = b_oo.as<int64_t>();
    a.retries = b_ret.as<int32_t>();
    if (nmax > TK_MAXN) {
      a.U = 0;
      int brc = big_tasks_inverse(&c->big, a, offs, out_offs, nullptr, c->st, &c->launches);
      if (brc != 0) { cleanup(); return brc; }
      e = cudaSuccess;
    } else {
      int g = c->sm_count * task_ctas_per_sm_any(nmax, d);
      if (g > n_tasks) g = (int)n_tasks;
      e = launch_task_inverse_any(a, g, c->st);
    }
  }
  c->launches++;
  if (e == cudaSuccess) e = cudaMemcpyAsync(out, b_out.p, total * 8, cudaMemcpyDeviceToHost, c->st);
  std::vector<int32_t> hret;
  if (e == cudaSuccess && inverse) {
    hret.resize(n_tasks);
    e = cudaMemcpyAsync(hret.data(), b_ret.p, n_tasks * 4, cudaMemcpyDeviceToHost, c->st);
  }
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
  cleanup();
  if (e != cudaSuccess) { mb_set_error(__FILE__, __LINE__, cudaGetErrorString(e)); return -2; }
  if (inverse) {
    int bad = 0;
    for (int64_t t = 0; t < n_tasks; ++t) { if (retries) retries[t] = hret[t]; if (hret[t] < 0) bad = 1; }
    if (bad) MB_FAIL(1, "chol_inv_jitter: no positive-definite jitter found for at least one task");
  }
  return 0;
}

// ================================================================================================
// Device-resident building blocks of the EM / VEM steps
// ================================================================================================

// logL_GP_mod (+ gradient) of one n x n problem whose data already sit on the device.
// Enqueues everything on `s` and leaves res[0..P] in ws.res (device) + an async copy in
// hres (pinned, 1 + MB_MAX_HP doubles) and the jitter info in hinfo; the caller synchronises.
static int dense_obj_enqueue(mb_context* c, DenseWs& ws, const DevKern& kd, const double* hp_host,
                             const double* X, int n, int D, const double* y, const double* mean,
                             const double* S, double pen, int want_grad, cudaStream_t s,
                             double* hres, int* hinfo) {
  LAUNCH(c, launch_dense_cov(kd, hp_host, X, n, X, n, D, -1, 1, ws.Kmat.as<double>(), n, 1, s));
  if (dense_use_sm(n))
    LAUNCH(c, launch_dense_cholinv_sm(ws.Kmat.as<double>(), n, ws.A.as<double>(), n, pen, 0, ws.info.as<int>(),
                                      ws.res.as<double>() + 16, c->split_lauum ? ws.Bp.as<double>() : nullptr, s));
  else
    MB_TRY(big_cholinv_sync(c, ws, ws.Kmat.as<double>(), n, pen, 0, ws.A.as<double>(), ws.info.as<int>(),
                            ws.res.as<double>() + 16, s));
  const double* logdet;
  if (c->logdet_chol) {
    logdet = ws.res.as<double>() + 17;
  } else {
    LAUNCH(c, launch_dense_lu_logdet(ws.A.as<double>(), ws.B.as<double>(), n, ws.done.as<int>(),
                                     ws.res.as<double>() + 18, s));
    logdet = ws.res.as<double>() + 18;
  }
  LAUNCH(c, launch_resid_gemv(n, ws.A.as<double>(), y, mean, ws.r(), ws.al(), s));
  const double* T = nullptr;
  if (want_grad && S) {
    LAUNCH(c, gemm_any(n, n, n, S, n, 1, ws.A.as<double>(), n, 1, ws.B.as<double>(), n, 1.0, nullptr, 0.0, s));
    LAUNCH(c, gemm_any(n, n, n, ws.A.as<double>(), n, 1, ws.B.as<double>(), n, 1, ws.T.as<double>(), n, 1.0, nullptr, 0.0, s));
    T = ws.T.as<double>();
  }
  LAUNCH(c, launch_dense_obj(kd, hp_host, X, n, D, ws.A.as<double>(), S, T, ws.al(), ws.r(), logdet,
                             c->logdet_chol, want_grad, ws.part.as<double>(), ws.res.as<double>(), s));
  c->launches++;  // launch_dense_obj issues two kernels
  MB_TRY(d2h(hres, ws.res.p, (1 + MB_MAX_HP) * sizeof(double), s));
  MB_TRY(d2h(hinfo, ws.info.p, sizeof(int), s));
  return 0;
}

struct EStepOut { double* post_mean; double* post_cov; };  // device, K x U and K x U x U

// E step / VE step core on the resident dataset.  All pointers are device pointers except the
// HP vectors of the mean processes.  info3: retries inv_k (max), post (max), tasks (max).
static int dev_e_step(mb_context* c, int K, const mb_kernel* kern_k, const double* hp_k_host,
                      const DevKern& kdi, const double* hp_i_dev, int64_t hp_stride,
                      const double* m_k_dev, const double* tau_dev, const double* gridX, int U,
                      const int32_t* gidx_override, double pen, double* post_mean, double* post_cov,
                      int32_t* info3) {
  DevKern kdk = make_devkern(kern_k);
  const size_t UU = (size_t)U * U;
  MB_TRY(c->ws1.ensure(U));
  MB_TRY(c->inv0.ensure(K * UU * 8));
  MB_TRY(c->post_inv.ensure(K * UU * 8));
  MB_TRY(c->retries.ensure(c->db.n_tasks * 4));
  int max_r0 = 0, max_rp = 0;
  // task inverses, added leaf by leaf in task order (estep3.cu); one accumulator [K x U x U | K x U] per local leaf
  const bool big = c->db.nmax > TK_MAXN;
  MB_TRY(ctx_leaves(c, U));
  const int nl = c->lv_n;
  const int64_t pstride = (int64_t)K * (UU + U);
  MB_TRY(c->leafbuf.ensure((size_t)nl * pstride * 8));
  MB_TRY(c->tickets.ensure((size_t)(nl + 2) * sizeof(int)));   // [nl] = the E-step kernel's work counter
  auto mark = [&](int i) -> int {
    if (!c->prof_on) return 0;
    if (!c->ev_mark[i]) CU(cudaEventCreate(&c->ev_mark[i]));
    CU(cudaEventRecord(c->ev_mark[i], c->st));
    return 0;
  };
  MB_TRY(mark(0));
  CU(cudaMemsetAsync(c->leafbuf.p, 0, (size_t)nl * pstride * 8, c->st));
  CU(cudaMemsetAsync(c->tickets.p, 0, (size_t)(nl + 2) * sizeof(int), c->st));
  TaskInvArgs a;
  memset(&a, 0, sizeof(a));
  a.db = c->db;
  if (gidx_override) a.db.gidx = gidx_override;
  a.kd = kdi; a.hp = hp_i_dev; a.hp_stride = hp_stride; a.pen = pen; a.K = K; a.tau = tau_dev; a.U = U;
  a.part_inv = c->leafbuf.as<double>(); a.part_w = c->leafbuf.as<double>() + K * UU; a.part_stride = pstride;
  a.leaf_offs = c->d_lv_offs.as<int64_t>(); a.n_leaves = nl; a.tickets = c->tickets.as<int>();
  a.retries = c->retries.as<int32_t>();
  { static const int diag = getenv("MB_ESTEP_DIAG") ? atoi(getenv("MB_ESTEP_DIAG")) : 0; a.diag_skip = diag; }   // timing decomposition only
  if (big) {
    MB_TRY(big_tasks_inverse(&c->big, a, c->h_offs.data(), nullptr, c->lv_offs.data(), c->st, &c->launches));
  } else {
    // Dealing of the (leaf, position) items to persistent CTAs (estep3.cu).  Default: items pulled in order from a device
    // counter -- any grid size is safe and every CTA slot is used.  MB_ESTEP_STATIC: CTA b serves leaf b % nl, positions
    // b / nl, b / nl + per, ...; the CTAs of a leaf wait for each other, so the whole grid has to be resident.  Same bits
    // and, at C2, the same time (0.83 ms) either way.
    int64_t cap = (int64_t)c->sm_count * task_ctas_per_sm_any(c->db.nmax, c->db.D);
    if (cap > c->db.n_tasks) cap = c->db.n_tasks;
    int64_t longest = 1;
    for (int l = 0; l < nl; ++l) if (c->lv_offs[l + 1] - c->lv_offs[l] > longest) longest = c->lv_offs[l + 1] - c->lv_offs[l];
    a.leaf_max = (int)longest;
    static const bool static_deal = getenv("MB_ESTEP_STATIC") != nullptr;
    int64_t per = cap / nl;
    if (static_deal && per >= 1) {
      if (per > longest) per = longest;
      a.queue = nullptr;
      a.leaf_max = (int)(((longest + per - 1) / per) * per);
      LAUNCH(c, launch_task_inverse_any(a, (int)(nl * per), c->st));
    } else {
      a.queue = c->tickets.as<int>() + nl;
      // a few SMs stay free for the inv_0 chain on st2 (union-grid kernels; the single-CTA factorisation needs a whole SM's
      // shared memory and would otherwise wait for the persistent task CTAs to drain)
      static const bool reserve = getenv("MB_ESTEP_NORESERVE") == nullptr;
      if (reserve && c->reserve_sms > 0 && cap == (int64_t)c->sm_count * task_ctas_per_sm_any(c->db.nmax, c->db.D))
        a.sm_limit = c->sm_count - c->reserve_sms;
      LAUNCH(c, launch_task_inverse_any(a, (int)cap, c->st));
    }
  }
  MB_TRY(mark(1));
  // inverse prior covariances on st2, enqueued AFTER the task kernel so that their host synchronisations (retry counts)
  // are spent while the tasks are being factorised on st
  for (int k = 0; k < K; ++k) {
    LAUNCH(c, launch_dense_cov(kdk, hp_k_host + (size_t)k * kdk.n_hp, gridX, U, gridX, U, c->db.D, -1, 1,
                               c->ws1.Kmat.as<double>(), U, 1, c->st2));
    int r = 0;
    int rc = dev_cholinv(c, c->ws1, c->ws1.Kmat.as<double>(), U, pen, c->st2, &r, nullptr, nullptr);
    if (rc != 0) {   // the mean process is replicated: every rank fails here alike, nobody is left in a collective
      cudaStreamSynchronize(c->st);
      return rc;
    }
    if (r > max_r0) max_r0 = r;
    CU(cudaMemcpyAsync(c->inv0.as<double>() + k * UU, c->ws1.A.p, UU * 8, cudaMemcpyDeviceToDevice, c->st2));
  }
  CU(cudaStreamSynchronize(c->st2));
  // [post_inv_k | weighted_k] = [inv_k | inv_k m_k] + tree over the leaves of all ranks
  MB_TRY(c->sums.ensure((size_t)pstride * 8 + 256));
  double* base = c->sums.as<double>();
  CU(cudaMemcpyAsync(base, c->inv0.p, K * UU * 8, cudaMemcpyDeviceToDevice, c->st));
  for (int k = 0; k < K; ++k)
    LAUNCH(c, launch_gemv(U, U, c->inv0.as<double>() + k * UU, U, m_k_dev + (size_t)k * U, nullptr, base + K * UU + (size_t)k * U, c->st));
  MB_TRY(c->post_inv.ensure((size_t)pstride * 8));
  MB_TRY(mark(2));
  MB_TRY(leaf_reduce(c, c->leafbuf.as<double>(), pstride, pstride, base, c->post_inv.as<double>(), c->st));
  // the per-task kernel adds upper triangles only (estep3.cu)
  LAUNCH(c, launch_mirror_upper(c->post_inv.as<double>(), U, K, (int64_t)UU, c->st));
  MB_TRY(mark(3));
  double* weighted = c->post_inv.as<double>() + K * UU;
  for (int k = 0; k < K; ++k) {
    int r = 0;
    int rc = dev_cholinv(c, c->ws1, c->post_inv.as<double>() + k * UU, U, pen, c->st, &r, nullptr, nullptr);
    if (rc != 0) return rc;
    if (r > max_rp) max_rp = r;
    CU(cudaMemcpyAsync(post_cov + k * UU, c->ws1.A.p, UU * 8, cudaMemcpyDeviceToDevice, c->st));
    LAUNCH(c, launch_gemv(U, U, post_cov + k * UU, U, weighted + (size_t)k * U, nullptr,
                          post_mean + (size_t)k * U, c->st));
  }
  // task retry statistics: the largest count and the "never factored" flag are reduced on the device
  MB_TRY(mark(4));
  LAUNCH(c, launch_retry_stats(c->retries.as<int32_t>(), c->db.n_tasks, c->tickets.as<int>(), c->st));
  MB_TRY(d2h(c->h_pin_i + 16, c->tickets.p, 2 * sizeof(int), c->st));
  MB_TRY(mark(5));
  CU(cudaStreamSynchronize(c->st));
  if (c->prof_on) {
    // stream time of the segments: [4] memsets + task kernel, [5] idle until inv_0 is there (host wait) + base copy,
    // [6] leaf reduction (+ all-reduce), [7] inverse of post_inv + mean, [8] retry statistics
    for (int i = 0; i < 5; ++i) {
      float ms = 0.f;
      CU(cudaEventElapsedTime(&ms, c->ev_mark[i], c->ev_mark[i + 1]));
      c->prof_phase[4 + i] += ms;
    }
  }
  int rc_tasks = 0;
  if (c->h_pin_i[17] != 0) { mb_set_error(__FILE__, __LINE__, "chol_inv_jitter: no positive-definite jitter found for a task"); rc_tasks = 1; }
  rc_tasks = ctx_agree(c, rc_tasks, c->st);
  if (rc_tasks) return rc_tasks;
  if (info3) { info3[0] = max_r0; info3[1] = max_rp; info3[2] = c->h_pin_i[16]; }
  return 0;
}

// sum log diag chol(src) without jitter (likelihoods.R:303-307, elbos.R:259-264)
static cudaError_t launch_chol_logdiag_any(mb_context* c, DenseWs& ws, const double* src, int n, int* info,
                                           double* out, cudaStream_t s) {
  if (dense_use_sm(n)) return launch_dense_cholinv_sm(src, n, nullptr, n, 0.0, 1, info, out, nullptr, s);
  const int64_t ldw = ((int64_t)n + 7) / 8 * 8;
  return launch_big_cholinv_attempt(src, n, ws.Ap.as<double>(), ws.Bp.as<double>(), ldw, nullptr, n, 0.0, 0, 1,
                                    ws.info.as<int>() + 4, info, out, s, &c->launches);
}

static TaskObjArgs make_obj_args(mb_context* c, const DevKern& kd, const double* hp_dev, int64_t stride,
                                 int mode, int want_grad, double pen, int K, const double* mean,
                                 const double* cov, const double* tau, double* f, double* g) {
  TaskObjArgs a;
  memset(&a, 0, sizeof(a));
  a.db = c->db; a.kd = kd; a.hp = hp_dev; a.hp_stride = stride; a.mode = mode; a.want_grad = want_grad;
  a.logdet_chol = c->logdet_chol; a.pen = pen;
  a.hpst.K = K; a.hpst.U = c->U; a.hpst.mean = mean; a.hpst.cov = cov; a.hpst.tau = tau;
  a.f = f; a.g = g; a.retries = nullptr;
  return a;
}

// Sum over tasks of a task objective (+ gradient) with one shared HP vector: returns f and g on
// the host.  Used by the shared-HP M steps and by the monitoring sums.
static int shared_obj_eval(mb_context* c, TaskObjArgs a, const double* hp_host, int P, int want_grad,
                           double* f_out, double* g_out) {
  MB_TRY(c->hp_i.ensure(MB_MAX_HP * 8));
  MB_TRY(h2d(c->hp_i.p, hp_host, P * 8, c->st));
  a.hp = c->hp_i.as<double>(); a.hp_stride = 0; a.want_grad = want_grad;
  MB_TRY(launch_obj(c, a, (int)c->db.n_tasks, c->st));
  MB_TRY(c->sums.ensure(64 * 8));
  MB_TRY(leaf_sum_columns(c, a.f, 1, c->sums.as<double>(), c->st));
  if (want_grad) MB_TRY(leaf_sum_columns(c, a.g, P, c->sums.as<double>() + 1, c->st));
  MB_TRY(d2h(c->h_pin, c->sums.p, (1 + P) * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  *f_out = c->h_pin[0];
  if (want_grad) for (int p = 0; p < P; ++p) g_out[p] = c->h_pin[1 + p];
  return 0;
}

// Lock-step batched L-BFGS-B over all tasks on the device (`a` describes the objective) while an
// optional list of single problems (mean / cluster processes) is optimised by host state
// machines whose evaluations run on st2, overlapped with the task launches.
struct SingleProblem {
  LbfgsbState s;
  DevKern kd;
  const double* X; int n; int D;
  const double* y; const double* mean; const double* S;
  int evals = 0;
};

static int run_single_eval(mb_context* c, SingleProblem& sp, double pen, double* f, double* g) {
  double* hres = c->h_pin + 32;
  int* hinfo = c->h_pin_i + 8;
  MB_TRY(dense_obj_enqueue(c, c->ws0, sp.kd, sp.s.x, sp.X, sp.n, sp.D, sp.y, sp.mean, sp.S, pen, 1,
                           c->st2, hres, hinfo));
  CU(cudaStreamSynchronize(c->st2));
  sp.evals++;
  *f = (hinfo[0] < 0) ? NAN : hres[0];
  for (int p = 0; p < sp.kd.n_hp; ++p) g[p] = hres[1 + p];
  return 0;
}

static int optimise(mb_context* c, std::vector<SingleProblem>& singles, bool tasks_on,
                    TaskObjArgs a, double* hp_i_dev, int P_i, double pen, int64_t* stats,
                    const int32_t* skip_dev = nullptr) {
  const int64_t M = c->db.n_tasks;
  int rounds = 0;
  bool tasks_active = tasks_on;
  if (tasks_on) {
    MB_TRY(c->lb.ensure(M * sizeof(LbfgsbState)));
    MB_TRY(c->counters.ensure(4096 * sizeof(int) + 2 * sizeof(unsigned long long)));
    CU(cudaMemsetAsync(c->counters.p, 0, 4096 * sizeof(int) + 2 * sizeof(unsigned long long), c->st));
    LAUNCH(c, launch_lbfgsb_init(c->lb.as<LbfgsbState>(), M, P_i, hp_i_dev, P_i, 1e13, 0.0, 25, c->st));
    if (skip_dev) LAUNCH(c, launch_lbfgsb_mask(c->lb.as<LbfgsbState>(), M, skip_dev, c->st));
    a.lb = c->lb.as<LbfgsbState>();
    a.hp = nullptr;
    a.want_grad = 1;
  }
  // persistent-kernel path (mstep3.cu): objective modes of the DMMA kernel, tasks that fit shared memory
  const int64_t fused_below = c->fused_below >= 0 ? c->fused_below
                                                  : (int64_t)3 * c->sm_count * task_ctas_per_sm_any(c->db.nmax, c->db.D);
  const bool fused_ok = tasks_on && fused_below > 0 && c->db.nmax <= TK_MAXN &&
                        (a.mode == TK_OBJ_MOD || a.mode == TK_OBJ_ELBO);
  bool fused_pending = fused_ok;
  bool fused_now = fused_ok && M <= fused_below;     // few tasks (e.g. a shard of 8 GPUs): no lock-step round at all
  auto launch_fused = [&]() -> int {
    TaskObjArgs af = a;
    const int n8 = 0; (void)n8;
    af.kcache_stride = (int64_t)(a.db.nmax <= 64 ? 36 : 78) * 64;
    MB_TRY(c->kcache.ensure((size_t)a.db.n_tasks * af.kcache_stride * 8));
    af.kcache = c->kcache.as<double>();
    int* queue = c->counters.as<int>() + 4095;
    unsigned long long* cyc = reinterpret_cast<unsigned long long*>(c->counters.as<int>() + 4096);
    const int reserve = singles.empty() ? 0 : c->reserve_sms;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (c->prof_on) {
      if (c->prof_used == c->prof_ev.size()) {
        CU(cudaEventCreate(&e0));
        CU(cudaEventCreate(&e1));
        c->prof_ev.push_back({e0, e1});
      }
      if (c->prof_kind.size() <= c->prof_used) c->prof_kind.resize(c->prof_used + 1);
      c->prof_kind[c->prof_used] = 1;
      auto& ev = c->prof_ev[c->prof_used++];
      e0 = ev.first; e1 = ev.second;
      CU(cudaEventRecord(e0, c->st));
    }
    LAUNCH(c, launch_task_optimise3(af, queue, c->sm_count, reserve, cyc, c->st));
    if (e1) CU(cudaEventRecord(e1, c->st));
    return 0;
  };
  size_t cur = 0;  // singles are optimised one after the other (they share ws0)
  // The host runs LB_LAG round(s) ahead of the device: round r + LB_LAG is enqueued before the active count of
  // round r is read, so the stream never drains between rounds.  The price is LB_LAG empty rounds at the end
  // (every CTA / thread sees an inactive problem and returns).
  const int LB_LAG = 1, RING = 8;
  int enq = 0;          // rounds enqueued
  for (int i = 0; i < RING; ++i) {
    if (!c->ev_ring[i]) CU(cudaEventCreateWithFlags(&c->ev_ring[i], cudaEventDisableTiming));
    if (!c->ev_cp[i]) CU(cudaEventCreateWithFlags(&c->ev_cp[i], cudaEventDisableTiming));
  }
  if (!c->st3) CU(cudaStreamCreateWithFlags(&c->st3, cudaStreamNonBlocking));
  // Event-driven host loop: neither the task rounds nor the single (mean / cluster process) problem ever block the other.
  // The single problem's evaluation is enqueued on st2 and polled; a blocking wait there (~225 us per evaluation of the
  // 201-point problem) would pace the task rounds at one per evaluation, which starves the task stream as soon as a round
  // is shorter than that (tail rounds, and every round once the tasks are sharded over 4-8 GPUs).
  if (!c->ev_single) CU(cudaEventCreateWithFlags(&c->ev_single, cudaEventDisableTiming));
  int read = 0;                 // rounds whose active count has been read
  bool single_inflight = false;
  double* hres = c->h_pin + 32;
  int* hinfo = c->h_pin_i + 8;
  while (tasks_active || cur < singles.size()) {
    bool progressed = false;
    if (tasks_active && fused_pending && fused_now) {
      // hand the remaining trajectories to the persistent kernel (it skips the problems that have stopped)
      MB_TRY(launch_fused());
      fused_pending = false;
      tasks_active = false;
      rounds = enq;
      progressed = true;
    }
    if (tasks_active && enq - read <= LB_LAG) {
      if (enq >= 4095) MB_FAIL(2, "L-BFGS-B lock-step did not terminate");
      MB_TRY(launch_obj(c, a, (int)M, c->st, true));
      LAUNCH(c, launch_lbfgsb_advance(c->lb.as<LbfgsbState>(), M, P_i, a.f, a.g, c->counters.as<int>() + enq, c->st));
      // the active count travels on a side stream: a copy between two kernels of the main stream would put a
      // kernel -> copy engine -> kernel hand-over (~20 us) into every round
      CU(cudaEventRecord(c->ev_ring[enq % RING], c->st));
      CU(cudaStreamWaitEvent(c->st3, c->ev_ring[enq % RING], 0));
      MB_TRY(d2h(c->h_pin_i + 64 + (enq % RING), c->counters.as<int>() + enq, sizeof(int), c->st3));
      CU(cudaEventRecord(c->ev_cp[enq % RING], c->st3));
      enq++;
      progressed = true;
    }
    if (cur < singles.size()) {
      SingleProblem& sp = singles[cur];
      if (!single_inflight) {
        if (!sp.s.active) { cur++; }
        else if (dense_use_big(sp.n) || c->single_blocking) {   // the large-matrix path synchronises inside (host jitter loop): blocking evaluation
          double f, g[MB_MAX_HP];
          MB_TRY(run_single_eval(c, sp, pen, &f, g));
          lb_advance(&sp.s, f, g);
          if (!sp.s.active) cur++;
        } else {
          MB_TRY(dense_obj_enqueue(c, c->ws0, sp.kd, sp.s.x, sp.X, sp.n, sp.D, sp.y, sp.mean, sp.S, pen, 1, c->st2, hres, hinfo));
          CU(cudaEventRecord(c->ev_single, c->st2));
          single_inflight = true;
        }
        progressed = true;
      } else {
        const cudaError_t q = tasks_active ? cudaEventQuery(c->ev_single) : cudaEventSynchronize(c->ev_single);
        if (q == cudaSuccess) {
          sp.evals++;
          double g[MB_MAX_HP];
          const double f = (hinfo[0] < 0) ? NAN : hres[0];
          for (int p = 0; p < sp.kd.n_hp; ++p) g[p] = hres[1 + p];
          lb_advance(&sp.s, f, g);
          if (!sp.s.active) cur++;
          single_inflight = false;
          progressed = true;
        } else if (q != cudaErrorNotReady) {
          mb_set_error(__FILE__, __LINE__, cudaGetErrorString(q));
          return -2;
        }
      }
    }
    if (tasks_active && read < enq) {
      // with nothing else to do the host may as well block on the oldest outstanding count
      const bool idle = !progressed && !(cur < singles.size());
      const cudaError_t q = idle ? cudaEventSynchronize(c->ev_cp[read % RING]) : cudaEventQuery(c->ev_cp[read % RING]);
      if (q == cudaSuccess) {
        const int still = c->h_pin_i[64 + (read % RING)];
        if (still == 0) { tasks_active = false; rounds = read + 1; }
        else if (fused_pending && still <= fused_below) fused_now = true;   // the tail: one task's latency per round from here on
        read++;
      } else if (q != cudaErrorNotReady) {
        mb_set_error(__FILE__, __LINE__, cudaGetErrorString(q));
        return -2;
      }
    }
  }
  if (tasks_on) {
    MB_TRY(c->fails.ensure(M * 4));
    MB_TRY(c->nfgv.ensure(M * 4));
    // a.f <- the objective at the final iterate of every task (what a re-evaluation at the exported HPs would return)
    LAUNCH(c, launch_lbfgsb_export(c->lb.as<LbfgsbState>(), M, P_i, hp_i_dev, P_i, c->fails.as<int>(), c->nfgv.as<int>(),
                                   skip_dev ? nullptr : a.f, c->st));
    std::vector<int> hf(M), hn(M);
    MB_TRY(d2h(hf.data(), c->fails.p, M * 4, c->st));
    MB_TRY(d2h(hn.data(), c->nfgv.p, M * 4, c->st));
    if (fused_ok && !fused_pending) MB_TRY(d2h(c->h_cycles, c->counters.as<int>() + 4096, 2 * sizeof(unsigned long long), c->st));
    CU(cudaStreamSynchronize(c->st));
    int64_t tot = 0, bad = 0;
    for (int64_t t = 0; t < M; ++t) { tot += hn[t]; if (hf[t] == 99) bad++; }
    if (stats) { stats[1] = rounds; stats[2] = tot; stats[3] = bad; }
    c->prof_evals += tot;
  }
  return 0;
}

// ---- Magma -----------------------------------------------------------------------------------
// device state of one EM step: hp_i table, m_0, post_mean, post_cov
static int ensure_em_buffers(mb_context* c, int K) {
  const size_t UU = (size_t)c->U * c->U;
  MB_TRY(c->pm.ensure((size_t)K * c->U * 8));
  MB_TRY(c->pc.ensure(K * UU * 8));
  MB_TRY(c->m0.ensure((size_t)K * c->U * 8));
  MB_TRY(c->f_t.ensure((size_t)c->db.n_tasks * (K > 1 ? K : 1) * 8));
  MB_TRY(c->g_t.ensure((size_t)c->db.n_tasks * MB_MAX_HP * 8));
  MB_TRY(c->ws0.ensure(c->U));
  return 0;
}

extern "C" int mb_e_step(mb_context* c, const mb_kernel* kern_0, const double* hp_0,
                         const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                         const double* m_0, double pen_diag, double* post_mean, double* post_cov,
                         int32_t* info) {
  MB_TRY(need_db(c, true));
  if (!kern_0 || !hp_0 || !kern_i || !hp_i || !m_0 || !post_mean || !post_cov) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(h2d(c->m0.p, m_0, (size_t)c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  MB_TRY(dev_e_step(c, 1, kern_0, hp_0, make_devkern(kern_i), c->hp_i.as<double>(), stride,
                    c->m0.as<double>(), nullptr, c->d_gridX.as<double>(), c->U, nullptr, pen_diag,
                    c->pm.as<double>(), c->pc.as<double>(), info));
  MB_TRY(d2h(post_mean, c->pm.p, (size_t)c->U * 8, c->st));
  MB_TRY(d2h(post_cov, c->pc.p, (size_t)c->U * c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// hyperposterior / hyperposterior_clust core: (V)E step on a caller-given grid (data U grid_inputs)
static int hyperposterior_any(mb_context* c, int K, const mb_kernel* kern_k, const double* hp_k,
                              const mb_kernel* kern_i, const double* hp_i, int hp_i_shared, int32_t n_grid,
                              const double* grid_X, const int32_t* grid_index, const double* m_k,
                              const double* tau, double pen_diag, double* post_mean, double* post_cov) {
  MB_TRY(check_ctx(c));
  if (!c->have_db) MB_FAIL(-1, "no dataset uploaded");
  if (!kern_k || !hp_k || !kern_i || !hp_i || !grid_X || !grid_index || n_grid < 1 || !m_k ||
      ((!post_mean || !post_cov) && !c->hpr_keep))
    MB_FAIL(-1, "bad argument");
  if (K < 1 || K > MB_MAX_CLUSTERS || (K > 1 && !tau)) MB_FAIL(-1, "bad cluster count / missing mixture");
  MB_TRY(check_grid_index(c->h_offs.data(), c->db.n_tasks, grid_index, n_grid));
  const size_t UU = (size_t)n_grid * n_grid;
  DBuf gX, gi, pm, pc, m0, dtau;
  int rc = 0;
  if (gX.ensure((size_t)n_grid * c->db.D * 8) || gi.ensure(c->n_rows * 4) || pm.ensure((size_t)K * n_grid * 8) ||
      pc.ensure(K * UU * 8) || m0.ensure((size_t)K * n_grid * 8) ||
      (tau && dtau.ensure((size_t)c->db.n_tasks * K * 8))) rc = -3;
  if (!rc) {
    cudaMemcpyAsync(gX.p, grid_X, (size_t)n_grid * c->db.D * 8, cudaMemcpyHostToDevice, c->st);
    cudaMemcpyAsync(gi.p, grid_index, c->n_rows * 4, cudaMemcpyHostToDevice, c->st);
    cudaMemcpyAsync(m0.p, m_k, (size_t)K * n_grid * 8, cudaMemcpyHostToDevice, c->st);
    if (tau) cudaMemcpyAsync(dtau.p, tau, (size_t)c->db.n_tasks * K * 8, cudaMemcpyHostToDevice, c->st);
    int64_t stride;
    rc = put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride);
    if (!rc) rc = (cudaStreamSynchronize(c->st) == cudaSuccess) ? 0 : -2;
    if (!rc) rc = dev_e_step(c, K, kern_k, hp_k, make_devkern(kern_i), c->hp_i.as<double>(), stride,
                             m0.as<double>(), tau ? dtau.as<double>() : nullptr, gX.as<double>(), n_grid,
                             gi.as<int32_t>(), pen_diag, pm.as<double>(), pc.as<double>(), nullptr);
    if (!rc) {
      if (post_mean) cudaMemcpyAsync(post_mean, pm.p, (size_t)K * n_grid * 8, cudaMemcpyDeviceToHost, c->st);
      if (post_cov) cudaMemcpyAsync(post_cov, pc.p, K * UU * 8, cudaMemcpyDeviceToHost, c->st);
      if (cudaStreamSynchronize(c->st) != cudaSuccess) { mb_set_error(__FILE__, __LINE__, "copy back failed"); rc = -2; }
    }
    if (!rc && c->hpr_keep) {   // the result stays on the device for the new-task entry points (no G x G round trip)
      c->hpr_mean.release(); c->hpr_cov.release();
      c->hpr_mean = pm; c->hpr_cov = pc;
      pm = DBuf(); pc = DBuf();
      c->hpr_K = K; c->hpr_U = n_grid; c->hpr_has_cov = true;
    }
  }
  gX.release(); gi.release(); pm.release(); pc.release(); m0.release(); dtau.release();
  return rc;
}

extern "C" int mb_hyperposterior(mb_context* c, const mb_kernel* kern_0, const double* hp_0,
                                 const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                                 int32_t n_grid, const double* grid_X, const int32_t* grid_index,
                                 const double* m_0, double pen_diag, double* post_mean,
                                 double* post_cov) {
  return hyperposterior_any(c, 1, kern_0, hp_0, kern_i, hp_i, hp_i_shared, n_grid, grid_X, grid_index, m_0,
                            nullptr, pen_diag, post_mean, post_cov);
}

extern "C" int mb_hyperposterior_clust(mb_context* c, int n_clusters, const mb_kernel* kern_k, const double* hp_k,
                                       const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                                       int32_t n_grid, const double* grid_X, const int32_t* grid_index,
                                       const double* m_k, const double* tau, double pen_diag,
                                       double* post_mean, double* post_cov) {
  return hyperposterior_any(c, n_clusters, kern_k, hp_k, kern_i, hp_i, hp_i_shared, n_grid, grid_X, grid_index,
                            m_k, tau, pen_diag, post_mean, post_cov);
}

// ================================================================================================
// New tasks against a resident hyper-posterior: logL_GP / gr_GP (marginal forms), train_gp and pred_magma(clust)
// batched over the tasks of the resident dataset.
// ================================================================================================
extern "C" int mb_hyperpost_keep(mb_context* c, int keep) {
  MB_TRY(check_ctx(c));
  c->hpr_keep = keep != 0;
  if (keep < 0) { c->hpr_mean.release(); c->hpr_cov.release(); c->hpr_K = c->hpr_U = 0; c->hpr_keep = false; }
  return 0;
}

extern "C" int mb_hyperpost_set(mb_context* c, int n_clusters, int32_t n_grid, const double* post_mean,
                                const double* post_cov) {
  MB_TRY(check_ctx(c));
  if (n_clusters < 1 || n_clusters > MB_MAX_CLUSTERS || n_grid < 1 || !post_mean) MB_FAIL(-1, "bad argument");
  const size_t UU = (size_t)n_grid * n_grid;
  MB_TRY(c->hpr_mean.ensure((size_t)n_clusters * n_grid * 8));
  MB_TRY(h2d(c->hpr_mean.p, post_mean, (size_t)n_clusters * n_grid * 8, c->st));
  if (post_cov) {
    MB_TRY(c->hpr_cov.ensure(n_clusters * UU * 8));
    MB_TRY(h2d(c->hpr_cov.p, post_cov, n_clusters * UU * 8, c->st));
  }
  CU(cudaStreamSynchronize(c->st));
  c->hpr_K = n_clusters; c->hpr_U = n_grid; c->hpr_has_cov = post_cov != nullptr;
  return 0;
}

static int need_hyperpost(mb_context* c, int cluster, bool need_cov) {
  MB_TRY(need_db(c, true));
  if (c->hpr_K < 1) MB_FAIL(-1, "no resident hyper-posterior (mb_hyperpost_set / mb_hyperpost_keep)");
  if (c->hpr_U != c->U) MB_FAIL(-1, "the resident dataset's grid is not the hyper-posterior's grid");
  if (cluster < 0 || cluster >= c->hpr_K) MB_FAIL(-1, "cluster out of range");
  if (need_cov && !c->hpr_has_cov) MB_FAIL(-1, "this entry point needs the hyper-posterior covariance");
  if (c->db.nmax > TK_MAXN) MB_FAIL(-4, "new tasks with more than 96 points are not supported by the batched new-task path");
  return 0;
}

static TaskObjArgs marg_args(mb_context* c, const DevKern& kd, const double* hp_dev, int64_t stride, int cluster,
                             int want_grad, double pen) {
  const size_t UU = (size_t)c->U * c->U;
  return make_obj_args(c, kd, hp_dev, stride, TK_OBJ_MARG, want_grad, pen, 1,
                       c->hpr_mean.as<double>() + (size_t)cluster * c->U,
                       c->hpr_has_cov ? c->hpr_cov.as<double>() + (size_t)cluster * UU : nullptr, nullptr,
                       c->f_t.as<double>(), c->g_t.as<double>());
}

extern "C" int mb_logL_GP_tasks(mb_context* c, const mb_kernel* kern, const double* hp, int hp_shared, int cluster,
                                double pen_diag, double* f, double* g, int32_t* retries) {
  MB_TRY(need_hyperpost(c, cluster, false));
  if (!kern || !hp || !f) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  int64_t stride;
  MB_TRY(put_hp_i(c, kern, hp, hp_shared, &stride));
  DevKern kd = make_devkern(kern);
  TaskObjArgs a = marg_args(c, kd, c->hp_i.as<double>(), stride, cluster, g != nullptr, pen_diag);
  MB_TRY(c->retries.ensure(c->db.n_tasks * 4));
  a.retries = c->retries.as<int32_t>();
  MB_TRY(launch_obj(c, a, (int)c->db.n_tasks, c->st));
  MB_TRY(d2h(f, c->f_t.p, c->db.n_tasks * 8, c->st));
  if (g) MB_TRY(d2h(g, c->g_t.p, (size_t)c->db.n_tasks * kd.n_hp * 8, c->st));
  if (retries) MB_TRY(d2h(retries, c->retries.p, c->db.n_tasks * 4, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// train_shared_gp (/root/reference/R/training.R:352-446; the default initialisation through precondition_hp,
// R/preconditioning.R:24-39): ONE HP vector shared by all resident tasks, optim(par, fn = sum_logL_GP, method = "L-BFGS-B",
// factr = 1e13, maxit = 100) with NO analytic gradient -- optim's central differences (R src/library/stats/src/optim.c
// fmingr: ndeps = 1e-3, df_i = (fn(x + eps e_i) - fn(x - eps e_i)) / (2 eps)), i.e. 1 + 2 P objective evaluations per
// optimiser evaluation, each ONE launch over all tasks.  sum_logL_GP's terms are logL_GP with post_cov = 0 and a constant
// mean (likelihoods.R:113-129), which is logL_GP_mod with S = 0: the DMMA kernel of the M step evaluates them.
extern "C" int mb_train_shared_gp(mb_context* c, const mb_kernel* kern, double* hp, double mean, double pen_diag,
                                  int maxit, int64_t* stats) {
  MB_TRY(need_db(c, true));
  if (!kern || !hp || maxit < 1) MB_FAIL(-1, "bad argument");
  MB_TRY(ensure_em_buffers(c, 1));
  const int P = kern->n_hp;
  DevKern kd = make_devkern(kern);
  CU(cudaMemsetAsync(c->pc.p, 0, (size_t)c->U * c->U * 8, c->st));
  LAUNCH(c, launch_fill(c->U, mean, c->pm.as<double>(), c->st));
  TaskObjArgs a = make_obj_args(c, kd, nullptr, 0, TK_OBJ_MOD, 0, pen_diag, 1, c->pm.as<double>(), c->pc.as<double>(),
                                nullptr, c->f_t.as<double>(), c->g_t.as<double>());
  LbfgsbState s;
  lb_init(&s, P, hp, 1e13, 0.0, maxit);
  const double ndeps = 1e-3;
  int64_t n_fn = 0;
  while (s.active) {
    double f, g[MB_MAX_HP], x[MB_MAX_HP], v1, v2;
    for (int p = 0; p < P; ++p) x[p] = s.x[p];
    MB_TRY(shared_obj_eval(c, a, x, P, 0, &f, nullptr));
    for (int p = 0; p < P; ++p) {
      x[p] = s.x[p] + ndeps;
      MB_TRY(shared_obj_eval(c, a, x, P, 0, &v1, nullptr));
      x[p] = s.x[p] - ndeps;
      MB_TRY(shared_obj_eval(c, a, x, P, 0, &v2, nullptr));
      g[p] = (v1 - v2) / (2 * ndeps);
      x[p] = s.x[p];
    }
    n_fn += 1 + 2 * P;
    lb_advance(&s, f, g);
  }
  bool bad = false;
  for (int p = 0; p < P; ++p) if (!std::isfinite(s.x[p])) bad = true;
  if (!bad) for (int p = 0; p < P; ++p) hp[p] = s.x[p];     // training.R:437-443: NA -> the initial values are returned
  if (stats) { stats[0] = s.nfgv; stats[1] = s.fail; stats[2] = n_fn * c->db.n_tasks; stats[3] = bad; }
  return 0;
}

extern "C" int mb_train_gp_tasks(mb_context* c, const mb_kernel* kern, double* hp, int cluster, double pen_diag,
                                 int64_t* stats) {
  MB_TRY(need_hyperpost(c, cluster, false));
  if (!kern || !hp) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  const int P = kern->n_hp;
  const int64_t M = c->db.n_tasks;
  MB_TRY(c->hp_k.ensure((size_t)M * P * 8));
  MB_TRY(h2d(c->hp_k.p, hp, (size_t)M * P * 8, c->st));
  DevKern kd = make_devkern(kern);
  TaskObjArgs a = marg_args(c, kd, c->hp_k.as<double>(), P, cluster, 1, pen_diag);
  std::vector<SingleProblem> none;
  int64_t st4[4] = {0, 0, 0, 0};
  MB_TRY(optimise(c, none, true, a, c->hp_k.as<double>(), P, pen_diag, st4));
  // training.R:277-284: a task whose optimisation produced a non-finite HP returns its initial values
  std::vector<double> out((size_t)M * P);
  MB_TRY(d2h(out.data(), c->hp_k.p, (size_t)M * P * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  for (int64_t t = 0; t < M; ++t) {
    bool bad = false;
    for (int p = 0; p < P; ++p) if (!std::isfinite(out[t * P + p])) bad = true;
    if (!bad) for (int p = 0; p < P; ++p) hp[t * P + p] = out[t * P + p];
  }
  if (stats) for (int i = 0; i < 4; ++i) stats[i] = st4[i];
  return 0;
}

// train_gp_clust (/root/reference/R/training.R:1923-2217) for every resident task at once.  Each task runs the
// reference's own EM loop -- M step: optim(L-BFGS-B) of sum_logL_GP_clust / gr_sum_logL_GP_clust with its current mixture;
// E step: update_mixture with prop_mixture; monitoring: sum_logL_GP_clust(new_hp, new_mixture, prop_mixture) -- and stops
// at its own iteration (|eps| < cv_threshold, a non-finite HP, or n_iter_max); tasks that have stopped are masked out of
// the following lock-step M steps, so every task's iterates are those of a stand-alone call.
extern "C" int mb_train_gp_clust_tasks(mb_context* c, const mb_kernel* kern, double* hp, const double* prop_mixture,
                                       double pen_diag, int n_iter_max, double cv_threshold, double* mixture_out,
                                       int32_t* n_iter_out, int32_t* converged_out, int64_t* stats);

extern "C" int mb_pred_magma_tasks(mb_context* c, const mb_kernel* kern, const double* hp, const double* tau,
                                   int32_t n_pred, const double* x_pred, const int32_t* pred_index, double pen_diag,
                                   double* pred_mean, double* pred_var, double* pred_mean_k, double* pred_var_k,
                                   int32_t* retries) {
  if (kern && kern->n_terms == 1 && kern->types[0] == MB_KERN_CONV) MB_FAIL(-4, "prediction with the multi-output convolution kernel is not on the accelerated path yet");
  MB_TRY(need_hyperpost(c, 0, true));
  const int K = c->hpr_K, D = c->db.D, P = kern ? kern->n_hp : 0;
  const int64_t M = c->db.n_tasks;
  if (!kern || !hp || !x_pred || !pred_index || !pred_mean || !pred_var || n_pred < 1 || (K > 1 && !tau) ||
      ((pred_mean_k == nullptr) != (pred_var_k == nullptr)))
    MB_FAIL(-1, "bad argument");
  for (int32_t g = 0; g < n_pred; ++g)
    if (pred_index[g] < 0 || pred_index[g] >= c->U) MB_FAIL(-1, "pred_index out of range");
  const int nmax = c->db.nmax;
  const size_t per_task = (size_t)K * ((size_t)nmax * nmax + nmax);
  // tasks are processed in chunks so that the Gamma^-1 table and the outputs stay bounded
  int64_t chunk = (int64_t)(((size_t)1 << 30) / (per_task * 8 + (size_t)n_pred * 8 * (2 + (pred_mean_k ? 2 * K : 0))));
  if (const char* ec = getenv("MB_PRED_CHUNK")) { if (atoll(ec) >= 1) chunk = atoll(ec); }   // test hook: force several chunks
  if (chunk < 1) chunk = 1;
  if (chunk > M) chunk = M;
  if (chunk > 65535) chunk = 65535;
  DBuf d_hp, d_tau, d_xp, d_pi, d_g, d_m, d_v, d_mk, d_vk, d_ret;
  int rc = 0;
  if (d_hp.ensure((size_t)M * P * 8) || (K > 1 && d_tau.ensure((size_t)M * K * 8)) || d_xp.ensure((size_t)n_pred * D * 8) ||
      d_pi.ensure((size_t)n_pred * 4) || d_g.ensure((size_t)chunk * per_task * 8) || d_m.ensure((size_t)M * n_pred * 8) ||
      d_v.ensure((size_t)M * n_pred * 8) || d_ret.ensure((size_t)M * K * 4) ||
      (pred_mean_k && (d_mk.ensure((size_t)M * K * n_pred * 8) || d_vk.ensure((size_t)M * K * n_pred * 8)))) rc = -3;
  auto cleanup = [&]() { d_hp.release(); d_tau.release(); d_xp.release(); d_pi.release(); d_g.release(); d_m.release();
                         d_v.release(); d_mk.release(); d_vk.release(); d_ret.release(); };
  if (rc) { cleanup(); return rc; }
  cudaMemcpyAsync(d_hp.p, hp, (size_t)M * P * 8, cudaMemcpyHostToDevice, c->st);
  if (K > 1) cudaMemcpyAsync(d_tau.p, tau, (size_t)M * K * 8, cudaMemcpyHostToDevice, c->st);
  cudaMemcpyAsync(d_xp.p, x_pred, (size_t)n_pred * D * 8, cudaMemcpyHostToDevice, c->st);
  cudaMemcpyAsync(d_pi.p, pred_index, (size_t)n_pred * 4, cudaMemcpyHostToDevice, c->st);
  TaskPredArgs a;
  memset(&a, 0, sizeof(a));
  a.db = c->db; a.kd = make_devkern(kern); a.hp = d_hp.as<double>(); a.hp_stride = P;
  a.hpst.K = K; a.hpst.U = c->U; a.hpst.mean = c->hpr_mean.as<double>(); a.hpst.cov = c->hpr_cov.as<double>();
  a.tau = K > 1 ? d_tau.as<double>() : nullptr; a.pen = pen_diag; a.n_pred = n_pred;
  a.x_pred = d_xp.as<double>(); a.pred_index = d_pi.as<int32_t>(); a.ginv = d_g.as<double>();
  a.out_mean = d_m.as<double>(); a.out_var = d_v.as<double>();
  a.out_mean_k = pred_mean_k ? d_mk.as<double>() : nullptr; a.out_var_k = pred_var_k ? d_vk.as<double>() : nullptr;
  a.retries = d_ret.as<int32_t>();
  cudaError_t e = cudaSuccess;
  for (int64_t t0 = 0; t0 < M && e == cudaSuccess; t0 += chunk) {
    a.t0 = t0;
    e = launch_task_pred(a, (int)std::min<int64_t>(chunk, M - t0), c->st);
    c->launches += 2;
  }
  std::vector<int32_t> hret((size_t)M * K);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pred_mean, d_m.p, (size_t)M * n_pred * 8, cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(pred_var, d_v.p, (size_t)M * n_pred * 8, cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess && pred_mean_k) e = cudaMemcpyAsync(pred_mean_k, d_mk.p, (size_t)M * K * n_pred * 8, cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess && pred_var_k) e = cudaMemcpyAsync(pred_var_k, d_vk.p, (size_t)M * K * n_pred * 8, cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess) e = cudaMemcpyAsync(hret.data(), d_ret.p, (size_t)M * K * 4, cudaMemcpyDeviceToHost, c->st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->st);
  cleanup();
  if (e != cudaSuccess) { mb_set_error(__FILE__, __LINE__, cudaGetErrorString(e)); return -2; }
  int bad = 0;
  for (size_t i = 0; i < hret.size(); ++i) { if (retries) retries[i] = hret[i]; if (hret[i] < 0) bad = 1; }
  if (bad) MB_FAIL(1, "chol_inv_jitter: no positive-definite jitter found for at least one new task");
  return 0;
}

extern "C" int mb_logL_GP_mod_tasks(mb_context* c, const mb_kernel* kern_i, const double* hp_i,
                                    int hp_i_shared, const double* post_mean, const double* post_cov,
                                    double pen_diag, double* f, double* g, int32_t* retries) {
  MB_TRY(need_db(c, true));
  if (!kern_i || !hp_i || !post_mean || !post_cov || !f) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(h2d(c->pm.p, post_mean, (size_t)c->U * 8, c->st));
  MB_TRY(h2d(c->pc.p, post_cov, (size_t)c->U * c->U * 8, c->st));
  DevKern kd = make_devkern(kern_i);
  TaskObjArgs a = make_obj_args(c, kd, c->hp_i.as<double>(), stride, TK_OBJ_MOD, g != nullptr, pen_diag, 1,
                                c->pm.as<double>(), c->pc.as<double>(), nullptr, c->f_t.as<double>(),
                                c->g_t.as<double>());
  MB_TRY(c->retries.ensure(c->db.n_tasks * 4));
  a.retries = c->retries.as<int32_t>();
  MB_TRY(launch_obj(c, a, (int)c->db.n_tasks, c->st));
  MB_TRY(d2h(f, c->f_t.p, c->db.n_tasks * 8, c->st));
  if (g) MB_TRY(d2h(g, c->g_t.p, (size_t)c->db.n_tasks * kd.n_hp * 8, c->st));
  if (retries) MB_TRY(d2h(retries, c->retries.p, c->db.n_tasks * 4, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int mb_logL_GP_mod(mb_context* c, const mb_kernel* k, const double* hp, const double* x,
                              int n, int d, const double* y, const double* mean,
                              const double* post_cov, double pen_diag, double* f, double* g) {
  MB_TRY(check_ctx(c));
  if (!k || !hp || !x || !y || !mean || !post_cov || !f || n < 1 || d < 1) MB_FAIL(-1, "bad argument");
  MB_TRY(c->ws0.ensure(n));
  MB_TRY(c->scratch.ensure(((size_t)n * d + 2 * (size_t)n) * 8 + 64));
  double* dX = c->scratch.as<double>();
  double* dy = dX + (size_t)n * d;
  double* dm = dy + n;
  MB_TRY(h2d(dX, x, (size_t)n * d * 8, c->st2));
  MB_TRY(h2d(dy, y, (size_t)n * 8, c->st2));
  MB_TRY(h2d(dm, mean, (size_t)n * 8, c->st2));
  MB_TRY(h2d(c->ws0.S.p, post_cov, (size_t)n * n * 8, c->st2));
  double* hres = c->h_pin + 32;
  int* hinfo = c->h_pin_i + 8;
  MB_TRY(dense_obj_enqueue(c, c->ws0, make_devkern(k), hp, dX, n, d, dy, dm, c->ws0.S.as<double>(),
                           pen_diag, g != nullptr, c->st2, hres, hinfo));
  CU(cudaStreamSynchronize(c->st2));
  if (hinfo[0] < 0) MB_FAIL(1, "chol_inv_jitter: no positive-definite jitter found");
  *f = hres[0];
  if (g) for (int p = 0; p < k->n_hp; ++p) g[p] = hres[1 + p];
  return 0;
}

// M step on device state: hp_0 (host vector, in/out), hp_i table on device (in/out)
static int dev_m_step(mb_context* c, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                      double* hp_i_dev, int shared_hp, const double* m0_dev, const double* pm_dev,
                      const double* pc_dev, double pen, int64_t* stats) {
  DevKern kd0 = make_devkern(kern_0), kdi = make_devkern(kern_i);
  if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
  std::vector<SingleProblem> singles(1);
  SingleProblem& sp = singles[0];
  sp.kd = kd0; sp.X = c->d_gridX.as<double>(); sp.n = c->U; sp.D = c->db.D;
  sp.y = pm_dev; sp.mean = m0_dev; sp.S = pc_dev;
  lb_init(&sp.s, kd0.n_hp, hp_0, 1e13, 0.0, 25);
  TaskObjArgs a = make_obj_args(c, kdi, hp_i_dev, kdi.n_hp, TK_OBJ_MOD, 1, pen, 1, pm_dev, pc_dev, nullptr,
                                c->f_t.as<double>(), c->g_t.as<double>());
  c->mstep_values_valid = false;
  if (!shared_hp) {
    MB_TRY(optimise(c, singles, true, a, hp_i_dev, kdi.n_hp, pen, stats));
    c->mstep_values_valid = sp.evals > 0;
    c->m_step_f0 = sp.s.f;
  } else {
    MB_TRY(optimise(c, singles, false, a, hp_i_dev, kdi.n_hp, pen, stats));
    // one optimisation of the summed objective (em-magma.R:172-196), HPs of the first task
    LbfgsbState s;
    double x0[MB_MAX_HP];
    MB_TRY(d2h(x0, hp_i_dev, kdi.n_hp * 8, c->st));
    CU(cudaStreamSynchronize(c->st));
    lb_init(&s, kdi.n_hp, x0, 1e13, 0.0, 25);
    int evals = 0;
    while (s.active) {
      double f, g[MB_MAX_HP];
      MB_TRY(shared_obj_eval(c, a, s.x, kdi.n_hp, 1, &f, g));
      lb_advance(&s, f, g);
      evals++;
    }
    std::vector<double> tab((size_t)c->db.n_tasks * kdi.n_hp);
    for (int64_t t = 0; t < c->db.n_tasks; ++t)
      for (int p = 0; p < kdi.n_hp; ++p) tab[t * kdi.n_hp + p] = s.x[p];
    MB_TRY(h2d(hp_i_dev, tab.data(), tab.size() * 8, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (stats) { stats[1] = evals; stats[2] = (int64_t)evals * c->db.n_tasks; stats[3] = s.fail != 0; }
  }
  for (int p = 0; p < kd0.n_hp; ++p) hp_0[p] = sp.s.x[p];
  if (stats) stats[0] = sp.evals;
  return 0;
}

// reuse_m_step: the caller has just run the per-task M step on the SAME hyper-posterior and passes its results on: the
// values logL_GP_mod(hp_i_new) and logL_GP_mod(hp_0_new) that likelihoods.R:262-300 evaluates again are the objective values
// the optimisers ended with (lbfgsb.h keeps (x, f) consistent on every exit), so they are summed as they are instead of
// being recomputed -- same numbers, one all-task launch and one union-grid evaluation less per EM iteration.
static int dev_logL_monitoring(mb_context* c, const mb_kernel* kern_0, const double* hp_0,
                               const mb_kernel* kern_i, const double* hp_i_dev, int64_t stride,
                               const double* m0_dev, const double* pm_dev, const double* pc_dev,
                               double pen, double* logL, bool reuse_m_step = false) {
  DevKern kd0 = make_devkern(kern_0), kdi = make_devkern(kern_i);
  double* hres = c->h_pin + 32;
  int* hinfo = c->h_pin_i + 8;
  TaskObjArgs a = make_obj_args(c, kdi, hp_i_dev, stride, TK_OBJ_MOD, 0, pen, 1, pm_dev, pc_dev, nullptr,
                                c->f_t.as<double>(), c->g_t.as<double>());
  if (!reuse_m_step) {
    MB_TRY(dense_obj_enqueue(c, c->ws0, kd0, hp_0, c->d_gridX.as<double>(), c->U, c->db.D, pm_dev, m0_dev,
                             pc_dev, pen, 0, c->st2, hres, hinfo));
    MB_TRY(launch_obj(c, a, (int)c->db.n_tasks, c->st));
  }
  MB_TRY(c->sums.ensure(64 * 8));
  MB_TRY(leaf_sum_columns(c, a.f, 1, c->sums.as<double>(), c->st));
  // det term: chol(post_cov) without jitter (likelihoods.R:303-307)
  LAUNCH(c, launch_chol_logdiag_any(c, c->ws1, pc_dev, c->U, c->ws1.info.as<int>(), c->sums.as<double>() + 8, c->st));
  MB_TRY(d2h(c->h_pin, c->sums.p, 16 * 8, c->st));
  MB_TRY(d2h(c->h_pin_i, c->ws1.info.p, sizeof(int), c->st));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaStreamSynchronize(c->st2));
  if (c->h_pin_i[0] != 0) MB_FAIL(1, "chol(post_cov) failed in logL_monitoring (no jitter is applied there)");
  double ll_0 = reuse_m_step ? c->m_step_f0 : ((hinfo[0] < 0) ? NAN : hres[0]);
  *logL = -ll_0 - c->h_pin[0] + c->h_pin[8];
  return 0;
}

extern "C" int mb_m_step(mb_context* c, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                         double* hp_i, int shared_hp, const double* m_0, const double* post_mean,
                         const double* post_cov, double pen_diag, int64_t* stats) {
  MB_TRY(need_db(c, true));
  if (!kern_0 || !hp_0 || !kern_i || !hp_i || !m_0 || !post_mean || !post_cov) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  const int P = kern_i->n_hp;
  MB_TRY(c->hp_k.ensure((size_t)c->db.n_tasks * P * 8));
  MB_TRY(h2d(c->hp_k.p, hp_i, (size_t)c->db.n_tasks * P * 8, c->st));
  MB_TRY(h2d(c->m0.p, m_0, (size_t)c->U * 8, c->st));
  MB_TRY(h2d(c->pm.p, post_mean, (size_t)c->U * 8, c->st));
  MB_TRY(h2d(c->pc.p, post_cov, (size_t)c->U * c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  MB_TRY(dev_m_step(c, kern_0, hp_0, kern_i, c->hp_k.as<double>(), shared_hp, c->m0.as<double>(),
                    c->pm.as<double>(), c->pc.as<double>(), pen_diag, stats));
  MB_TRY(d2h(hp_i, c->hp_k.p, (size_t)c->db.n_tasks * P * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int mb_last_task_status(mb_context* c, int32_t* convergence, int32_t* n_eval) {
  MB_TRY(need_db(c, false));
  if (!c->fails.p || !c->nfgv.p) MB_FAIL(-1, "no per-task M step has run yet");
  if (convergence) MB_TRY(d2h(convergence, c->fails.p, c->db.n_tasks * 4, c->st));
  if (n_eval) MB_TRY(d2h(n_eval, c->nfgv.p, c->db.n_tasks * 4, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int mb_logL_monitoring(mb_context* c, const mb_kernel* kern_0, const double* hp_0,
                                  const mb_kernel* kern_i, const double* hp_i, int hp_i_shared,
                                  const double* m_0, const double* post_mean, const double* post_cov,
                                  double pen_diag, double* logL) {
  MB_TRY(need_db(c, true));
  if (!kern_0 || !hp_0 || !kern_i || !hp_i || !m_0 || !post_mean || !post_cov || !logL) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  MB_TRY(c->ws1.ensure(c->U));
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(h2d(c->m0.p, m_0, (size_t)c->U * 8, c->st));
  MB_TRY(h2d(c->pm.p, post_mean, (size_t)c->U * 8, c->st));
  MB_TRY(h2d(c->pc.p, post_cov, (size_t)c->U * c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return dev_logL_monitoring(c, kern_0, hp_0, kern_i, c->hp_i.as<double>(), stride, c->m0.as<double>(),
                             c->pm.as<double>(), c->pc.as<double>(), pen_diag, logL);
}

// e_step + m_step + logL_monitoring with every intermediate kept on the device
static int dev_em_step(mb_context* c, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                       double* hp_i_dev, int shared_hp, const double* m0_dev, double pen, double* logL,
                       int64_t* stats) {
  DevKern kdi = make_devkern(kern_i);
  // every phase ends with the host waiting for its result, so host time stamps are stream time (mb_profile only)
  auto now = [] { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  const double t0 = c->prof_on ? now() : 0.0;
  MB_TRY(dev_e_step(c, 1, kern_0, hp_0, kdi, hp_i_dev, kdi.n_hp, m0_dev, nullptr, c->d_gridX.as<double>(),
                    c->U, nullptr, pen, c->pm.as<double>(), c->pc.as<double>(), nullptr));
  const double t1 = c->prof_on ? now() : 0.0;
  MB_TRY(dev_m_step(c, kern_0, hp_0, kern_i, hp_i_dev, shared_hp, m0_dev, c->pm.as<double>(),
                    c->pc.as<double>(), pen, stats));
  const double t2 = c->prof_on ? now() : 0.0;
  MB_TRY(dev_logL_monitoring(c, kern_0, hp_0, kern_i, hp_i_dev, kdi.n_hp, m0_dev, c->pm.as<double>(),
                             c->pc.as<double>(), pen, logL, c->mstep_values_valid && !c->monitor_recompute));
  c->mstep_values_valid = false;
  if (c->prof_on) {
    const double t3 = now();
    c->prof_phase[0] += t1 - t0; c->prof_phase[1] += t2 - t1; c->prof_phase[2] += t3 - t2;
  }
  return 0;
}

extern "C" int mb_em_step(mb_context* c, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                          double* hp_i, int shared_hp, const double* m_0, double pen_diag, double* logL,
                          double* post_mean, double* post_cov, int64_t* stats) {
  MB_TRY(need_db(c, true));
  if (!kern_0 || !hp_0 || !kern_i || !hp_i || !m_0 || !logL) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  const int P = kern_i->n_hp;
  MB_TRY(c->hp_k.ensure((size_t)c->db.n_tasks * P * 8));
  MB_TRY(h2d(c->hp_k.p, hp_i, (size_t)c->db.n_tasks * P * 8, c->st));
  MB_TRY(h2d(c->m0.p, m_0, (size_t)c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  MB_TRY(dev_em_step(c, kern_0, hp_0, kern_i, c->hp_k.as<double>(), shared_hp, c->m0.as<double>(), pen_diag, logL, stats));
  MB_TRY(d2h(hp_i, c->hp_k.p, (size_t)c->db.n_tasks * P * 8, c->st));
  if (post_mean) MB_TRY(d2h(post_mean, c->pm.p, (size_t)c->U * 8, c->st));
  if (post_cov) MB_TRY(d2h(post_cov, c->pc.p, (size_t)c->U * c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int mb_train_magma(mb_context* c, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                              double* hp_i, int shared_hp, const double* m_0, double pen_diag,
                              int n_iter_max, double cv_threshold, double* seq_logL, int32_t* n_iter,
                              int32_t* converged, double* post_mean, double* post_cov) {
  MB_TRY(need_db(c, true));
  if (!kern_0 || !hp_0 || !kern_i || !hp_i || !m_0 || !seq_logL || !n_iter || !converged) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  const int P = kern_i->n_hp, P0 = kern_0->n_hp;
  const size_t tab = (size_t)c->db.n_tasks * P * 8;
  MB_TRY(c->hp_k.ensure(2 * tab));
  double* cur = c->hp_k.as<double>();
  double* nxt = cur + (size_t)c->db.n_tasks * P;
  MB_TRY(h2d(cur, hp_i, tab, c->st));
  MB_TRY(h2d(c->m0.p, m_0, (size_t)c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  double ll = -INFINITY;
  *converged = 0;
  *n_iter = 0;
  std::vector<double> htab((size_t)c->db.n_tasks * P);
  for (int it = 1; it <= n_iter_max; ++it) {
    // work on a copy so that a NaN M step leaves the last valid HPs in place (training.R:956-963)
    CU(cudaMemcpyAsync(nxt, cur, tab, cudaMemcpyDeviceToDevice, c->st));
    double h0[MB_MAX_HP];
    for (int p = 0; p < P0; ++p) h0[p] = hp_0[p];
    double new_ll;
    MB_TRY(dev_em_step(c, kern_0, h0, kern_i, nxt, shared_hp, c->m0.as<double>(), pen_diag, &new_ll, nullptr));
    MB_TRY(d2h(htab.data(), nxt, tab, c->st));
    CU(cudaStreamSynchronize(c->st));
    bool bad = false;
    for (int p = 0; p < P0; ++p) if (isnan(h0[p])) bad = true;
    for (double v : htab) if (isnan(v)) { bad = true; break; }
    // every rank takes the same branch: a NaN anywhere stops the fit everywhere (the ranks meet again in the next all-reduce)
    if (c->world > 1) {
      MB_TRY(c->sums.ensure(64 * 8));
      const double flag = bad ? 1.0 : 0.0;
      double tot = 0.0;
      CU(cudaMemcpyAsync(c->sums.as<double>() + 33, &flag, 8, cudaMemcpyHostToDevice, c->st));
      CU(cudaStreamSynchronize(c->st));
      MB_TRY(ctx_allreduce(c, c->sums.as<double>() + 33, 1, c->st));
      CU(cudaMemcpyAsync(&tot, c->sums.as<double>() + 33, 8, cudaMemcpyDeviceToHost, c->st));
      CU(cudaStreamSynchronize(c->st));
      bad = tot != 0.0;
    }
    if (bad) break;
    double diff = new_ll - ll;
    if (isnan(diff)) diff = -INFINITY;
    for (int p = 0; p < P0; ++p) hp_0[p] = h0[p];
    std::swap(cur, nxt);
    ll = new_ll;
    seq_logL[it - 1] = ll;
    *n_iter = it;
    double eps = diff / fabs(ll);
    if (isnan(eps)) eps = 1.0;
    if (fabs(eps) < cv_threshold) { *converged = 1; break; }
  }
  MB_TRY(d2h(hp_i, cur, tab, c->st));
  if (post_mean) MB_TRY(d2h(post_mean, c->pm.p, (size_t)c->U * 8, c->st));
  if (post_cov) MB_TRY(d2h(post_cov, c->pc.p, (size_t)c->U * c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// ---- resident state, timers, profiling -----------------------------------------------------------
extern "C" int mb_set_state(mb_context* c, const mb_kernel* kern_0, const double* hp_0,
                            const mb_kernel* kern_i, const double* hp_i, const double* m_0) {
  MB_TRY(need_db(c, true));
  if (!kern_0 || !hp_0 || !kern_i || !hp_i || !m_0) MB_FAIL(-1, "null argument");
  MB_TRY(ensure_em_buffers(c, 1));
  const size_t tab = (size_t)c->db.n_tasks * kern_i->n_hp * 8;
  MB_TRY(c->st_hp_i0.ensure(tab));
  MB_TRY(c->st_hp_i.ensure(tab));
  MB_TRY(h2d(c->st_hp_i0.p, hp_i, tab, c->st));
  MB_TRY(h2d(c->st_hp_i.p, hp_i, tab, c->st));
  MB_TRY(h2d(c->m0.p, m_0, (size_t)c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  for (int p = 0; p < kern_0->n_hp; ++p) c->st_hp_00[p] = c->st_hp_0[p] = hp_0[p];
  c->have_state = true;
  return 0;
}

extern "C" int mb_em_step_resident(mb_context* c, const mb_kernel* kern_0, const mb_kernel* kern_i,
                                   int shared_hp, double pen_diag, int restart, double* logL,
                                   int64_t* stats) {
  MB_TRY(need_db(c, true));
  if (!c->have_state) MB_FAIL(-1, "mb_set_state was not called");
  if (!kern_0 || !kern_i || !logL) MB_FAIL(-1, "null argument");
  const size_t tab = (size_t)c->db.n_tasks * kern_i->n_hp * 8;
  if (restart) {
    CU(cudaMemcpyAsync(c->st_hp_i.p, c->st_hp_i0.p, tab, cudaMemcpyDeviceToDevice, c->st));
    for (int p = 0; p < kern_0->n_hp; ++p) c->st_hp_0[p] = c->st_hp_00[p];
  }
  return dev_em_step(c, kern_0, c->st_hp_0, kern_i, c->st_hp_i.as<double>(), shared_hp, c->m0.as<double>(),
                     pen_diag, logL, stats);
}

extern "C" int mb_get_state(mb_context* c, const mb_kernel* kern_0, double* hp_0, const mb_kernel* kern_i,
                            double* hp_i) {
  MB_TRY(need_db(c, true));
  if (!c->have_state) MB_FAIL(-1, "mb_set_state was not called");
  for (int p = 0; p < kern_0->n_hp; ++p) hp_0[p] = c->st_hp_0[p];
  MB_TRY(d2h(hp_i, c->st_hp_i.p, (size_t)c->db.n_tasks * kern_i->n_hp * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

extern "C" int mb_timer(mb_context* c, int which) {
  MB_TRY(check_ctx(c));
  CU(cudaEventRecord(which ? c->ev_t1 : c->ev_t0, c->st));
  return 0;
}
extern "C" int mb_timer_ms(mb_context* c, double* ms) {
  MB_TRY(check_ctx(c));
  float f = 0;
  CU(cudaEventSynchronize(c->ev_t1));
  CU(cudaEventElapsedTime(&f, c->ev_t0, c->ev_t1));
  *ms = f;
  return 0;
}
extern "C" int mb_profile(mb_context* c, int enable) {
  MB_TRY(check_ctx(c));
  c->prof_on = enable != 0;
  return 0;
}
extern "C" int mb_profile_read(mb_context* c, double* out3) {
  MB_TRY(check_ctx(c));
  MB_TRY(prof_collect(c));
  out3[0] = c->prof_ms; out3[1] = (double)c->prof_launches; out3[2] = (double)c->prof_evals;
  c->prof_ms = 0; c->prof_launches = 0; c->prof_evals = 0;
  return 0;
}
extern "C" int mb_profile_read_gaps(mb_context* c, double* out2) {
  MB_TRY(check_ctx(c));
  MB_TRY(prof_collect(c));
  out2[0] = c->prof_gap_ms; out2[1] = c->prof_other_ms;
  c->prof_gap_ms = 0; c->prof_other_ms = 0;
  return 0;
}
// Call AFTER mb_profile_read: host-clock milliseconds dev_em_step spent in the E step, the M step and the monitoring
// (each ends with a synchronisation) and the device milliseconds of the persistent M-step kernel, since the last read
extern "C" int mb_profile_read_phases(mb_context* c, double* out12) {
  MB_TRY(check_ctx(c));
  for (int i = 0; i < 12; ++i) { out12[i] = c->prof_phase[i]; c->prof_phase[i] = 0.0; }
  return 0;
}
// SM cycles the persistent M-step kernel of the LAST per-task M step spent in objective evaluations (out2[0]) and in
// optimiser steps incl. the staging of the state (out2[1]), summed over its CTAs
extern "C" int mb_mstep_cycles(mb_context* c, double* out2) {
  MB_TRY(check_ctx(c));
  out2[0] = (double)c->h_cycles[0]; out2[1] = (double)c->h_cycles[1];
  return 0;
}
// diagnostics (tests / scripts/debug): the E step's last [post_inv | weighted] (which = 0) or its per-leaf sums (which = 1)
extern "C" int mb_debug_estep_buffers(mb_context* c, int which, double* out, int64_t count) {
  MB_TRY(check_ctx(c));
  DBuf& b = which == 0 ? c->post_inv : c->leafbuf;
  if (!b.p || (size_t)count * 8 > b.cap) MB_FAIL(-1, "buffer not available");
  MB_TRY(d2h(out, b.p, (size_t)count * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}
extern "C" int mb_flush_l2(mb_context* c, int64_t bytes) {
  MB_TRY(check_ctx(c));
  static thread_local DBuf flush;
  MB_TRY(flush.ensure((size_t)bytes));
  CU(cudaMemsetAsync(flush.p, 1, (size_t)bytes, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// R's `%*%` / crossprod / tcrossprod on host column-major matrices (-> dgemm in the reference): C = op(A) op(B).
// Computed as C' (row-major) = op(B)' op(A)' so that the device kernel's row-major output IS column-major C.
extern "C" int mb_matmul(mb_context* c, int transa, int transb, int m, int n, int k, const double* A, int lda,
                         const double* B, int ldb, double* C, int ldc) {
  MB_TRY(check_ctx(c));
  if (!A || !B || !C || m < 1 || n < 1 || k < 1 || ldc < m) MB_FAIL(-1, "bad argument");
  const int a_rows = transa ? k : m, a_cols = transa ? m : k, b_rows = transb ? n : k, b_cols = transb ? k : n;
  if (lda < a_rows || ldb < b_rows) MB_FAIL(-1, "bad leading dimension");
  static thread_local DBuf dA, dB, dC;
  MB_TRY(dA.ensure((size_t)lda * a_cols * 8));
  MB_TRY(dB.ensure((size_t)ldb * b_cols * 8));
  MB_TRY(dC.ensure((size_t)ldc * n * 8));
  MB_TRY(h2d(dA.p, A, (size_t)lda * a_cols * 8, c->st));
  MB_TRY(h2d(dB.p, B, (size_t)ldb * b_cols * 8, c->st));
  if (ldc != m) CU(cudaMemsetAsync(dC.p, 0, (size_t)ldc * n * 8, c->st));
  // C'(j, i) = sum_k opB(k, j) opA(i, k): left operand (j, k), right operand (k, i)
  const int64_t lrs = transb ? 1 : ldb, lcs = transb ? ldb : 1;
  const int64_t rrs = transa ? 1 : lda, rcs = transa ? lda : 1;
  LAUNCH(c, gemm_any(n, m, k, dB.as<double>(), lrs, lcs, dA.as<double>(), rrs, rcs, dC.as<double>(), ldc, 1.0, nullptr,
                     0.0, c->st));
  MB_TRY(d2h(C, dC.p, (size_t)ldc * n * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// Device-only timing of the large-matrix kernels on a synthetic SE + noise covariance of n equispaced points
// (what = 0: C = A'B n^3 product, 1: Cholesky, 2: Cholesky + inverse).  ms_out[0] = mean ms per repetition.
extern "C" int mb_bench_dense(mb_context* c, int what, int n, int reps, double* ms_out) {
  MB_TRY(check_ctx(c));
  if (n < 1 || reps < 1 || !ms_out) MB_FAIL(-1, "bad argument");
  MB_TRY(c->ws1.ensure(n));
  DenseWs& ws = c->ws1;
  std::vector<double> x(n);
  for (int i = 0; i < n; ++i) x[i] = -1.0 + 2.0 * i / (n > 1 ? n - 1 : 1);
  static thread_local DBuf dx;
  MB_TRY(dx.ensure((size_t)n * 8));
  MB_TRY(h2d(dx.p, x.data(), (size_t)n * 8, c->st));
  mb_kernel k;
  MB_TRY(mb_kernel_parse("SE", 1, &k));
  DevKern kd = make_devkern(&k);
  const double hp[3] = {0.3, what == 3 ? -2.0 : -5.0, -3.0};
  LAUNCH(c, launch_dense_cov(kd, hp, dx.as<double>(), n, dx.as<double>(), n, 1, -1, 1, ws.Kmat.as<double>(), n, 1, c->st));
  const int64_t ldw = ((int64_t)n + 7) / 8 * 8;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  for (int it = -1; it < reps; ++it) {   // it = -1: warm-up
    if (it == 0) CU(cudaEventRecord(e0, c->st));
    if (what == 0) {
      LAUNCH(c, launch_big_gemm_plain(n, n, n, ws.Kmat.as<double>(), 1, n, ws.Kmat.as<double>(), n, 1, ws.T.as<double>(), n,
                                      1.0, nullptr, 0.0, c->st));
    } else if (what == 3) {   // the single-CTA shared-memory kernel of the union grid (n <= 224)
      if (!dense_sm_fits(n)) MB_FAIL(-1, "what = 3 needs n <= 224");
      LAUNCH(c, launch_dense_cholinv_sm(ws.Kmat.as<double>(), n, ws.A.as<double>(), n, 1e-10, 0, ws.info.as<int>(),
                                        ws.res.as<double>() + 16, c->split_lauum ? ws.Bp.as<double>() : nullptr, c->st));
    } else {
      CU(launch_big_cholinv_attempt(ws.Kmat.as<double>(), n, ws.Ap.as<double>(), ws.Bp.as<double>(), ldw, ws.A.as<double>(),
                                    n, 1e-10, 0, what == 1 ? 1 : 0, ws.info.as<int>() + 4, ws.info.as<int>(),
                                    ws.res.as<double>() + 16, c->st, &c->launches));
    }
  }
  CU(cudaEventRecord(e1, c->st));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  MB_TRY(d2h(c->h_pin_i, ws.info.p, sizeof(int), c->st));
  CU(cudaStreamSynchronize(c->st));
  if (what != 0 && what != 3 && c->h_pin_i[0] != 0) MB_FAIL(1, "bench matrix did not factor");
  ms_out[0] = ms / reps;
  if (what == 3) {   // phase cycles of the last launch: load, Cholesky, log-det, triangular inverse, X X', store
    MB_TRY(d2h(c->h_pin, ws.res.as<double>() + 16, 16 * sizeof(double), c->st));
    CU(cudaStreamSynchronize(c->st));
    for (int q = 0; q < 6; ++q) ms_out[1 + q] = c->h_pin[4 + q];
  }
  return 0;
}

// Device-only timing of the batched per-task factorisation (BASELINE.json: "batched chol FP64 TFLOP/s"): every
// task's K(x; hp_i) + jitter -> Cholesky -> inverse with the resident dataset and the pristine HP table of
// mb_set_state, no accumulation or output (the E step's kernel without its scatter).  flops_out = sum N_i^3.
extern "C" int mb_bench_task_inverse(mb_context* c, const mb_kernel* kern_i, double pen_diag, int reps, double* ms_out,
                                     double* flops_out) {
  MB_TRY(need_db(c, false));
  if (!c->have_state || !kern_i || reps < 1 || !ms_out) MB_FAIL(-1, "mb_set_state first");
  if (c->db.nmax > TK_MAXN) MB_FAIL(-4, "use mb_bench_dense for large tasks");
  TaskInvArgs a;
  memset(&a, 0, sizeof(a));
  a.db = c->db; a.kd = make_devkern(kern_i); a.hp = c->st_hp_i0.as<double>(); a.hp_stride = kern_i->n_hp;
  a.pen = pen_diag; a.K = 1; a.U = c->U;
  MB_TRY(c->retries.ensure(c->db.n_tasks * 4));
  a.retries = c->retries.as<int32_t>();
  int G = c->sm_count * task_ctas_per_sm_any(c->db.nmax, c->db.D);
  if (G > c->db.n_tasks) G = (int)c->db.n_tasks;
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  LAUNCH(c, launch_task_inverse_any(a, G, c->st));
  CU(cudaEventRecord(e0, c->st));
  for (int r = 0; r < reps; ++r) LAUNCH(c, launch_task_inverse_any(a, G, c->st));
  CU(cudaEventRecord(e1, c->st));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  ms_out[0] = ms / reps;
  if (flops_out) {
    double fl = 0;
    for (int64_t t = 0; t < c->db.n_tasks; ++t) { const double n = (double)(c->h_offs[t + 1] - c->h_offs[t]); fl += n * n * n; }
    flops_out[0] = fl;
  }
  return 0;
}

// Device-only timing of the HBM-bound kernels of the path on the resident dataset:
//   what = 0  list_kern_to_cov: every task's covariance matrix written to HBM (bytes = 8 sum n_i^2 + table reads)
//   what = 1  the E step's fixed-tree reduction of the per-leaf precision sums (bytes = 8 (leaves + 2) U^2)
// ms_out[0] = mean ms per repetition, bytes_out[0] = algorithmic bytes per repetition.
extern "C" int mb_bench_hbm(mb_context* c, int what, const mb_kernel* kern_i, int reps, double* ms_out, double* bytes_out) {
  MB_TRY(need_db(c, what == 1));
  if (!c->have_state || !kern_i || reps < 1 || !ms_out || !bytes_out) MB_FAIL(-1, "mb_set_state first");
  cudaEvent_t e0, e1;
  CU(cudaEventCreate(&e0));
  CU(cudaEventCreate(&e1));
  static thread_local DBuf d_out, d_oo;
  if (what == 0) {
    const int64_t M = c->db.n_tasks;
    std::vector<int64_t> oo(M);
    size_t total = 0;
    for (int64_t t = 0; t < M; ++t) { const size_t n = (size_t)(c->h_offs[t + 1] - c->h_offs[t]); oo[t] = (int64_t)total; total += n * n; }
    MB_TRY(d_out.ensure(total * 8));
    MB_TRY(d_oo.ensure(M * 8));
    MB_TRY(h2d(d_oo.p, oo.data(), M * 8, c->st));
    DevKern kd = make_devkern(kern_i);
    for (int it = -1; it < reps; ++it) {
      if (it == 0) CU(cudaEventRecord(e0, c->st));
      cudaError_t e = launch_task_cov_sym(c->db, kd, c->st_hp_i0.as<double>(), kern_i->n_hp, -1, d_out.as<double>(),
                                          d_oo.as<int64_t>(), c->sm_count, c->st);
      if (e != cudaSuccess) { mb_set_error(__FILE__, __LINE__, cudaGetErrorString(e)); return -2; }
      c->launches++;
    }
    bytes_out[0] = 8.0 * (double)total + 8.0 * (double)c->h_offs[M] * c->db.D + 8.0 * (double)M * kern_i->n_hp;
  } else if (what == 1) {
    const size_t UU = (size_t)c->U * c->U;
    MB_TRY(ctx_leaves(c, c->U));
    const int G = c->lv_n;
    MB_TRY(c->leafbuf.ensure((size_t)G * UU * 8));
    MB_TRY(c->inv0.ensure(UU * 8));
    MB_TRY(c->post_inv.ensure(UU * 8));
    CU(cudaMemsetAsync(c->leafbuf.p, 0, (size_t)G * UU * 8, c->st));
    CU(cudaMemsetAsync(c->inv0.p, 0, UU * 8, c->st));
    for (int it = -1; it < reps; ++it) {
      if (it == 0) CU(cudaEventRecord(e0, c->st));
      LAUNCH(c, launch_reduce_tree(UU, G, c->inv0.as<double>(), c->leafbuf.as<double>(), UU, c->post_inv.as<double>(), c->st));
    }
    bytes_out[0] = 8.0 * (double)UU * (G + 2);
  } else {
    MB_FAIL(-1, "what must be 0 or 1");
  }
  CU(cudaEventRecord(e1, c->st));
  CU(cudaEventSynchronize(e1));
  float ms = 0;
  CU(cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  if (what == 0) { d_out.release(); d_oo.release(); }
  ms_out[0] = ms / reps;
  return 0;
}

// ---- pred_magma --------------------------------------------------------------------------------
cudaError_t launch_pred_var(const DevKern& kd, const double* hp_host, const double* xp, int np, int D,
                            const double* sdiag, const double* C, const double* GC, int no,
                            double noise, double* out, cudaStream_t st);

extern "C" int mb_pred_magma(mb_context* c, const mb_kernel* k, const double* hp, const double* x_obs,
                             int n_obs, const double* y_obs, const double* x_pred, int n_pred, int d,
                             const double* mean_obs, const double* mean_pred, const double* cov_obs,
                             const double* cov_pred, const double* cov_crossed, double pen_diag,
                             double* pred_mean, double* pred_var, double* pred_cov) {
  if (k && k->n_terms == 1 && k->types[0] == MB_KERN_CONV) MB_FAIL(-4, "prediction with the multi-output convolution kernel is not on the accelerated path yet");
  MB_TRY(check_ctx(c));
  if (!k || !hp || !x_obs || !y_obs || !x_pred || !mean_obs || !mean_pred || !cov_obs || !cov_pred ||
      !cov_crossed || !pred_mean || !pred_var || n_obs < 1 || n_pred < 1 || d < 1)
    MB_FAIL(-1, "bad argument");
  DevKern kd = make_devkern(k);
  const int no = n_obs, np = n_pred;
  const bool full = pred_cov != nullptr;
  MB_TRY(c->ws1.ensure(no));
  size_t need = ((size_t)no * d + (size_t)np * d + 3 * (size_t)no + 4 * (size_t)np + (size_t)no * no +
                 3 * (size_t)no * np + (full ? 2 * (size_t)np * np : 0)) * 8 + 256;
  MB_TRY(c->scratch.ensure(need));
  double* p = c->scratch.as<double>();
  double* dXo = p; p += (size_t)no * d;
  double* dXp = p; p += (size_t)np * d;
  double* dyo = p; p += no;
  double* dmo = p; p += no;
  double* dvec = p; p += no;
  double* dmp = p; p += np;
  double* dmean = p; p += np;
  double* dbase = p; p += np;                 // diag(K_pred + S_pred)
  double* dvar = p; p += np;
  double* dSo = p; p += (size_t)no * no;
  double* dC = p; p += (size_t)no * np;       // C(i,j) at i + j*no (column-major, like the host block)
  double* dCh = p; p += (size_t)no * np;
  double* dGC = p; p += (size_t)no * np;      // Gamma^{-1} C, row-major no x np
  double* dKp = full ? p : nullptr; p += full ? (size_t)np * np : 0;
  double* dSp = full ? p : nullptr;
  cudaStream_t s = c->st;
  MB_TRY(h2d(dXo, x_obs, (size_t)no * d * 8, s));
  MB_TRY(h2d(dXp, x_pred, (size_t)np * d * 8, s));
  MB_TRY(h2d(dyo, y_obs, (size_t)no * 8, s));
  MB_TRY(h2d(dmo, mean_obs, (size_t)no * 8, s));
  MB_TRY(h2d(dmp, mean_pred, (size_t)np * 8, s));
  MB_TRY(h2d(dSo, cov_obs, (size_t)no * no * 8, s));
  MB_TRY(h2d(dCh, cov_crossed, (size_t)no * np * 8, s));
  std::vector<double> sdiag;
  if (full) MB_TRY(h2d(dSp, cov_pred, (size_t)np * np * 8, s));
  else {
    sdiag.resize(np);
    for (int j = 0; j < np; ++j) sdiag[j] = cov_pred[(size_t)j * np + j];
    MB_TRY(h2d(dvar, sdiag.data(), (size_t)np * 8, s));
  }
  // Gamma = K(obs; hp with noise) + S_obs  (prediction.R:1147)
  LAUNCH(c, launch_dense_cov(kd, hp, dXo, no, dXo, no, d, -1, 1, c->ws1.Kmat.as<double>(), no, 1, s));
  LAUNCH(c, launch_axpby((int64_t)no * no, 1.0, c->ws1.Kmat.as<double>(), 1.0, dSo, c->ws1.Kmat.as<double>(), s));
  MB_TRY(dev_cholinv(c, c->ws1, c->ws1.Kmat.as<double>(), no, pen_diag, s, nullptr, nullptr, nullptr));
  const double* Ginv = c->ws1.A.as<double>();
  // C = K(obs, pred; noise removed) + S_crossed  (prediction.R:1174-1181)
  LAUNCH(c, launch_dense_cov(kd, hp, dXo, no, dXp, np, d, -1, 0, dC, 1, no, s));
  LAUNCH(c, launch_axpby((int64_t)no * np, 1.0, dC, 1.0, dCh, dC, s));
  // mean = mu_pred + C' Gamma^{-1} (y - mu_obs)  (prediction.R:1184-1186)
  LAUNCH(c, launch_axpby(no, 1.0, dyo, -1.0, dmo, dmo, s));
  LAUNCH(c, launch_gemv(no, no, Ginv, no, dmo, nullptr, dvec, s));
  LAUNCH(c, gemm_any(np, 1, no, dC, no, 1, dvec, 1, 1, dmean, 1, 1.0, dmp, 1.0, s));
  LAUNCH(c, gemm_any(no, np, no, Ginv, no, 1, dC, 1, no, dGC, np, 1.0, nullptr, 0.0, s));
  const double noise = kd.has_noise ? exp(hp[kd.n_hp - 1]) : 0.0;
  if (full) {
    // cov = K(pred) + S_pred - C' Gamma^{-1} C  (prediction.R:1189)
    LAUNCH(c, launch_dense_cov(kd, hp, dXp, np, dXp, np, d, -1, 0, dKp, np, 1, s));
    LAUNCH(c, launch_axpby((int64_t)np * np, 1.0, dKp, 1.0, dSp, dKp, s));
    LAUNCH(c, gemm_any(np, np, no, dC, no, 1, dGC, np, 1, dKp, np, -1.0, dKp, 1.0, s));
    LAUNCH(c, launch_diag(np, dKp, np, noise, dvar, s));
    MB_TRY(d2h(pred_cov, dKp, (size_t)np * np * 8, s));
  } else {
    // only the diagonal of the same expression
    LAUNCH(c, launch_pred_var(kd, hp, dXp, np, d, dvar, dC, dGC, no, noise, dbase, s));
    dvar = dbase;
  }
  MB_TRY(d2h(pred_mean, dmean, (size_t)np * 8, s));
  MB_TRY(d2h(pred_var, dvar, (size_t)np * 8, s));
  CU(cudaStreamSynchronize(s));
  return 0;
}

// pred_magmaclust (prediction.R:2298-2413) for one new task: the GP conditional of mb_pred_magma per cluster
// with that cluster's hyper-posterior blocks, then the mixture mean / variance.  The final combination of
// K short vectors (prediction.R:2380,2396-2404) is host arithmetic in the reference's order.
extern "C" int mb_pred_magmaclust(mb_context* c, int n_clusters, const mb_kernel* k, const double* hp,
                                  const double* x_obs, int n_obs, const double* y_obs, const double* x_pred,
                                  int n_pred, int d, const double* mean_obs, const double* mean_pred,
                                  const double* cov_obs, const double* cov_pred, const double* cov_crossed,
                                  const double* proba, double pen_diag, double* pred_mean, double* pred_var,
                                  double* pred_cov, double* mix_mean, double* mix_var) {
  if (n_clusters < 1 || n_clusters > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  if (!proba || !mix_mean || !mix_var || !pred_mean || !pred_var) MB_FAIL(-1, "null argument");
  const size_t no = n_obs, np = n_pred;
  for (int q = 0; q < n_clusters; ++q)
    MB_TRY(mb_pred_magma(c, k, hp, x_obs, n_obs, y_obs, x_pred, n_pred, d, mean_obs + q * no, mean_pred + q * np,
                         cov_obs + q * no * no, cov_pred + q * np * np, cov_crossed + q * no * np, pen_diag,
                         pred_mean + q * np, pred_var + q * np, pred_cov ? pred_cov + q * np * np : nullptr));
  for (size_t j = 0; j < np; ++j) {
    double m = 0.0;
    for (int q = 0; q < n_clusters; ++q) m = m + proba[q] * pred_mean[q * np + j];
    mix_mean[j] = m;
  }
  for (size_t j = 0; j < np; ++j) {
    double v = 0.0;
    for (int q = 0; q < n_clusters; ++q) {
      const double dm = pred_mean[q * np + j] - mix_mean[j];
      v = v + proba[q] * (pred_var[q * np + j] + dm * dm);
    }
    mix_var[j] = v;
  }
  return 0;
}

// ================================================================================================
// MagmaClust: ve_step / update_mixture / vm_step / elbo_monitoring_VEM
// ================================================================================================

// R's round(x, digits) for digits > 0 (nmath/fround.c, R >= 4.0.0): the closer of the two
// neighbouring decimals, ties to the even one.  vem-magmaclust.R:392 rounds tau to 5 decimals.
static double r_round(double x, int dig) {
  if (!std::isfinite(x)) return x;
  double sgn = 1.0;
  if (x < 0) { sgn = -1.0; x = -x; }
  long double p10 = 1.0L;
  for (int i = 0; i < dig; ++i) p10 *= 10.0L;
  long double x10 = p10 * (long double)x;
  long double i10 = floorl(x10);
  double xd = (double)(i10 / p10), xu = (double)(ceill(x10) / p10);
  double du = xu - x, dd = x - xd;
  bool even = fmodl(i10, 2.0L) == 0.0L;
  return sgn * ((dd < du || (dd == du && even)) ? xd : xu);
}

// R's mean(): long double sum, then one correction pass (summary.c rsum / mean)
static double r_mean(const double* v, int n) {
  long double s = 0.0L;
  for (int i = 0; i < n; ++i) s += v[i];
  s /= n;
  long double t = 0.0L;
  for (int i = 0; i < n; ++i) t += (v[i] - s);
  return (double)(s + t / n);
}

// upload K hyper-posteriors (mean K x U, cov K x U x U) and tau (n_tasks x K)
static int put_hyperpost(mb_context* c, int K, const double* post_mean, const double* post_cov,
                         const double* tau) {
  const size_t UU = (size_t)c->U * c->U;
  MB_TRY(ensure_em_buffers(c, K));
  MB_TRY(h2d(c->pm.p, post_mean, (size_t)K * c->U * 8, c->st));
  MB_TRY(h2d(c->pc.p, post_cov, K * UU * 8, c->st));
  if (tau) {
    MB_TRY(c->tau.ensure((size_t)c->db.n_tasks * K * 8));
    MB_TRY(h2d(c->tau.p, tau, (size_t)c->db.n_tasks * K * 8, c->st));
  }
  return 0;
}

// update_mixture on device state -> host tau_out (n_tasks x K)
static int dev_update_mixture(mb_context* c, int K, const DevKern& kdi, const double* hp_i_dev,
                              int64_t stride, const double* pm_dev, const double* pc_dev,
                              const double* prop, double pen, double* tau_out) {
  const int64_t M = c->db.n_tasks;
  MB_TRY(c->f_t.ensure((size_t)M * K * 8));
  TaskObjArgs a = make_obj_args(c, kdi, hp_i_dev, stride, TK_OBJ_MIX, 0, pen, K, pm_dev, pc_dev, nullptr,
                                c->f_t.as<double>(), c->g_t.as<double>());
  MB_TRY(launch_obj(c, a, (int)M, c->st));
  std::vector<double> f((size_t)M * K);
  MB_TRY(d2h(f.data(), c->f_t.p, f.size() * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  // vem-magmaclust.R:383-392: exp(L - max L) * pi, normalised per task, round(5)
  for (int64_t t = 0; t < M; ++t) {
    double mx = -INFINITY;
    for (int k = 0; k < K; ++k) mx = fmax(mx, -f[t * K + k]);
    double w[MB_MAX_CLUSTERS], sum = 0.0;
    for (int k = 0; k < K; ++k) { w[k] = prop[k] * exp(-f[t * K + k] - mx); sum += w[k]; }
    for (int k = 0; k < K; ++k) tau_out[t * K + k] = r_round(w[k] / sum, 5);
  }
  return 0;
}

extern "C" int mb_update_mixture(mb_context* c, int n_clusters, const mb_kernel* kern_i, const double* hp_i,
                                 int hp_i_shared, const double* post_mean, const double* post_cov,
                                 const double* prop_mixture, double pen_diag, double* tau_out) {
  MB_TRY(need_db(c, true));
  if (!kern_i || !hp_i || !post_mean || !post_cov || !prop_mixture || !tau_out) MB_FAIL(-1, "null argument");
  if (n_clusters < 1 || n_clusters > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(put_hyperpost(c, n_clusters, post_mean, post_cov, nullptr));
  return dev_update_mixture(c, n_clusters, make_devkern(kern_i), c->hp_i.as<double>(), stride, c->pm.as<double>(),
                            c->pc.as<double>(), prop_mixture, pen_diag, tau_out);
}

extern "C" int mb_train_gp_clust_tasks(mb_context* c, const mb_kernel* kern, double* hp, const double* prop_mixture,
                                       double pen_diag, int n_iter_max, double cv_threshold, double* mixture_out,
                                       int32_t* n_iter_out, int32_t* converged_out, int64_t* stats) {
  MB_TRY(need_hyperpost(c, 0, true));
  if (!kern || !hp || !prop_mixture || !mixture_out || n_iter_max < 1) MB_FAIL(-1, "bad argument");
  const int K = c->hpr_K, P = kern->n_hp;
  const int64_t M = c->db.n_tasks;
  MB_TRY(ensure_em_buffers(c, K));
  DevKern kd = make_devkern(kern);
  DBuf d_hp, d_tau, d_skip, d_fk;
  int rc = 0;
  if (d_hp.ensure((size_t)M * P * 8) || d_tau.ensure((size_t)M * K * 8) || d_skip.ensure((size_t)M * 4) ||
      d_fk.ensure((size_t)M * K * 8)) rc = -3;
  auto cleanup = [&]() { d_hp.release(); d_tau.release(); d_skip.release(); d_fk.release(); };
  if (rc) { cleanup(); return rc; }
  std::vector<double> cur_hp(hp, hp + (size_t)M * P), new_hp((size_t)M * P), mix((size_t)M * K), new_mix((size_t)M * K),
      fk((size_t)M * K), LL(M, -INFINITY);
  std::vector<int32_t> done(M, 0), iters(M, 0), conv(M, 0);
  for (int64_t t = 0; t < M; ++t) for (int k = 0; k < K; ++k) mix[t * K + k] = prop_mixture[k];
  int64_t tot_evals = 0, tot_rounds = 0;
  auto fail = [&](int code) { cleanup(); return code; };
  for (int it = 1; it <= n_iter_max; ++it) {
    int64_t live = 0;
    for (int64_t t = 0; t < M; ++t) live += !done[t];
    if (!live) break;
    // ---- M step on the live tasks
    if (h2d(d_hp.p, cur_hp.data(), (size_t)M * P * 8, c->st) || h2d(d_tau.p, mix.data(), (size_t)M * K * 8, c->st) ||
        h2d(d_skip.p, done.data(), (size_t)M * 4, c->st)) return fail(-2);
    TaskObjArgs a = make_obj_args(c, kd, d_hp.as<double>(), P, TK_OBJ_MARG_MIX, 1, pen_diag, K, c->hpr_mean.as<double>(),
                                  c->hpr_cov.as<double>(), d_tau.as<double>(), c->f_t.as<double>(), c->g_t.as<double>());
    std::vector<SingleProblem> none;
    int64_t st4[4] = {0, 0, 0, 0};
    rc = optimise(c, none, true, a, d_hp.as<double>(), P, pen_diag, st4, d_skip.as<int32_t>());
    if (rc) return fail(rc);
    tot_rounds += st4[1]; tot_evals += st4[2];
    if (d2h(new_hp.data(), d_hp.p, (size_t)M * P * 8, c->st) || cudaStreamSynchronize(c->st) != cudaSuccess) return fail(-2);
    for (int64_t t = 0; t < M; ++t) {
      if (done[t]) { for (int p = 0; p < P; ++p) new_hp[t * P + p] = cur_hp[t * P + p]; continue; }
      bool bad = false;
      for (int p = 0; p < P; ++p) if (!std::isfinite(new_hp[t * P + p])) bad = true;
      if (bad) {   // training.R:2122-2129: stop, keep the last valid values
        done[t] = 1;
        for (int p = 0; p < P; ++p) new_hp[t * P + p] = cur_hp[t * P + p];
      } else {
        iters[t] = it;
      }
    }
    // ---- E step (update_mixture with the new HPs) and the per-cluster marginal values for the monitoring
    if (h2d(d_hp.p, new_hp.data(), (size_t)M * P * 8, c->st)) return fail(-2);
    rc = dev_update_mixture(c, K, kd, d_hp.as<double>(), P, c->hpr_mean.as<double>(), c->hpr_cov.as<double>(),
                            prop_mixture, pen_diag, new_mix.data());
    if (rc) return fail(rc);
    TaskObjArgs am = make_obj_args(c, kd, d_hp.as<double>(), P, TK_OBJ_MARG_MIX, 0, pen_diag, K, c->hpr_mean.as<double>(),
                                   c->hpr_cov.as<double>(), d_tau.as<double>(), c->f_t.as<double>(), c->g_t.as<double>());
    am.fk = d_fk.as<double>();
    rc = launch_obj(c, am, (int)M, c->st);
    if (rc) return fail(rc);
    if (d2h(fk.data(), d_fk.p, (size_t)M * K * 8, c->st) || cudaStreamSynchronize(c->st) != cudaSuccess) return fail(-2);
    for (int64_t t = 0; t < M; ++t) {
      if (done[t]) continue;
      // sum_logL_GP_clust with prop_mixture (likelihoods.R:378-396): sum_k [-tau_k logL_GP_k + tau_k log(pi_k / tau_k)]
      long double acc = 0.0L;
      for (int k = 0; k < K; ++k) {
        const double tk = new_mix[t * K + k], pk = prop_mixture[k];
        const double lf = (tk == 0.0 || pk == 0.0) ? 0.0 : log(pk / tk);
        acc += (long double)(-(tk * fk[t * K + k]) + tk * lf);
      }
      const double nl = (double)acc;
      double diff = nl - LL[t];
      if (std::isnan(diff)) diff = -INFINITY;
      for (int p = 0; p < P; ++p) cur_hp[t * P + p] = new_hp[t * P + p];
      for (int k = 0; k < K; ++k) mix[t * K + k] = new_mix[t * K + k];
      LL[t] = nl;
      double eps = diff / fabs(nl);
      if (std::isnan(eps)) eps = 1.0;
      if (fabs(eps) < cv_threshold) { conv[t] = 1; done[t] = 1; }
    }
  }
  cleanup();
  for (size_t i = 0; i < cur_hp.size(); ++i) hp[i] = cur_hp[i];
  for (size_t i = 0; i < mix.size(); ++i) mixture_out[i] = mix[i];
  if (n_iter_out) for (int64_t t = 0; t < M; ++t) n_iter_out[t] = iters[t];
  if (converged_out) for (int64_t t = 0; t < M; ++t) converged_out[t] = conv[t];
  if (stats) { stats[0] = 0; stats[1] = tot_rounds; stats[2] = tot_evals; stats[3] = 0; }
  return 0;
}

extern "C" int mb_ve_step(mb_context* c, int n_clusters, const mb_kernel* kern_k, const double* hp_k,
                          const mb_kernel* kern_i, const double* hp_i, int hp_i_shared, const double* m_k,
                          const double* tau, const double* prop_mixture, int update, double pen_diag,
                          double* post_mean, double* post_cov, double* tau_out) {
  MB_TRY(need_db(c, true));
  if (!kern_k || !hp_k || !kern_i || !hp_i || !m_k || !tau || !post_mean || !post_cov) MB_FAIL(-1, "null argument");
  const int K = n_clusters;
  if (K < 1 || K > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  if (update && (!prop_mixture || !tau_out)) MB_FAIL(-1, "update_mixture needs prop_mixture and tau_out");
  const size_t UU = (size_t)c->U * c->U;
  MB_TRY(ensure_em_buffers(c, K));
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(c->tau.ensure((size_t)c->db.n_tasks * K * 8));
  MB_TRY(h2d(c->tau.p, tau, (size_t)c->db.n_tasks * K * 8, c->st));
  MB_TRY(h2d(c->m0.p, m_k, (size_t)K * c->U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  DevKern kdi = make_devkern(kern_i);
  MB_TRY(dev_e_step(c, K, kern_k, hp_k, kdi, c->hp_i.as<double>(), stride, c->m0.as<double>(), c->tau.as<double>(),
                    c->d_gridX.as<double>(), c->U, nullptr, pen_diag, c->pm.as<double>(), c->pc.as<double>(), nullptr));
  MB_TRY(d2h(post_mean, c->pm.p, (size_t)K * c->U * 8, c->st));
  MB_TRY(d2h(post_cov, c->pc.p, K * UU * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  if (tau_out) {
    if (update) {
      MB_TRY(dev_update_mixture(c, K, kdi, c->hp_i.as<double>(), stride, c->pm.as<double>(), c->pc.as<double>(),
                                prop_mixture, pen_diag, tau_out));
    } else {
      memcpy(tau_out, tau, (size_t)c->db.n_tasks * K * 8);   // vem-magmaclust.R:130-131
    }
  }
  return 0;
}

extern "C" int mb_elbo_clust_tasks(mb_context* c, int n_clusters, const mb_kernel* kern_i, const double* hp_i,
                                   int hp_i_shared, const double* post_mean, const double* post_cov,
                                   const double* tau, double pen_diag, double* f, double* g) {
  MB_TRY(need_db(c, true));
  if (!kern_i || !hp_i || !post_mean || !post_cov || !tau || !f) MB_FAIL(-1, "null argument");
  if (n_clusters < 1 || n_clusters > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(put_hyperpost(c, n_clusters, post_mean, post_cov, tau));
  DevKern kd = make_devkern(kern_i);
  TaskObjArgs a = make_obj_args(c, kd, c->hp_i.as<double>(), stride, TK_OBJ_ELBO, g != nullptr, pen_diag, n_clusters,
                                c->pm.as<double>(), c->pc.as<double>(), c->tau.as<double>(), c->f_t.as<double>(),
                                c->g_t.as<double>());
  MB_TRY(launch_obj(c, a, (int)c->db.n_tasks, c->st));
  MB_TRY(d2h(f, c->f_t.p, c->db.n_tasks * 8, c->st));
  if (g) MB_TRY(d2h(g, c->g_t.p, (size_t)c->db.n_tasks * kd.n_hp * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// elbo_GP_mod_shared_hp_k + gr_GP_mod_shared_hp_k (elbos.R:78-100, gradients-elbos.R:94-150):
// sum over clusters of logL_GP_mod with ONE hyper-parameter vector.
static int shared_k_eval(mb_context* c, int K, const DevKern& kdk, const double* hp, const double* mk_dev,
                         double pen, int want_grad, double* f, double* g) {
  const size_t UU = (size_t)c->U * c->U;
  double* hres = c->h_pin + 32;
  int* hinfo = c->h_pin_i + 8;
  double ll = 0.0;
  *f = 0.0;
  for (int p = 0; p < kdk.n_hp; ++p) g[p] = 0.0;
  for (int k = 0; k < K; ++k) {
    MB_TRY(dense_obj_enqueue(c, c->ws0, kdk, hp, c->d_gridX.as<double>(), c->U, c->db.D,
                             c->pm.as<double>() + (size_t)k * c->U, mk_dev + (size_t)k * c->U,
                             c->pc.as<double>() + k * UU, pen, want_grad, c->st2, hres, hinfo));
    CU(cudaStreamSynchronize(c->st2));
    if (hinfo[0] < 0) { *f = NAN; return 0; }
    ll += hres[0];
    if (want_grad) for (int p = 0; p < kdk.n_hp; ++p) g[p] += hres[1 + p];
  }
  *f = ll;
  return 0;
}

// vm_step on device-resident state: c->pm / c->pc / c->tau hold the K hyper-posteriors and the mixture, hp_i_dev the task HP
// table (in/out).  post_mean_host (K x U) and tau_host (local tasks x K) are the host copies the O(K U + M K) scalar glue of
// the reference reads (vem-magmaclust.R:229-232, :265-266).  m_k (host, K x U) and prop_mixture are outputs.
static int dev_vm_step(mb_context* c, int K, const mb_kernel* kern_k, double* hp_k, const mb_kernel* kern_i,
                       double* hp_i_dev, int shared_hp_k, int shared_hp_i, double* m_k, const double* post_mean_host,
                       const double* tau_host, double pen_diag, double* prop_mixture, int64_t* stats) {
  const int U = c->U;
  const int64_t M = c->db.n_tasks;
  DevKern kdk = make_devkern(kern_k), kdi = make_devkern(kern_i);
  const int Pi = kdi.n_hp, Pk = kdk.n_hp;
  if (stats) stats[0] = stats[1] = stats[2] = stats[3] = 0;
  // m_k <- mean of the hyper-posterior mean, replicated (vem-magmaclust.R:229-232)
  for (int k = 0; k < K; ++k) {
    double mu = r_mean(post_mean_host + (size_t)k * U, U);
    for (int u = 0; u < U; ++u) m_k[(size_t)k * U + u] = mu;
  }
  MB_TRY(h2d(c->m0.p, m_k, (size_t)K * U * 8, c->st));
  // pi_k = colMeans(tau) (:265-266) over the tasks of every rank
  {
    double colsum[MB_MAX_CLUSTERS];
    MB_TRY(leaf_sum_columns_host(c, tau_host, K, colsum));
    const double m_all = (double)(c->shard_set ? c->g_total : M);
    for (int k = 0; k < K; ++k) prop_mixture[k] = colsum[k] / m_all;
  }
  CU(cudaStreamSynchronize(c->st));
  TaskObjArgs a = make_obj_args(c, kdi, hp_i_dev, Pi, TK_OBJ_ELBO, 1, pen_diag, K, c->pm.as<double>(),
                                c->pc.as<double>(), c->tau.as<double>(), c->f_t.as<double>(), c->g_t.as<double>());
  // cluster processes: K independent problems, or one shared problem handled after the tasks
  std::vector<SingleProblem> singles;
  if (!shared_hp_k) {
    singles.resize(K);
    for (int k = 0; k < K; ++k) {
      SingleProblem& sp = singles[k];
      sp.kd = kdk; sp.X = c->d_gridX.as<double>(); sp.n = U; sp.D = c->db.D;
      sp.y = c->pm.as<double>() + (size_t)k * U; sp.mean = c->m0.as<double>() + (size_t)k * U;
      sp.S = c->pc.as<double>() + (size_t)k * U * U;
      lb_init(&sp.s, Pk, hp_k + (size_t)k * Pk, 1e13, 0.0, 25);
    }
  }
  c->vm_valid = false;
  if (!shared_hp_i) {
    MB_TRY(optimise(c, singles, true, a, hp_i_dev, Pi, pen_diag, stats));
    c->vm_tasks_in_f_t = true;
  } else {
    MB_TRY(optimise(c, singles, false, a, hp_i_dev, Pi, pen_diag, stats));
    c->vm_tasks_in_f_t = false;
    LbfgsbState s;
    double x0[MB_MAX_HP];
    MB_TRY(d2h(x0, hp_i_dev, Pi * 8, c->st));     // block of the first task (:236-238): the same on every rank
    CU(cudaStreamSynchronize(c->st));
    lb_init(&s, Pi, x0, 1e13, 0.0, 25);
    int evals = 0;
    while (s.active) {
      double f, g[MB_MAX_HP];
      MB_TRY(shared_obj_eval(c, a, s.x, Pi, 1, &f, g));
      lb_advance(&s, f, g);
      evals++;
    }
    std::vector<double> tab((size_t)M * Pi);
    for (int64_t t = 0; t < M; ++t)
      for (int p = 0; p < Pi; ++p) tab[t * Pi + p] = s.x[p];
    MB_TRY(h2d(hp_i_dev, tab.data(), tab.size() * 8, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (stats) { stats[1] = evals; stats[2] = (int64_t)evals * M; stats[3] = (s.fail == 99); }
    c->vm_sum_ll_i = s.f;
  }
  bool k_ok = true;
  if (!shared_hp_k) {
    int ev = 0;
    double sum_k = 0.0;
    for (int k = 0; k < K; ++k) {
      for (int p = 0; p < Pk; ++p) hp_k[(size_t)k * Pk + p] = singles[k].s.x[p];
      ev += singles[k].evals;
      sum_k += singles[k].s.f;          // in cluster order, like elbo_monitoring_VEM's loop
      if (singles[k].evals == 0) k_ok = false;
    }
    c->vm_sum_ll_k = sum_k;
    if (stats) stats[0] = ev;
  } else {
    LbfgsbState s;
    lb_init(&s, Pk, hp_k, 1e13, 0.0, 25);       // block of the first cluster (:270-272)
    int evals = 0;
    while (s.active) {
      double f, g[MB_MAX_HP];
      MB_TRY(shared_k_eval(c, K, kdk, s.x, c->m0.as<double>(), pen_diag, 1, &f, g));
      lb_advance(&s, f, g);
      evals++;
    }
    for (int k = 0; k < K; ++k)
      for (int p = 0; p < Pk; ++p) hp_k[(size_t)k * Pk + p] = s.x[p];
    if (stats) stats[0] = evals;
    c->vm_sum_ll_k = s.f;
  }
  c->vm_valid = k_ok;
  return 0;
}

extern "C" int mb_vm_step(mb_context* c, int n_clusters, const mb_kernel* kern_k, double* hp_k,
                          const mb_kernel* kern_i, double* hp_i, int shared_hp_k, int shared_hp_i, double* m_k,
                          const double* post_mean, const double* post_cov, const double* tau, double pen_diag,
                          double* prop_mixture, int64_t* stats) {
  MB_TRY(need_db(c, true));
  if (!kern_k || !hp_k || !kern_i || !hp_i || !m_k || !post_mean || !post_cov || !tau || !prop_mixture)
    MB_FAIL(-1, "null argument");
  const int K = n_clusters;
  if (K < 1 || K > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  const int64_t M = c->db.n_tasks;
  const int Pi = kern_i->n_hp;
  MB_TRY(put_hyperpost(c, K, post_mean, post_cov, tau));
  MB_TRY(c->hp_k.ensure((size_t)M * Pi * 8));
  MB_TRY(h2d(c->hp_k.p, hp_i, (size_t)M * Pi * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  MB_TRY(dev_vm_step(c, K, kern_k, hp_k, kern_i, c->hp_k.as<double>(), shared_hp_k, shared_hp_i, m_k, post_mean, tau,
                     pen_diag, prop_mixture, stats));
  MB_TRY(d2h(hp_i, c->hp_k.p, (size_t)M * Pi * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// elbo_monitoring_VEM on device-resident state (c->pm / c->pc / c->tau / c->m0 = m_k); hp_i_dev with the given stride
static int dev_elbo_monitoring(mb_context* c, int K, const mb_kernel* kern_k, const double* hp_k, const mb_kernel* kern_i,
                               const double* hp_i_dev, int64_t stride, const double* tau_host, const double* prop_mixture,
                               double pen_diag, double* elbo, bool reuse_vm_step = false) {
  // reuse_vm_step: called right after dev_vm_step on the same hyper-posterior and mixture -- the task and cluster terms
  // (elbos.R:209-235) are the objective values the VM step's optimisers ended with (see dev_logL_monitoring)
  const int U = c->U;
  const int64_t M = c->db.n_tasks;
  const size_t UU = (size_t)U * U;
  DevKern kdk = make_devkern(kern_k), kdi = make_devkern(kern_i);
  MB_TRY(c->ws1.ensure(U));
  // task sum (elbos.R:221-235)
  TaskObjArgs a = make_obj_args(c, kdi, hp_i_dev, stride, TK_OBJ_ELBO, 0, pen_diag, K, c->pm.as<double>(),
                                c->pc.as<double>(), c->tau.as<double>(), c->f_t.as<double>(), c->g_t.as<double>());
  double sum_ll_i;
  if (reuse_vm_step && !c->vm_tasks_in_f_t) {
    sum_ll_i = c->vm_sum_ll_i;
  } else {
    if (!reuse_vm_step) MB_TRY(launch_obj(c, a, (int)M, c->st));
    MB_TRY(c->sums.ensure(64 * 8));
    MB_TRY(leaf_sum_columns(c, a.f, 1, c->sums.as<double>(), c->st));
    MB_TRY(d2h(c->h_pin, c->sums.p, 8, c->st));
    CU(cudaStreamSynchronize(c->st));
    sum_ll_i = c->h_pin[0];
  }
  // sum_i tau_ik log(pi_k / tau_ik) over the tasks of every rank (elbos.R:249-256)
  double sum_tau_k[MB_MAX_CLUSTERS];
  {
    std::vector<double> terms((size_t)M * K);
    for (int64_t t = 0; t < M; ++t)
      for (int k = 0; k < K; ++k) {
        const double tk = tau_host[t * K + k], pi_k = prop_mixture[k];
        terms[t * K + k] = (tk == 0.0 || pi_k == 0.0) ? 0.0 : tk * log(pi_k / tk);
      }
    MB_TRY(leaf_sum_columns_host(c, terms.data(), K, sum_tau_k));
  }
  // cluster terms (elbos.R:209-219, 237-268)
  double sum_ll_k = 0.0, sum_corr = 0.0;
  double* hres = c->h_pin + 32;
  int* hinfo = c->h_pin_i + 8;
  if (reuse_vm_step) sum_ll_k = c->vm_sum_ll_k;
  for (int k = 0; k < K; ++k) {
    if (!reuse_vm_step) {
      MB_TRY(dense_obj_enqueue(c, c->ws0, kdk, hp_k + (size_t)k * kdk.n_hp, c->d_gridX.as<double>(), U, c->db.D,
                               c->pm.as<double>() + (size_t)k * U, c->m0.as<double>() + (size_t)k * U,
                               c->pc.as<double>() + k * UU, pen_diag, 0, c->st2, hres, hinfo));
      CU(cudaStreamSynchronize(c->st2));
      sum_ll_k += (hinfo[0] < 0) ? NAN : hres[0];
    }
    LAUNCH(c, launch_chol_logdiag_any(c, c->ws1, c->pc.as<double>() + k * UU, U, c->ws1.info.as<int>(),
                                      c->sums.as<double>() + 8, c->st));
    MB_TRY(d2h(c->h_pin + 8, c->sums.as<double>() + 8, 8, c->st));
    MB_TRY(d2h(c->h_pin_i, c->ws1.info.p, sizeof(int), c->st));
    CU(cudaStreamSynchronize(c->st));
    if (c->h_pin_i[0] != 0) MB_FAIL(1, "chol(cov_k) failed in elbo_monitoring_VEM (no jitter is applied there)");
    sum_corr += sum_tau_k[k] + c->h_pin[8];
  }
  *elbo = -sum_ll_k - sum_ll_i + sum_corr;
  return 0;
}

extern "C" int mb_elbo_monitoring(mb_context* c, int n_clusters, const mb_kernel* kern_k, const double* hp_k,
                                  const mb_kernel* kern_i, const double* hp_i, int hp_i_shared, const double* m_k,
                                  const double* post_mean, const double* post_cov, const double* tau,
                                  const double* prop_mixture, double pen_diag, double* elbo) {
  MB_TRY(need_db(c, true));
  if (!kern_k || !hp_k || !kern_i || !hp_i || !m_k || !post_mean || !post_cov || !tau || !prop_mixture || !elbo)
    MB_FAIL(-1, "null argument");
  const int K = n_clusters, U = c->U;
  if (K < 1 || K > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  int64_t stride;
  MB_TRY(put_hp_i(c, kern_i, hp_i, hp_i_shared, &stride));
  MB_TRY(put_hyperpost(c, K, post_mean, post_cov, tau));
  MB_TRY(h2d(c->m0.p, m_k, (size_t)K * U * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  return dev_elbo_monitoring(c, K, kern_k, hp_k, kern_i, c->hp_i.as<double>(), stride, tau, prop_mixture, pen_diag, elbo);
}

// train_magmaclust's VEM loop (/root/reference/R/training.R:1586-1758) with the state resident on the device: the K
// hyper-posteriors (K x U x U), the task HP table and the mixture never cross PCIe inside the loop; per iteration the host
// reads the K x U posterior means (m_k = their averages, vem-magmaclust.R:229-232) and the M x K table of per-cluster
// likelihoods for the reference's softmax + round(5) (vem-magmaclust.R:383-392).
// hp_k: K x P_k, hp_i: local tasks x P_i, m_k: K x U (all in/out); ini_mixture: local tasks x K.  fast_approx != 0 stops
// after the first VE step with the ELBO of the initial HPs (training.R:1633-1647).  Outputs may be NULL.
extern "C" int mb_train_magmaclust(mb_context* c, int n_clusters, const mb_kernel* kern_k, double* hp_k,
                                   const mb_kernel* kern_i, double* hp_i, int shared_hp_k, int shared_hp_i, double* m_k,
                                   const double* ini_mixture, double pen_diag, int n_iter_max, double cv_threshold,
                                   int fast_approx, double* seq_elbo, int32_t* n_iter, int32_t* converged,
                                   double* mixture_out, double* prop_mixture_out, double* post_mean, double* post_cov) {
  MB_TRY(need_db(c, true));
  if (!kern_k || !hp_k || !kern_i || !hp_i || !m_k || !ini_mixture || !seq_elbo || !n_iter || !converged || n_iter_max < 1)
    MB_FAIL(-1, "bad argument");
  const int K = n_clusters, U = c->U;
  if (K < 1 || K > MB_MAX_CLUSTERS) MB_FAIL(-1, "bad cluster count");
  const int64_t M = c->db.n_tasks;
  const size_t UU = (size_t)U * U;
  const int Pi = kern_i->n_hp, Pk = kern_k->n_hp;
  DevKern kdi = make_devkern(kern_i);
  MB_TRY(ensure_em_buffers(c, K));
  const size_t tab = (size_t)M * Pi * 8;
  MB_TRY(c->hp_k.ensure(2 * tab));
  double* cur = c->hp_k.as<double>();
  double* nxt = cur + (size_t)M * Pi;
  MB_TRY(h2d(cur, hp_i, tab, c->st));
  MB_TRY(c->tau.ensure((size_t)M * K * 8));
  std::vector<double> mixture(ini_mixture, ini_mixture + (size_t)M * K), post_mix((size_t)M * K), pm_host((size_t)K * U),
      new_m_k((size_t)K * U), new_hp_k(hp_k, hp_k + (size_t)K * Pk), htab((size_t)M * Pi);
  double prop[MB_MAX_CLUSTERS], new_prop[MB_MAX_CLUSTERS];
  {
    double colsum[MB_MAX_CLUSTERS];
    MB_TRY(leaf_sum_columns_host(c, mixture.data(), K, colsum));      // prop_mixture = colMeans(ini_mixture) (:1609-1612)
    const double m_all = (double)(c->shard_set ? c->g_total : M);
    for (int k = 0; k < K; ++k) prop[k] = colsum[k] / m_all;
  }
  double elbo = -INFINITY;
  *converged = 0;
  *n_iter = 0;
  for (int it = 1; it <= n_iter_max; ++it) {
    // ---- VE step (vem-magmaclust.R:32-150): hyper-posteriors with the current mixture, then the mixture update (not at
    // iteration 1, :130-131)
    MB_TRY(h2d(c->tau.p, mixture.data(), (size_t)M * K * 8, c->st));
    MB_TRY(h2d(c->m0.p, m_k, (size_t)K * U * 8, c->st));
    CU(cudaStreamSynchronize(c->st));
    MB_TRY(dev_e_step(c, K, kern_k, hp_k, kdi, cur, Pi, c->m0.as<double>(), c->tau.as<double>(), c->d_gridX.as<double>(), U,
                      nullptr, pen_diag, c->pm.as<double>(), c->pc.as<double>(), nullptr));
    if (it > 1) {
      MB_TRY(dev_update_mixture(c, K, kdi, cur, Pi, c->pm.as<double>(), c->pc.as<double>(), prop, pen_diag, post_mix.data()));
      MB_TRY(h2d(c->tau.p, post_mix.data(), (size_t)M * K * 8, c->st));
    } else {
      post_mix = mixture;
    }
    MB_TRY(d2h(pm_host.data(), c->pm.p, (size_t)K * U * 8, c->st));
    CU(cudaStreamSynchronize(c->st));
    if (fast_approx) {
      MB_TRY(dev_elbo_monitoring(c, K, kern_k, hp_k, kern_i, cur, Pi, post_mix.data(), prop, pen_diag, &elbo));
      seq_elbo[0] = elbo;
      *n_iter = 1;
      mixture = post_mix;
      break;
    }
    // ---- VM step on a copy of the HP table: a NaN leaves the last valid values in place (training.R:1672-1679)
    CU(cudaMemcpyAsync(nxt, cur, tab, cudaMemcpyDeviceToDevice, c->st));
    for (size_t q = 0; q < new_hp_k.size(); ++q) new_hp_k[q] = hp_k[q];
    MB_TRY(dev_vm_step(c, K, kern_k, new_hp_k.data(), kern_i, nxt, shared_hp_k, shared_hp_i, new_m_k.data(), pm_host.data(),
                       post_mix.data(), pen_diag, new_prop, nullptr));
    MB_TRY(d2h(htab.data(), nxt, tab, c->st));
    CU(cudaStreamSynchronize(c->st));
    bool bad = false;
    for (double v : new_hp_k) if (isnan(v)) bad = true;
    for (double v : htab) if (isnan(v)) { bad = true; break; }
    if (c->world > 1) {
      MB_TRY(c->sums.ensure(64 * 8));
      const double flag = bad ? 1.0 : 0.0;
      double tot = 0.0;
      CU(cudaMemcpyAsync(c->sums.as<double>() + 33, &flag, 8, cudaMemcpyHostToDevice, c->st));
      CU(cudaStreamSynchronize(c->st));
      MB_TRY(ctx_allreduce(c, c->sums.as<double>() + 33, 1, c->st));
      CU(cudaMemcpyAsync(&tot, c->sums.as<double>() + 33, 8, cudaMemcpyDeviceToHost, c->st));
      CU(cudaStreamSynchronize(c->st));
      bad = tot != 0.0;
    }
    if (bad) break;
    double new_elbo;
    MB_TRY(dev_elbo_monitoring(c, K, kern_k, new_hp_k.data(), kern_i, nxt, Pi, post_mix.data(), new_prop, pen_diag, &new_elbo,
                               c->vm_valid && !c->monitor_recompute));
    c->vm_valid = false;
    double diff = new_elbo - elbo;
    if (isnan(diff)) diff = -INFINITY;
    for (size_t q = 0; q < new_hp_k.size(); ++q) hp_k[q] = new_hp_k[q];
    for (size_t q = 0; q < new_m_k.size(); ++q) m_k[q] = new_m_k[q];
    for (int k = 0; k < K; ++k) prop[k] = new_prop[k];
    std::swap(cur, nxt);
    mixture = post_mix;
    elbo = new_elbo;
    seq_elbo[it - 1] = elbo;
    *n_iter = it;
    double eps = diff / fabs(elbo);
    if (isnan(eps)) eps = 1.0;
    if (fabs(eps) < cv_threshold) { *converged = 1; break; }
  }
  MB_TRY(d2h(hp_i, cur, tab, c->st));
  if (post_mean) MB_TRY(d2h(post_mean, c->pm.p, (size_t)K * U * 8, c->st));
  if (post_cov) MB_TRY(d2h(post_cov, c->pc.p, K * UU * 8, c->st));
  CU(cudaStreamSynchronize(c->st));
  if (mixture_out) memcpy(mixture_out, mixture.data(), (size_t)M * K * 8);
  if (prop_mixture_out) for (int k = 0; k < K; ++k) prop_mixture_out[k] = prop[k];
  return 0;
}
