// bigtasks.cu -- per-task work for tasks too large for one CTA's shared memory (N > 96; BASELINE.json
// config 4: 256 tasks on a shared 4096-point grid).  Each task is a pipeline of multi-CTA kernels on
// HBM-resident N x N matrices (dense_big.cu); several tasks are in flight on independent streams so that
// the latency-bound panel factorisations of one task overlap the DMMA updates of the others ("batched"
// Cholesky).  Same arguments and results as the fused kernels of tasks3.cu:
//   objective : logL_GP_mod / gr_GP_mod        /root/reference/R/likelihoods.R:155-174, R/gradients-likelihoods.R:71-97
//               elbo_clust_multi_GP + gradient   R/elbos.R:18-55, R/gradients-elbos.R
//               update_mixture likelihoods       R/vem-magmaclust.R:344-381
//   inverse   : e_step / ve_step task loop       R/em-magma.R:36-46,61-70 ; list_kern_to_inv R/kernel-to-matrices.R:586-652
// Jitter escalation (R/kernel-to-matrices.R:670-680) is optimistic: every task is run with its current
// retry count, the host reads the per-task status once per pass and re-runs the (rare) failures.
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tasks.cuh"
#include "dense_big.cuh"
#include "tiles.cuh"

#define BT_LOG_2PI 1.8378770664093453
#define BT_NSP 296          /* partial sums of tr(A o S) */

struct BigSlot {
  cudaStream_t st = nullptr;
  cudaEvent_t ev = nullptr, ev_sc = nullptr;   // ev_sc: this slot's last E-step scatter
  double *W = nullptr, *Y = nullptr, *A = nullptr, *Bm = nullptr, *Sg = nullptr, *vec = nullptr, *gpart = nullptr,
         *spart = nullptr, *out = nullptr;
  int *fail = nullptr, *info = nullptr;
};
struct BigTaskPool {
  std::vector<BigSlot> slots;
  int nmax = 0;
  int64_t ldw = 0;
  size_t n_gpart = 0;
  cudaEvent_t ev_in = nullptr;
  int* d_status = nullptr;   // per task: retries used, or -1 = failed with the current jitter
  int* d_mask = nullptr;
  int* h_ints = nullptr;     // pinned, 2 * cap_tasks
  int64_t cap_tasks = 0;
};

static void slot_free(BigSlot& s) {
  cudaFree(s.W); cudaFree(s.Y); cudaFree(s.A); cudaFree(s.Bm); cudaFree(s.Sg); cudaFree(s.vec); cudaFree(s.gpart);
  cudaFree(s.spart); cudaFree(s.out); cudaFree(s.fail); cudaFree(s.info);
  if (s.ev) cudaEventDestroy(s.ev);
  if (s.ev_sc) cudaEventDestroy(s.ev_sc);
  if (s.st) cudaStreamDestroy(s.st);
  s = BigSlot();
}
void big_pool_destroy(BigTaskPool* p) {
  if (!p) return;
  for (auto& s : p->slots) slot_free(s);
  if (p->ev_in) cudaEventDestroy(p->ev_in);
  cudaFree(p->d_status); cudaFree(p->d_mask);
  if (p->h_ints) cudaFreeHost(p->h_ints);
  delete p;
}

#define BT_CU(call) MB_CUDA_OK(call)

static int pool_ensure(BigTaskPool** pp, int nmax, int64_t n_tasks, bool need_grad_bufs) {
  BigTaskPool* p = *pp;
  if (p && (p->nmax < nmax || p->cap_tasks < n_tasks)) { big_pool_destroy(p); p = nullptr; *pp = nullptr; }
  if (p) return 0;
  p = new BigTaskPool();
  *pp = p;
  p->nmax = nmax;
  p->ldw = ((int64_t)nmax + 7) / 8 * 8;
  p->cap_tasks = n_tasks;
  int S = 4;
  if (const char* e = getenv("MB_BIG_STREAMS")) { S = atoi(e); if (S < 1) S = 1; if (S > 16) S = 16; }
  if ((int64_t)S > n_tasks) S = (int)n_tasks;
  // at most ~40 GB of task workspaces
  const size_t per_slot = (size_t)5 * p->ldw * p->ldw * 8;
  while (S > 1 && per_slot * S > ((size_t)40 << 30)) --S;
  int gx, gy;
  big_gemm_grid(nmax, nmax, &gx, &gy);
  p->n_gpart = (size_t)gx * gy * MB_MAX_HP;
  BT_CU(cudaEventCreateWithFlags(&p->ev_in, cudaEventDisableTiming));
  BT_CU(cudaMalloc(&p->d_status, n_tasks * sizeof(int)));
  BT_CU(cudaMalloc(&p->d_mask, n_tasks * sizeof(int)));
  BT_CU(cudaMallocHost((void**)&p->h_ints, 2 * n_tasks * sizeof(int)));
  p->slots.resize(S);
  const size_t mat = (size_t)p->ldw * p->ldw * 8;
  for (auto& s : p->slots) {
    BT_CU(cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking));
    BT_CU(cudaEventCreateWithFlags(&s.ev, cudaEventDisableTiming));
    BT_CU(cudaEventCreateWithFlags(&s.ev_sc, cudaEventDisableTiming));
    BT_CU(cudaMalloc(&s.W, mat));
    BT_CU(cudaMalloc(&s.Y, mat));
    BT_CU(cudaMalloc(&s.A, mat));
    BT_CU(cudaMalloc(&s.Bm, mat));
    BT_CU(cudaMalloc(&s.Sg, mat));
    BT_CU(cudaMalloc(&s.vec, (size_t)8 * p->ldw * 8));
    BT_CU(cudaMalloc(&s.gpart, p->n_gpart * 8));
    BT_CU(cudaMalloc(&s.spart, BT_NSP * 8));
    BT_CU(cudaMalloc(&s.out, 8 * 8));
    BT_CU(cudaMalloc(&s.fail, sizeof(int)));
    BT_CU(cudaMalloc(&s.info, 4 * sizeof(int)));
  }
  (void)need_grad_bufs;
  return 0;
}

// ---- kernels -----------------------------------------------------------------------------------------
// upper triangle of K(x; hp) + noise + the escalated jitter, 32 x 32 elements per CTA
__global__ void __launch_bounds__(256) k_bt_cov(DevKern kd, const double* __restrict__ hp, const double* __restrict__ X,
                                                int n, int D, double* W, int64_t ldw, double pen, int retries) {
  const int j0 = blockIdx.x * 32, i0 = blockIdx.y * 32;
  if (j0 + 31 < i0) return;
  double hpe[MB_MAX_HP];
  prep_hp(kd, hp, hpe);
  const int j = j0 + threadIdx.x;
  for (int ii = threadIdx.y; ii < 32; ii += 8) {
    const int i = i0 + ii;
    if (i >= n || j >= n || j < i) continue;
    double v = cov_eval<false>(kd, hp, hpe, X + (int64_t)i * D, X + (int64_t)j * D, D, true, nullptr);
    if (i == j) { double q = pen; for (int r = 0; r <= retries; ++r) { v += q; q = 10 * q; } }
    W[(int64_t)i * ldw + j] = v;
  }
}

// r and c1 of the objective (tasks3.cu:k_task_objective3); kmix >= 0 selects one cluster (update_mixture)
__global__ void k_bt_vectors(int mode, int kmix, HyperPost h, int64_t task, const double* __restrict__ y,
                             const int32_t* __restrict__ gidx, int n, double* r, double* c1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int g = gidx ? gidx[i] : i;
  double rv = 0.0, cv = 0.0;
  if (mode == TK_OBJ_ELBO) {
    for (int k = 0; k < h.K; ++k) cv = cv + h.tau[task * h.K + k] * h.mean[(size_t)k * h.U + g];
    rv = y[i];
  } else if (mode == TK_OBJ_MIX) {
    rv = y[i] - h.mean[(size_t)kmix * h.U + g];
  } else {
    rv = y[i] - h.mean[g];
  }
  r[i] = rv;
  c1[i] = cv;
}

// y1 = A x1 (and y2 = A x2), one warp per row, fixed summation order
__global__ void __launch_bounds__(256) k_bt_gemv(const double* __restrict__ A, int64_t ld, int n, const double* __restrict__ x1,
                                                 double* y1, const double* __restrict__ x2, double* y2, const int* fail) {
  if (*fail) return;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= n) return;
  const double* Ar = A + (int64_t)row * ld;
  double s1 = 0.0, s2 = 0.0;
  for (int j = lane; j < n; j += 32) {
    const double a = Ar[j];
    s1 += a * x1[j];
    if (x2) s2 += a * x2[j];
  }
  for (int off = 16; off > 0; off >>= 1) {
    s1 += __shfl_xor_sync(FULLMASK, s1, off);
    s2 += __shfl_xor_sync(FULLMASK, s2, off);
  }
  if (lane == 0) { y1[row] = s1; if (x2) y2[row] = s2; }
}

__global__ void k_bt_lr(const double* al, const double* be, double* lr, int n, int elbo) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) lr[i] = elbo ? al[i] - 2 * be[i] : al[i];
}

// S of the objective: gathered post_cov block (MOD), sum_k tau_k (m_k m_k' + C_k) (ELBO), C_kmix (MIX)
__global__ void __launch_bounds__(256) k_bt_gather_S(int mode, int kmix, HyperPost h, int64_t task,
                                                     const int32_t* __restrict__ gidx, int n, double* Sg, int64_t lds) {
  const int j = blockIdx.x * 32 + threadIdx.x;
  if (j >= n) return;
  const int gj = gidx ? gidx[j] : j;
  const size_t U = h.U;
  for (int ii = threadIdx.y; ii < 32; ii += 8) {
    const int i = blockIdx.y * 32 + ii;
    if (i >= n) continue;
    const int gi = gidx ? gidx[i] : i;
    double v;
    if (mode == TK_OBJ_ELBO) {
      v = 0.0;
      for (int k = 0; k < h.K; ++k) {
        const double* mk = h.mean + (size_t)k * U;
        const double* ck = h.cov + (size_t)k * U * U;
        v = v + h.tau[task * h.K + k] * (mk[gi] * mk[gj] + ck[(size_t)gi * U + gj]);
      }
    } else if (mode == TK_OBJ_MIX) {
      v = h.cov[(size_t)kmix * U * U + (size_t)gi * U + gj];
    } else {
      v = h.cov[(size_t)gi * U + gj];
    }
    Sg[(int64_t)i * lds + j] = v;
  }
}

// partial sums of sum_ij A_ij S_ij, fixed assignment of elements to CTAs and a fixed tree inside
__global__ void __launch_bounds__(256) k_bt_trace_part(const double* __restrict__ A, int64_t lda, const double* __restrict__ S,
                                                       int64_t lds, int n, double* spart, const int* fail) {
  if (*fail) return;
  __shared__ double red[256];
  double v = 0.0;
  for (int i = blockIdx.x; i < n; i += gridDim.x) {
    const double* Ar = A + (int64_t)i * lda;
    const double* Sr = S + (int64_t)i * lds;
    for (int j = threadIdx.x; j < n; j += 256) v += Ar[j] * Sr[j];
  }
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) spart[blockIdx.x] = red[0];
}

__device__ inline double bt_block_sum(double v, double* red) {
  red[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) red[threadIdx.x] += red[threadIdx.x + s];
    __syncthreads();
  }
  const double t = red[0];
  __syncthreads();
  return t;
}

// value and gradient of one task from the pieces; one CTA of 256 threads
__global__ void __launch_bounds__(256) k_bt_finish(int64_t task, int mode, int kmix, int K, int n, int P, int want_grad,
                                                   const int* fail, const double* out, const double* spart, int nsp,
                                                   const double* r, const double* al, const double* c1,
                                                   const double* gpart, int ngp, int retries, double* f, double* g,
                                                   int32_t* retries_out, int* status) {
  __shared__ double red[256];
  const bool bad = *fail != 0;
  if (threadIdx.x == 0) {
    status[task] = bad ? -1 : retries;
    if (retries_out && (mode != TK_OBJ_MIX || kmix == 0)) retries_out[task] = bad ? -1 : retries;
  }
  if (bad) {
    if (threadIdx.x == 0) {
      if (mode == TK_OBJ_MIX) f[task * K + kmix] = NAN; else f[task] = NAN;
      if (want_grad) for (int p = 0; p < P; ++p) g[task * P + p] = NAN;
    }
    return;
  }
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) { v0 += r[i] * al[i]; if (mode == TK_OBJ_ELBO) v2 += al[i] * c1[i]; }
  for (int i = threadIdx.x; i < nsp; i += 256) v1 += spart[i];
  const double quad = bt_block_sum(v0, red), tr = bt_block_sum(v1, red), ac = bt_block_sum(v2, red);
  if (threadIdx.x == 0) {
    const double logdetA = -2.0 * out[1];
    double fv = 0.5 * (n * BT_LOG_2PI - logdetA + quad);
    if (mode == TK_OBJ_MOD || mode == TK_OBJ_MIX) fv = fv + 0.5 * tr;
    if (mode == TK_OBJ_ELBO) fv = fv - ac + 0.5 * tr;
    if (mode == TK_OBJ_MIX) f[task * K + kmix] = fv; else f[task] = fv;
  }
  if (!want_grad) return;
  for (int p = 0; p < P; ++p) {
    double v = 0.0;
    for (int i = threadIdx.x; i < ngp; i += 256) v += gpart[(size_t)i * MB_MAX_HP + p];
    const double t = bt_block_sum(v, red);
    if (threadIdx.x == 0) g[task * P + p] = -0.5 * t;
  }
}

// E step: slot accumulators  part_inv[k][g_i, g_j] += tau_k A_ij ,  part_w[k][g_i] += tau_k (A y)_i
__global__ void __launch_bounds__(256) k_bt_scatter(const double* __restrict__ A, int64_t lda, int n,
                                                    const int32_t* __restrict__ gidx, const double* tau, int64_t task, int K,
                                                    int U, double* part_inv, const double* al, double* part_w,
                                                    const int* fail) {
  if (*fail) return;
  const int j = blockIdx.x * 32 + threadIdx.x;
  if (j >= n) return;
  const int gj = gidx ? gidx[j] : j;
  for (int ii = threadIdx.y; ii < 32; ii += 8) {
    const int i = blockIdx.y * 32 + ii;
    if (i >= n) continue;
    const int gi = gidx ? gidx[i] : i;
    const double a = A[(int64_t)i * lda + j];
    for (int k = 0; k < K; ++k) {
      const double w = tau ? tau[task * K + k] : 1.0;
      part_inv[(size_t)k * U * U + (size_t)gi * U + gj] += w * a;
      if (j == 0 && part_w) part_w[(size_t)k * U + gi] += w * al[i];
    }
  }
}
__global__ void k_bt_copy_out(const double* __restrict__ A, int64_t lda, int n, double* out, const int* fail) {
  const int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (e >= (int64_t)n * n) return;
  const int i = (int)(e / n), j = (int)(e - (int64_t)i * n);
  out[e] = *fail ? NAN : A[(int64_t)i * lda + j];
}
__global__ void k_bt_status(const int* fail, int retries, int64_t task, int* status, int32_t* retries_out) {
  const int v = *fail ? -1 : retries;
  status[task] = v;
  if (retries_out) retries_out[task] = v;
}
__global__ void k_bt_active(const LbfgsbState* lb, int64_t m, int* mask) {
  const int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < m) mask[t] = lb[t].active;
}

// ---- host orchestration ------------------------------------------------------------------------------
static double total_jitter(double pen, int retries) {
  double tot = 0.0, p = pen;
  for (int r = 0; r <= retries; ++r) { tot += p; p = 10 * p; }
  return tot;
}

// covariance -> Cholesky -> inverse of one task on its slot; leaves A (ld = ldw) and out[1] = sum log diag R
static int enqueue_task_inverse(BigTaskPool* p, BigSlot& s, const TaskData& db, const DevKern& kd, const double* hp_dev,
                                int64_t row0, int n, double pen, int retries, int64_t* launches) {
  BT_CU(launch_big_clear_flag(s.fail, s.st));
  dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 8);
  k_bt_cov<<<grid, block, 0, s.st>>>(kd, hp_dev, db.X + row0 * db.D, n, db.D, s.W, p->ldw, pen, retries);
  *launches += 2;
  BT_CU(launch_big_cholinv_inplace(s.W, s.Y, p->ldw, s.A, p->ldw, n, total_jitter(pen, retries), retries, 0, s.fail, s.info,
                                   s.out, s.st, launches));
  return 0;
}

static int join_in(BigTaskPool* p, cudaStream_t st) {
  BT_CU(cudaEventRecord(p->ev_in, st));
  for (auto& s : p->slots) BT_CU(cudaStreamWaitEvent(s.st, p->ev_in, 0));
  return 0;
}
static int join_out(BigTaskPool* p, cudaStream_t st) {
  for (auto& s : p->slots) {
    BT_CU(cudaEventRecord(s.ev, s.st));
    BT_CU(cudaStreamWaitEvent(st, s.ev, 0));
  }
  return 0;
}

// The objective of every (active) task.  Results land in a.f / a.g exactly like launch_task_objective3.
int big_tasks_objective(BigTaskPool** pp, const TaskObjArgs& a, const int64_t* h_offs, cudaStream_t st, int64_t* launches) {
  const int64_t M = a.db.n_tasks;
  int rc = pool_ensure(pp, a.db.nmax, M, a.want_grad != 0);
  if (rc) return rc;
  BigTaskPool* p = *pp;
  if (a.mode == TK_OBJ_MARG) { mb_set_error(__FILE__, __LINE__, "marginal objective is not available on the large-task path"); return -4; }
  std::vector<int64_t> todo;
  if (a.lb) {
    k_bt_active<<<(unsigned)((M + 255) / 256), 256, 0, st>>>(a.lb, M, p->d_mask);
    ++*launches;
    BT_CU(cudaMemcpyAsync(p->h_ints, p->d_mask, M * sizeof(int), cudaMemcpyDeviceToHost, st));
    BT_CU(cudaStreamSynchronize(st));
    for (int64_t t = 0; t < M; ++t) if (p->h_ints[t]) todo.push_back(t);
  } else {
    for (int64_t t = 0; t < M; ++t) todo.push_back(t);
  }
  std::vector<int> rr(M, 0);
  const int P = a.kd.n_hp, K = a.hpst.K;
  const int S = (int)p->slots.size();
  while (!todo.empty()) {
    rc = join_in(p, st);
    if (rc) return rc;
    for (size_t q = 0; q < todo.size(); ++q) {
      const int64_t task = todo[q];
      BigSlot& s = p->slots[q % S];
      const int64_t row0 = h_offs[task];
      const int n = (int)(h_offs[task + 1] - row0);
      const double* hp_dev = a.lb ? (const double*)((const char*)(a.lb + task) + offsetof(LbfgsbState, x))
                                  : a.hp + task * a.hp_stride;
      rc = enqueue_task_inverse(p, s, a.db, a.kd, hp_dev, row0, n, a.pen, rr[task], launches);
      if (rc) return rc;
      const int32_t* gi = a.db.gidx ? a.db.gidx + row0 : nullptr;
      double *r = s.vec, *c1 = s.vec + p->ldw, *al = s.vec + 2 * p->ldw, *be = s.vec + 3 * p->ldw, *lr = s.vec + 4 * p->ldw;
      const int nv = (n + 255) / 256;
      dim3 g2((n + 31) / 32, (n + 31) / 32), b2(32, 8);
      const int nmix = (a.mode == TK_OBJ_MIX) ? K : 1;
      for (int km = 0; km < nmix; ++km) {
        k_bt_vectors<<<nv, 256, 0, s.st>>>(a.mode, km, a.hpst, task, a.db.y + row0, gi, n, r, c1);
        const bool elbo = a.mode == TK_OBJ_ELBO;
        k_bt_gemv<<<(n + 7) / 8, 256, 0, s.st>>>(s.A, p->ldw, n, r, al, elbo ? c1 : nullptr, be, s.fail);
        k_bt_gather_S<<<g2, b2, 0, s.st>>>(a.mode, km, a.hpst, task, gi, n, s.Sg, p->ldw);
        k_bt_trace_part<<<BT_NSP, 256, 0, s.st>>>(s.A, p->ldw, s.Sg, p->ldw, n, s.spart, s.fail);
        *launches += 4;
        int ngp = 0;
        if (a.want_grad && a.mode != TK_OBJ_MIX) {
          k_bt_lr<<<nv, 256, 0, s.st>>>(al, be, lr, n, elbo ? 1 : 0);
          // B = S A  (both symmetric: read as k-major operands)
          BigGemmArgs ga;
          memset(&ga, 0, sizeof(ga));
          ga.A = s.Sg; ga.lda = p->ldw; ga.B = s.A; ga.ldb = p->ldw; ga.C = s.Bm; ga.ldc = p->ldw;
          ga.M = n; ga.N = n; ga.K = n; ga.alpha = 1.0; ga.fail = s.fail;
          BT_CU(launch_big_gemm(ga, 1, 1, s.st));
          // upper tiles of A B contracted with dK
          int gx, gy;
          big_gemm_grid(n, n, &gx, &gy);
          ngp = gx * gy;
          BT_CU(cudaMemsetAsync(s.gpart, 0, (size_t)ngp * MB_MAX_HP * 8, s.st));
          memset(&ga, 0, sizeof(ga));
          ga.A = s.A; ga.lda = p->ldw; ga.B = s.Bm; ga.ldb = p->ldw; ga.M = n; ga.N = n; ga.K = n; ga.alpha = 1.0;
          ga.flags = BG_UPPER | BG_EPI_GRAD; ga.fail = s.fail;
          ga.kd = a.kd; ga.hp = hp_dev; ga.X = a.db.X + row0 * a.db.D; ga.D = a.db.D;
          ga.Ainv = s.A; ga.ldainv = p->ldw; ga.al = al; ga.lr = lr; ga.gpart = s.gpart;
          BT_CU(launch_big_gemm(ga, 1, 1, s.st));
          *launches += 3;
        }
        k_bt_finish<<<1, 256, 0, s.st>>>(task, a.mode, km, K, n, P, (a.want_grad && a.mode != TK_OBJ_MIX) ? 1 : 0, s.fail,
                                         s.out, s.spart, BT_NSP, r, al, c1, s.gpart, ngp, rr[task], a.f, a.g, a.retries,
                                         p->d_status);
        ++*launches;
      }
    }
    BT_CU(cudaGetLastError());
    rc = join_out(p, st);
    if (rc) return rc;
    BT_CU(cudaMemcpyAsync(p->h_ints + M, p->d_status, M * sizeof(int), cudaMemcpyDeviceToHost, st));
    BT_CU(cudaStreamSynchronize(st));
    std::vector<int64_t> again;
    for (int64_t task : todo)
      if (p->h_ints[M + task] < 0 && rr[task] < TK_MAX_RETRY) { rr[task]++; again.push_back(task); }
    todo.swap(again);
  }
  return 0;
}

// E step / list_kern_to_inv.  With accumulators (a.part_inv: one zero-initialised [K x U x U | K x U] block per leaf,
// a.part_stride apart; h_leaf_offs = first task of every leaf) the factorisations of consecutive tasks overlap on the pool's
// streams while their scatters are chained by events, so every leaf receives its tasks in task order (deterministic, and
// the same order as the shared-memory kernel of estep3.cu).
int big_tasks_inverse(BigTaskPool** pp, const TaskInvArgs& a, const int64_t* h_offs, const int64_t* h_inv_offs,
                      const int64_t* h_leaf_offs, cudaStream_t st, int64_t* launches) {
  const int64_t M = a.db.n_tasks;
  int rc = pool_ensure(pp, a.db.nmax, M, false);
  if (rc) return rc;
  BigTaskPool* p = *pp;
  const int S = (int)p->slots.size();
  std::vector<int64_t> todo;
  for (int64_t t = 0; t < M; ++t) todo.push_back(t);
  std::vector<int> rr(M, 0);
  std::vector<int> leaf_of(M, 0);
  if (a.part_inv && h_leaf_offs)
    for (int l = 0; l < a.n_leaves; ++l)
      for (int64_t t = h_leaf_offs[l]; t < h_leaf_offs[l + 1]; ++t) leaf_of[t] = l;
  while (!todo.empty()) {
    rc = join_in(p, st);
    if (rc) return rc;
    for (size_t q = 0; q < todo.size(); ++q) {
      const int64_t task = todo[q];
      const int si = (int)(q % S);
      BigSlot& s = p->slots[si];
      const int64_t row0 = h_offs[task];
      const int n = (int)(h_offs[task + 1] - row0);
      rc = enqueue_task_inverse(p, s, a.db, a.kd, a.hp + task * a.hp_stride, row0, n, a.pen, rr[task], launches);
      if (rc) return rc;
      const int32_t* gi = a.db.gidx ? a.db.gidx + row0 : nullptr;
      if (a.inv_out) {
        const int64_t tot = (int64_t)n * n;
        k_bt_copy_out<<<(unsigned)((tot + 255) / 256), 256, 0, s.st>>>(s.A, p->ldw, n, a.inv_out + h_inv_offs[task], s.fail);
        ++*launches;
      }
      if (a.part_inv) {
        double* al = s.vec;
        k_bt_gemv<<<(n + 7) / 8, 256, 0, s.st>>>(s.A, p->ldw, n, a.db.y + row0, al, nullptr, nullptr, s.fail);
        dim3 g2((n + 31) / 32, (n + 31) / 32), b2(32, 8);
        if (q > 0) BT_CU(cudaStreamWaitEvent(s.st, p->slots[(q - 1) % S].ev_sc, 0));   // scatters in task order
        k_bt_scatter<<<g2, b2, 0, s.st>>>(s.A, p->ldw, n, gi, a.tau, task, a.K, a.U,
                                          a.part_inv + (size_t)leaf_of[task] * a.part_stride, al,
                                          a.part_w ? a.part_w + (size_t)leaf_of[task] * a.part_stride : nullptr, s.fail);
        BT_CU(cudaEventRecord(s.ev_sc, s.st));
        *launches += 2;
      }
      k_bt_status<<<1, 1, 0, s.st>>>(s.fail, rr[task], task, p->d_status, a.retries);
      ++*launches;
    }
    BT_CU(cudaGetLastError());
    rc = join_out(p, st);
    if (rc) return rc;
    BT_CU(cudaMemcpyAsync(p->h_ints + M, p->d_status, M * sizeof(int), cudaMemcpyDeviceToHost, st));
    BT_CU(cudaStreamSynchronize(st));
    std::vector<int64_t> again;
    for (int64_t task : todo)
      if (p->h_ints[M + task] < 0 && rr[task] < TK_MAX_RETRY) { rr[task]++; again.push_back(task); }
    todo.swap(again);
  }
  return 0;
}
