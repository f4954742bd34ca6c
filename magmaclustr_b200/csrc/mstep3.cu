// mstep3.cu -- the per-task M step as ONE persistent kernel (em-magma.R:200-230, vem-magmaclust.R:249-261).
//
// The reference runs one stats::optim(method = "L-BFGS-B", factr = 1e13, maxit = 25) per task; the trajectories are
// independent.  Each persistent CTA (4 warps, four per SM) pulls a task from an atomic queue and runs that task's WHOLE
// trajectory: objective + gradient (tasks3.cuh, the very code of k_task_objective3) -> warp-cooperative L-BFGS-B step
// (lbfgsb_dev.cu:lbw_advance, compiled with -fmad=false and linked as relocatable device code so that the iterates stay
// bit-identical to the host build and to oracle/lbfgsb.py) -> next evaluation.  The task's rows, outputs, grid indices and
// job tables are loaded once per task; the optimiser state lives in HBM/L2 (mb LbfgsbState table) and is staged through the
// shared memory that the evaluation has just released.  No host in the loop, no launch per round, and a straggler task
// only occupies its own CTA while the others move on to the next tasks.
// The kernel resumes whatever state it finds (active problems only), so a few lock-step rounds may precede it.
#include "tasks3.cuh"

// lbfgsb_dev.cu: one L-BFGS-B step of the problem whose state / scratch sit at the given offsets (in doubles) of the
// calling kernel's dynamic shared memory; executed by ONE warp
extern __device__ void lbw_advance_smem(int state_off, int scratch_off, double f, const double* g, int lane);

extern __shared__ double4 task3_smem_raw[];

struct TaskOptArgs {
  TaskObjArgs obj;          // obj.lb = optimiser states (in/out), obj.f / obj.g = per-task scratch
  int* queue;               // zero-initialised work counter
  int64_t n_items;          // tasks in the queue
  int sm_limit;             // CTAs that land on an SM with %smid >= sm_limit leave at once (SMs kept free for the
                            // union-grid kernels of the hp_0 / cluster problems running on the other stream)
  unsigned long long* cycles;   // optional [2]: SM cycles spent in evaluations / in optimiser steps (thread 0 of every CTA)
};

#define LB_STATE_DOUBLES 274
#define LB_SCRATCH_OFF 280    /* doubles after the state */

template <int KS, int N8C, bool D1, int MODE_T>
__global__ void __launch_bounds__(T3_THREADS, (N8C <= 8) ? 4 : 2) k_task_optimise3(const __grid_constant__ TaskOptArgs o) {
  static_assert(sizeof(LbfgsbState) == LB_STATE_DOUBLES * 8, "LbfgsbState layout changed");
  {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid >= o.sm_limit) return;
  }
  const TaskObjArgs& a = o.obj;
  const Sm3 s = carve3<N8C>((double*)task3_smem_raw, a.db.D);
  const Ln L;
  const int tid = threadIdx.x;
  // the evaluation leaves T2 free: optimiser state + scratch are staged there
  double* stage = s.T2;
  const int state_off = (int)(stage - (double*)task3_smem_raw);
  // per-CTA bookkeeping lives in 64 extra bytes of shared memory behind the evaluation's layout, not in registers (the
  // evaluation needs all 128 of them): [0] queue item, [1] rows of the task, [2..3] first row, [4..7] cycle counters
  int* book = reinterpret_cast<int*>(reinterpret_cast<char*>(task3_smem_raw) + t3_smem_bytes<N8C>(a.db.D));
  long long* book_ll = reinterpret_cast<long long*>(book);
  if (tid == 0) { book_ll[2] = 0; book_ll[3] = 0; }
  for (;;) {
    if (tid == 0) book[0] = atomicAdd(o.queue, 1);
    __syncthreads();
    const int item = book[0];
    if (item >= o.n_items) break;
    {
      const LbfgsbState* gs0 = a.lb + item;
      if (!gs0->active) { __syncthreads(); continue; }
      if (tid == 0) {
        const int64_t r0 = a.db.offs[item];
        book_ll[1] = r0;
        book[1] = (int)(a.db.offs[item + 1] - r0);
      }
    }
    __syncthreads();
    bool first = true;
    for (;;) {
      const int64_t task = book[0];
      if (o.cycles && tid == 0) book_ll[4] = clock64();
      {
        const double* hp_src = a.lb[task].x;
        double hp_pre[3] = {0.0, 0.0, 0.0};
        if (KS == KS_SE) { hp_pre[0] = hp_src[0]; hp_pre[1] = hp_src[1]; if (a.kd.has_noise) hp_pre[2] = hp_src[2]; }
        task_objective_eval3<KS, N8C, D1, MODE_T, false>(a, task, KS == KS_SE ? hp_pre : hp_src, book_ll[1], book[1], s, L, first);
      }
      first = false;
      __syncthreads();
      {
        const int64_t task2 = book[0];
        LbfgsbState* gs = const_cast<LbfgsbState*>(a.lb) + task2;
        if (o.cycles && tid == 0) { const long long c1 = clock64(); book_ll[2] += c1 - book_ll[4]; book_ll[4] = c1; }
        const double* src = reinterpret_cast<const double*>(gs);
        for (int e = tid; e < LB_STATE_DOUBLES; e += T3_THREADS) stage[e] = src[e];
        __syncthreads();
        if (L.warp == 0) lbw_advance_smem(state_off, state_off + LB_SCRATCH_OFF, a.f[task2], a.g + task2 * a.kd.n_hp, L.lane);
        __syncthreads();
        double* dst = reinterpret_cast<double*>(gs);
        for (int e = tid; e < LB_STATE_DOUBLES; e += T3_THREADS) dst[e] = stage[e];
      }
      const int still = reinterpret_cast<const LbfgsbState*>(stage)->active;
      __syncthreads();
      if (o.cycles && tid == 0) book_ll[3] += clock64() - book_ll[4];
      if (!still) break;
    }
  }
  if (o.cycles && tid == 0) {
    atomicAdd(o.cycles, (unsigned long long)book_ll[2]);
    atomicAdd(o.cycles + 1, (unsigned long long)book_ll[3]);
  }
}

// ---- launcher ---------------------------------------------------------------------------------------
static inline bool is_se(const DevKern& kd) { return kd.n_terms == 1 && kd.types[0] == MB_KERN_SE; }

template <int KS, int N8C, bool D1, int MODE_T>
static cudaError_t launch_opt3_m(const TaskOptArgs& o, int n_ctas, cudaStream_t st) {
  const size_t smem = t3_smem_bytes<N8C>(o.obj.db.D) + 64;   // + the kernel's bookkeeping words
  cudaError_t e = cudaFuncSetAttribute(k_task_optimise3<KS, N8C, D1, MODE_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_task_optimise3<KS, N8C, D1, MODE_T><<<n_ctas, T3_THREADS, smem, st>>>(o);
  return cudaGetLastError();
}
template <int KS, int N8C, bool D1>
static cudaError_t launch_opt3_t(const TaskOptArgs& o, int n_ctas, cudaStream_t st) {
  if (KS == KS_SE && D1 && o.obj.mode == TK_OBJ_MOD) return launch_opt3_m<KS, N8C, D1, (KS == KS_SE && D1) ? TK_OBJ_MOD : -1>(o, n_ctas, st);
  return launch_opt3_m<KS, N8C, D1, -1>(o, n_ctas, st);
}

// a: the objective (a.lb = initialised states); queue: device int, zeroed by the caller on the same stream.
// Modes TK_OBJ_MOD / TK_OBJ_ELBO, tasks of at most TK_MAXN points.
cudaError_t launch_task_optimise3(TaskObjArgs a, int* queue, int sm_count, int reserve_sms, unsigned long long* cycles,
                                  cudaStream_t st) {
  if (a.db.n_tasks <= 0) return cudaSuccess;
  if (a.mode != TK_OBJ_MOD && a.mode != TK_OBJ_ELBO) return cudaErrorInvalidValue;
  TaskOptArgs o;
  o.obj = a;
  o.obj.want_grad = 1;
  o.queue = queue;
  o.n_items = a.db.n_tasks;
  o.sm_limit = sm_count - reserve_sms;
  if (o.sm_limit < 1) o.sm_limit = 1;
  o.cycles = cycles;
  const bool small = a.db.nmax <= 64, d1 = a.db.D == 1;
  int n_ctas = sm_count * (small ? 4 : 2);
  if ((int64_t)n_ctas > a.db.n_tasks + (int64_t)reserve_sms * (small ? 4 : 2)) n_ctas = (int)a.db.n_tasks + reserve_sms * (small ? 4 : 2);
  if (is_se(a.kd)) {
    if (d1) return small ? launch_opt3_t<KS_SE, 8, true>(o, n_ctas, st) : launch_opt3_t<KS_SE, 12, true>(o, n_ctas, st);
    return small ? launch_opt3_t<KS_SE, 8, false>(o, n_ctas, st) : launch_opt3_t<KS_SE, 12, false>(o, n_ctas, st);
  }
  return small ? launch_opt3_t<KS_GEN, 8, false>(o, n_ctas, st) : launch_opt3_t<KS_GEN, 12, false>(o, n_ctas, st);
}
