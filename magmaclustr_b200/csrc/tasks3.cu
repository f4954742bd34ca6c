// tasks3.cu -- the all-task objective kernel (logL_GP_mod + gr_GP_mod, ELBO pair, mixture likelihoods): one CTA of
// 4 warps per task, four tasks resident per SM.  Device code in tasks3.cuh.
#include "tasks3.cuh"

extern __shared__ double4 task3_smem_raw[];

template <int KS, int N8C, bool D1, int MODE_T>
__global__ void __launch_bounds__(T3_THREADS, (N8C <= 8) ? 4 : 2) k_task_objective3(const __grid_constant__ TaskObjArgs a) {
  const int64_t task = a.active ? a.active[blockIdx.x] : blockIdx.x;
  // every load the prologue needs is issued before the first one is waited for: the optimiser state's `active` flag and
  // HPs, the task's row range (three dependent L2 round trips otherwise), and the HP exponentials are evaluated while the
  // table rows are on their way
  const double* hp_src = a.lb ? a.lb[task].x : a.hp + task * a.hp_stride;
  double hp_pre[3] = {0.0, 0.0, 0.0};
  if (KS == KS_SE) { hp_pre[0] = hp_src[0]; hp_pre[1] = hp_src[1]; if (a.kd.has_noise) hp_pre[2] = hp_src[2]; }
  const int is_active = a.lb ? a.lb[task].active : 1;
  const int64_t row0 = a.db.offs[task];
  const int n = (int)(a.db.offs[task + 1] - row0);
  if (!is_active) return;
  const Sm3 s = carve3<N8C>((double*)task3_smem_raw, a.db.D);
  const Ln L;
  task_objective_eval3<KS, N8C, D1, MODE_T>(a, task, KS == KS_SE ? hp_pre : hp_src, row0, n, s, L, true);
}

// ---- launchers --------------------------------------------------------------------------------------
static inline bool is_se(const DevKern& kd) { return kd.n_terms == 1 && kd.types[0] == MB_KERN_SE; }

template <int KS, int N8C, bool D1, int MODE_T>
static cudaError_t launch_obj3_m(const TaskObjArgs& a, int n_launch, cudaStream_t st) {
  const size_t smem = t3_smem_bytes<N8C>(a.db.D);
  cudaError_t e = cudaFuncSetAttribute(k_task_objective3<KS, N8C, D1, MODE_T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_task_objective3<KS, N8C, D1, MODE_T><<<n_launch, T3_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}
template <int KS, int N8C, bool D1>
static cudaError_t launch_obj3_t(const TaskObjArgs& a, int n_launch, cudaStream_t st) {
  // the M step / monitoring objective of train_magma on the SE fast path gets its own instantiation
  if (KS == KS_SE && D1 && a.mode == TK_OBJ_MOD) return launch_obj3_m<KS, N8C, D1, (KS == KS_SE && D1) ? TK_OBJ_MOD : -1>(a, n_launch, st);
  return launch_obj3_m<KS, N8C, D1, -1>(a, n_launch, st);
}

cudaError_t launch_task_objective3(TaskObjArgs a, int n_launch, cudaStream_t st) {
  if (n_launch <= 0) return cudaSuccess;
  const bool small = a.db.nmax <= 64, d1 = a.db.D == 1;
  if (is_se(a.kd)) {
    if (d1) return small ? launch_obj3_t<KS_SE, 8, true>(a, n_launch, st) : launch_obj3_t<KS_SE, 12, true>(a, n_launch, st);
    return small ? launch_obj3_t<KS_SE, 8, false>(a, n_launch, st) : launch_obj3_t<KS_SE, 12, false>(a, n_launch, st);
  }
  return small ? launch_obj3_t<KS_GEN, 8, false>(a, n_launch, st) : launch_obj3_t<KS_GEN, 12, false>(a, n_launch, st);
}

