// estep3.cu -- E step / VE step over all tasks (em-magma.R:25-82, vem-magmaclust.R:32-150) and list_kern_to_inv
// (kernel-to-matrices.R:586-652): every task's chol_inv_jitter(K_i) in shared memory (device code in tasks3.cuh), then
//   post_inv_k (+)= tau_ik K_i^{-1}   and   weighted_k (+)= tau_ik K_i^{-1} y_i     scattered onto the union grid.
//
// The reference adds the tasks one after the other (em-magma.R:39-46).  Here the GLOBAL task list is cut into a fixed number
// of LEAVES (contiguous task ranges, api.cu:leaf_count -- fixed by the number of tasks and the grid size alone, not by the
// number of GPUs, SMs or CTAs); every leaf owns one accumulator and receives its tasks IN TASK ORDER: persistent CTAs
// pull (leaf, position) items from a device counter, factorise concurrently and take turns at the leaf's accumulator through a
// ticket counter.  The leaf sums are then added in a fixed binary tree (dense.cu:k_reduce_tree).  Results are therefore
// bit-identical whatever the launch geometry and however the leaves are dealt to GPUs.  The adds inside the ordered section
// are red.global.add.f64 -- one addend per element and ticket, so the order of the floating-point sums is still fixed.
// With one accumulator per leaf (<= 128 x 323 KB at U = 201) that traffic stays in L2.
#include "tasks3.cuh"

extern __shared__ double4 task3_smem_raw[];

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_gpu(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int KS, int N8C, bool D1>
__global__ void __launch_bounds__(T3_THREADS, (N8C <= 8) ? 4 : 2) k_task_inverse3(const __grid_constant__ TaskInvArgs a) {
  const int D = a.db.D, K = a.K, U = a.U;
  const Sm3 s = carve3<N8C>((double*)task3_smem_raw, D);
  const Ln L;
  const int tid = threadIdx.x;
  if (a.sm_limit > 0 && a.queue) {   // dynamic dealing only: the remaining CTAs simply take more items
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)smid >= a.sm_limit) return;
  }
  // accumulating launches: the CTAs pull work items from an atomic counter, item g = (leaf g % n_leaves, position
  // g / n_leaves in that leaf) -- consecutive items go to different leaves, the tasks of one leaf are handed out in task
  // order, and whoever holds position p waits (ticket) for position p - 1, which was handed out EARLIER to a CTA that is
  // running: no deadlock whatever the grid size, and a CTA that meets a slow task (jitter retries) or a busy SM simply takes
  // fewer items.  list_kern_to_inv launches (no accumulators): plain grid-stride walk.
  int leaf = 0;
  int64_t leaf_t0 = 0;
  int64_t stride_task = a.part_inv ? 0 : (int64_t)blockIdx.x;
  double* pinv = nullptr;
  double* pw = nullptr;
  for (;;) {
    int64_t task;
    if (a.part_inv) {
      int g;
      if (a.queue) {
        if (tid == 0) s.flag[2] = atomicAdd(a.queue, 1);
        __syncthreads();
        g = s.flag[2];
        __syncthreads();
      } else {   // static dealing (grid = n_leaves * ctas_per_leaf, all resident): CTA -> leaf, its positions round-robin
        g = (int)stride_task * (int)gridDim.x + (int)blockIdx.x;
        ++stride_task;
      }
      if (g >= a.n_leaves * a.leaf_max) break;
      leaf = g % a.n_leaves;
      const int pos = g / a.n_leaves;
      leaf_t0 = a.leaf_offs[leaf];
      if (pos >= (int)(a.leaf_offs[leaf + 1] - leaf_t0)) continue;
      task = leaf_t0 + pos;
      pinv = a.part_inv + (size_t)leaf * a.part_stride;
      pw = a.part_w + (size_t)leaf * a.part_stride;
    } else {
      task = stride_task;
      if (task >= a.db.n_tasks) break;
      stride_task += gridDim.x;
    }
    const int64_t row0 = a.db.offs[task];
    const int n = (int)(a.db.offs[task + 1] - row0);
    const int n8 = t3_n8(n), np = 8 * n8;
    Cov<KS, D1> cov;
    cov.init(a.kd, a.hp + task * a.hp_stride, s.xs, D, s.red + 32);
    const int nd_t0 = cov.bind_pairs(a.db, task, true);
    pairs_prefetch3<N8C>(s, cov, n8);
    load_task3<N8C>(s, a.db, row0, n, np, nd_t0);
    double sumlog;
    T3_MARK_DECL;
    const int retries = task_inverse3<N8C>(s, n, n8, cov, a.pen, &sumlog, L, nullptr, true T3_MARK_PASS);
    if (a.retries && tid == 0) a.retries[task] = retries;
    const bool bad = retries < 0;
    if (a.inv_out) {
      double* o = a.inv_out + a.inv_offs[task];
      for (int e = tid; e < n * n; e += T3_THREADS) o[e] = bad ? NAN : symA_at<N8C>(s.T1, e % n, e / n);
    }
    if (pinv) {
      double* al = s.vout;
      gemv3<N8C, 1>(s, n8, s.yv, al, L);
      // The leaf's accumulator lives in L2; the ticket admits one task at a time and the grid positions of one task are
      // distinct, so every accumulator element receives at most one addend per ticket: a fire-and-forget reduction
      // (red.global.add.f64 -- no load round trip) adds exactly what a read-modify-write would, in the same order.
      // Only the upper triangle is added (half the L2 reductions -- their rate is what bounds this kernel); the lower
      // triangle of the final sum is its mirror image (api.cu: launch_mirror_upper), bit for bit what adding it would give.
      // Values and target offsets are prepared BEFORE the ticket wait, tile by tile in the fragment layout the inverse
      // already sits in (one 16-byte shared-memory load per tile and lane, no index arithmetic beyond the job table), nine
      // tiles per warp and pass -- the 36 upper tiles of a 64-point task in one pass: the ordered section, a chain through
      // all tasks of the leaf, is at most 18 back-to-back reductions per thread, the fence that completes them, and the
      // hand-over.
      constexpr int TPW = 9;
      const int ntri = t3_ntri(n8);
      double xv[2 * TPW];
      unsigned off[2 * TPW];
      auto prep = [&](int base) {
#pragma unroll
        for (int u = 0; u < TPW; ++u) {
          const int idx = base + L.warp + NW3 * u;
          off[2 * u] = 0xffffffffu; off[2 * u + 1] = 0xffffffffu;
          xv[2 * u] = 0.0; xv[2 * u + 1] = 0.0;
          if (idx < ntri) {
            const int e = s.tab[idx], i = (e >> 4) & 15, j = e & 15;
            const int r = 8 * i + L.g, c = 8 * j + 2 * L.t;
            const double2 av = *reinterpret_cast<const double2*>(s.T1 + (size_t)(e >> 8) * 64 + L.o_c);
            if (r < n) {
              const unsigned gr = (unsigned)s.gi[r];
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                if (c + h < n && c + h >= r) {   // upper triangle, diagonal included
                  const unsigned gc = (unsigned)s.gi[c + h];
                  xv[2 * u + h] = bad ? NAN : (h ? av.y : av.x);
                  off[2 * u + h] = (gr <= gc) ? gr * (unsigned)U + gc : gc * (unsigned)U + gr;
                }
              }
            }
          }
        }
      };
      prep(0);
      // ---- ordered section: wait for the previous task of this leaf
      const int ticket = (int)(task - leaf_t0);
      if (tid == 0 && !(a.diag_skip & 1)) {
        while (ld_acquire_gpu(a.tickets + leaf) != ticket) __nanosleep(32);
      }
      __syncthreads();
      for (int base = 0; base < ntri; base += NW3 * TPW) {
        if (base) prep(base);
        for (int k = 0; k < K; ++k) {
          const double tau = a.tau ? a.tau[task * K + k] : 1.0;
          double* pk = pinv + (size_t)k * U * U;
#pragma unroll
          for (int q = 0; q < 2 * TPW; ++q)
            if (off[q] != 0xffffffffu && !(a.diag_skip & 2)) atomicAdd(pk + off[q], a.tau ? tau * xv[q] : xv[q]);
        }
      }
      for (int k = 0; k < K; ++k) {
        const double tau = a.tau ? a.tau[task * K + k] : 1.0;
        double* wk = pw + (size_t)k * U;
        for (int i = tid; i < n; i += T3_THREADS) {
          const double v = bad ? NAN : al[i];
          atomicAdd(wk + s.gi[i], a.tau ? tau * v : v);
        }
      }
      __threadfence();
      __syncthreads();
      if (tid == 0) st_release_gpu(a.tickets + leaf, ticket + 1);
    }
    __syncthreads();
  }
}

// ---- launchers --------------------------------------------------------------------------------------
static inline bool is_se(const DevKern& kd) { return kd.n_terms == 1 && kd.types[0] == MB_KERN_SE; }

template <int KS, int N8C, bool D1>
static cudaError_t launch_inv3_t(const TaskInvArgs& a, int n_ctas, cudaStream_t st) {
  const size_t smem = t3_smem_bytes<N8C>(a.db.D);
  cudaError_t e = cudaFuncSetAttribute(k_task_inverse3<KS, N8C, D1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_task_inverse3<KS, N8C, D1><<<n_ctas, T3_THREADS, smem, st>>>(a);
  return cudaGetLastError();
}

// n_ctas: any positive count; with accumulators (a.part_inv) the resident capacity (task3_max_ctas_per_sm()) is the useful one.
cudaError_t launch_task_inverse3(TaskInvArgs a, int n_ctas, cudaStream_t st) {
  const bool small = a.db.nmax <= 64, d1 = a.db.D == 1;
  if (is_se(a.kd)) {
    if (d1) return small ? launch_inv3_t<KS_SE, 8, true>(a, n_ctas, st) : launch_inv3_t<KS_SE, 12, true>(a, n_ctas, st);
    return small ? launch_inv3_t<KS_SE, 8, false>(a, n_ctas, st) : launch_inv3_t<KS_SE, 12, false>(a, n_ctas, st);
  }
  return small ? launch_inv3_t<KS_GEN, 8, false>(a, n_ctas, st) : launch_inv3_t<KS_GEN, 12, false>(a, n_ctas, st);
}

// resident CTAs per SM of the E-step kernel: the smallest figure over its instantiations for this task size
int task3_max_ctas_per_sm(int nmax, int D) {
  int best = 1 << 30;
  auto probe = [&](auto kern, size_t smem) {
    int n = 0;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, T3_THREADS, smem) != cudaSuccess) n = 1;
    if (n < best) best = n;
  };
  if (nmax <= 64) {
    const size_t smem = t3_smem_bytes<8>(D);
    probe(k_task_inverse3<KS_SE, 8, true>, smem);
    probe(k_task_inverse3<KS_SE, 8, false>, smem);
    probe(k_task_inverse3<KS_GEN, 8, false>, smem);
  } else {
    const size_t smem = t3_smem_bytes<12>(D);
    probe(k_task_inverse3<KS_SE, 12, true>, smem);
    probe(k_task_inverse3<KS_SE, 12, false>, smem);
    probe(k_task_inverse3<KS_GEN, 12, false>, smem);
  }
  cudaGetLastError();
  return best > 0 && best < (1 << 30) ? best : 1;
}
