// tasks2.cu -- version 2 of the per-task fused kernels: 8x8-tiled blocked algorithms whose
// tile products run on the FP64 tensor cores (mma.sync.m8n8k4.f64 -> SASS DMMA; tcgen05 has no
// FP64 operand type).  One CTA per task; the task's matrices live in two shared-memory buffers
// (row-major, leading dimension == 4 mod 8 so that every DMMA fragment load is bank-conflict
// free); the gathered post_cov block is consumed straight from L2 as A-fragments and the
// A S A product is contracted with dK/dtheta in registers, so neither S nor A S A is ever stored.
//
//   K = cov(x; theta) + jitter          build_cov_tiles      kernel-to-matrices.R:71-503, :672
//   K = R'R  (blocked, NB = 8)          chol_tiles           chol()      -> dpotrf('U')
//   X = R^{-1}                          trtri_tiles          chol2inv()  -> dpotri('U')
//   A = X X'                            lauum_tiles
//   log|det A| = -2 sum log R_jj        (or LU of A when the context asks for the reference's
//                                        determinant() path, likelihoods.R:37-41)
//   B = S A ; T = A B (upper tiles)     objective            likelihoods.R:155-174
//   g_p = -1/2 sum (aa' + T - A) o dK_p                      gradients-likelihoods.R:71-97
//
// Tasks are padded to a multiple of 8 points with an identity block (its Cholesky factor,
// inverse and log-determinant contributions are exactly I, I and 0).
#include "linalg.cuh"
#include "tasks.cuh"
#include "tiles.cuh"

#define NW2 8 /* warps per CTA of the v2 kernels (blockDim.x == 256) */

#define LOG_2PI 1.8378770664093453

struct Sm2 {
  double *A, *B, *Xd, *xs, *yv, *rv, *al, *c1, *be, *red;
  int *gi, *done, *flag;
};

__host__ __device__ inline int t2_np(int nmax) { return ((nmax + 7) / 8) * 8; }
__host__ __device__ inline int t2_ld(int nmax) { return t2_np(nmax) + 4; }

__host__ __device__ inline size_t t2_smem_bytes(int nmax, int D) {
  const int np = t2_np(nmax), ld = t2_ld(nmax), n8 = np / 8;
  size_t dbl = (size_t)2 * np * ld + (size_t)n8 * XD_STRIDE + (size_t)np * D + 5 * (size_t)np + 16 * 32;
  return dbl * sizeof(double) + (2 * (size_t)np + 8) * sizeof(int) + 16;
}

__device__ inline Sm2 carve2(double* base, int nmax, int D) {
  const int np = t2_np(nmax), ld = t2_ld(nmax), n8 = np / 8;
  Sm2 s;
  s.A = base;
  s.B = s.A + (size_t)np * ld;
  s.Xd = s.B + (size_t)np * ld;
  s.xs = s.Xd + (size_t)n8 * XD_STRIDE;
  s.yv = s.xs + (size_t)np * D;
  s.rv = s.yv + np;
  s.al = s.rv + np;
  s.c1 = s.al + np;
  s.be = s.c1 + np;
  s.red = s.be + np;
  s.gi = (int*)(s.red + 16 * 32);
  s.done = s.gi + np;
  s.flag = s.done + np;
  return s;
}

// Upper triangle of K (tiles i <= j) + cumulative jitter; identity on the padding.
__device__ inline void build_cov_tiles(double* A, int n, int n8, int ld, const double* xs, int D,
                                       const DevKern& kd, const double* hp, double pen, int retries) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int ntiles = (n8 * (n8 + 1)) >> 1;
  double hpe[MB_MAX_HP];
  prep_hp(kd, hp, hpe);
  for (int idx = warp; idx < ntiles; idx += NW2) {
    int i, j;
    tri_ij(idx, n8, i, j);
    {
      const int r = 8 * i + g;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = 8 * j + 2 * t + q;
        double v = 0.0;
        if (r < n && c < n) {
          if (c >= r) {
            v = cov_eval<false>(kd, hp, hpe, xs + r * D, xs + c * D, D, true, nullptr);
            if (r == c) {
              double p = pen;
              for (int k = 0; k <= retries; ++k) { v += p; p = 10 * p; }
            }
          }
        } else if (r == c) {
          v = 1.0;
        }
        A[r * ld + c] = v;
      }
    }
  }
  __syncthreads();
}

// chol_inv_jitter(K(x; hp)): A <- inverse (full, exactly symmetric).  Returns retries or -1.
// *sumlog = sum_j log R_jj (uniform).
__device__ inline int task_inverse2(Sm2& s, int n, int n8, int ld, int D, const DevKern& kd,
                                    const double* hp, double pen, double* sumlog) {
  int retries = 0;
  for (;;) {
    build_cov_tiles(s.A, n, n8, ld, s.xs, D, kd, hp, pen, retries);
    if (chol_tiles<NW2>(s.A, s.Xd, n8, ld, s.flag) == 0) break;
    __syncthreads();
    if (++retries > TK_MAX_RETRY) return -1;
  }
  double v[1] = {0.0};
  for (int j = threadIdx.x; j < n; j += blockDim.x) v[0] += log(s.A[j * ld + j]);
  blk_reduce_sum<1>(v, s.red);
  *sumlog = v[0];
  trtri_tiles<NW2>(s.A, s.Xd, s.B, n8, ld);
  lauum_tiles<NW2>(s.B, s.A, n8, ld);
  return retries;
}

__device__ inline void load_task2(Sm2& s, const TaskData& db, int64_t row0, int n, int np) {
  for (int e = threadIdx.x; e < np * db.D; e += blockDim.x)
    s.xs[e] = (e < n * db.D) ? db.X[row0 * db.D + e] : 0.0;
  for (int i = threadIdx.x; i < np; i += blockDim.x) {
    s.yv[i] = (i < n) ? db.y[row0 + i] : 0.0;
    s.gi[i] = (i < n) ? (db.gidx ? db.gidx[row0 + i] : i) : 0;
  }
  __syncthreads();
}

// y = A x for the padded np x np matrix (one warp per row, fixed reduction tree)
__device__ inline void gemv2(double* y, const double* A, const double* x, int np, int ld) {
  const int warp = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
  for (int r = warp; r < np; r += nw) {
    double v = 0.0;
    for (int k = lane; k < np; k += 32) v += A[r * ld + k] * x[k];
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULLMASK, v, off);
    if (lane == 0) y[r] = v;
  }
  __syncthreads();
}

extern __shared__ double4 task2_smem_raw[];

// S(r, c): gathered post_cov block / corr2 of the ELBO (elbos.R:44-53)
__device__ __forceinline__ double gather_S(const TaskObjArgs& a, int64_t task, const int* gi, int r,
                                           int c, int n) {
  if (r >= n || c >= n) return 0.0;
  const int U = a.hpst.U;
  if (a.mode == TK_OBJ_ELBO) {
    double v = 0.0;
    for (int k = 0; k < a.hpst.K; ++k) {
      const double* mk = a.hpst.mean + (size_t)k * U;
      const double* ck = a.hpst.cov + (size_t)k * U * U;
      v = v + a.hpst.tau[task * a.hpst.K + k] * (mk[gi[r]] * mk[gi[c]] + ck[(size_t)gi[r] * U + gi[c]]);
    }
    return v;
  }
  return a.hpst.cov[(size_t)gi[r] * U + gi[c]];
}

__global__ void __launch_bounds__(256, 2) k_task_objective2(TaskObjArgs a) {
  const int64_t task = a.active ? a.active[blockIdx.x] : blockIdx.x;
  if (a.lb && !a.lb[task].active) return;
  const int64_t row0 = a.db.offs[task];
  const int n = (int)(a.db.offs[task + 1] - row0);
  const int D = a.db.D, P = a.kd.n_hp, K = a.hpst.K, U = a.hpst.U;
  const int np = t2_np(n), n8 = np / 8, ld = t2_ld(a.db.nmax);
  Sm2 s = carve2((double*)task2_smem_raw, a.db.nmax, D);
  const int tid = threadIdx.x, nt = blockDim.x, warp = tid >> 5, nw = nt >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  double hp[MB_MAX_HP];
  {
    const double* src = a.lb ? a.lb[task].x : a.hp + task * a.hp_stride;
    for (int p = 0; p < P; ++p) hp[p] = src[p];
  }
  load_task2(s, a.db, row0, n, np);
  for (int i = tid; i < np; i += nt) {
    double r = 0.0, c = 0.0;
    if (i < n) {
      if (a.mode == TK_OBJ_ELBO) {
        for (int k = 0; k < K; ++k) c = c + a.hpst.tau[task * K + k] * a.hpst.mean[(size_t)k * U + s.gi[i]];
        r = s.yv[i];
      } else if (a.mode != TK_OBJ_MIX) {
        r = s.yv[i] - a.hpst.mean[s.gi[i]];
      }
    }
    s.rv[i] = r;
    s.c1[i] = c;
  }
  double sumlog;
  int retries = task_inverse2(s, n, n8, ld, D, a.kd, hp, a.pen, &sumlog);
  if (a.retries && tid == 0) a.retries[task] = retries;
  if (retries < 0) {
    if (tid == 0) {
      if (a.mode == TK_OBJ_MIX) for (int k = 0; k < K; ++k) a.f[task * K + k] = NAN;
      else a.f[task] = NAN;
      if (a.want_grad) for (int p = 0; p < P; ++p) a.g[task * P + p] = NAN;
    }
    return;
  }
  double logdetA;
  if (a.logdet_chol) {
    logdetA = -2.0 * sumlog;
  } else {
    for (int e = tid; e < n * n; e += nt) s.B[(e / n) * ld + e % n] = s.A[(e / n) * ld + e % n];
    __syncthreads();
    logdetA = blk_lu_logabsdet(s.B, n, ld, s.done);
  }

  if (a.mode == TK_OBJ_MIX) {
    for (int k = 0; k < K; ++k) {
      const double* mk = a.hpst.mean + (size_t)k * U;
      const double* ck = a.hpst.cov + (size_t)k * U * U;
      for (int i = tid; i < np; i += nt) s.rv[i] = (i < n) ? s.yv[i] - mk[s.gi[i]] : 0.0;
      __syncthreads();
      gemv2(s.al, s.A, s.rv, np, ld);
      double v[2] = {0.0, 0.0};
      for (int i = tid; i < n; i += nt) v[0] += s.rv[i] * s.al[i];
      for (int e = tid; e < n * n; e += nt) {
        int i = e / n, j = e % n;
        v[1] += s.A[i * ld + j] * ck[(size_t)s.gi[i] * U + s.gi[j]];
      }
      blk_reduce_sum<2>(v, s.red);
      if (tid == 0) a.f[task * K + k] = 0.5 * (n * LOG_2PI - logdetA + v[0]) + 0.5 * v[1];
    }
    return;
  }

  gemv2(s.al, s.A, s.rv, np, ld);
  if (a.mode == TK_OBJ_ELBO) gemv2(s.be, s.A, s.c1, np, ld);

  // B = S A with S taken from L2 as A-fragments; tr(A o S) on the way
  double tr = 0.0;
  const bool need_S = a.mode != TK_OBJ_MARG;
  if (need_S) {
    for (int i = warp; i < n8; i += NW2) {
      double sa[2 * (TK_MAXN / 8)];
      for (int k = 0; k < n8; ++k) {
        sa[2 * k] = gather_S(a, task, s.gi, 8 * i + g, 8 * k + t, n);
        sa[2 * k + 1] = gather_S(a, task, s.gi, 8 * i + g, 8 * k + t + 4, n);
        tr += sa[2 * k] * s.A[(8 * i + g) * ld + 8 * k + t] + sa[2 * k + 1] * s.A[(8 * i + g) * ld + 8 * k + t + 4];
      }
      if (a.want_grad) {
        for (int j = 0; j < n8; ++j) {
          double c0 = 0.0, c1 = 0.0;
          for (int k = 0; k < n8; ++k) {
            const double* Bt = s.A + (8 * k) * ld + 8 * j;
            dmma884(c0, c1, sa[2 * k], Bt[t * ld + g]);
            dmma884(c0, c1, sa[2 * k + 1], Bt[(t + 4) * ld + g]);
          }
          *reinterpret_cast<double2*>(s.B + (8 * i + g) * ld + 8 * j + 2 * t) = make_double2(c0, c1);
        }
      }
    }
  }
  {
    double v[3] = {0.0, tr, 0.0};
    for (int i = tid; i < n; i += nt) {
      v[0] += s.rv[i] * s.al[i];
      if (a.mode == TK_OBJ_ELBO) v[2] += s.al[i] * s.c1[i];
    }
    blk_reduce_sum<3>(v, s.red);   // also orders the B writes before the reads below
    if (tid == 0) {
      double f = 0.5 * (n * LOG_2PI - logdetA + v[0]);
      if (a.mode == TK_OBJ_MOD) f = f + 0.5 * v[1];
      if (a.mode == TK_OBJ_ELBO) f = f - v[2] + 0.5 * v[1];
      a.f[task] = f;
    }
  }
  if (!a.want_grad) return;

  // gradient: upper tiles of T = A B contracted with dK on the fly
  double gacc[MB_MAX_HP], hpe[MB_MAX_HP];
  prep_hp(a.kd, hp, hpe);
#pragma unroll
  for (int p = 0; p < MB_MAX_HP; ++p) gacc[p] = 0.0;
  const int ntiles = (n8 * (n8 + 1)) >> 1;
  for (int idx = warp; idx < ntiles; idx += NW2) {
    int i, j;
    tri_ij(idx, n8, i, j);
    {
      double c0 = 0.0, c1 = 0.0;
      if (need_S)
        for (int k = 0; k < n8; ++k)
          tile_mma<false, false, false>(c0, c1, s.A + (8 * i) * ld + 8 * k, ld, s.B + (8 * k) * ld + 8 * j, ld, g, t);
      const int r = 8 * i + g;
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int c = 8 * j + 2 * t + q;
        if (r < n && c < n && c >= r) {
          double tij = q ? c1 : c0;
          double lr = s.al[r], lc = s.al[c];
          if (a.mode == TK_OBJ_ELBO) { lr -= 2 * s.be[r]; lc -= 2 * s.be[c]; }
          double w;
          if (r == c) w = lr * s.al[r] + tij - s.A[r * ld + r];
          else w = (lr * s.al[c] + lc * s.al[r]) + 2.0 * (tij - s.A[r * ld + c]);
          double dv[MB_MAX_HP];
          cov_eval<true>(a.kd, hp, hpe, s.xs + r * D, s.xs + c * D, D, true, dv);
          for (int p = 0; p < P; ++p) gacc[p] += w * dv[p];
        }
      }
    }
  }
  for (int p0 = 0; p0 < P; p0 += 4) {
    double v[4];
    for (int q = 0; q < 4; ++q) v[q] = (p0 + q < P) ? gacc[p0 + q] : 0.0;
    blk_reduce_sum<4>(v, s.red);
    if (tid == 0)
      for (int q = 0; q < 4 && p0 + q < P; ++q) a.g[task * P + p0 + q] = -0.5 * v[q];
  }
}

// E step, persistent CTAs (see tasks.cu k_task_inverse for the accumulation contract)
__global__ void __launch_bounds__(256, 2) k_task_inverse2(TaskInvArgs a) {
  const int D = a.db.D, P = a.kd.n_hp, K = a.K, U = a.U;
  const int ld = t2_ld(a.db.nmax);
  Sm2 s = carve2((double*)task2_smem_raw, a.db.nmax, D);
  const int tid = threadIdx.x, nt = blockDim.x;
  const int64_t per = (a.db.n_tasks + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * per;
  const int64_t t1 = (t0 + per < a.db.n_tasks) ? t0 + per : a.db.n_tasks;
  double* pinv = a.part_inv ? a.part_inv + (size_t)blockIdx.x * K * U * U : nullptr;
  double* pw = a.part_w ? a.part_w + (size_t)blockIdx.x * K * U : nullptr;
  for (int64_t task = t0; task < t1; ++task) {
    const int64_t row0 = a.db.offs[task];
    const int n = (int)(a.db.offs[task + 1] - row0);
    const int np = t2_np(n), n8 = np / 8;
    double hp[MB_MAX_HP];
    for (int p = 0; p < P; ++p) hp[p] = a.hp[task * a.hp_stride + p];
    load_task2(s, a.db, row0, n, np);
    double sumlog;
    int retries = task_inverse2(s, n, n8, ld, D, a.kd, hp, a.pen, &sumlog);
    if (a.retries && tid == 0) a.retries[task] = retries;
    if (retries < 0) {
      for (int e = tid; e < n * n; e += nt) s.A[(e / n) * ld + e % n] = NAN;
      __syncthreads();
    }
    if (a.inv_out) {
      double* o = a.inv_out + a.inv_offs[task];
      for (int e = tid; e < n * n; e += nt) o[e] = s.A[(e % n) * ld + e / n];
    }
    if (pinv) {
      gemv2(s.al, s.A, s.yv, np, ld);
      for (int k = 0; k < K; ++k) {
        const double tau = a.tau ? a.tau[task * K + k] : 1.0;
        double* pk = pinv + (size_t)k * U * U;
        for (int e = tid; e < n * n; e += nt) {
          int i = e / n, j = e % n;
          size_t c = (size_t)s.gi[i] * U + s.gi[j];
          pk[c] = pk[c] + (a.tau ? tau * s.A[i * ld + j] : s.A[i * ld + j]);
        }
        double* wk = pw + (size_t)k * U;
        for (int i = tid; i < n; i += nt) wk[s.gi[i]] = wk[s.gi[i]] + (a.tau ? tau * s.al[i] : s.al[i]);
      }
    }
    __syncthreads();
  }
}

// ---- launchers --------------------------------------------------------------------------------
cudaError_t launch_task_objective2(TaskObjArgs a, int n_launch, cudaStream_t st) {
  if (n_launch <= 0) return cudaSuccess;
  size_t smem = t2_smem_bytes(a.db.nmax, a.db.D);
  cudaError_t e = cudaFuncSetAttribute(k_task_objective2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_task_objective2<<<n_launch, 256, smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_task_inverse2(TaskInvArgs a, int n_ctas, cudaStream_t st) {
  size_t smem = t2_smem_bytes(a.db.nmax, a.db.D);
  cudaError_t e = cudaFuncSetAttribute(k_task_inverse2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_task_inverse2<<<n_ctas, 256, smem, st>>>(a);
  return cudaGetLastError();
}

int task2_max_ctas_per_sm(int nmax, int D) {
  size_t smem = t2_smem_bytes(nmax, D);
  int n = 0;
  cudaFuncSetAttribute(k_task_inverse2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, k_task_inverse2, 256, smem);
  return n > 0 ? n : 1;
}
