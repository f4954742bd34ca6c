// tasks.cu -- per-task fused kernels: one CTA per task, the task's N x N matrices resident in
// shared memory from covariance construction to the objective/gradient scalars (nothing
// but x, y, the gathered hyper-posterior block and P+1 doubles crosses HBM).
//
// Reference path replaced: list_kern_to_inv -> kern_to_inv -> kern_to_cov -> chol_inv_jitter
// (/root/reference/R/kernel-to-matrices.R:553-562,633-652,670-680), the task loop of e_step
// (R/em-magma.R:36-46,61-70) and ve_step (R/vem-magmaclust.R:68-98,113-127), logL_GP_mod +
// gr_GP_mod (R/likelihoods.R:155-174, R/gradients-likelihoods.R:71-97), elbo_clust_multi_GP
// + gr_clust_multi_GP (R/elbos.R:18-55, R/gradients-elbos.R:17-70) and update_mixture's
// likelihood table (R/vem-magmaclust.R:344-381).
#include "linalg.cuh"
#include "tasks.cuh"

#define LOG_2PI 1.8378770664093453

struct TaskSmem {
  double *A, *B, *C, *xs, *yv, *rv, *al, *c1, *be, *red;
  int *gi, *done;
};

__host__ __device__ inline size_t task_smem_bytes(int nmax, int ld, int D) {
  size_t dbl = (size_t)3 * nmax * ld + (size_t)nmax * D + 5 * (size_t)nmax + (MB_MAX_HP + 4) * 32;
  return dbl * sizeof(double) + 2 * (size_t)nmax * sizeof(int) + 16;
}

__device__ inline TaskSmem carve(double* base, int nmax, int ld, int D) {
  TaskSmem s;
  s.A = base;
  s.B = s.A + (size_t)nmax * ld;
  s.C = s.B + (size_t)nmax * ld;
  s.xs = s.C + (size_t)nmax * ld;
  s.yv = s.xs + (size_t)nmax * D;
  s.rv = s.yv + nmax;
  s.al = s.rv + nmax;
  s.c1 = s.al + nmax;
  s.be = s.c1 + nmax;
  s.red = s.be + nmax;
  s.gi = (int*)(s.red + (MB_MAX_HP + 4) * 32);
  s.done = s.gi + nmax;
  return s;
}

// Upper triangle of K(x, x; hp) [+ Sadd] with the cumulative jitter of `retries` failed
// attempts already on the diagonal (kernel-to-matrices.R:672-678: pen, then 10*pen, ... are
// added one after the other to the same matrix).
__device__ inline void build_cov_upper(double* A, int n, int ld, const double* xs, int D,
                                       const DevKern& kd, const double* hp, const double* Sadd,
                                       double pen, int retries) {
  double hpe[MB_MAX_HP];
  prep_hp(kd, hp, hpe);
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int i = e / n, j = e % n;
    if (j < i) continue;
    double v = cov_eval<false>(kd, hp, hpe, xs + i * D, xs + j * D, D, true, nullptr);
    if (Sadd) v += Sadd[i * ld + j];
    if (i == j) {
      double p = pen;
      for (int r = 0; r <= retries; ++r) { v += p; p = 10 * p; }
    }
    A[i * ld + j] = v;
  }
  __syncthreads();
}

// chol_inv_jitter on K(x; hp) [+ Sadd]: A <- full symmetric inverse, B destroyed.
// Returns the retry count, or -1 after TK_MAX_RETRY failures.  *logdet_chol receives
// sum_j log R_jj of the successful factorisation (same value in every thread).
__device__ inline int task_inverse(TaskSmem& s, int n, int ld, int D, const DevKern& kd,
                                   const double* hp, const double* Sadd, double pen,
                                   double* logdet_chol) {
  int retries = 0;
  for (;;) {
    build_cov_upper(s.A, n, ld, s.xs, D, kd, hp, Sadd, pen, retries);
    if (blk_chol_upper(s.A, n, ld) == 0) break;
    __syncthreads();
    if (++retries > TK_MAX_RETRY) return -1;
  }
  double v[1] = {0.0};
  for (int j = threadIdx.x; j < n; j += blockDim.x) v[0] += log(s.A[j * ld + j]);
  blk_reduce_sum<1>(v, s.red);
  *logdet_chol = v[0];
  blk_tri_inv_upper(s.A, s.B, n, ld);
  blk_lauum_upper(s.B, s.A, n, ld);
  return retries;
}

__device__ inline void load_task(TaskSmem& s, const TaskData& db, int64_t row0, int n) {
  for (int e = threadIdx.x; e < n * db.D; e += blockDim.x) s.xs[e] = db.X[row0 * db.D + e];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s.yv[i] = db.y[row0 + i];
    s.gi[i] = db.gidx ? db.gidx[row0 + i] : i;
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------
// Objective + gradient, one CTA per (active) task.
// ---------------------------------------------------------------------------------------
extern __shared__ double4 task_smem_raw[];

__global__ void __launch_bounds__(256) k_task_objective(TaskObjArgs a) {
  const int64_t t = a.active ? a.active[blockIdx.x] : blockIdx.x;
  if (a.lb && !a.lb[t].active) return;
  const int64_t row0 = a.db.offs[t];
  const int n = (int)(a.db.offs[t + 1] - row0);
  const int ld = a.ld, D = a.db.D, P = a.kd.n_hp, K = a.hpst.K, U = a.hpst.U;
  TaskSmem s = carve((double*)task_smem_raw, a.db.nmax, ld, D);
  double hp[MB_MAX_HP];
  {
    const double* src = a.lb ? a.lb[t].x : a.hp + t * a.hp_stride;
    for (int p = 0; p < P; ++p) hp[p] = src[p];
  }
  load_task(s, a.db, row0, n);
  const int tid = threadIdx.x, nt = blockDim.x;

  // TK_OBJ_MARG_MIX: the marginal value / gradient against each of the K cluster hyper-posteriors in turn, weighted by
  // the task's mixture probabilities (likelihoods.R:349-398, gradients-likelihoods.R:192-227); every other mode: one pass.
  const bool mixm = a.mode == TK_OBJ_MARG_MIX;
  const int npass = mixm ? K : 1;
  double f_tot = 0.0, g_tot[MB_MAX_HP];
#pragma unroll
  for (int p = 0; p < MB_MAX_HP; ++p) g_tot[p] = 0.0;
  int retries_max = 0;
  for (int kc = 0; kc < npass; ++kc) {
  const double* hmean = a.hpst.mean + (mixm ? (size_t)kc * U : 0);
  const double* hcov = a.hpst.cov ? a.hpst.cov + (mixm ? (size_t)kc * U * U : 0) : nullptr;
  const double wk = mixm ? a.hpst.tau[t * K + kc] : 1.0;
  const bool marg = a.mode == TK_OBJ_MARG || mixm;
  __syncthreads();
  // gather S (post_cov block / c2) into C; r and c1 vectors
  const double* Sadd = nullptr;
  if (a.mode != TK_OBJ_MIX) {
    for (int e = tid; e < n * n; e += nt) {
      int i = e / n, j = e % n;
      double v;
      if (a.mode == TK_OBJ_ELBO) {
        v = 0.0;  // corr2 += tau_k * (mu_k mu_k' + cov_k[t_i, t_i])   elbos.R:44-53
        for (int k = 0; k < K; ++k) {
          const double* mk = a.hpst.mean + (size_t)k * U;
          const double* ck = a.hpst.cov + (size_t)k * U * U;
          double tau = a.hpst.tau[t * K + k];
          v = v + tau * (mk[s.gi[i]] * mk[s.gi[j]] + ck[(size_t)s.gi[i] * U + s.gi[j]]);
        }
      } else {
        v = hcov ? hcov[(size_t)s.gi[i] * U + s.gi[j]] : 0.0;   // NULL: post_cov = 0 (sum_logL_GP)
      }
      s.C[i * ld + j] = v;
    }
    for (int i = tid; i < n; i += nt) {
      if (a.mode == TK_OBJ_ELBO) {
        double c = 0.0;
        for (int k = 0; k < K; ++k)
          c = c + a.hpst.tau[t * K + k] * a.hpst.mean[(size_t)k * U + s.gi[i]];
        s.c1[i] = c;
        s.rv[i] = s.yv[i];
      } else {
        s.rv[i] = s.yv[i] - hmean[s.gi[i]];
      }
    }
    __syncthreads();
    if (marg) Sadd = s.C;
  }

  double ldchol;
  int retries = task_inverse(s, n, ld, D, a.kd, hp, Sadd, a.pen, &ldchol);
  if (retries > retries_max) retries_max = retries;
  if (a.retries && tid == 0) a.retries[t] = retries < 0 ? retries : retries_max;
  if (retries < 0) {
    if (tid == 0) {
      if (a.mode == TK_OBJ_MIX) for (int k = 0; k < K; ++k) a.f[t * K + k] = NAN;
      else a.f[t] = NAN;
      if (a.want_grad) for (int p = 0; p < P; ++p) a.g[t * P + p] = NAN;
      if (mixm && a.fk) for (int k = 0; k < K; ++k) a.fk[t * K + k] = NAN;
    }
    return;
  }
  // log|det A|
  double logdetA;
  if (a.logdet_chol) {
    logdetA = -2.0 * ldchol;
  } else {
    for (int e = tid; e < n * n; e += nt) {
      int i = e / n, j = e % n;
      s.B[i * ld + j] = s.A[i * ld + j];
    }
    __syncthreads();
    logdetA = blk_lu_logabsdet(s.B, n, ld, s.done);
  }

  if (a.mode == TK_OBJ_MIX) {
    // one inverse, K Gaussian terms (vem-magmaclust.R:360-379 recomputes it K times)
    for (int k = 0; k < K; ++k) {
      const double* mk = a.hpst.mean + (size_t)k * U;
      const double* ck = a.hpst.cov + (size_t)k * U * U;
      for (int i = tid; i < n; i += nt) s.rv[i] = s.yv[i] - mk[s.gi[i]];
      __syncthreads();
      blk_gemv(s.al, s.A, s.rv, n, ld);
      double v[2] = {0.0, 0.0};
      for (int i = tid; i < n; i += nt) v[0] += s.rv[i] * s.al[i];
      for (int e = tid; e < n * n; e += nt) {
        int i = e / n, j = e % n;
        v[1] += s.A[i * ld + j] * ck[(size_t)s.gi[i] * U + s.gi[j]];
      }
      blk_reduce_sum<2>(v, s.red);
      if (tid == 0) a.f[t * K + k] = 0.5 * (n * LOG_2PI - logdetA + v[0]) + 0.5 * v[1];
    }
    return;
  }

  blk_gemv(s.al, s.A, s.rv, n, ld);  // alpha = A r
  if (a.mode == TK_OBJ_ELBO) blk_gemv(s.be, s.A, s.c1, n, ld);
  {
    double v[3] = {0.0, 0.0, 0.0};
    for (int i = tid; i < n; i += nt) {
      v[0] += s.rv[i] * s.al[i];
      if (a.mode == TK_OBJ_ELBO) v[2] += s.al[i] * s.c1[i];
    }
    if (!marg)
      for (int e = tid; e < n * n; e += nt) {
        int i = e / n, j = e % n;
        v[1] += s.A[i * ld + j] * s.C[i * ld + j];
      }
    blk_reduce_sum<3>(v, s.red);
    if (tid == 0) {
      double f = 0.5 * (n * LOG_2PI - logdetA + v[0]);
      if (a.mode == TK_OBJ_MOD) f = f + 0.5 * v[1];
      if (a.mode == TK_OBJ_ELBO) f = f - v[2] + 0.5 * v[1];
      if (mixm) {
        if (a.fk) a.fk[t * K + kc] = f;
        f_tot = f_tot + wk * f;
        if (kc == npass - 1) a.f[t] = f_tot;
      } else {
        a.f[t] = f;
      }
    }
  }
  if (!a.want_grad) { if (mixm) continue; return; }

  // common term W:  MOD  aa' + A S A - A ;  ELBO (a - 2b) a' + A c2 A - A ;  MARG aa' - A
  if (!marg) {
    blk_gemm_nn(s.B, s.C, s.A, n, ld);  // B = S A
    blk_gemm_nn(s.C, s.A, s.B, n, ld);  // C = A S A
  }
  double gacc[MB_MAX_HP], hpe[MB_MAX_HP];
  prep_hp(a.kd, hp, hpe);
#pragma unroll
  for (int p = 0; p < MB_MAX_HP; ++p) gacc[p] = 0.0;
  for (int e = tid; e < n * n; e += nt) {
    int i = e / n, j = e % n;
    if (j < i) continue;
    double lhs_i = s.al[i], lhs_j = s.al[j];
    if (a.mode == TK_OBJ_ELBO) { lhs_i -= 2 * s.be[i]; lhs_j -= 2 * s.be[j]; }
    double w = lhs_i * s.al[j] - s.A[i * ld + j];
    if (!marg) w += s.C[i * ld + j];
    if (i != j) {
      double w2 = lhs_j * s.al[i] - s.A[j * ld + i];
      if (!marg) w2 += s.C[j * ld + i];
      w += w2;
    }
    double dv[MB_MAX_HP];
    cov_eval<true>(a.kd, hp, hpe, s.xs + i * D, s.xs + j * D, D, true, dv);
    for (int p = 0; p < P; ++p) gacc[p] += w * dv[p];
  }
  // reduce P sums (in chunks of 4 to bound the scratch)
  for (int p0 = 0; p0 < P; p0 += 4) {
    double v[4];
    for (int q = 0; q < 4; ++q) v[q] = (p0 + q < P) ? gacc[p0 + q] : 0.0;
    blk_reduce_sum<4>(v, s.red);
    if (tid == 0)
      for (int q = 0; q < 4 && p0 + q < P; ++q) {
        if (mixm) {
          g_tot[p0 + q] = g_tot[p0 + q] + wk * (-0.5 * v[q]);
          if (kc == npass - 1) a.g[t * P + p0 + q] = g_tot[p0 + q];
        } else {
          a.g[t * P + p0 + q] = -0.5 * v[q];
        }
      }
  }
  }  // cluster passes
}


// list_kern_to_cov: covariance or one derivative matrix per task, written column-major.
__global__ void k_task_cov(TaskData db, DevKern kd, const double* hp, int64_t hp_stride,
                           int deriv, double* out, const int64_t* out_offs) {
  const int64_t t = blockIdx.x;
  const int64_t row0 = db.offs[t];
  const int n = (int)(db.offs[t + 1] - row0);
  const int D = db.D;
  double hpl[MB_MAX_HP];
  for (int p = 0; p < kd.n_hp; ++p) hpl[p] = hp[t * hp_stride + p];
  double hpe[MB_MAX_HP];
  prep_hp(kd, hpl, hpe);
  const double* x = db.X + row0 * D;
  double* o = out + out_offs[t];
  for (int e = threadIdx.x; e < n * n; e += blockDim.x) {
    int i = e % n, j = e / n;  // column-major output, consecutive threads -> consecutive rows
    double dv[MB_MAX_HP];
    double v;
    if (deriv < 0) v = cov_eval<false>(kd, hpl, hpe, x + i * D, x + j * D, D, true, nullptr);
    else { cov_eval<true>(kd, hpl, hpe, x + i * D, x + j * D, D, true, dv); v = dv[deriv]; }
    o[e] = v;
  }
}

// ---------------------------------------------------------------------------------------
// pred_magma / pred_magmaclust batched over new tasks (/root/reference/R/prediction.R:1146-1197, 2319-2404).
// Step 1, one CTA per (task, cluster): Gamma = K(x; hp) + cov_k[t, t] (+ jitter), Gamma^-1 by chol_inv_jitter,
// alpha = Gamma^-1 (y - mean_k[t]); both go to a scratch table in HBM (n_tasks x K x (nmax^2 + nmax) doubles).
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_task_pred_prepare(TaskPredArgs a) {
  const int64_t t = a.t0 + blockIdx.x / a.hpst.K;
  const int k = blockIdx.x % a.hpst.K;
  const int64_t row0 = a.db.offs[t];
  const int n = (int)(a.db.offs[t + 1] - row0);
  const int ld = a.ld, D = a.db.D, P = a.kd.n_hp, U = a.hpst.U, nmax = a.db.nmax;
  TaskSmem s = carve((double*)task_smem_raw, nmax, ld, D);
  double hp[MB_MAX_HP];
  for (int p = 0; p < P; ++p) hp[p] = a.hp[t * a.hp_stride + p];
  load_task(s, a.db, row0, n);
  const int tid = threadIdx.x, nt = blockDim.x;
  const double* mk = a.hpst.mean + (size_t)k * U;
  const double* ck = a.hpst.cov + (size_t)k * U * U;
  for (int e = tid; e < n * n; e += nt) {
    const int i = e / n, j = e % n;
    s.C[i * ld + j] = ck[(size_t)s.gi[i] * U + s.gi[j]];
  }
  for (int i = tid; i < n; i += nt) s.rv[i] = s.yv[i] - mk[s.gi[i]];
  __syncthreads();
  double ldchol;
  const int retries = task_inverse(s, n, ld, D, a.kd, hp, s.C, a.pen, &ldchol);
  if (a.retries && tid == 0) a.retries[t * a.hpst.K + k] = retries;
  double* G = a.ginv + ((size_t)(blockIdx.x)) * ((size_t)nmax * nmax + nmax);
  if (retries < 0) {
    for (int e = tid; e < n * n + n; e += nt) G[e] = NAN;
    return;
  }
  blk_gemv(s.al, s.A, s.rv, n, ld);
  for (int e = tid; e < n * n; e += nt) { const int i = e / n, j = e % n; G[e] = s.A[i * ld + j]; }
  for (int i = tid; i < n; i += nt) G[(size_t)n * n + i] = s.al[i];
}

// Step 2, one CTA per (block of PRED_T grid points, task): every thread owns one prediction point g.
//   kv_i  = k(x_i, xp_g) without noise (prediction.R:1155-1156)                       -- once, shared by the clusters
//   c_i   = kv_i + cov_k[t_i, g]                    mean_k = mean_k[g] + c' alpha_k    (prediction.R:1174-1186)
//   var_k = k(xp_g, xp_g) + cov_k[g, g] - c' Gamma_k^-1 c + exp(noise)                 (prediction.R:1189-1197)
//   K > 1: mean = sum_k tau_k mean_k ; var = sum_k tau_k (var_k + (mean_k - mean)^2)   (prediction.R:2386-2404)
// Gamma^-1 and alpha sit in shared memory (every lane reads the same entry: broadcast), kv / c are per-thread
// arrays in local memory (lane-interleaved by the hardware, i.e. unit stride across the warp).
#define PRED_T 128
__global__ void __launch_bounds__(PRED_T) k_task_pred_apply(TaskPredArgs a) {
  extern __shared__ double pred_sh[];
  const int64_t t = a.t0 + blockIdx.y;
  const int64_t row0 = a.db.offs[t];
  const int n = (int)(a.db.offs[t + 1] - row0);
  const int D = a.db.D, P = a.kd.n_hp, U = a.hpst.U, K = a.hpst.K, nmax = a.db.nmax;
  const int lda = n | 1;
  double* A = pred_sh;                       // n x lda
  double* al = A + (size_t)nmax * (nmax | 1);
  double* xo = al + nmax;                    // n x D
  int* gi = (int*)(xo + (size_t)nmax * D);
  const int tid = threadIdx.x;
  const int g = blockIdx.x * PRED_T + tid;
  const bool live = g < a.n_pred;
  double hp[MB_MAX_HP], hpe[MB_MAX_HP];
  for (int p = 0; p < P; ++p) hp[p] = a.hp[t * a.hp_stride + p];
  prep_hp(a.kd, hp, hpe);
  for (int e = tid; e < n * D; e += PRED_T) xo[e] = a.db.X[row0 * D + e];
  for (int i = tid; i < n; i += PRED_T) gi[i] = a.db.gidx[row0 + i];
  const int pg = live ? a.pred_index[g] : 0;
  const double* xp = a.x_pred + (size_t)(live ? g : 0) * D;
  __syncthreads();
  double kv[TK_MAXN], cv[TK_MAXN];
  for (int i = 0; i < n; ++i) kv[i] = cov_eval<false>(a.kd, hp, hpe, xo + i * D, xp, D, false, nullptr);
  const double kpp = cov_eval<false>(a.kd, hp, hpe, xp, xp, D, false, nullptr);
  const double noise = a.kd.has_noise ? hpe[P - 1] : 0.0;
  double mean_k[MB_MAX_CLUSTERS], var_k[MB_MAX_CLUSTERS];
  for (int k = 0; k < K; ++k) {
    const double* G = a.ginv + ((size_t)(blockIdx.y) * K + k) * ((size_t)nmax * nmax + nmax);
    __syncthreads();
    for (int e = tid; e < n * n; e += PRED_T) { const int i = e / n, j = e - i * n; A[i * lda + j] = G[e]; }
    for (int i = tid; i < n; i += PRED_T) al[i] = G[(size_t)n * n + i];
    __syncthreads();
    const double* mk = a.hpst.mean + (size_t)k * U;
    const double* ck = a.hpst.cov + (size_t)k * U * U;
    double m = 0.0;
    for (int i = 0; i < n; ++i) {
      const double c = kv[i] + ck[(size_t)gi[i] * U + pg];   // symmetric: row t_i, unit stride across the warp in g
      cv[i] = c;
      m += c * al[i];
    }
    double q = 0.0;
    for (int i = 0; i < n; ++i) {
      double w = 0.0;
      const double* Ai = A + i * lda;
      for (int j = 0; j < n; ++j) w += Ai[j] * cv[j];
      q += cv[i] * w;
    }
    mean_k[k] = mk[pg] + m;
    var_k[k] = ((kpp + ck[(size_t)pg * U + pg]) - q) + noise;
    if (live && a.out_mean_k) {
      a.out_mean_k[((size_t)t * K + k) * a.n_pred + g] = mean_k[k];
      a.out_var_k[((size_t)t * K + k) * a.n_pred + g] = var_k[k];
    }
  }
  if (!live) return;
  if (K == 1) {
    a.out_mean[(size_t)t * a.n_pred + g] = mean_k[0];
    a.out_var[(size_t)t * a.n_pred + g] = var_k[0];
  } else {
    double mm = 0.0, vv = 0.0;
    for (int k = 0; k < K; ++k) mm = mm + a.tau[t * K + k] * mean_k[k];
    for (int k = 0; k < K; ++k) {
      const double d = mean_k[k] - mm;
      vv = vv + a.tau[t * K + k] * (var_k[k] + d * d);
    }
    a.out_mean[(size_t)t * a.n_pred + g] = mm;
    a.out_var[(size_t)t * a.n_pred + g] = vv;
  }
}

// ---------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------
static int pick_threads(int nmax) { return nmax <= 16 ? 64 : (nmax <= 32 ? 128 : 256); }

int task_pick_ld(int nmax) { return nmax | 1; }

cudaError_t launch_task_objective(TaskObjArgs a, int n_launch, cudaStream_t st) {
  if (n_launch <= 0) return cudaSuccess;
  a.ld = task_pick_ld(a.db.nmax);
  size_t smem = task_smem_bytes(a.db.nmax, a.ld, a.db.D);
  cudaError_t e = cudaFuncSetAttribute(k_task_objective,
                                       cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  k_task_objective<<<n_launch, pick_threads(a.db.nmax), smem, st>>>(a);
  return cudaGetLastError();
}

cudaError_t launch_task_cov(TaskData db, DevKern kd, const double* hp, int64_t hp_stride,
                            int deriv, double* out, const int64_t* out_offs, cudaStream_t st) {
  k_task_cov<<<(unsigned)db.n_tasks, 256, 0, st>>>(db, kd, hp, hp_stride, deriv, out, out_offs);
  return cudaGetLastError();
}

cudaError_t launch_task_pred(TaskPredArgs a, int n_launch_tasks, cudaStream_t st) {
  a.ld = task_pick_ld(a.db.nmax);
  const size_t smem1 = task_smem_bytes(a.db.nmax, a.ld, a.db.D);
  cudaError_t e = cudaFuncSetAttribute(k_task_pred_prepare, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1);
  if (e != cudaSuccess) return e;
  k_task_pred_prepare<<<(unsigned)(n_launch_tasks * a.hpst.K), pick_threads(a.db.nmax), smem1, st>>>(a);
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  const size_t smem2 = ((size_t)a.db.nmax * (a.db.nmax | 1) + a.db.nmax + (size_t)a.db.nmax * a.db.D) * sizeof(double) +
                       (size_t)a.db.nmax * sizeof(int) + 16;
  e = cudaFuncSetAttribute(k_task_pred_apply, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
  if (e != cudaSuccess) return e;
  dim3 grid((unsigned)((a.n_pred + PRED_T - 1) / PRED_T), (unsigned)n_launch_tasks);
  k_task_pred_apply<<<grid, PRED_T, smem2, st>>>(a);
  return cudaGetLastError();
}
