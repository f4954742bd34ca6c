// lbfgsb_dev.cu -- device side of the lock-step batched L-BFGS-B: one thread per problem
// advances the reverse-communication state machine of lbfgsb.h; states stay in HBM between
// objective launches, so the M step never ships gradients to the host.
// Compiled with -fmad=false so the sums match the host build and the Python oracle bit for bit.
//
// Reference: the per-task stats::optim(..., method = "L-BFGS-B", control = list(factr = 1e13,
// maxit = 25)) calls of m_step / vm_step (/root/reference/R/em-magma.R:200-230,
// R/vem-magmaclust.R:249-261).
#include <cuda_runtime.h>

#include "lbfgsb.h"

__global__ void k_lbfgsb_init(LbfgsbState* st, int64_t m, int n, const double* x0,
                              int64_t x0_stride, double factr, double pgtol, int maxit) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  lb_init(&st[t], n, x0 + t * x0_stride, factr, pgtol, maxit);
}

// ---------------------------------------------------------------------------------------------------------
// Warp-cooperative advance: ONE WARP PER PROBLEM.  A single thread walking the whole state machine needs
// ~100 us for a step with five stored corrections (a few thousand dependent FP64 operations, divisions and
// square roots at ~25 cycles each), and that latency is paid once per lock-step round whatever the number of
// problems.  Here lane 0 runs the control flow and the inherently serial pieces (dcsrch, the three small
// Cholesky factorisations, the two 2col x 2col triangular solves); every independent element -- the dot
// products that fill WN and WT, the col triangular solves of formk, the entries of the Schur complement, the
// components of the new direction -- is computed by its own lane with EXACTLY the expression and summation
// order of lbfgsb.h, so the iterates are bit-identical to the host build and to oracle/lbfgsb.py.
// ---------------------------------------------------------------------------------------------------------
struct LbScratch {            // per-problem shared-memory scratch
  double wn[4 * LB_M * LB_M]; // 2m x 2m, leading dimension 2m
  double wt[LB_M * LB_M];
  double wv[2 * LB_M];
  double r[LB_NP], d[LB_NP];
  double rr, dr;
  int code, pad;
};
#define LBW_LD (2 * LB_M)
#define LBW_SYNC() __syncwarp()

// lb_matupd with the dot products spread over the lanes
__device__ __forceinline__ void lbw_matupd(LbfgsbState* s, LbScratch* w, int lane) {
  const int n = s->n;
  if (s->col == LB_M) {
    // shift the memory by one: every lane reads its elements first, then all write
    double v0 = 0.0, v1 = 0.0, v2 = 0.0, v3 = 0.0;
    const int tot = (LB_M - 1) * n;
    const int e0 = lane, e1 = lane + 32;
    if (e0 < tot) { v0 = s->ws[e0 / n + 1][e0 % n]; v1 = s->wy[e0 / n + 1][e0 % n]; }
    if (e1 < tot) { v2 = s->ws[e1 / n + 1][e1 % n]; v3 = s->wy[e1 / n + 1][e1 % n]; }
    double q0 = 0.0, q1 = 0.0;
    const int j = lane / (LB_M - 1), k = lane % (LB_M - 1);
    if (lane < (LB_M - 1) * (LB_M - 1)) { q0 = s->sy[j + 1][k + 1]; q1 = s->ss[j + 1][k + 1]; }
    LBW_SYNC();
    if (e0 < tot) { s->ws[e0 / n][e0 % n] = v0; s->wy[e0 / n][e0 % n] = v1; }
    if (e1 < tot) { s->ws[e1 / n][e1 % n] = v2; s->wy[e1 / n][e1 % n] = v3; }
    if (lane < (LB_M - 1) * (LB_M - 1)) { s->sy[j][k] = q0; s->ss[j][k] = q1; }
    if (lane == 0) s->col -= 1;
    LBW_SYNC();
  }
  const int col = s->col;
  if (lane < col) s->sy[col][lane] = lb_ddot(n, w->d, s->wy[lane]);
  else if (lane >= 16 && lane - 16 < col) s->ss[lane - 16][col] = lb_ddot(n, s->ws[lane - 16], w->d);
  LBW_SYNC();
  if (lane < n) { s->ws[col][lane] = w->d[lane]; s->wy[col][lane] = w->r[lane]; }
  if (lane == 0) {
    s->ss[col][col] = (s->stp == 1.0) ? s->dtd : s->stp * s->stp * s->dtd;
    s->sy[col][col] = w->dr;
    s->col = col + 1;
    s->theta = w->rr / w->dr;
  }
  LBW_SYNC();
}

// lb_formt: the entries of WT by the lanes, dpofa by lane 0.  Returns 1 when positive definite (uniform).
__device__ __forceinline__ int lbw_formt(LbfgsbState* s, LbScratch* w, int lane) {
  const int col = s->col;
  const double theta = s->theta;
  if (lane < LB_M * LB_M) {
    const int i = lane / LB_M, j = lane % LB_M;
    if (i == 0 && j < col) w->wt[j] = theta * s->ss[0][j];
    else if (i >= 1 && i < col && j >= i && j < col) {
      double ddum = 0.0;
      for (int k = 0; k < i; ++k) ddum += s->sy[i][k] * s->sy[j][k] / s->sy[k][k];
      w->wt[i * LB_M + j] = ddum + theta * s->ss[i][j];
    }
  }
  LBW_SYNC();
  int ok = 0;
  if (lane == 0) ok = lb_dpofa(w->wt, LB_M, col) == 0;
  LBW_SYNC();
  return __shfl_sync(0xffffffffu, ok, 0);
}

// lb_sd_core (col > 0): returns 1 when a factorisation fails (uniform)
__device__ __forceinline__ int lbw_sd_core(LbfgsbState* s, LbScratch* w, int lane) {
  const int n = s->n, col = s->col, m2 = 2 * col, ld = LBW_LD;
  const double theta = s->theta;
  double* wn = w->wn;
  for (int e = lane; e < 4 * LB_M * LB_M; e += 32) wn[e] = 0.0;
  LBW_SYNC();
  {
    // lane -> (iy, jy) of the upper triangle; lanes 0..14: WY'WY / theta block, lanes 16..30: the WS'WY block
    const int q = lane & 15;
    int iy = 0, rem = q;
    while (iy < LB_M && rem > iy) { rem -= iy + 1; ++iy; }   // q = iy (iy + 1) / 2 + jy, jy <= iy
    const int jy = rem;
    if (q < LB_M * (LB_M + 1) / 2 && iy < col) {
      if (lane < 16) {
        double v = lb_ddot(n, s->wy[iy], s->wy[jy]) / theta;
        if (iy == jy) v += s->sy[iy][iy];
        wn[jy * ld + iy] = v;
      } else {
        // pairs (iy2 <= jy2): element wn[jy2 * ld + col + iy2] = ws[iy2] . wy[jy2]
        const int iy2 = jy, jy2 = iy;
        wn[jy2 * ld + col + iy2] = lb_ddot(n, s->ws[iy2], s->wy[jy2]);
      }
    }
  }
  LBW_SYNC();
  int info = 0;
  if (lane == 0) info = lb_dpofa(wn, ld, col);
  LBW_SYNC();
  info = __shfl_sync(0xffffffffu, info, 0);
  if (info != 0) return 1;
  if (lane < col) lb_dtrsl(wn, ld, col, &wn[col + lane], ld, 1);
  LBW_SYNC();
  {
    const int q = lane;
    int a = 0, rem = q;
    while (a < LB_M && rem >= LB_M - a) { rem -= LB_M - a; ++a; }   // q -> (a, b = a + rem), rows of length LB_M - a
    const int b = a + rem;
    if (q < LB_M * (LB_M + 1) / 2 && a < col && b < col) {
      const int is = col + a, js = col + b;
      double sum = 0.0;
      for (int k = 0; k < col; ++k) sum += wn[k * ld + is] * wn[k * ld + js];
      wn[is * ld + js] += sum;
    }
  }
  LBW_SYNC();
  if (lane == 0) info = lb_dpofa(&wn[col * ld + col], ld, col);
  LBW_SYNC();
  info = __shfl_sync(0xffffffffu, info, 0);
  if (info != 0) return 1;
  if (lane < m2) {
    const int i = lane < col ? lane : lane - col;
    double t = 0.0;
    if (lane < col) { for (int k = 0; k < n; ++k) t += s->wy[i][k] * (-s->g[k]); w->wv[i] = t; }
    else { for (int k = 0; k < n; ++k) t += s->ws[i][k] * (-s->g[k]); w->wv[col + i] = theta * t; }
  }
  LBW_SYNC();
  if (lane == 0) {
    info = lb_dtrsl(wn, ld, m2, w->wv, 1, 1);
    if (info == 0) {
      for (int i = 0; i < col; ++i) w->wv[i] = -w->wv[i];
      info = lb_dtrsl(wn, ld, m2, w->wv, 1, 0);
    }
  }
  LBW_SYNC();
  info = __shfl_sync(0xffffffffu, info, 0);
  if (info != 0) return 1;
  if (lane < n) {
    double r = -s->g[lane];
    for (int jy = 0; jy < col; ++jy) r += s->wy[jy][lane] * w->wv[jy] / theta + s->ws[jy][lane] * w->wv[col + jy];
    r /= theta;
    s->z[lane] = s->x[lane] + 1.0 * r;
  }
  LBW_SYNC();
  return 0;
}

// lb_begin_iteration
__device__ __forceinline__ void lbw_begin_iteration(LbfgsbState* s, LbScratch* w, int lane) {
  for (;;) {
    // lb_search_direction
    for (;;) {
      if (s->col == 0) {
        if (lane == 0) {
          const int n = s->n;
          double f1 = 0.0;
          for (int i = 0; i < n; ++i) {
            double neggi = -s->g[i];
            s->d[i] = neggi;
            f1 -= neggi * neggi;
          }
          double f2 = -s->theta * f1;
          double dtm = -f1 / f2;
          if (dtm <= 0.0) dtm = 0.0;
          for (int i = 0; i < n; ++i) s->z[i] = s->x[i] + dtm * s->d[i];
        }
        LBW_SYNC();
        break;
      }
      if (lbw_sd_core(s, w, lane) == 0) break;
      if (lane == 0) lb_reset_memory(s);
      LBW_SYNC();
    }
    int again = 0;
    if (lane == 0) {
      const int n = s->n;
      for (int i = 0; i < n; ++i) s->d[i] = s->z[i] - s->x[i];
      lb_lnsrlb_start(s);
      const int res = lb_lnsrlb(s);
      if (res == 0) s->stage = LB_STAGE_LNSRCH;
      else again = lb_line_search_failed(s);
    }
    LBW_SYNC();
    again = __shfl_sync(0xffffffffu, again, 0);
    if (!again) return;
  }
}

// lb_advance; g points at the problem's gradient in global memory
__device__ __forceinline__ void lbw_advance(LbfgsbState* s, LbScratch* w, double f, const double* g, int lane) {
  // code: 0 = done, 1 = begin a new iteration, 2 = update the memory first
  if (lane == 0) {
    int code = 0;
    const int n = s->n;
    if (!isfinite(f)) lb_fail_nonfinite(s, f);
    else {
      s->f = f;
      for (int i = 0; i < n; ++i) s->g[i] = g[i];
      if (s->stage == LB_STAGE_START) {
        s->nfgv = 1;
        double sbgnrm = 0.0;
        for (int i = 0; i < n; ++i) sbgnrm = fmax(sbgnrm, fabs(s->g[i]));
        if (sbgnrm <= s->pgtol) s->active = 0;
        else code = 1;
      } else {
        const int res = lb_lnsrlb(s);
        if (res == 0 && s->iback < 20) code = 0;
        else if (res == 2 || res == 0) code = lb_line_search_failed(s) ? 1 : 0;
        else {
          s->iter += 1;
          double sbgnrm = 0.0;
          for (int i = 0; i < n; ++i) sbgnrm = fmax(sbgnrm, fabs(s->g[i]));
          s->niter += 1;
          const double ddum0 = lb_max3(fabs(s->fold), fabs(s->f), 1.0);
          if (s->niter > s->maxit) { s->fail = 1; s->active = 0; }
          else if (sbgnrm <= s->pgtol) s->active = 0;
          else if (s->fold - s->f <= s->tol * ddum0) s->active = 0;
          else {
            double ddum;
            for (int i = 0; i < n; ++i) w->r[i] = s->g[i] - s->r[i];
            w->rr = lb_ddot(n, w->r, w->r);
            if (s->stp == 1.0) {
              w->dr = s->gd - s->gdold;
              ddum = -s->gdold;
              for (int i = 0; i < n; ++i) w->d[i] = s->d[i];
            } else {
              w->dr = (s->gd - s->gdold) * s->stp;
              for (int i = 0; i < n; ++i) w->d[i] = s->stp * s->d[i];
              ddum = -s->gdold * s->stp;
            }
            if (w->dr <= LB_EPS * ddum) { s->updatd = 0; code = 1; }
            else { s->updatd = 1; code = 2; }
          }
        }
      }
    }
    w->code = code;
  }
  LBW_SYNC();
  const int code = w->code;
  if (code == 0) return;
  if (code == 2) {
    lbw_matupd(s, w, lane);
    if (!lbw_formt(s, w, lane)) {
      if (lane == 0) lb_reset_memory(s);
      LBW_SYNC();
    }
  }
  lbw_begin_iteration(s, w, lane);
}

#define LB_STRIDE_D 275   /* doubles per staged state: 274 + 1 pad */
#define LB_ADV_WARPS 8
#define LB_ADV_THREADS (32 * LB_ADV_WARPS)
// four CTAs (32 warps) per SM: 64 registers with 60 bytes of spills beat three CTAs at 77 registers (full rounds are bound by
// instruction issue over too few warps; measured at C2: -0.1 ms per EM step)
__global__ void __launch_bounds__(LB_ADV_THREADS, 4) k_lbfgsb_advance(LbfgsbState* st, int64_t m, int n, const double* f,
                                                                   const double* g, int* n_active) {
  extern __shared__ double lb_smem[];
  static_assert(sizeof(LbfgsbState) == 274 * 8, "LbfgsbState layout changed: update LB_STRIDE_D");
  static_assert(sizeof(LbScratch) % 8 == 0, "LbScratch must be a whole number of doubles");
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t t = (int64_t)blockIdx.x * LB_ADV_WARPS + warp;
  if (t >= m) return;
  if (!st[t].active) return;          // warp-uniform
  double* mine = lb_smem + (size_t)warp * (LB_STRIDE_D + sizeof(LbScratch) / 8 + 1);
  LbfgsbState* s = reinterpret_cast<LbfgsbState*>(mine);
  LbScratch* w = reinterpret_cast<LbScratch*>(mine + LB_STRIDE_D + 1);
  {
    const double* src = reinterpret_cast<const double*>(&st[t]);
    double v[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int e = lane + 32 * i; v[i] = (e < 274) ? src[e] : 0.0; }
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int e = lane + 32 * i; if (e < 274) mine[e] = v[i]; }
  }
  __syncwarp();
  lbw_advance(s, w, f[t], g + t * n, lane);
  __syncwarp();
  {
    double* dst = reinterpret_cast<double*>(&st[t]);
#pragma unroll
    for (int i = 0; i < 9; ++i) { const int e = lane + 32 * i; if (e < 274) dst[e] = mine[e]; }
  }
  if (lane == 0 && s->active) atomicAdd(n_active, 1);
}

// Entry point for the fused M-step kernel (mstep3.cu, linked as relocatable device code): the problem's state and scratch
// sit at the given offsets (in doubles) of the CALLING kernel's dynamic shared memory -- `extern __shared__` arrays of all
// translation units alias that one window, so the accesses below compile to shared-memory loads / stores, not generic ones.
// Executed by one full warp; this unit is built with -fmad=false, which is the point of the separate compilation.
__device__ __noinline__ void lbw_advance_smem(int state_off, int scratch_off, double f, const double* g, int lane) {
  extern __shared__ double lb_dyn_smem[];
  LbfgsbState* s = reinterpret_cast<LbfgsbState*>(lb_dyn_smem + state_off);
  LbScratch* w = reinterpret_cast<LbScratch*>(lb_dyn_smem + scratch_off);
  lbw_advance(s, w, f, g, lane);
  __syncwarp();
}

// x -> HP table; fail codes and evaluation counts out.
__global__ void k_lbfgsb_export(const LbfgsbState* st, int64_t m, int n, double* x,
                                int64_t x_stride, int* fail, int* nfgv, double* fval) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t >= m) return;
  for (int p = 0; p < n; ++p) x[t * x_stride + p] = st[t].x[p];
  if (fail) fail[t] = st[t].fail;
  if (nfgv) nfgv[t] = st[t].nfgv;
  if (fval) fval[t] = st[t].f;   // the objective at the exported x (every exit of the state machine keeps the pair consistent)
}

cudaError_t launch_lbfgsb_init(LbfgsbState* st, int64_t m, int n, const double* x0,
                               int64_t x0_stride, double factr, double pgtol, int maxit,
                               cudaStream_t s) {
  k_lbfgsb_init<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(st, m, n, x0, x0_stride, factr, pgtol,
                                                            maxit);
  return cudaGetLastError();
}

cudaError_t launch_lbfgsb_advance(LbfgsbState* st, int64_t m, int n, const double* f,
                                  const double* g, int* n_active, cudaStream_t s) {
  const size_t smem = (size_t)LB_ADV_WARPS * (LB_STRIDE_D + sizeof(LbScratch) / 8 + 1) * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k_lbfgsb_advance, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  k_lbfgsb_advance<<<(unsigned)((m + LB_ADV_WARPS - 1) / LB_ADV_WARPS), LB_ADV_THREADS, smem, s>>>(st, m, n, f, g, n_active);
  return cudaGetLastError();
}

// Problems whose skip flag is set never start: their state stays at x0 (what k_lbfgsb_export writes back) with no
// evaluation counted.  Used by the batched train_gp_clust, where every new task runs its own EM loop and stops at its
// own iteration (training.R:2098-2215).
__global__ void k_lbfgsb_mask(LbfgsbState* st, int64_t m, const int32_t* skip) {
  int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (t < m && skip[t]) st[t].active = 0;
}
cudaError_t launch_lbfgsb_mask(LbfgsbState* st, int64_t m, const int32_t* skip, cudaStream_t s) {
  k_lbfgsb_mask<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(st, m, skip);
  return cudaGetLastError();
}

cudaError_t launch_lbfgsb_export(const LbfgsbState* st, int64_t m, int n, double* x,
                                 int64_t x_stride, int* fail, int* nfgv, double* fval, cudaStream_t s) {
  k_lbfgsb_export<<<(unsigned)((m + 127) / 128), 128, 0, s>>>(st, m, n, x, x_stride, fail, nfgv, fval);
  return cudaGetLastError();
}

// ---- host self test (mb_lbfgsb_selftest) ----------------------------------------------------
extern "C" int mb_lbfgsb_selftest(int which, double* x, int n, double factr, int maxit,
                                  double* fval, int32_t* n_eval, int32_t* fail) {
  if (which != 0 || n != 2) return -1;
  LbfgsbState s;
  lb_init(&s, n, x, factr, 0.0, maxit);
  while (s.active) {
    const double* v = s.x;
    double f = 100 * (v[1] - v[0] * v[0]) * (v[1] - v[0] * v[0]) + (1 - v[0]) * (1 - v[0]);
    double g[2] = {-400 * v[0] * (v[1] - v[0] * v[0]) - 2 * (1 - v[0]),
                   200 * (v[1] - v[0] * v[0])};
    lb_advance(&s, f, g);
  }
  for (int i = 0; i < n; ++i) x[i] = s.x[i];
  *fval = s.f;
  *n_eval = s.nfgv;
  *fail = s.fail;
  return 0;
}
