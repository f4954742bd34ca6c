// lbfgsb.h -- R's optim(method = "L-BFGS-B") as a reverse-communication state machine,
// unbounded case, usable from host code and from a CUDA kernel (one thread per problem).
//
// Stands for stats::optim(par, fn, gr, method = "L-BFGS-B", control = list(factr = 1e13,
// maxit = 25)) as the reference calls it with no bounds
// (/root/reference/R/em-magma.R:157-168,178-189,215-226; vem-magmaclust.R:239-294).
// R's optimiser is its C translation of the Netlib L-BFGS-B 2.x code (third-party, not in
// /root/reference).  The algorithm is restated here with that code's subroutine structure
// (mainlb / cauchy / formk / cmprlb / subsm / lnsrlb / matupd / formt / dcsrch / dcstep /
// dpofa / dtrsl) specialised to nbd == 0:  lmm = 5, pgtol = 0, stpmx = 1e10,
// ftol = 1e-3, gtol = 0.9, xtol = 0.1; first step 1/||d||; memory refresh on a failed
// factorisation or line search; optim's driver stops once iter > maxit (fail = 1).
// All sums run sequentially in index order (no FMA contraction: compile this unit with
// -fmad=false) so host, device and the Python oracle agree step for step.
#pragma once
#include <math.h>
#include <stdint.h>

#if defined(__CUDACC__)
#define MB_HD __host__ __device__ __forceinline__
#else
#define MB_HD inline
#endif

#define LB_NP 12 /* = MB_MAX_HP */
#define LB_M 5

enum { LB_STAGE_START = 0, LB_STAGE_LNSRCH = 1 };
enum { LB_LS_START = 0, LB_LS_FG = 1, LB_LS_CONV = 2, LB_LS_WARN = 3, LB_LS_ERROR = 4 };

struct LbfgsbState {
  int n, col, iter, niter, nfgv, ifun, iback, stage, fail, active, updatd, maxit;
  int brackt, ls_stage, ls_task, pad_;
  double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;
  double theta, f, fold, stp, stpmx, dtd, dnorm, gd, gdold, tol, pgtol;
  double x[LB_NP], g[LB_NP], d[LB_NP], z[LB_NP], t[LB_NP], r[LB_NP];
  double ws[LB_M][LB_NP], wy[LB_M][LB_NP];
  double sy[LB_M][LB_M], ss[LB_M][LB_M];
};

#define LB_EPS 2.220446049250313e-16

// Loops whose bounds are template constants in the device specialisations (n = number of HPs, col = stored
// corrections) are fully unrolled there, so the small work arrays live in registers; the arithmetic and its
// order are those of the generic code.
#ifdef __CUDA_ARCH__
#define LB_UNROLL _Pragma("unroll")
#else
#define LB_UNROLL
#endif

MB_HD double lb_ddot(int n, const double* a, const double* b) {
  double s = 0.0;
  LB_UNROLL
  for (int i = 0; i < n; ++i) s += a[i] * b[i];
  return s;
}

// LINPACK dpofa on the leading n x n upper triangle of a (row stride lda).
MB_HD int lb_dpofa(double* a, int lda, int n) {
  LB_UNROLL
  for (int j = 0; j < n; ++j) {
    double s = 0.0;
    LB_UNROLL
    for (int k = 0; k < j; ++k) {
      double t = a[k * lda + j];
      LB_UNROLL
      for (int i = 0; i < k; ++i) t -= a[i * lda + k] * a[i * lda + j];
      t = t / a[k * lda + k];
      a[k * lda + j] = t;
      s += t * t;
    }
    s = a[j * lda + j] - s;
    if (s <= 0.0) return j + 1;
    a[j * lda + j] = sqrt(s);
  }
  return 0;
}

// LINPACK dtrsl with t upper triangular: trans = 0 solves t x = b (job 01), trans = 1
// solves t' x = b (job 11).  b has stride bs.
MB_HD int lb_dtrsl(const double* t, int ldt, int n, double* b, int bs, int trans) {
  LB_UNROLL
  for (int j = 0; j < n; ++j)
    if (t[j * ldt + j] == 0.0) return j + 1;
  if (!trans) {
    b[(n - 1) * bs] = b[(n - 1) * bs] / t[(n - 1) * ldt + (n - 1)];
    LB_UNROLL
    for (int jj = 2; jj <= n; ++jj) {
      int j = n - jj;
      double temp = -b[(j + 1) * bs];
      LB_UNROLL
      for (int i = 0; i <= j; ++i) b[i * bs] += temp * t[i * ldt + (j + 1)];
      b[j * bs] = b[j * bs] / t[j * ldt + j];
    }
  } else {
    b[0] = b[0] / t[0];
    LB_UNROLL
    for (int j = 1; j < n; ++j) {
      double s = 0.0;
      LB_UNROLL
      for (int i = 0; i < j; ++i) s += t[i * ldt + j] * b[i * bs];
      b[j * bs] = b[j * bs] - s;
      b[j * bs] = b[j * bs] / t[j * ldt + j];
    }
  }
  return 0;
}

MB_HD double lb_max3(double a, double b, double c) { return fmax(fmax(a, b), c); }

MB_HD void lb_dcstep(double* stx, double* fx, double* dx, double* sty, double* fy, double* dy,
                     double* stp, double fp, double dp, int* brackt, double stpmin,
                     double stpmax) {
  double gamma, p, q, r, s, sgnd, stpc, stpf, stpq, theta;
  sgnd = dp * (*dx / fabs(*dx));
  if (fp > *fx) {
    theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
    s = lb_max3(fabs(theta), fabs(*dx), fabs(dp));
    gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
    if (*stp < *stx) gamma = -gamma;
    p = (gamma - *dx) + theta;
    q = ((gamma - *dx) + gamma) + dp;
    r = p / q;
    stpc = *stx + r * (*stp - *stx);
    stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
    if (fabs(stpc - *stx) < fabs(stpq - *stx)) stpf = stpc;
    else stpf = stpc + (stpq - stpc) / 2.0;
    *brackt = 1;
  } else if (sgnd < 0.0) {
    theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
    s = lb_max3(fabs(theta), fabs(*dx), fabs(dp));
    gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
    if (*stp > *stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = ((gamma - dp) + gamma) + *dx;
    r = p / q;
    stpc = *stp + r * (*stx - *stp);
    stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
    if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc;
    else stpf = stpq;
    *brackt = 1;
  } else if (fabs(dp) < fabs(*dx)) {
    theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
    s = lb_max3(fabs(theta), fabs(*dx), fabs(dp));
    gamma = s * sqrt(fmax(0.0, (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
    if (*stp > *stx) gamma = -gamma;
    p = (gamma - dp) + theta;
    q = (gamma + (*dx - dp)) + gamma;
    r = p / q;
    if (r < 0.0 && gamma != 0.0) stpc = *stp + r * (*stx - *stp);
    else if (*stp > *stx) stpc = stpmax;
    else stpc = stpmin;
    stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
    if (*brackt) {
      if (fabs(stpc - *stp) < fabs(stpq - *stp)) stpf = stpc;
      else stpf = stpq;
      if (*stp > *stx) stpf = fmin(*stp + 0.66 * (*sty - *stp), stpf);
      else stpf = fmax(*stp + 0.66 * (*sty - *stp), stpf);
    } else {
      if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc;
      else stpf = stpq;
      stpf = fmin(stpmax, stpf);
      stpf = fmax(stpmin, stpf);
    }
  } else {
    if (*brackt) {
      theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
      s = lb_max3(fabs(theta), fabs(*dy), fabs(dp));
      gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
      if (*stp > *sty) gamma = -gamma;
      p = (gamma - dp) + theta;
      q = ((gamma - dp) + gamma) + *dy;
      r = p / q;
      stpc = *stp + r * (*sty - *stp);
      stpf = stpc;
    } else if (*stp > *stx) stpf = stpmax;
    else stpf = stpmin;
  }
  if (fp > *fx) {
    *sty = *stp; *fy = fp; *dy = dp;
  } else {
    if (sgnd < 0.0) { *sty = *stx; *fy = *fx; *dy = *dx; }
    *stx = *stp; *fx = fp; *dx = dp;
  }
  *stp = stpf;
}

// dcsrch with ftol 1e-3, gtol 0.9, xtol 0.1, stpmin 0, stpmax = s->stpmx.
MB_HD void lb_dcsrch(LbfgsbState* s, double f, double g, double* stp) {
  const double ftol = 1e-3, gtol = 0.9, xtol = 0.1, stpmin = 0.0;
  const double stpmax = s->stpmx;
  const double xtrapl = 1.1, xtrapu = 4.0, p5 = 0.5, p66 = 0.66;
  if (s->ls_task == LB_LS_START) {
    if (*stp < stpmin || *stp > stpmax || g >= 0.0) { s->ls_task = LB_LS_ERROR; return; }
    s->brackt = 0;
    s->ls_stage = 1;
    s->finit = f;
    s->ginit = g;
    s->gtest = ftol * s->ginit;
    s->width = stpmax - stpmin;
    s->width1 = s->width / p5;
    s->stx = 0.0; s->fx = s->finit; s->gx = s->ginit;
    s->sty = 0.0; s->fy = s->finit; s->gy = s->ginit;
    s->stmin = 0.0;
    s->stmax = *stp + xtrapu * *stp;
    s->ls_task = LB_LS_FG;
    return;
  }
  double ftest = s->finit + *stp * s->gtest;
  if (s->ls_stage == 1 && f <= ftest && g >= 0.0) s->ls_stage = 2;
  int task = LB_LS_FG;
  if (s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) task = LB_LS_WARN;
  if (s->brackt && s->stmax - s->stmin <= xtol * s->stmax) task = LB_LS_WARN;
  if (*stp == stpmax && f <= ftest && g <= s->gtest) task = LB_LS_WARN;
  if (*stp == stpmin && (f > ftest || g >= s->gtest)) task = LB_LS_WARN;
  if (f <= ftest && fabs(g) <= gtol * (-s->ginit)) task = LB_LS_CONV;
  if (task != LB_LS_FG) { s->ls_task = task; return; }
  if (s->ls_stage == 1 && f <= s->fx && f > ftest) {
    double fm = f - *stp * s->gtest;
    double fxm = s->fx - s->stx * s->gtest;
    double fym = s->fy - s->sty * s->gtest;
    double gm = g - s->gtest;
    double gxm = s->gx - s->gtest;
    double gym = s->gy - s->gtest;
    lb_dcstep(&s->stx, &fxm, &gxm, &s->sty, &fym, &gym, stp, fm, gm, &s->brackt, s->stmin,
              s->stmax);
    s->fx = fxm + s->stx * s->gtest;
    s->fy = fym + s->sty * s->gtest;
    s->gx = gxm + s->gtest;
    s->gy = gym + s->gtest;
  } else {
    lb_dcstep(&s->stx, &s->fx, &s->gx, &s->sty, &s->fy, &s->gy, stp, f, g, &s->brackt,
              s->stmin, s->stmax);
  }
  if (s->brackt) {
    if (fabs(s->sty - s->stx) >= p66 * s->width1) *stp = s->stx + p5 * (s->sty - s->stx);
    s->width1 = s->width;
    s->width = fabs(s->sty - s->stx);
  }
  if (s->brackt) {
    s->stmin = fmin(s->stx, s->sty);
    s->stmax = fmax(s->stx, s->sty);
  } else {
    s->stmin = *stp + xtrapl * (*stp - s->stx);
    s->stmax = *stp + xtrapu * (*stp - s->stx);
  }
  *stp = fmax(*stp, stpmin);
  *stp = fmin(*stp, stpmax);
  if ((s->brackt && (*stp <= s->stmin || *stp >= s->stmax)) ||
      (s->brackt && s->stmax - s->stmin <= xtol * s->stmax))
    *stp = s->stx;
  s->ls_task = LB_LS_FG;
}

MB_HD void lb_reset_memory(LbfgsbState* s) {
  s->col = 0;
  s->theta = 1.0;
  s->updatd = 0;
}

MB_HD void lb_init(LbfgsbState* s, int n, const double* x0, double factr, double pgtol,
                   int maxit) {
  s->n = n;
  for (int i = 0; i < n; ++i) s->x[i] = x0[i];
  s->tol = factr * LB_EPS;
  s->pgtol = pgtol;
  s->maxit = maxit;
  s->active = (n > 0 && maxit > 0) ? 1 : 0;
  s->fail = 0;
  s->iter = 0;
  s->niter = 0;
  s->nfgv = 0;
  s->ifun = 0;
  s->iback = 0;
  s->f = 0.0;
  s->fold = 0.0;
  s->stage = LB_STAGE_START;
  s->ls_task = LB_LS_START;
  lb_reset_memory(s);
}

// The col > 0 branch of the search direction (formk + cmprlb + subsm of mainlb without bounds): z from the
// compact L-BFGS matrices.  NT / CT > 0 fix n / col at compile time.  Returns 1 when a factorisation fails
// (the caller resets the memory and starts over).
template <int NT, int CT>
MB_HD int lb_sd_core(LbfgsbState* s) {
  const int n = NT > 0 ? NT : s->n;
  const int col = CT > 0 ? CT : s->col, m2 = 2 * col;
  const double theta = s->theta;
  double wn[2 * LB_M * 2 * LB_M];
  const int ld = 2 * LB_M;
  LB_UNROLL
  for (int i = 0; i < m2; ++i) {
    LB_UNROLL
    for (int j = 0; j < m2; ++j) wn[i * ld + j] = 0.0;
  }
  LB_UNROLL
  for (int iy = 0; iy < col; ++iy) {
    LB_UNROLL
    for (int jy = 0; jy <= iy; ++jy)
      wn[jy * ld + iy] = lb_ddot(n, s->wy[iy], s->wy[jy]) / theta;
    LB_UNROLL
    for (int jy = iy; jy < col; ++jy)
      wn[jy * ld + col + iy] = lb_ddot(n, s->ws[iy], s->wy[jy]);
    wn[iy * ld + iy] += s->sy[iy][iy];
  }
  int info = lb_dpofa(wn, ld, col);
  if (info == 0) {
    LB_UNROLL
    for (int js = col; js < m2; ++js) lb_dtrsl(wn, ld, col, &wn[js], ld, 1);
    LB_UNROLL
    for (int is = col; is < m2; ++is) {
      LB_UNROLL
      for (int js = is; js < m2; ++js) {
        double sum = 0.0;
        LB_UNROLL
        for (int k = 0; k < col; ++k) sum += wn[k * ld + is] * wn[k * ld + js];
        wn[is * ld + js] += sum;
      }
    }
    info = lb_dpofa(&wn[col * ld + col], ld, col);
  }
  if (info != 0) return 1;
  double r[LB_NP], wv[2 * LB_M];
  LB_UNROLL
  for (int i = 0; i < n; ++i) r[i] = -s->g[i];
  LB_UNROLL
  for (int i = 0; i < col; ++i) {
    double t1 = 0.0, t2 = 0.0;
    LB_UNROLL
    for (int k = 0; k < n; ++k) {
      t1 += s->wy[i][k] * r[k];
      t2 += s->ws[i][k] * r[k];
    }
    wv[i] = t1;
    wv[col + i] = theta * t2;
  }
  info = lb_dtrsl(wn, ld, m2, wv, 1, 1);
  if (info == 0) {
    LB_UNROLL
    for (int i = 0; i < col; ++i) wv[i] = -wv[i];
    info = lb_dtrsl(wn, ld, m2, wv, 1, 0);
  }
  if (info != 0) return 1;
  LB_UNROLL
  for (int jy = 0; jy < col; ++jy) {
    const int js = col + jy;
    LB_UNROLL
    for (int i = 0; i < n; ++i) r[i] += s->wy[jy][i] * wv[jy] / theta + s->ws[jy][i] * wv[js];
  }
  LB_UNROLL
  for (int i = 0; i < n; ++i) r[i] /= theta;
  LB_UNROLL
  for (int i = 0; i < n; ++i) s->z[i] = s->x[i] + 1.0 * r[i];
  return 0;
}

template <int NT>
MB_HD int lb_sd_col(LbfgsbState* s) {
  switch (s->col) {
    case 1: return lb_sd_core<NT, 1>(s);
    case 2: return lb_sd_core<NT, 2>(s);
    case 3: return lb_sd_core<NT, 3>(s);
    case 4: return lb_sd_core<NT, 4>(s);
    default: return lb_sd_core<NT, 5>(s);
  }
}
MB_HD int lb_sd_dispatch(LbfgsbState* s) {
#ifdef __CUDA_ARCH__
  static_assert(LB_M == 5, "lb_sd_col enumerates col = 1..5");
  if (s->n == 3) return lb_sd_col<3>(s);
  if (s->n == 4) return lb_sd_col<4>(s);
#endif
  return lb_sd_core<0, 0>(s);
}

// L222..L555 of mainlb: search direction d = z - x.  Returns after refreshing the memory
// as often as the factorisations demand.
MB_HD void lb_search_direction(LbfgsbState* s) {
  const int n = s->n;
  for (;;) {
    if (s->col == 0) {
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
    } else {
      if (lb_sd_dispatch(s)) { lb_reset_memory(s); continue; }
    }
    for (int i = 0; i < n; ++i) s->d[i] = s->z[i] - s->x[i];
    return;
  }
}

MB_HD void lb_lnsrlb_start(LbfgsbState* s) {
  const int n = s->n;
  s->dtd = lb_ddot(n, s->d, s->d);
  s->dnorm = sqrt(s->dtd);
  s->stpmx = 1e10;
  if (s->iter == 0) s->stp = (s->dnorm > 0.0) ? fmin(1.0 / s->dnorm, s->stpmx) : s->stpmx;
  else s->stp = 1.0;
  for (int i = 0; i < n; ++i) { s->t[i] = s->x[i]; s->r[i] = s->g[i]; }
  s->fold = s->f;
  s->ifun = 0;
  s->iback = 0;
  s->ls_task = LB_LS_START;
}

// returns 0 = FG wanted at the new x, 1 = NEW_X, 2 = failure (info != 0)
MB_HD int lb_lnsrlb(LbfgsbState* s) {
  const int n = s->n;
  s->gd = lb_ddot(n, s->g, s->d);
  if (s->ifun == 0) {
    s->gdold = s->gd;
    if (s->gd >= 0.0) return 2;
  }
  lb_dcsrch(s, s->f, s->gd, &s->stp);
  if (s->ls_task == LB_LS_ERROR) return 2;
  if (s->ls_task != LB_LS_CONV && s->ls_task != LB_LS_WARN) {
    s->ifun += 1;
    s->nfgv += 1;
    s->iback = s->ifun - 1;
    if (s->stp == 1.0) {
      for (int i = 0; i < n; ++i) s->x[i] = s->z[i];
    } else {
      for (int i = 0; i < n; ++i) s->x[i] = s->stp * s->d[i] + s->t[i];
    }
    return 0;
  }
  return 1;
}

// restore the previous iterate; returns 1 when the iteration restarts with an empty memory
MB_HD int lb_line_search_failed(LbfgsbState* s) {
  for (int i = 0; i < s->n; ++i) { s->x[i] = s->t[i]; s->g[i] = s->r[i]; }
  s->f = s->fold;
  if (s->col == 0) {
    s->fail = 52; /* ABNORMAL_TERMINATION_IN_LNSRCH */
    s->active = 0;
    return 0;
  }
  lb_reset_memory(s);
  return 1;
}

MB_HD void lb_begin_iteration(LbfgsbState* s) {
  for (;;) {
    lb_search_direction(s);
    lb_lnsrlb_start(s);
    int res = lb_lnsrlb(s);
    if (res == 0) { s->stage = LB_STAGE_LNSRCH; return; }
    if (!lb_line_search_failed(s)) return;
  }
}

MB_HD void lb_matupd(LbfgsbState* s, const double* d, const double* r, double rr, double dr) {
  const int n = s->n;
  if (s->col == LB_M) {
    for (int j = 0; j < LB_M - 1; ++j) {
      for (int i = 0; i < n; ++i) { s->ws[j][i] = s->ws[j + 1][i]; s->wy[j][i] = s->wy[j + 1][i]; }
      for (int k = 0; k < LB_M - 1; ++k) {
        s->sy[j][k] = s->sy[j + 1][k + 1];
        s->ss[j][k] = s->ss[j + 1][k + 1];
      }
    }
    s->col -= 1;
  }
  const int col = s->col;
  for (int j = 0; j < col; ++j) {
    s->sy[col][j] = lb_ddot(n, d, s->wy[j]);
    s->ss[j][col] = lb_ddot(n, s->ws[j], d);
  }
  for (int i = 0; i < n; ++i) { s->ws[col][i] = d[i]; s->wy[col][i] = r[i]; }
  s->ss[col][col] = (s->stp == 1.0) ? s->dtd : s->stp * s->stp * s->dtd;
  s->sy[col][col] = dr;
  s->col = col + 1;
  s->theta = rr / dr;
}

template <int CT>
MB_HD int lb_formt_core(LbfgsbState* s) {
  const int col = CT > 0 ? CT : s->col;
  const double theta = s->theta;
  double wt[LB_M * LB_M];
  LB_UNROLL
  for (int j = 0; j < col; ++j) wt[j] = theta * s->ss[0][j];
  LB_UNROLL
  for (int i = 1; i < col; ++i) {
    LB_UNROLL
    for (int j = i; j < col; ++j) {
      const int k1 = (i < j) ? i : j;
      double ddum = 0.0;
      LB_UNROLL
      for (int k = 0; k < k1; ++k) ddum += s->sy[i][k] * s->sy[j][k] / s->sy[k][k];
      wt[i * LB_M + j] = ddum + theta * s->ss[i][j];
    }
  }
  return lb_dpofa(wt, LB_M, col) == 0;
}
MB_HD int lb_formt(LbfgsbState* s) {
#ifdef __CUDA_ARCH__
  switch (s->col) {
    case 1: return lb_formt_core<1>(s);
    case 2: return lb_formt_core<2>(s);
    case 3: return lb_formt_core<3>(s);
    case 4: return lb_formt_core<4>(s);
    case 5: return lb_formt_core<5>(s);
    default: break;
  }
#endif
  return lb_formt_core<0>(s);
}

// A non-finite objective (a trial point of the line search at which the covariance overflows or never factors): R's optim
// stops with the error "L-BFGS-B needs finite values of 'fn'" and train_magma aborts.  Here the problem ends with fail = 99
// at its LAST ACCEPTED iterate (the start of the running line search; the initial point when the very first evaluation
// fails), so that the caller neither trains on from the failing trial point nor loses the other tasks' work; the count of
// such tasks is returned in stats[3] and their code by mb_last_task_status.
MB_HD void lb_fail_nonfinite(LbfgsbState* s, double f) {
  s->fail = 99;
  s->active = 0;
  if (s->stage == LB_STAGE_LNSRCH) {
    for (int i = 0; i < s->n; ++i) { s->x[i] = s->t[i]; s->g[i] = s->r[i]; }
    s->f = s->fold;
  } else {
    s->f = f;   // the initial point itself has no finite objective: s->f stays "the objective at s->x"
  }
}

// Feed (f, g) evaluated at s->x.
MB_HD void lb_advance(LbfgsbState* s, double f, const double* g) {
  if (!s->active) return;
  const int n = s->n;
  if (!isfinite(f)) { lb_fail_nonfinite(s, f); return; }
  s->f = f;
  for (int i = 0; i < n; ++i) s->g[i] = g[i];
  if (s->stage == LB_STAGE_START) {
    s->nfgv = 1;
    double sbgnrm = 0.0;
    for (int i = 0; i < n; ++i) sbgnrm = fmax(sbgnrm, fabs(s->g[i]));
    if (sbgnrm <= s->pgtol) { s->active = 0; return; }
    lb_begin_iteration(s);
    return;
  }
  int res = lb_lnsrlb(s);
  if (res == 0 && s->iback < 20) return;
  if (res == 2 || res == 0) {
    if (lb_line_search_failed(s)) lb_begin_iteration(s);
    return;
  }
  s->iter += 1;
  double sbgnrm = 0.0;
  for (int i = 0; i < n; ++i) sbgnrm = fmax(sbgnrm, fabs(s->g[i]));
  s->niter += 1;
  if (s->niter > s->maxit) { s->fail = 1; s->active = 0; return; }
  if (sbgnrm <= s->pgtol) { s->active = 0; return; }
  double ddum = lb_max3(fabs(s->fold), fabs(s->f), 1.0);
  if (s->fold - s->f <= s->tol * ddum) { s->active = 0; return; }
  double r[LB_NP], d[LB_NP];
  for (int i = 0; i < n; ++i) r[i] = s->g[i] - s->r[i];
  double rr = lb_ddot(n, r, r);
  double dr;
  if (s->stp == 1.0) {
    dr = s->gd - s->gdold;
    ddum = -s->gdold;
    for (int i = 0; i < n; ++i) d[i] = s->d[i];
  } else {
    dr = (s->gd - s->gdold) * s->stp;
    for (int i = 0; i < n; ++i) d[i] = s->stp * s->d[i];
    ddum = -s->gdold * s->stp;
  }
  if (dr <= LB_EPS * ddum) {
    s->updatd = 0;
  } else {
    s->updatd = 1;
    lb_matupd(s, d, r, rr, dr);
    if (!lb_formt(s)) lb_reset_memory(s);
  }
  lb_begin_iteration(s);
}
