/* magma_b200_shim.c -- the `.Call` layer that puts libmagma_b200.so under MagmaClustR.
 *
 * Drop-in for /root/reference/src/init.c:17-30 + src/RcppExports.cpp:14-75: it registers the five native builders under
 * their ORIGINAL names and arities (so R/RcppExports.R is untouched) and one routine per accelerated R function
 * (INTEGRATION.md 1).  Every routine only unpacks REAL()/INTEGER() pointers, allocates the result with R's
 * allocator, calls ONE C-ABI entry point of include/magma_b200.h and turns a non-zero status into Rf_error() AFTER the
 * library has returned (nothing longjmps through library frames).  R is single threaded: one context per session.
 *
 * Array conventions (the R wrappers of shim/magma_b200.R produce them):
 *   matrices of the five builders / kern_to_cov      R numeric matrices as they are (column-major n x d)
 *   X, grid_X of a resident dataset                  t(as.matrix(.)): d x rows column-major == rows x d row-major
 *   HP tables                                        t(as.matrix(hp[, -ID])): P x M column-major == M x P row-major
 *   mixtures (tau)                                   t(as.matrix(mixture[, -ID])): K x M column-major == M x K row-major
 *   post_mean of K clusters                          U x K matrix (column k = cluster k) == K x U row-major
 *   post_cov of K clusters                           U x U x K array (every slice symmetric)
 *   grid_index                                       match(db$Reference, all_input$Reference) - 1L, computed IN R
 *
 * Built only where R's headers exist (src/Makevars in INTEGRATION.md 2).  In this repository it is compiled and exercised
 * against the stand-in runtime of shim/mock/ (tests/test_shim.py) -- R is not installable in the build image.
 */
#include <R.h>
#include <Rinternals.h>
#include <R_ext/Rdynload.h>
#include <stdint.h>
#include <string.h>

#include "magma_b200.h"

static mb_context* g_ctx = NULL; /* one context per R session */

static mb_context* ctx(void) {
  if (!g_ctx) {
    int dev = 0;
    const char* e = getenv("MAGMA_B200_DEVICE");
    if (e) dev = atoi(e);
    if (mb_create(&g_ctx, dev) != 0) { g_ctx = NULL; Rf_error("magma_b200: %s", mb_last_error()); }
  }
  return g_ctx;
}
#define MB_CHECK(call)                                                              \
  do {                                                                              \
    int rc_ = (call);                                                               \
    if (rc_ != 0) Rf_error("magma_b200 (%d): %s", rc_, mb_last_error());            \
  } while (0)

static mb_kernel parse(SEXP kern, int has_noise) {
  mb_kernel k;
  MB_CHECK(mb_kernel_parse(CHAR(STRING_ELT(kern, 0)), has_noise, &k));
  return k;
}
static SEXP named_list(int n, const char** names, SEXP* vals) {
  SEXP out = PROTECT(Rf_allocVector(VECSXP, n));
  SEXP nm = PROTECT(Rf_allocVector(STRSXP, n));
  for (int i = 0; i < n; ++i) {
    SET_VECTOR_ELT(out, i, vals[i]);
    SET_STRING_ELT(nm, i, Rf_mkChar(names[i]));
  }
  Rf_setAttrib(out, R_NamesSymbol, nm);
  UNPROTECT(2);
  return out;
}
static SEXP copy_real(SEXP x) { /* in/out arguments are copied: R objects are never modified in place */
  SEXP out = PROTECT(Rf_allocVector(REALSXP, XLENGTH(x)));
  memcpy(REAL(out), REAL(x), sizeof(double) * (size_t)XLENGTH(x));
  UNPROTECT(1);
  return out;
}
static int64_t* offsets(SEXP offs, R_xlen_t* m) { /* double vector M + 1 -> int64 */
  *m = XLENGTH(offs) - 1;
  int64_t* o = (int64_t*)R_alloc((size_t)(*m + 1), sizeof(int64_t));
  for (R_xlen_t i = 0; i <= *m; ++i) o[i] = (int64_t)REAL(offs)[i];
  return o;
}
static SEXP stats_vec(const int64_t* st) {
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 4));
  for (int i = 0; i < 4; ++i) REAL(out)[i] = (double)st[i];
  UNPROTECT(1);
  return out;
}

/* ---- the five native builders: same names and arities as src/init.c:17-24 -------------------------------- */
#define BUILDER2(NAME, FN)                                                                     \
  SEXP NAME(SEXP m1, SEXP m2) {                                                                \
    int n1 = Rf_nrows(m1), n2 = Rf_nrows(m2), d = Rf_ncols(m1);                                \
    if (Rf_ncols(m2) != d) Rf_error("Incompatible number of dimensions"); /* code.cpp:11-13 */ \
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n1, n2));                                       \
    MB_CHECK(FN(ctx(), REAL(m1), n1, REAL(m2), n2, d, REAL(out)));                             \
    UNPROTECT(1);                                                                              \
    return out;                                                                                \
  }
#define BUILDER3(NAME, FN)                                                                     \
  SEXP NAME(SEXP m1, SEXP m2, SEXP par) {                                                      \
    int n1 = Rf_nrows(m1), n2 = Rf_nrows(m2), d = Rf_ncols(m1);                                \
    if (Rf_ncols(m2) != d) Rf_error("Incompatible number of dimensions");                      \
    SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n1, n2));                                       \
    MB_CHECK(FN(ctx(), REAL(m1), n1, REAL(m2), n2, d, Rf_asReal(par), REAL(out)));             \
    UNPROTECT(1);                                                                              \
    return out;                                                                                \
  }
BUILDER2(_MagmaClustR_cpp_dist, mb_cpp_dist)
BUILDER2(_MagmaClustR_cpp_prod, mb_cpp_prod)
BUILDER3(_MagmaClustR_cpp_perio, mb_cpp_perio)
BUILDER3(_MagmaClustR_cpp_perio_deriv, mb_cpp_perio_deriv)
BUILDER3(_MagmaClustR_cpp_noise, mb_cpp_noise)

/* ---- kern_to_cov / list_kern_to_cov / list_kern_to_inv / chol_inv_jitter (R/kernel-to-matrices.R) ---------- */
/* x1, x2: n x d R matrices (x2 = NULL: x1 against itself); deriv: 0-based HP index or -1 */
SEXP mbR_kern_to_cov(SEXP kern, SEXP has_noise, SEXP hp, SEXP x1, SEXP x2, SEXP deriv) {
  mb_kernel k = parse(kern, Rf_asLogical(has_noise));
  int n1 = Rf_nrows(x1), d = Rf_ncols(x1), n2 = Rf_isNull(x2) ? n1 : Rf_nrows(x2);
  /* the library reads points row-major: transpose the column-major R matrices */
  double* a = (double*)R_alloc((size_t)n1 * d, sizeof(double));
  for (int i = 0; i < n1; ++i) for (int q = 0; q < d; ++q) a[(size_t)i * d + q] = REAL(x1)[i + (size_t)q * n1];
  double* b = NULL;
  if (!Rf_isNull(x2)) {
    if (Rf_ncols(x2) != d) Rf_error("Incompatible number of dimensions");
    b = (double*)R_alloc((size_t)n2 * d, sizeof(double));
    for (int i = 0; i < n2; ++i) for (int q = 0; q < d; ++q) b[(size_t)i * d + q] = REAL(x2)[i + (size_t)q * n2];
  }
  SEXP out = PROTECT(Rf_allocMatrix(REALSXP, n1, n2));
  MB_CHECK(mb_kern_to_cov(ctx(), &k, REAL(hp), a, n1, b, n2, d, Rf_asInteger(deriv), REAL(out)));
  UNPROTECT(1);
  return out;
}
/* all tasks at once: offs (M + 1), X (d x rows), hp (P x M or P x 1 when shared); returns a list of M matrices */
SEXP mbR_list_kern_to_mat(SEXP kern, SEXP has_noise, SEXP offs, SEXP X, SEXP d_, SEXP hp, SEXP shared, SEXP deriv,
                          SEXP inverse, SEXP pen_diag) {
  mb_kernel k = parse(kern, Rf_asLogical(has_noise));
  R_xlen_t m;
  int64_t* o = offsets(offs, &m);
  int64_t* oo = (int64_t*)R_alloc((size_t)m, sizeof(int64_t));
  size_t total = 0;
  for (R_xlen_t t = 0; t < m; ++t) { size_t n = (size_t)(o[t + 1] - o[t]); oo[t] = (int64_t)total; total += n * n; }
  double* buf = (double*)R_alloc(total ? total : 1, sizeof(double));
  int32_t* retries = (int32_t*)R_alloc((size_t)m, sizeof(int32_t));
  MB_CHECK(mb_list_kern_to_mat(ctx(), &k, (int64_t)m, o, Rf_asInteger(d_), REAL(X), REAL(hp), Rf_asLogical(shared),
                               Rf_asInteger(deriv), Rf_asLogical(inverse), Rf_asReal(pen_diag), oo, buf, retries));
  SEXP out = PROTECT(Rf_allocVector(VECSXP, m));
  for (R_xlen_t t = 0; t < m; ++t) {
    int n = (int)(o[t + 1] - o[t]);
    SEXP mat = PROTECT(Rf_allocMatrix(REALSXP, n, n));
    memcpy(REAL(mat), buf + oo[t], sizeof(double) * (size_t)n * n);
    SET_VECTOR_ELT(out, t, mat);
    UNPROTECT(1);
  }
  UNPROTECT(1);
  return out;
}
SEXP mbR_chol_inv_jitter(SEXP mat, SEXP pen_diag) {
  int n = Rf_nrows(mat);
  if (Rf_ncols(mat) != n) Rf_error("chol_inv_jitter: a square matrix is expected");
  SEXP inv = PROTECT(Rf_allocMatrix(REALSXP, n, n));
  int32_t r = 0;
  double jit = 0;
  MB_CHECK(mb_chol_inv_jitter(ctx(), REAL(mat), n, n, Rf_asReal(pen_diag), REAL(inv), &r, &jit));
  UNPROTECT(1);
  return inv;
}

/* ---- resident dataset ----------------------------------------------------------------------------------- */
SEXP mbR_upload_dataset(SEXP offs, SEXP X, SEXP y, SEXP gidx, SEXP gridX, SEXP d) {
  R_xlen_t m;
  int64_t* o = offsets(offs, &m);
  int D = Rf_asInteger(d);
  const int has_grid = !Rf_isNull(gidx) && !Rf_isNull(gridX);
  MB_CHECK(mb_upload_dataset(ctx(), (int64_t)m, o, D, REAL(X), REAL(y), has_grid ? INTEGER(gidx) : NULL,
                             has_grid ? (int32_t)(XLENGTH(gridX) / D) : 0, has_grid ? REAL(gridX) : NULL));
  return R_NilValue;
}

/* ---- Magma: e_step / m_step / logL_monitoring / objective (R/em-magma.R, R/likelihoods.R) ---------------- */
SEXP mbR_e_step(SEXP kern_0, SEXP hp_0, SEXP kern_i, SEXP hp_i, SEXP shared, SEXP m_0, SEXP pen_diag) {
  mb_kernel k0 = parse(kern_0, 0), ki = parse(kern_i, 1); /* hp_0 never carries noise (training.R:856) */
  int U = (int)XLENGTH(m_0);
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, U));
  SEXP cov = PROTECT(Rf_allocMatrix(REALSXP, U, U));
  MB_CHECK(mb_e_step(ctx(), &k0, REAL(hp_0), &ki, REAL(hp_i), Rf_asLogical(shared), REAL(m_0), Rf_asReal(pen_diag),
                     REAL(mean), REAL(cov), NULL));
  const char* nm[] = {"mean", "cov"};
  SEXP v[] = {mean, cov};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}
SEXP mbR_m_step(SEXP kern_0, SEXP hp_0, SEXP kern_i, SEXP hp_i, SEXP shared_hp, SEXP m_0, SEXP post_mean, SEXP post_cov,
                SEXP pen_diag) {
  mb_kernel k0 = parse(kern_0, 0), ki = parse(kern_i, 1);
  SEXP h0 = PROTECT(copy_real(hp_0));
  SEXP hi = PROTECT(copy_real(hp_i));
  int64_t st[4] = {0, 0, 0, 0};
  MB_CHECK(mb_m_step(ctx(), &k0, REAL(h0), &ki, REAL(hi), Rf_asLogical(shared_hp), REAL(m_0), REAL(post_mean),
                     REAL(post_cov), Rf_asReal(pen_diag), st));
  SEXP stv = PROTECT(stats_vec(st));
  const char* nm[] = {"hp_0", "hp_i", "stats"};
  SEXP v[] = {h0, hi, stv};
  SEXP out = named_list(3, nm, v);
  UNPROTECT(3);
  return out;
}
SEXP mbR_logL_monitoring(SEXP kern_0, SEXP hp_0, SEXP kern_i, SEXP hp_i, SEXP shared, SEXP m_0, SEXP post_mean,
                         SEXP post_cov, SEXP pen_diag) {
  mb_kernel k0 = parse(kern_0, 0), ki = parse(kern_i, 1);
  double ll = 0;
  MB_CHECK(mb_logL_monitoring(ctx(), &k0, REAL(hp_0), &ki, REAL(hp_i), Rf_asLogical(shared), REAL(m_0), REAL(post_mean),
                              REAL(post_cov), Rf_asReal(pen_diag), &ll));
  return Rf_ScalarReal(ll);
}
/* logL_GP_mod + gr_GP_mod of ONE problem (x: d x n); returns c(value, gradient...) */
SEXP mbR_logL_GP_mod(SEXP kern, SEXP has_noise, SEXP hp, SEXP x, SEXP d_, SEXP y, SEXP mean, SEXP post_cov, SEXP pen_diag) {
  mb_kernel k = parse(kern, Rf_asLogical(has_noise));
  int n = (int)XLENGTH(y);
  SEXP out = PROTECT(Rf_allocVector(REALSXP, 1 + k.n_hp));
  MB_CHECK(mb_logL_GP_mod(ctx(), &k, REAL(hp), REAL(x), n, Rf_asInteger(d_), REAL(y), REAL(mean), REAL(post_cov),
                          Rf_asReal(pen_diag), REAL(out), REAL(out) + 1));
  UNPROTECT(1);
  return out;
}
/* every resident task: list(f = M, g = P x M, retries = M) */
SEXP mbR_logL_GP_mod_tasks(SEXP kern_i, SEXP hp_i, SEXP shared, SEXP post_mean, SEXP post_cov, SEXP pen_diag, SEXP n_tasks) {
  mb_kernel ki = parse(kern_i, 1);
  int M = Rf_asInteger(n_tasks);
  SEXP f = PROTECT(Rf_allocVector(REALSXP, M));
  SEXP g = PROTECT(Rf_allocMatrix(REALSXP, ki.n_hp, M));
  SEXP r = PROTECT(Rf_allocVector(INTSXP, M));
  MB_CHECK(mb_logL_GP_mod_tasks(ctx(), &ki, REAL(hp_i), Rf_asLogical(shared), REAL(post_mean), REAL(post_cov),
                                Rf_asReal(pen_diag), REAL(f), REAL(g), INTEGER(r)));
  const char* nm[] = {"f", "g", "retries"};
  SEXP v[] = {f, g, r};
  SEXP out = named_list(3, nm, v);
  UNPROTECT(3);
  return out;
}
/* the whole EM loop of train_magma (R/training.R:896-1065) on device-resident state */
SEXP mbR_train_magma(SEXP kern_0, SEXP hp_0, SEXP kern_i, SEXP hp_i, SEXP shared_hp, SEXP m_0, SEXP pen_diag, SEXP n_iter_max,
                     SEXP cv_threshold) {
  mb_kernel k0 = parse(kern_0, 0), ki = parse(kern_i, 1);
  int U = (int)XLENGTH(m_0), nmax = Rf_asInteger(n_iter_max);
  SEXP h0 = PROTECT(copy_real(hp_0));
  SEXP hi = PROTECT(copy_real(hp_i));
  SEXP seq = PROTECT(Rf_allocVector(REALSXP, nmax));
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, U));
  SEXP cov = PROTECT(Rf_allocMatrix(REALSXP, U, U));
  int32_t n_iter = 0, conv = 0;
  MB_CHECK(mb_train_magma(ctx(), &k0, REAL(h0), &ki, REAL(hi), Rf_asLogical(shared_hp), REAL(m_0), Rf_asReal(pen_diag), nmax,
                          Rf_asReal(cv_threshold), REAL(seq), &n_iter, &conv, REAL(mean), REAL(cov)));
  SEXP ni = PROTECT(Rf_ScalarInteger(n_iter));
  SEXP cv = PROTECT(Rf_ScalarLogical(conv));
  const char* nm[] = {"hp_0", "hp_i", "seq_loglikelihood", "n_iter", "converged", "mean", "cov"};
  SEXP v[] = {h0, hi, seq, ni, cv, mean, cov};
  SEXP out = named_list(7, nm, v);
  UNPROTECT(7);
  return out;
}

/* ---- hyper-posterior and prediction (R/prediction.R) ------------------------------------------------------ */
SEXP mbR_hyperposterior(SEXP kern_0, SEXP hp_0, SEXP kern_i, SEXP hp_i, SEXP shared, SEXP grid_X, SEXP grid_index, SEXP m_0,
                        SEXP pen_diag) {
  mb_kernel k0 = parse(kern_0, 0), ki = parse(kern_i, 1);
  int U = (int)XLENGTH(m_0);
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, U));
  SEXP cov = PROTECT(Rf_allocMatrix(REALSXP, U, U));
  MB_CHECK(mb_hyperposterior(ctx(), &k0, REAL(hp_0), &ki, REAL(hp_i), Rf_asLogical(shared), U, REAL(grid_X),
                             INTEGER(grid_index), REAL(m_0), Rf_asReal(pen_diag), REAL(mean), REAL(cov)));
  const char* nm[] = {"mean", "cov"};
  SEXP v[] = {mean, cov};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}
/* pred_magma's algebra (prediction.R:1146-1197); x_obs: d x n_obs, x_pred: d x n_pred; full_cov: return the G x G matrix */
SEXP mbR_pred_magma(SEXP kern, SEXP hp, SEXP x_obs, SEXP y_obs, SEXP x_pred, SEXP d_, SEXP mean_obs, SEXP mean_pred,
                    SEXP cov_obs, SEXP cov_pred, SEXP cov_crossed, SEXP pen_diag, SEXP full_cov) {
  mb_kernel k = parse(kern, 1);
  int no = (int)XLENGTH(y_obs), np = (int)XLENGTH(mean_pred), full = Rf_asLogical(full_cov);
  SEXP mean = PROTECT(Rf_allocVector(REALSXP, np));
  SEXP var = PROTECT(Rf_allocVector(REALSXP, np));
  SEXP cov = PROTECT(full ? Rf_allocMatrix(REALSXP, np, np) : Rf_allocVector(REALSXP, 0));
  MB_CHECK(mb_pred_magma(ctx(), &k, REAL(hp), REAL(x_obs), no, REAL(y_obs), REAL(x_pred), np, Rf_asInteger(d_),
                         REAL(mean_obs), REAL(mean_pred), REAL(cov_obs), REAL(cov_pred), REAL(cov_crossed),
                         Rf_asReal(pen_diag), REAL(mean), REAL(var), full ? REAL(cov) : NULL));
  const char* nm[] = {"Mean", "Var", "cov"};
  SEXP v[] = {mean, var, cov};
  SEXP out = named_list(3, nm, v);
  UNPROTECT(3);
  return out;
}
SEXP mbR_hyperposterior_clust(SEXP kern_k, SEXP hp_k, SEXP kern_i, SEXP hp_i, SEXP shared, SEXP grid_X, SEXP grid_index,
                              SEXP m_k, SEXP tau, SEXP pen_diag) {
  mb_kernel kk = parse(kern_k, 0), ki = parse(kern_i, 1);
  int U = Rf_nrows(m_k), K = Rf_ncols(m_k);
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, U, K));
  SEXP cov = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)U * U * K));
  MB_CHECK(mb_hyperposterior_clust(ctx(), K, &kk, REAL(hp_k), &ki, REAL(hp_i), Rf_asLogical(shared), U, REAL(grid_X),
                                   INTEGER(grid_index), REAL(m_k), REAL(tau), Rf_asReal(pen_diag), REAL(mean), REAL(cov)));
  const char* nm[] = {"mean", "cov"};
  SEXP v[] = {mean, cov};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}
/* pred_magmaclust's algebra (prediction.R:2298-2413): per-cluster conditionals and the mixture mean / variance */
SEXP mbR_pred_magmaclust(SEXP kern, SEXP hp, SEXP x_obs, SEXP y_obs, SEXP x_pred, SEXP d_, SEXP mean_obs, SEXP mean_pred,
                         SEXP cov_obs, SEXP cov_pred, SEXP cov_crossed, SEXP proba, SEXP pen_diag) {
  mb_kernel k = parse(kern, 1);
  int K = (int)XLENGTH(proba), no = (int)XLENGTH(y_obs), np = (int)(XLENGTH(mean_pred) / K);
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, np, K));
  SEXP var = PROTECT(Rf_allocMatrix(REALSXP, np, K));
  SEXP mm = PROTECT(Rf_allocVector(REALSXP, np));
  SEXP mv = PROTECT(Rf_allocVector(REALSXP, np));
  MB_CHECK(mb_pred_magmaclust(ctx(), K, &k, REAL(hp), REAL(x_obs), no, REAL(y_obs), REAL(x_pred), np, Rf_asInteger(d_),
                              REAL(mean_obs), REAL(mean_pred), REAL(cov_obs), REAL(cov_pred), REAL(cov_crossed), REAL(proba),
                              Rf_asReal(pen_diag), REAL(mean), REAL(var), NULL, REAL(mm), REAL(mv)));
  const char* nm[] = {"Mean_k", "Var_k", "Mean", "Var"};
  SEXP v[] = {mean, var, mm, mv};
  SEXP out = named_list(4, nm, v);
  UNPROTECT(4);
  return out;
}
/* tuning / parity switches of the library (include/magma_b200.h: mb_set_option), e.g. ("logdet_chol", 0) for the
 * reference's LU log-determinant (R/likelihoods.R:37-41) or ("pair_tables", 0) for the element-by-element SE kernel */
SEXP mbR_set_option(SEXP name, SEXP value) {
  MB_CHECK(mb_set_option(ctx(), CHAR(STRING_ELT(name, 0)), (int64_t)Rf_asReal(value)));
  return R_NilValue;
}
/* keep the next hyper-posterior on the device / install one (new-task path) */
SEXP mbR_hyperpost_keep(SEXP keep) {
  MB_CHECK(mb_hyperpost_keep(ctx(), Rf_asLogical(keep)));
  return R_NilValue;
}
SEXP mbR_hyperpost_set(SEXP post_mean, SEXP post_cov) {
  int U = Rf_nrows(post_mean), K = Rf_ncols(post_mean);
  MB_CHECK(mb_hyperpost_set(ctx(), K, U, REAL(post_mean), Rf_isNull(post_cov) ? NULL : REAL(post_cov)));
  return R_NilValue;
}
/* logL_GP + gr_GP of every resident (new) task; train_gp; train_gp_clust; pred_magma(clust) of all new tasks */
SEXP mbR_logL_GP_tasks(SEXP kern, SEXP hp, SEXP shared, SEXP cluster, SEXP pen_diag, SEXP n_tasks) {
  mb_kernel k = parse(kern, 1);
  int M = Rf_asInteger(n_tasks);
  SEXP f = PROTECT(Rf_allocVector(REALSXP, M));
  SEXP g = PROTECT(Rf_allocMatrix(REALSXP, k.n_hp, M));
  MB_CHECK(mb_logL_GP_tasks(ctx(), &k, REAL(hp), Rf_asLogical(shared), Rf_asInteger(cluster), Rf_asReal(pen_diag), REAL(f),
                            REAL(g), NULL));
  const char* nm[] = {"f", "g"};
  SEXP v[] = {f, g};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}
SEXP mbR_train_gp_tasks(SEXP kern, SEXP hp, SEXP cluster, SEXP pen_diag) {
  mb_kernel k = parse(kern, 1);
  SEXP h = PROTECT(copy_real(hp));
  int64_t st[4] = {0, 0, 0, 0};
  MB_CHECK(mb_train_gp_tasks(ctx(), &k, REAL(h), Rf_asInteger(cluster), Rf_asReal(pen_diag), st));
  UNPROTECT(1);
  return h;
}
SEXP mbR_train_shared_gp(SEXP kern, SEXP hp, SEXP mean, SEXP pen_diag, SEXP maxit) {
  mb_kernel k = parse(kern, 1);
  SEXP h = PROTECT(copy_real(hp));
  int64_t st[4] = {0, 0, 0, 0};
  MB_CHECK(mb_train_shared_gp(ctx(), &k, REAL(h), Rf_asReal(mean), Rf_asReal(pen_diag), Rf_asInteger(maxit), st));
  UNPROTECT(1);
  return h;
}
SEXP mbR_train_gp_clust_tasks(SEXP kern, SEXP hp, SEXP prop_mixture, SEXP pen_diag, SEXP n_iter_max, SEXP cv_threshold,
                              SEXP n_tasks) {
  mb_kernel k = parse(kern, 1);
  int M = Rf_asInteger(n_tasks), K = (int)XLENGTH(prop_mixture);
  SEXP h = PROTECT(copy_real(hp));
  SEXP mix = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  MB_CHECK(mb_train_gp_clust_tasks(ctx(), &k, REAL(h), REAL(prop_mixture), Rf_asReal(pen_diag), Rf_asInteger(n_iter_max),
                                   Rf_asReal(cv_threshold), REAL(mix), NULL, NULL, NULL));
  const char* nm[] = {"hp", "mixture"};
  SEXP v[] = {h, mix};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}
SEXP mbR_pred_magma_tasks(SEXP kern, SEXP hp, SEXP tau, SEXP x_pred, SEXP pred_index, SEXP pen_diag, SEXP n_tasks) {
  mb_kernel k = parse(kern, 1);
  int M = Rf_asInteger(n_tasks), G = (int)XLENGTH(pred_index);
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, G, M));
  SEXP var = PROTECT(Rf_allocMatrix(REALSXP, G, M));
  MB_CHECK(mb_pred_magma_tasks(ctx(), &k, REAL(hp), Rf_isNull(tau) ? NULL : REAL(tau), G, REAL(x_pred), INTEGER(pred_index),
                               Rf_asReal(pen_diag), REAL(mean), REAL(var), NULL, NULL, NULL));
  const char* nm[] = {"Mean", "Var"};
  SEXP v[] = {mean, var};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}

/* ---- MagmaClust: ve_step / update_mixture / vm_step / ELBOs / train_magmaclust (R/vem-magmaclust.R, R/elbos.R) ---- */
SEXP mbR_ve_step(SEXP kern_k, SEXP hp_k, SEXP kern_i, SEXP hp_i, SEXP shared, SEXP m_k, SEXP tau, SEXP prop_mixture,
                 SEXP update, SEXP pen_diag) {
  mb_kernel kk = parse(kern_k, 0), ki = parse(kern_i, 1);
  int U = Rf_nrows(m_k), K = Rf_ncols(m_k);
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, U, K));
  SEXP cov = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)U * U * K));
  SEXP mix = PROTECT(Rf_allocMatrix(REALSXP, K, (int)(XLENGTH(tau) / K)));
  MB_CHECK(mb_ve_step(ctx(), K, &kk, REAL(hp_k), &ki, REAL(hp_i), Rf_asLogical(shared), REAL(m_k), REAL(tau),
                      Rf_isNull(prop_mixture) ? NULL : REAL(prop_mixture), Rf_asLogical(update), Rf_asReal(pen_diag),
                      REAL(mean), REAL(cov), REAL(mix)));
  const char* nm[] = {"mean", "cov", "mixture"};
  SEXP v[] = {mean, cov, mix};
  SEXP out = named_list(3, nm, v);
  UNPROTECT(3);
  return out;
}
SEXP mbR_update_mixture(SEXP kern_i, SEXP hp_i, SEXP shared, SEXP post_mean, SEXP post_cov, SEXP prop_mixture, SEXP pen_diag,
                        SEXP n_tasks) {
  mb_kernel ki = parse(kern_i, 1);
  int K = (int)XLENGTH(prop_mixture), M = Rf_asInteger(n_tasks);
  SEXP mix = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  MB_CHECK(mb_update_mixture(ctx(), K, &ki, REAL(hp_i), Rf_asLogical(shared), REAL(post_mean), REAL(post_cov),
                             REAL(prop_mixture), Rf_asReal(pen_diag), REAL(mix)));
  UNPROTECT(1);
  return mix;
}
SEXP mbR_elbo_clust_tasks(SEXP kern_i, SEXP hp_i, SEXP shared, SEXP post_mean, SEXP post_cov, SEXP tau, SEXP pen_diag,
                          SEXP n_tasks) {
  mb_kernel ki = parse(kern_i, 1);
  int K = Rf_ncols(post_mean), M = Rf_asInteger(n_tasks);
  SEXP f = PROTECT(Rf_allocVector(REALSXP, M));
  SEXP g = PROTECT(Rf_allocMatrix(REALSXP, ki.n_hp, M));
  MB_CHECK(mb_elbo_clust_tasks(ctx(), K, &ki, REAL(hp_i), Rf_asLogical(shared), REAL(post_mean), REAL(post_cov), REAL(tau),
                               Rf_asReal(pen_diag), REAL(f), REAL(g)));
  const char* nm[] = {"f", "g"};
  SEXP v[] = {f, g};
  SEXP out = named_list(2, nm, v);
  UNPROTECT(2);
  return out;
}
SEXP mbR_vm_step(SEXP kern_k, SEXP hp_k, SEXP kern_i, SEXP hp_i, SEXP shared_hp_k, SEXP shared_hp_i, SEXP post_mean,
                 SEXP post_cov, SEXP tau, SEXP pen_diag) {
  mb_kernel kk = parse(kern_k, 0), ki = parse(kern_i, 1);
  int U = Rf_nrows(post_mean), K = Rf_ncols(post_mean);
  SEXP hk = PROTECT(copy_real(hp_k));
  SEXP hi = PROTECT(copy_real(hp_i));
  SEXP mk = PROTECT(Rf_allocMatrix(REALSXP, U, K));
  SEXP prop = PROTECT(Rf_allocVector(REALSXP, K));
  int64_t st[4] = {0, 0, 0, 0};
  MB_CHECK(mb_vm_step(ctx(), K, &kk, REAL(hk), &ki, REAL(hi), Rf_asLogical(shared_hp_k), Rf_asLogical(shared_hp_i), REAL(mk),
                      REAL(post_mean), REAL(post_cov), REAL(tau), Rf_asReal(pen_diag), REAL(prop), st));
  const char* nm[] = {"m_k", "hp_k", "hp_i", "prop_mixture"};
  SEXP v[] = {mk, hk, hi, prop};
  SEXP out = named_list(4, nm, v);
  UNPROTECT(4);
  return out;
}
SEXP mbR_elbo_monitoring(SEXP kern_k, SEXP hp_k, SEXP kern_i, SEXP hp_i, SEXP shared, SEXP m_k, SEXP post_mean, SEXP post_cov,
                         SEXP tau, SEXP prop_mixture, SEXP pen_diag) {
  mb_kernel kk = parse(kern_k, 0), ki = parse(kern_i, 1);
  double elbo = 0;
  MB_CHECK(mb_elbo_monitoring(ctx(), Rf_ncols(post_mean), &kk, REAL(hp_k), &ki, REAL(hp_i), Rf_asLogical(shared), REAL(m_k),
                              REAL(post_mean), REAL(post_cov), REAL(tau), REAL(prop_mixture), Rf_asReal(pen_diag), &elbo));
  return Rf_ScalarReal(elbo);
}
/* the whole VEM loop of train_magmaclust (R/training.R:1586-1758) on device-resident state */
SEXP mbR_train_magmaclust(SEXP kern_k, SEXP hp_k, SEXP kern_i, SEXP hp_i, SEXP shared_hp_k, SEXP shared_hp_i, SEXP m_k,
                          SEXP ini_mixture, SEXP pen_diag, SEXP n_iter_max, SEXP cv_threshold, SEXP fast_approx) {
  mb_kernel kk = parse(kern_k, 0), ki = parse(kern_i, 1);
  int U = Rf_nrows(m_k), K = Rf_ncols(m_k), nmax = Rf_asInteger(n_iter_max);
  int M = (int)(XLENGTH(ini_mixture) / K);
  SEXP hk = PROTECT(copy_real(hp_k));
  SEXP hi = PROTECT(copy_real(hp_i));
  SEXP mk = PROTECT(copy_real(m_k));
  SEXP seq = PROTECT(Rf_allocVector(REALSXP, nmax));
  SEXP mix = PROTECT(Rf_allocMatrix(REALSXP, K, M));
  SEXP prop = PROTECT(Rf_allocVector(REALSXP, K));
  SEXP mean = PROTECT(Rf_allocMatrix(REALSXP, U, K));
  SEXP cov = PROTECT(Rf_allocVector(REALSXP, (R_xlen_t)U * U * K));
  int32_t n_iter = 0, conv = 0;
  MB_CHECK(mb_train_magmaclust(ctx(), K, &kk, REAL(hk), &ki, REAL(hi), Rf_asLogical(shared_hp_k), Rf_asLogical(shared_hp_i),
                               REAL(mk), REAL(ini_mixture), Rf_asReal(pen_diag), nmax, Rf_asReal(cv_threshold),
                               Rf_asLogical(fast_approx), REAL(seq), &n_iter, &conv, REAL(mix), REAL(prop), REAL(mean),
                               REAL(cov)));
  SEXP ni = PROTECT(Rf_ScalarInteger(n_iter));
  SEXP cv = PROTECT(Rf_ScalarLogical(conv));
  const char* nm[] = {"hp_k", "hp_i", "m_k", "seq_elbo", "n_iter", "converged", "mixture", "prop_mixture", "mean", "cov"};
  SEXP v[] = {hk, hi, mk, seq, ni, cv, mix, prop, mean, cov};
  SEXP out = named_list(10, nm, v);
  UNPROTECT(10);
  return out;
}

/* ---- registration ------------------------------------------------------------------------------------------- */
#define ENTRY(name, n) {#name, (DL_FUNC)&name, n}
static const R_CallMethodDef CallEntries[] = {
    ENTRY(_MagmaClustR_cpp_dist, 2),        /* src/init.c:17-24: same names, same arities */
    ENTRY(_MagmaClustR_cpp_noise, 3),
    ENTRY(_MagmaClustR_cpp_perio, 3),
    ENTRY(_MagmaClustR_cpp_perio_deriv, 3),
    ENTRY(_MagmaClustR_cpp_prod, 2),
    ENTRY(mbR_kern_to_cov, 6),
    ENTRY(mbR_list_kern_to_mat, 10),
    ENTRY(mbR_chol_inv_jitter, 2),
    ENTRY(mbR_upload_dataset, 6),
    ENTRY(mbR_e_step, 7),
    ENTRY(mbR_m_step, 9),
    ENTRY(mbR_logL_monitoring, 9),
    ENTRY(mbR_logL_GP_mod, 9),
    ENTRY(mbR_logL_GP_mod_tasks, 7),
    ENTRY(mbR_train_magma, 9),
    ENTRY(mbR_hyperposterior, 9),
    ENTRY(mbR_pred_magma, 13),
    ENTRY(mbR_hyperposterior_clust, 10),
    ENTRY(mbR_pred_magmaclust, 13),
    ENTRY(mbR_set_option, 2),
    ENTRY(mbR_hyperpost_keep, 1),
    ENTRY(mbR_hyperpost_set, 2),
    ENTRY(mbR_logL_GP_tasks, 6),
    ENTRY(mbR_train_gp_tasks, 4),
    ENTRY(mbR_train_shared_gp, 5),
    ENTRY(mbR_train_gp_clust_tasks, 7),
    ENTRY(mbR_pred_magma_tasks, 7),
    ENTRY(mbR_ve_step, 10),
    ENTRY(mbR_update_mixture, 8),
    ENTRY(mbR_elbo_clust_tasks, 8),
    ENTRY(mbR_vm_step, 10),
    ENTRY(mbR_elbo_monitoring, 11),
    ENTRY(mbR_train_magmaclust, 12),
    {NULL, NULL, 0}};

void R_init_MagmaClustR(DllInfo* dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, FALSE);
}
void R_unload_MagmaClustR(DllInfo* dll) {
  (void)dll;
  if (g_ctx) { mb_destroy(g_ctx); g_ctx = NULL; }
}
