"""Magma (single mean process) EM numerics of the reference, restated (oracle).

TEST INFRASTRUCTURE.  Follows /root/reference/R/em-magma.R (e_step :25-82,
m_step :119-235), R/likelihoods.R (dmnorm :19-50, logL_GP :73-86, logL_GP_mod
:155-174, logL_GP_mod_shared_hp :201-225, logL_monitoring :253-312),
R/gradients-likelihoods.R (gr_GP :22-49, gr_GP_mod :71-97, gr_GP_mod_shared_hp
:121-165), the loop body of train_magma (R/training.R:690-720, 896-1065),
hyperposterior (R/prediction.R:670-726) and pred_magma (R/prediction.R:1035-1059,
1146-1197).  Parity for these functions is unpinned by the reference's tests
(SURVEY.md section 8c); they are validated by finite differences and identities.
"""
import math

import numpy as np

from . import kernels as K
from .lbfgsb import optim_lbfgsb
from .rformat import reference_key, c_locale_order, r_signif


class Dataset:
    """The tibble `data` after train_magma's preparation (training.R:690-701):
    inputs rounded by signif(6), Reference key, tasks ordered by ID string and rows
    by Reference string (C locale); union grid = unique keys in sorted order
    (em-magma.R:27-30)."""

    def __init__(self, ids, X, y):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        y = np.asarray(y, dtype=np.float64)
        ids = [str(i) for i in ids]
        Xr = np.vectorize(r_signif)(X) if X.size else X
        keys = [reference_key(row) for row in Xr]
        by_id = {}
        for r, i in enumerate(ids):
            by_id.setdefault(i, []).append(r)
        self.ids = sorted(by_id, key=lambda s: s.encode("utf-8"))
        self.tasks = {}
        for i in self.ids:
            rows = by_id[i]
            order = sorted(rows, key=lambda r: keys[r].encode("utf-8"))
            self.tasks[i] = ([keys[r] for r in order], Xr[order].copy(), y[order].copy())
        allk = {}
        for i in self.ids:
            ks, Xi, _ = self.tasks[i]
            for k, row in zip(ks, Xi):
                allk.setdefault(k, row)
        self.grid_keys = sorted(allk, key=lambda s: s.encode("utf-8"))
        self.grid_X = np.array([allk[k] for k in self.grid_keys]).reshape(len(self.grid_keys), X.shape[1])
        pos = {k: j for j, k in enumerate(self.grid_keys)}
        self.index = {i: np.array([pos[k] for k in self.tasks[i][0]], dtype=np.int64)
                      for i in self.ids}

    def extend_grid(self, grid_X):
        """Union of the observed keys and a prediction grid (prediction.R:551-567)."""
        g = np.asarray(grid_X, dtype=np.float64)
        if g.ndim == 1:
            g = g[:, None]
        g = np.vectorize(r_signif)(g)
        allk = {k: row for k, row in zip(self.grid_keys, self.grid_X)}
        for row in g:
            allk.setdefault(reference_key(row), row)
        keys = sorted(allk, key=lambda s: s.encode("utf-8"))
        Xg = np.array([allk[k] for k in keys]).reshape(len(keys), self.grid_X.shape[1])
        pos = {k: j for j, k in enumerate(keys)}
        index = {i: np.array([pos[k] for k in self.tasks[i][0]], dtype=np.int64) for i in self.ids}
        return keys, Xg, index


# ---- likelihoods ------------------------------------------------------------
def neg_dmnorm_log(x, mu, inv):
    """-dmnorm(x, mu, inv, log=TRUE) (likelihoods.R:19-50): log-det through an LU
    of the explicit inverse."""
    n = len(mu)
    z = np.asarray(x, dtype=np.float64) - np.asarray(mu, dtype=np.float64)
    logdetS = -K.logdet_lu(inv)
    ssq = float(z @ inv @ z)
    return (n * math.log(2 * math.pi) + logdetS + ssq) / 2


def _mean_vec(mean, n):
    mean = np.atleast_1d(np.asarray(mean, dtype=np.float64))
    return np.repeat(mean, n) if mean.size == 1 else mean


def logL_GP_mod(hp, X, y, mean, kern, post_cov, pen_diag):
    """likelihoods.R:155-174."""
    mean = _mean_vec(mean, len(y))
    inv = K.kern_to_inv(X, kern, hp, pen_diag)
    return neg_dmnorm_log(y, mean, inv) + 0.5 * float(np.sum(inv * post_cov))


def gr_GP_mod(hp, X, y, mean, kern, post_cov, pen_diag):
    """gradients-likelihoods.R:71-97; one entry per name of hp, in hp's order."""
    mean = _mean_vec(mean, len(y))
    inv = K.kern_to_inv(X, kern, hp, pen_diag)
    prod_inv = inv @ (np.asarray(y) - mean)
    common = np.outer(prod_inv, prod_inv) + inv @ (post_cov @ inv - np.eye(len(y)))
    return np.array([float(np.sum(np.diag(-0.5 * (common @ K.kern_to_cov(X, kern, hp, d)))))
                     for d in hp])


def logL_GP(hp, X, y, mean, kern, post_cov, pen_diag):
    """likelihoods.R:73-86 (marginal form: K + post_cov inverted)."""
    mean = _mean_vec(mean, len(y))
    cov = K.kern_to_cov(X, kern, hp) + post_cov
    inv = K.chol_inv_jitter(cov, pen_diag)
    return neg_dmnorm_log(y, mean, inv)


def gr_GP(hp, X, y, mean, kern, post_cov, pen_diag):
    """gradients-likelihoods.R:22-49."""
    mean = _mean_vec(mean, len(y))
    cov = K.kern_to_cov(X, kern, hp) + post_cov
    inv = K.chol_inv_jitter(cov, pen_diag)
    prod_inv = inv @ (np.asarray(y) - mean)
    common = np.outer(prod_inv, prod_inv) - inv
    return np.array([float(np.sum(np.diag(-0.5 * (common @ K.kern_to_cov(X, kern, hp, d)))))
                     for d in hp])


def _task_views(db, post_mean, post_cov, i, index=None):
    keys, Xi, yi = db.tasks[i]
    idx = (index or db.index)[i]
    return Xi, yi, post_mean[idx], post_cov[np.ix_(idx, idx)]


def logL_GP_mod_shared_hp(hp, db, post_mean, kern, post_cov, pen_diag):
    """likelihoods.R:201-225."""
    tot = 0.0
    for i in db.ids:
        Xi, yi, mi, Si = _task_views(db, post_mean, post_cov, i)
        tot += logL_GP_mod(hp, Xi, yi, mi, kern, Si, pen_diag)
    return tot


def gr_GP_mod_shared_hp(hp, db, post_mean, kern, post_cov, pen_diag):
    """gradients-likelihoods.R:121-165."""
    tot = np.zeros(len(hp))
    for i in db.ids:
        Xi, yi, mi, Si = _task_views(db, post_mean, post_cov, i)
        tot += gr_GP_mod(hp, Xi, yi, mi, kern, Si, pen_diag)
    return tot


# ---- E step -----------------------------------------------------------------
def e_step(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag, grid=None, return_info=False):
    """em-magma.R:25-82 (grid=None) / prediction.R:670-707 (grid = extended grid).

    hp_i: dict id -> hp dict.  Returns (post_mean[U], post_cov[U,U])."""
    if grid is None:
        grid_X, index = db.grid_X, db.index
    else:
        _, grid_X, index = grid
    U = grid_X.shape[0]
    m_0 = _mean_vec(m_0, U)
    inv_0, r0 = K.chol_inv_jitter(K.kern_to_cov(grid_X, kern_0, hp_0), pen_diag, True)
    post_inv = inv_0.copy()
    weighted_0 = inv_0 @ m_0
    retries = {}
    list_inv = {}
    for i in db.ids:
        _, Xi, yi = db.tasks[i]
        list_inv[i], retries[i] = K.chol_inv_jitter(K.kern_to_cov(Xi, kern_i, hp_i[i]), pen_diag, True)
    for i in db.ids:                                     # em-magma.R:39-46
        idx = index[i]
        post_inv[np.ix_(idx, idx)] += list_inv[i]
    post_cov, rp = K.chol_inv_jitter(post_inv, pen_diag, True)       # :48-57
    for i in db.ids:                                     # :61-70
        weighted_0[index[i]] += list_inv[i] @ db.tasks[i][2]
    post_mean = post_cov @ weighted_0
    if return_info:
        return post_mean, post_cov, {"inv_0": r0, "post": rp, "tasks": retries}
    return post_mean, post_cov


# ---- M step -----------------------------------------------------------------
def _optim(hp, fn, gr):
    names = list(hp)

    def fg(x):
        h = dict(zip(names, x))
        return fn(h), gr(h)

    res = optim_lbfgsb([hp[k] for k in names], fg, factr=1e13, maxit=25)
    return dict(zip(names, res["par"])), res


def m_step(db, m_0, kern_0, kern_i, old_hp_0, old_hp_i, post_mean, post_cov, shared_hp,
           pen_diag, stats=None):
    """em-magma.R:119-235."""
    U = db.grid_X.shape[0]
    m_0 = _mean_vec(m_0, U)
    new_hp_0, r0 = _optim(
        old_hp_0,
        lambda h: logL_GP_mod(h, db.grid_X, post_mean, m_0, kern_0, post_cov, pen_diag),
        lambda h: gr_GP_mod(h, db.grid_X, post_mean, m_0, kern_0, post_cov, pen_diag))
    if stats is not None:
        stats["hp_0"] = r0
        stats["tasks"] = {}
    if shared_hp:
        one, r = _optim(
            old_hp_i[db.ids[0]],
            lambda h: logL_GP_mod_shared_hp(h, db, post_mean, kern_i, post_cov, pen_diag),
            lambda h: gr_GP_mod_shared_hp(h, db, post_mean, kern_i, post_cov, pen_diag))
        new_hp_i = {i: dict(one) for i in db.ids}
        if stats is not None:
            stats["tasks"]["shared"] = r
    else:
        new_hp_i = {}
        for i in db.ids:
            Xi, yi, mi, Si = _task_views(db, post_mean, post_cov, i)
            new_hp_i[i], r = _optim(
                old_hp_i[i],
                lambda h: logL_GP_mod(h, Xi, yi, mi, kern_i, Si, pen_diag),
                lambda h: gr_GP_mod(h, Xi, yi, mi, kern_i, Si, pen_diag))
            if stats is not None:
                stats["tasks"][i] = r
    return new_hp_0, new_hp_i


def logL_monitoring(hp_0, hp_i, db, m_0, kern_0, kern_i, post_mean, post_cov, pen_diag):
    """likelihoods.R:253-312 (chol(post_cov) without jitter, :303-307)."""
    from scipy.linalg import lapack
    U = db.grid_X.shape[0]
    m_0 = _mean_vec(m_0, U)
    ll_0 = logL_GP_mod(hp_0, db.grid_X, post_mean, m_0, kern_0, post_cov, pen_diag)
    sum_ll_i = 0.0
    for i in db.ids:
        Xi, yi, mi, Si = _task_views(db, post_mean, post_cov, i)
        sum_ll_i += logL_GP_mod(hp_i[i], Xi, yi, mi, kern_i, Si, pen_diag)
    c, info = lapack.dpotrf(np.asfortranarray(post_cov), lower=0, clean=1)
    if info != 0:
        raise FloatingPointError("chol(post_cov) failed in logL_monitoring")
    det = float(np.sum(np.log(np.diag(c))))
    return -ll_0 - sum_ll_i + det


def train_magma(db, prior_mean, ini_hp_0, ini_hp_i, kern_0, kern_i, shared_hp=True,
                pen_diag=1e-10, n_iter_max=25, cv_threshold=1e-3, fast_approx=False,
                trace=None):
    """EM loop of train_magma, training.R:896-1065 (HPs are given explicitly)."""
    U = db.grid_X.shape[0]
    m_0 = np.zeros(U) if prior_mean is None else _mean_vec(prior_mean, U)
    hp_0 = dict(ini_hp_0)
    hp_i = {i: dict(ini_hp_i[i]) for i in db.ids}
    cv = False
    ll = -math.inf
    seq = []
    post_mean = post_cov = None
    for it in range(1, n_iter_max + 1):
        post_mean, post_cov = e_step(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag)
        if fast_approx:
            seq = [logL_monitoring(hp_0, hp_i, db, m_0, kern_0, kern_i, post_mean, post_cov, pen_diag)]
            break
        stats = {} if trace is not None else None
        new_hp_0, new_hp_i = m_step(db, m_0, kern_0, kern_i, hp_0, hp_i, post_mean, post_cov,
                                    shared_hp, pen_diag, stats)
        if any(math.isnan(v) for v in new_hp_0.values()) or \
           any(math.isnan(v) for h in new_hp_i.values() for v in h.values()):
            break
        new_ll = logL_monitoring(new_hp_0, new_hp_i, db, m_0, kern_0, kern_i, post_mean,
                                 post_cov, pen_diag)
        diff = new_ll - ll
        if math.isnan(diff):
            diff = -math.inf
        hp_0, hp_i, ll = new_hp_0, new_hp_i, new_ll
        seq.append(ll)
        eps = diff / abs(ll)
        if math.isnan(eps):
            eps = 1.0
        if trace is not None:
            trace.append({"iter": it, "logL": ll, "eps": eps, "stats": stats})
        if abs(eps) < cv_threshold:
            cv = True
            break
    return {"hp_0": hp_0, "hp_i": hp_i, "post_mean": post_mean, "post_cov": post_cov,
            "seq_loglikelihood": seq, "converged": cv, "m_0": m_0}


# ---- prediction -------------------------------------------------------------
def train_gp(X, y, mean, kern, ini_hp, post_cov, pen_diag=1e-10, return_info=False):
    """training.R:78-287 (the optim call :263-274 and the NA fallback :277-284): HPs of one (new) task by
    L-BFGS-B on logL_GP / gr_GP against the hyper-posterior (mean, post_cov) at the task's inputs."""
    hp, res = _optim(ini_hp, lambda h: logL_GP(h, X, y, mean, kern, post_cov, pen_diag),
                     lambda h: gr_GP(h, X, y, mean, kern, post_cov, pen_diag))
    if any(not np.isfinite(v) for v in hp.values()):
        hp = dict(ini_hp)
    return (hp, res) if return_info else hp


def sum_logL_GP(hp, db, mean, kern, pen_diag):
    """likelihoods.R:113-129: sapply over the tasks %>% sum() (long double accumulation)."""
    vals = [logL_GP(hp, db.tasks[i][1], db.tasks[i][2], mean, kern, 0.0, pen_diag) for i in db.ids]
    return float(np.sum(np.array(vals, dtype=np.longdouble)))


def train_shared_gp(db, mean, kern, ini_hp, pen_diag=1e-10, maxit=100, return_info=False):
    """training.R:352-446: optim(par, fn = sum_logL_GP, method = "L-BFGS-B", factr = 1e13, maxit = 100) WITHOUT gr --
    optim's own central differences (R src/library/stats/src/optim.c, fmingr: ndeps = 1e-3)."""
    from .lbfgsb import optim_lbfgsb
    names = list(ini_hp)
    ndeps = 1e-3
    n_fn = [0]

    def fn(x):
        n_fn[0] += 1
        return sum_logL_GP(dict(zip(names, x)), db, mean, kern, pen_diag)

    def fg(x):
        f = fn(x)
        g = []
        for i in range(len(x)):
            xp = list(x)
            xp[i] = x[i] + ndeps
            v1 = fn(xp)
            xp[i] = x[i] - ndeps
            v2 = fn(xp)
            g.append((v1 - v2) / (2 * ndeps))
        return f, g

    res = optim_lbfgsb([ini_hp[k] for k in names], fg, factr=1e13, maxit=maxit)
    hp = dict(zip(names, res["par"]))
    if any(not np.isfinite(v) for v in hp.values()):
        hp = dict(ini_hp)
    res["fn_evals"] = n_fn[0]
    return (hp, res) if return_info else hp


def hyperposterior(db, hp_0, hp_i, kern_0, kern_i, prior_mean, grid_X, pen_diag=1e-10):
    """prediction.R:467-727: E step re-evaluated on data U grid."""
    if grid_X is None:
        keys, Xg, index = db.grid_keys, db.grid_X, db.index
    else:
        keys, Xg, index = db.extend_grid(grid_X)
    m_0 = 0.0 if prior_mean is None else prior_mean
    mean, cov = e_step(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag, grid=(keys, Xg, index))
    return {"keys": keys, "X": Xg, "mean": mean, "cov": cov}


def pred_magma(X_obs, y_obs, X_pred, kern, hp, hyperpost, pen_diag=1e-10):
    """prediction.R:1035-1059,1146-1197.  hyperpost must cover obs U pred keys.
    Rows of obs and pred are processed in Reference (C-locale) order."""
    X_obs = np.vectorize(r_signif)(K._mat(X_obs))
    X_pred = np.vectorize(r_signif)(K._mat(X_pred))
    ko = [reference_key(r) for r in X_obs]
    kp = [reference_key(r) for r in X_pred]
    oo = c_locale_order(ko)
    X_obs, y_obs, ko = X_obs[oo], np.asarray(y_obs, dtype=np.float64)[oo], [ko[j] for j in oo]
    # .build_grid_inputs arranges the grid by Reference (utils.R:183-295)
    op = c_locale_order(kp)
    X_pred, kp = X_pred[op], [kp[j] for j in op]
    pos = {k: j for j, k in enumerate(hyperpost["keys"])}
    io = np.array([pos[k] for k in ko])
    ip = np.array([pos[k] for k in kp])
    mean_obs, mean_pred = hyperpost["mean"][io], hyperpost["mean"][ip]
    S = hyperpost["cov"]
    cov_obs = K.kern_to_cov(X_obs, kern, hp) + S[np.ix_(io, io)]
    inv_obs = K.chol_inv_jitter(cov_obs, pen_diag)
    hp_rm = {k: v for k, v in hp.items() if k != "noise"}
    noise = math.exp(hp["noise"]) if "noise" in hp else 0.0
    cov_pred = K.kern_to_cov(X_pred, kern, hp_rm) + S[np.ix_(ip, ip)]
    cov_crossed = K.kern_to_cov(X_obs, kern, hp_rm, input_2=X_pred) + S[np.ix_(io, ip)]
    pred_mean = mean_pred + cov_crossed.T @ inv_obs @ (y_obs - mean_obs)
    pred_cov = cov_pred - cov_crossed.T @ inv_obs @ cov_crossed
    return {"keys": kp, "X": X_pred, "Mean": pred_mean, "Var": np.diag(pred_cov) + noise,
            "cov": pred_cov}
