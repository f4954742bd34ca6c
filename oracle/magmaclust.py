"""MagmaClust (K cluster-specific mean processes, variational EM), restated (oracle).

TEST INFRASTRUCTURE.  Follows /root/reference/R/vem-magmaclust.R (ve_step :32-150, vm_step
:191-303, update_mixture :330-396), R/elbos.R (elbo_clust_multi_GP :18-55,
elbo_GP_mod_shared_hp_k :78-100, elbo_clust_multi_GP_shared_hp_i :120-172,
elbo_monitoring_VEM :200-272), R/gradients-elbos.R (gr_clust_multi_GP :17-70,
gr_GP_mod_shared_hp_k :73-150, gr_clust_multi_GP_shared_hp_i :153-241) and the VEM loop of
train_magmaclust (R/training.R:1622-1758).  Parity unpinned by the reference's own tests
(SURVEY.md 8c); validated by finite differences and identities.

Clusters are lists indexed 0..K-1 (the reference names them "K1".."Kk"); the mixture is an
(n_tasks x K) array whose rows follow db.ids.
"""
import math

import numpy as np

from . import kernels as K
from . import magma as M


def r_round(x, dig=5):
    """R >= 4.0.0 round(x, dig): the closer of the two neighbouring decimals, ties to even
    (nmath/fround.c); used on the mixture at vem-magmaclust.R:392."""
    x = float(x)
    if not math.isfinite(x):
        return x
    sgn = 1.0
    if x < 0:
        sgn, x = -1.0, -x
    p10 = np.longdouble(10) ** dig
    x10 = p10 * np.longdouble(x)
    i10 = np.floor(x10)
    xd, xu = float(i10 / p10), float(np.ceil(x10) / p10)
    du, dd = xu - x, x - xd
    even = float(i10 % 2) == 0.0
    return sgn * (xd if (dd < du or (dd == du and even)) else xu)


def r_mean(v):
    """R's mean(): long double sum / n plus one correction pass."""
    v = np.asarray(v, dtype=np.longdouble)
    s = v.sum() / len(v)
    return float(s + (v - s).sum() / len(v))


def ve_task_order(db):
    """ve_step re-sorts the whole table by Reference (vem-magmaclust.R:48), so its task loops
    run in order of first appearance in that sorted table."""
    first = {i: (db.tasks[i][0][0].encode("utf-8"), pos) for pos, i in enumerate(db.ids)}
    return sorted(db.ids, key=lambda i: first[i])


def ve_step(db, m_k, kern_k, kern_i, hp_k, hp_i, old_mixture, iteration, pen_diag, prop_mixture=None):
    """vem-magmaclust.R:32-150.  m_k: list of K vectors (U); hp_k: list of K dicts;
    old_mixture: (n_tasks x K), rows in db.ids order.  Returns (mean_k, cov_k, mixture)."""
    Kc = len(m_k)
    U = db.grid_X.shape[0]
    row = {i: t for t, i in enumerate(db.ids)}
    inv_k = [K.kern_to_inv(db.grid_X, kern_k, hp_k[k], pen_diag) for k in range(Kc)]
    order = ve_task_order(db)
    inv_i = {i: K.kern_to_inv(db.tasks[i][1], kern_i, hp_i[i], pen_diag) for i in order}
    cov_k, mean_k = [], []
    for k in range(Kc):
        post_inv = inv_k[k].copy()
        for i in order:
            idx = db.index[i]
            post_inv[np.ix_(idx, idx)] += old_mixture[row[i], k] * inv_i[i]
        cov_k.append(K.chol_inv_jitter(post_inv, pen_diag))
    for k in range(Kc):
        weighted = inv_k[k] @ M._mean_vec(m_k[k], U)
        for i in order:
            weighted[db.index[i]] += old_mixture[row[i], k] * (inv_i[i] @ db.tasks[i][2])
        mean_k.append(cov_k[k] @ weighted)
    if iteration == 1:
        mixture = np.array(old_mixture, dtype=np.float64, copy=True)
    else:
        mixture = update_mixture(db, mean_k, cov_k, hp_i, kern_i, prop_mixture, pen_diag)
    return mean_k, cov_k, mixture


def update_mixture(db, mean_k, cov_k, hp_i, kern_i, prop_mixture, pen_diag):
    """vem-magmaclust.R:330-396."""
    Kc = len(mean_k)
    out = np.zeros((len(db.ids), Kc))
    for t, i in enumerate(db.ids):
        _, Xi, yi = db.tasks[i]
        idx = db.index[i]
        L = np.array([-M.logL_GP_mod(hp_i[i], Xi, yi, mean_k[k][idx], kern_i, cov_k[k][np.ix_(idx, idx)], pen_diag)
                      for k in range(Kc)])
        w = np.asarray(prop_mixture, dtype=np.float64) * np.exp(L - L.max())
        w = w / w.sum()
        out[t] = [r_round(v, 5) for v in w]
    return out


def _corr_terms(db, i, t, mean_k, cov_k, mixture):
    idx = db.index[i]
    corr1 = 0.0
    corr2 = 0.0
    for k in range(len(mean_k)):
        mu = mean_k[k][idx]
        corr1 = corr1 + mixture[t, k] * mu
        corr2 = corr2 + mixture[t, k] * (np.outer(mu, mu) + cov_k[k][np.ix_(idx, idx)])
    return corr1, corr2


def elbo_clust_multi_GP(hp, db, i, t, mean_k, cov_k, mixture, kern, pen_diag):
    """elbos.R:18-55 for task i (row t of the mixture)."""
    _, Xi, yi = db.tasks[i]
    inv = K.kern_to_inv(Xi, kern, hp, pen_diag)
    ll = M.neg_dmnorm_log(yi, np.zeros(len(yi)), inv)
    c1, c2 = _corr_terms(db, i, t, mean_k, cov_k, mixture)
    return float(ll - yi @ inv @ c1 + 0.5 * np.sum(inv * c2))


def gr_clust_multi_GP(hp, db, i, t, mean_k, cov_k, mixture, kern, pen_diag):
    """gradients-elbos.R:17-70."""
    _, Xi, yi = db.tasks[i]
    c1, c2 = _corr_terms(db, i, t, mean_k, cov_k, mixture)
    inv = K.kern_to_inv(Xi, kern, hp, pen_diag)
    prod_inv = inv @ yi
    common = np.outer(prod_inv - 2 * inv @ c1, prod_inv) + inv @ (c2 @ inv - np.eye(len(yi)))
    return np.array([float(np.sum(np.diag(-0.5 * (common @ K.kern_to_cov(Xi, kern, hp, d))))) for d in hp])


def elbo_clust_multi_GP_shared_hp_i(hp, db, mean_k, cov_k, mixture, kern, pen_diag):
    """elbos.R:120-172."""
    s = 0.0
    for t, i in enumerate(db.ids):
        s = s + elbo_clust_multi_GP(hp, db, i, t, mean_k, cov_k, mixture, kern, pen_diag)
    return s


def gr_clust_multi_GP_shared_hp_i(hp, db, mean_k, cov_k, mixture, kern, pen_diag):
    """gradients-elbos.R:153-241."""
    g = np.zeros(len(hp))
    for t, i in enumerate(db.ids):
        g += gr_clust_multi_GP(hp, db, i, t, mean_k, cov_k, mixture, kern, pen_diag)
    return g


def elbo_GP_mod_shared_hp_k(hp, db, mean_k, m_k, cov_k, kern, pen_diag):
    """elbos.R:78-100: one inverse, K Gaussian + trace terms."""
    inv = K.kern_to_inv(db.grid_X, kern, hp, pen_diag)
    ll = 0.0
    cor = 0.0
    for k in range(len(mean_k)):
        ll = ll + M.neg_dmnorm_log(mean_k[k], M._mean_vec(m_k[k], len(mean_k[k])), inv)
        cor = cor + 0.5 * float(np.sum(inv * cov_k[k]))
    return ll + cor


def gr_GP_mod_shared_hp_k(hp, db, mean_k, m_k, cov_k, kern, pen_diag):
    """gradients-elbos.R:73-150."""
    inv = K.kern_to_inv(db.grid_X, kern, hp, pen_diag)
    g = np.zeros(len(hp))
    U = db.grid_X.shape[0]
    for k in range(len(mean_k)):
        prod_inv = inv @ (mean_k[k] - M._mean_vec(m_k[k], U))
        common = np.outer(prod_inv, prod_inv) + inv @ (cov_k[k] @ inv - np.eye(U))
        g += np.array([float(np.sum(np.diag(-0.5 * (common @ K.kern_to_cov(db.grid_X, kern, hp, d)))))
                       for d in hp])
    return g


def vm_step(db, old_hp_k, old_hp_i, mean_k, cov_k, mixture, kern_k, kern_i, m_k, shared_hp_k, shared_hp_i,
            pen_diag, stats=None):
    """vem-magmaclust.R:191-303.  Returns (new_m_k, new_hp_k, new_hp_i, prop_mixture)."""
    Kc = len(mean_k)
    U = db.grid_X.shape[0]
    new_m_k = [np.repeat(r_mean(mean_k[k]), U) for k in range(Kc)]
    if stats is not None:
        stats["tasks"] = {}
    if shared_hp_i:
        one, r = M._optim(
            old_hp_i[db.ids[0]],
            lambda h: elbo_clust_multi_GP_shared_hp_i(h, db, mean_k, cov_k, mixture, kern_i, pen_diag),
            lambda h: gr_clust_multi_GP_shared_hp_i(h, db, mean_k, cov_k, mixture, kern_i, pen_diag))
        new_hp_i = {i: dict(one) for i in db.ids}
        if stats is not None:
            stats["tasks"]["shared"] = r
    else:
        new_hp_i = {}
        for t, i in enumerate(db.ids):
            new_hp_i[i], r = M._optim(
                old_hp_i[i],
                lambda h: elbo_clust_multi_GP(h, db, i, t, mean_k, cov_k, mixture, kern_i, pen_diag),
                lambda h: gr_clust_multi_GP(h, db, i, t, mean_k, cov_k, mixture, kern_i, pen_diag))
            if stats is not None:
                stats["tasks"][i] = r
    prop = np.array([float(np.sum(np.asarray(mixture[:, k], dtype=np.longdouble)) / mixture.shape[0])
                     for k in range(Kc)])
    if shared_hp_k:
        one, _ = M._optim(
            old_hp_k[0],
            lambda h: elbo_GP_mod_shared_hp_k(h, db, mean_k, new_m_k, cov_k, kern_k, pen_diag),
            lambda h: gr_GP_mod_shared_hp_k(h, db, mean_k, new_m_k, cov_k, kern_k, pen_diag))
        new_hp_k = [dict(one) for _ in range(Kc)]
    else:
        new_hp_k = []
        for k in range(Kc):
            h, _ = M._optim(
                old_hp_k[k],
                lambda h: M.logL_GP_mod(h, db.grid_X, mean_k[k], new_m_k[k], kern_k, cov_k[k], pen_diag),
                lambda h: M.gr_GP_mod(h, db.grid_X, mean_k[k], new_m_k[k], kern_k, cov_k[k], pen_diag))
            new_hp_k.append(h)
    return new_m_k, new_hp_k, new_hp_i, prop


def elbo_monitoring_VEM(hp_k, hp_i, db, kern_i, kern_k, mean_k, cov_k, mixture, m_k, prop_mixture, pen_diag):
    """elbos.R:200-272."""
    from scipy.linalg import lapack
    Kc = len(mean_k)
    sum_ll_k = 0.0
    for k in range(Kc):
        sum_ll_k += M.logL_GP_mod(hp_k[k], db.grid_X, mean_k[k], m_k[k], kern_k, cov_k[k], pen_diag)
    sum_ll_i = 0.0
    for t, i in enumerate(db.ids):
        sum_ll_i += elbo_clust_multi_GP(hp_i[i], db, i, t, mean_k, cov_k, mixture, kern_i, pen_diag)
    sum_corr = 0.0
    for k in range(Kc):
        sum_tau = 0.0
        pi_k = prop_mixture[k]
        for t in range(mixture.shape[0]):
            tk = mixture[t, k]
            lf = 0.0 if (tk == 0 or pi_k == 0) else math.log(pi_k / tk)
            sum_tau = sum_tau + tk * lf
        c, info = lapack.dpotrf(np.asfortranarray(cov_k[k]), lower=0, clean=1)
        if info != 0:
            raise FloatingPointError("chol(cov_k) failed in elbo_monitoring_VEM")
        sum_corr += sum_tau + float(np.sum(np.log(np.diag(c))))
    return -sum_ll_k - sum_ll_i + sum_corr


def train_magmaclust(db, nb_cluster, ini_hp_k, ini_hp_i, kern_k, kern_i, ini_mixture, shared_hp_k=True,
                     shared_hp_i=True, pen_diag=1e-10, n_iter_max=25, cv_threshold=1e-3):
    """VEM loop of train_magmaclust, training.R:1586-1758 (explicit HPs / mixture, prior mean 0)."""
    Kc = nb_cluster
    U = db.grid_X.shape[0]
    m_k = [np.zeros(U) for _ in range(Kc)]
    hp_k = [dict(h) for h in ini_hp_k]
    hp_i = {i: dict(ini_hp_i[i]) for i in db.ids}
    mixture = np.array(ini_mixture, dtype=np.float64)
    prop = np.array([float(np.sum(np.asarray(mixture[:, k], dtype=np.longdouble)) / mixture.shape[0])
                     for k in range(Kc)])
    elbo = -math.inf
    seq = []
    cv = False
    mean_k = cov_k = None
    for it in range(1, n_iter_max + 1):
        mean_k, cov_k, post_mix = ve_step(db, m_k, kern_k, kern_i, hp_k, hp_i, mixture, it, pen_diag, prop)
        new_m_k, new_hp_k, new_hp_i, new_prop = vm_step(db, hp_k, hp_i, mean_k, cov_k, post_mix, kern_k, kern_i,
                                                        m_k, shared_hp_k, shared_hp_i, pen_diag)
        new_elbo = elbo_monitoring_VEM(new_hp_k, new_hp_i, db, kern_i, kern_k, mean_k, cov_k, post_mix, new_m_k,
                                       new_prop, pen_diag)
        diff = new_elbo - elbo
        if math.isnan(diff):
            diff = -math.inf
        m_k, hp_k, hp_i, mixture, prop, elbo = new_m_k, new_hp_k, new_hp_i, post_mix, new_prop, new_elbo
        seq.append(elbo)
        eps = diff / abs(elbo)
        if math.isnan(eps):
            eps = 1.0
        if abs(eps) < cv_threshold:
            cv = True
            break
    return {"hp_k": hp_k, "hp_i": hp_i, "m_k": m_k, "mixture": mixture, "prop_mixture": prop, "mean_k": mean_k,
            "cov_k": cov_k, "seq_elbo": seq, "converged": cv}


# ---- prediction -------------------------------------------------------------------------------------
def hyperposterior_clust(db, hp_k, hp_i, kern_k, kern_i, m_k, mixture, grid_X, pen_diag=1e-10):
    """prediction.R:1685-1764: the K weighted hyper-posteriors re-evaluated on data U grid (the
    mixture is taken as given).  m_k: list of K scalars/vectors.  Returns keys, X, mean[K], cov[K]."""
    if grid_X is None:
        keys, Xg, index = db.grid_keys, db.grid_X, db.index
    else:
        keys, Xg, index = db.extend_grid(grid_X)
    Kc, U = len(hp_k), Xg.shape[0]
    row = {i: t for t, i in enumerate(db.ids)}
    inv_k = [K.kern_to_inv(Xg, kern_k, hp_k[k], pen_diag) for k in range(Kc)]
    inv_i = {i: K.kern_to_inv(db.tasks[i][1], kern_i, hp_i[i], pen_diag) for i in db.ids}
    cov_k, mean_k = [], []
    for k in range(Kc):
        post_inv = inv_k[k].copy()
        for i in db.ids:
            idx = index[i]
            post_inv[np.ix_(idx, idx)] += mixture[row[i], k] * inv_i[i]
        cov_k.append(K.chol_inv_jitter(post_inv, pen_diag))
    for k in range(Kc):
        weighted = inv_k[k] @ M._mean_vec(m_k[k], U)
        for i in db.ids:
            weighted[index[i]] += mixture[row[i], k] * (inv_i[i] @ db.tasks[i][2])
        mean_k.append(cov_k[k] @ weighted)
    return {"keys": keys, "X": Xg, "mean": mean_k, "cov": cov_k}


def pred_magmaclust(X_obs, y_obs, X_pred, kern, hp, hyperpost, proba, pen_diag=1e-10):
    """prediction.R:2298-2413 for one new task whose mixture probabilities are `proba` (K)."""
    preds = [M.pred_magma(X_obs, y_obs, X_pred, kern, hp,
                          {"keys": hyperpost["keys"], "mean": hyperpost["mean"][k], "cov": hyperpost["cov"][k]}, pen_diag)
             for k in range(len(proba))]
    mixture_mean = 0
    for k, p in enumerate(preds):
        mixture_mean = mixture_mean + proba[k] * p["Mean"]
    mixture_var = 0
    for k, p in enumerate(preds):
        mixture_var = mixture_var + proba[k] * (p["Var"] + (p["Mean"] - mixture_mean) ** 2)
    return {"pred": preds, "Mean": mixture_mean, "Var": mixture_var, "keys": preds[0]["keys"], "X": preds[0]["X"]}


def sum_logL_GP_clust(hp, X, y, mixture, mean_k, cov_k, kern, pen_diag, prop_mixture=None):
    """likelihoods.R:349-398 for one task: mean_k / cov_k are the K hyper-posteriors at the task's inputs."""
    terms = []
    for k in range(len(mean_k)):
        tau_k = mixture[k]
        s = tau_k * M.logL_GP(hp, X, y, mean_k[k], kern, cov_k[k], pen_diag)
        if prop_mixture is not None:
            pi_k = prop_mixture[k]
            log_frac = 0.0 if (tau_k == 0 or pi_k == 0) else math.log(pi_k / tau_k)
            s = -s + tau_k * log_frac
        terms.append(s)
    return float(np.sum(np.array(terms, dtype=np.longdouble)))


def gr_sum_logL_GP_clust(hp, X, y, mixture, mean_k, cov_k, kern, pen_diag):
    """gradients-likelihoods.R:192-227."""
    g = np.zeros(len(hp), dtype=np.longdouble)
    for k in range(len(mean_k)):
        g = g + mixture[k] * M.gr_GP(hp, X, y, mean_k[k], kern, cov_k[k], pen_diag)
    return np.asarray(g, dtype=np.float64)


def train_gp_clust(X, y, mean_k, cov_k, kern, ini_hp, prop_mixture, pen_diag=1e-10, n_iter_max=25, cv_threshold=1e-3):
    """training.R:1923-2217 (EM loop :2098-2215) for one new task."""
    hp = dict(ini_hp)
    mixture = np.asarray(prop_mixture, dtype=np.float64).copy()
    prop = np.asarray(prop_mixture, dtype=np.float64)
    ll, cv, n_it, counts = -np.inf, False, 0, []
    Kc = len(mean_k)
    for it in range(1, n_iter_max + 1):
        new_hp, res = M._optim(hp, lambda h: sum_logL_GP_clust(h, X, y, mixture, mean_k, cov_k, kern, pen_diag),
                               lambda h: gr_sum_logL_GP_clust(h, X, y, mixture, mean_k, cov_k, kern, pen_diag))
        if any(not np.isfinite(v) for v in new_hp.values()):
            break
        counts.append(res["counts"])
        n_it = it
        L = np.array([-M.logL_GP_mod(new_hp, X, y, mean_k[k], kern, cov_k[k], pen_diag) for k in range(Kc)])
        w = prop * np.exp(L - L.max())
        w = w / w.sum()
        new_mixture = np.array([r_round(v, 5) for v in w])
        new_ll = sum_logL_GP_clust(new_hp, X, y, new_mixture, mean_k, cov_k, kern, pen_diag, prop)
        diff = new_ll - ll
        if np.isnan(diff):
            diff = -np.inf
        hp, mixture, ll = new_hp, new_mixture, new_ll
        eps = diff / abs(ll)
        if np.isnan(eps):
            eps = 1.0
        if abs(eps) < cv_threshold:
            cv = True
            break
    return {"hp": hp, "mixture": mixture, "n_iter": n_it, "converged": cv, "logL": ll, "counts": counts}
