"""select_nb_cluster (/root/reference/R/training.R:2286-2454) on top of the library's training loops.

One training per candidate K -- train_magma for K = 1, train_magmaclust otherwise, ``fast_approx`` by default (one (V)E step
and the monitoring value of the initial hyper-parameters, training.R:921-937,1633-1647) -- then the variational BIC
``elbo - 0.5 (nb_hp_k + nb_hp_i + K - 1) log(nb_tasks)`` with nb_hp = number of DISTINCT hyper-parameter values
(training.R:2411-2423).  The K trainings are independent: with several GPUs every rank keeps the WHOLE dataset resident,
takes the candidates ``grid[rank::world]`` and only the K scalars are exchanged (torch.distributed object gather) -- no
data-path collective, efficiency ~ 1.

``ini_mixture``: the reference clusters (min, mean, max) of every task's outputs with stats::kmeans (Hartigan-Wong,
nstart = 50, R's RNG; initialisation.R:192-268).  That generator is not reproducible outside R; `ini_mixture` below uses the
same three features with scikit-learn's k-means (n_init = 50).  Pass explicit mixtures for parity runs.
"""
import math

import numpy as np


def ini_mixture(prep, k, seed=0):
    """Hard initial mixture (tasks x k) from k-means on (min, mean, max) of each task's outputs."""
    from sklearn.cluster import KMeans
    feats = np.array([[prep.y[prep.offs[t]:prep.offs[t + 1]].min(), prep.y[prep.offs[t]:prep.offs[t + 1]].mean(),
                       prep.y[prep.offs[t]:prep.offs[t + 1]].max()] for t in range(len(prep.offs) - 1)])
    if k == 1:
        return np.ones((len(feats), 1))
    lab = KMeans(n_clusters=k, n_init=50, random_state=seed).fit(feats).labels_
    mix = np.zeros((len(feats), k))
    mix[np.arange(len(feats)), lab] = 1.0
    return mix


def _n_distinct(a):
    return int(len(np.unique(np.asarray(a, dtype=np.float64))))


def vbic(elbo, hp_k, hp_i, k, n_tasks):
    """training.R:2411-2425."""
    return elbo - 0.5 * (_n_distinct(hp_k) + _n_distinct(hp_i) + k - 1) * math.log(n_tasks)


def select_nb_cluster(ctx, prep, kern_k, kern_i, ini_hp_k, ini_hp_i, grid_nb_cluster=range(1, 11), fast_approx=True,
                      ini_mixtures=None, shared_hp_k=True, shared_hp_i=True, pen_diag=1e-10, dist=None, seed=0, **train_kw):
    """ctx: a Context with the WHOLE dataset resident (no mb_set_shard).  ini_hp_k: one HP row (P_k) attributed to all
    clusters (training.R:2362-2364); ini_hp_i: tasks x P_i.  ini_mixtures: optional {K: tasks x K array}.
    Returns {"best_k", "seq_vbic" {K: value}, "seq_elbo" {K: value}, "model"} -- `model` is the training result of best_k."""
    grid = [int(k) for k in grid_nb_cluster]
    rank, world = (dist.get_rank(), dist.get_world_size()) if dist is not None else (0, 1)
    n_tasks = len(prep.offs) - 1
    U = ctx.U
    ini_hp_k = np.asarray(ini_hp_k, dtype=np.float64).reshape(-1)
    ini_hp_i = np.ascontiguousarray(ini_hp_i, dtype=np.float64)
    mine = {}
    for k in grid[rank::world]:
        hp_k = np.tile(ini_hp_k, (k, 1))
        if k == 1:
            if fast_approx:
                pm, pc, _ = ctx.e_step(kern_k, hp_k[0], kern_i, ini_hp_i, np.zeros(U), pen_diag=pen_diag)
                elbo = ctx.logL_monitoring(kern_k, hp_k[0], kern_i, ini_hp_i, np.zeros(U), pm, pc, pen_diag=pen_diag)
                mod = {"hp_0": hp_k[0], "hp_i": ini_hp_i, "post_mean": pm, "post_cov": pc,
                       "seq_loglikelihood": np.array([elbo]), "converged": False}
            else:
                mod = ctx.train_magma(kern_k, hp_k[0], kern_i, ini_hp_i, np.zeros(U), shared_hp=shared_hp_i,
                                      pen_diag=pen_diag, **train_kw)
                elbo = float(mod["seq_loglikelihood"][-1])
        else:
            mix = ini_mixtures[k] if ini_mixtures is not None and k in ini_mixtures else ini_mixture(prep, k, seed)
            mod = ctx.train_magmaclust(k, kern_k, hp_k, kern_i, ini_hp_i, mix, shared_hp_k, shared_hp_i, pen_diag=pen_diag,
                                       fast_approx=fast_approx, **train_kw)
            elbo = float(mod["seq_elbo"][-1])
        mod["V-BIC"] = vbic(elbo, hp_k, ini_hp_i, k, n_tasks)
        mod["elbo"] = elbo
        mine[k] = mod
    table = {k: (m["elbo"], m["V-BIC"]) for k, m in mine.items()}
    if world > 1:
        parts = [None] * world
        dist.all_gather_object(parts, table)
        table = {k: v for p in parts for k, v in p.items()}
    seq_vbic = {k: table[k][1] for k in grid}
    best_k = max(grid, key=lambda k: seq_vbic[k])          # which.max: first maximum in grid order
    model = mine.get(best_k)
    if world > 1:
        box = [model if best_k in mine else None]
        dist.broadcast_object_list(box, src=grid.index(best_k) % world)
        model = box[0]
    return {"best_k": best_k, "seq_vbic": seq_vbic, "seq_elbo": {k: table[k][0] for k in grid}, "model": model}
