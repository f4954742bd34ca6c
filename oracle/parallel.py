"""All-host-cores driver of the oracle's EM step (TEST INFRASTRUCTURE / CPU baseline only).

The reference runs its per-task loops sequentially in one R thread
(/root/reference/R/em-magma.R:200-230, R/likelihoods.R:271-298).  For the `cpu_baseline`
and `--impl reference` legs of bench.py the same restated arithmetic (oracle/magma.py) is
spread over a fork pool, one task per work item, so the CPU figure uses every host core.
"""
import math
import multiprocessing as mp
import os

import numpy as np

from . import magma as M

_G = {}


def _limit_blas_threads():
    """One BLAS thread per worker: the pool already uses every core (no oversubscription)."""
    try:
        from threadpoolctl import threadpool_limits
        _G["_tp"] = threadpool_limits(1)
    except Exception:  # pragma: no cover
        os.environ["OPENBLAS_NUM_THREADS"] = "1"


def _task_opt(i):
    db, hp_i, kern_i, pm, pc, pen = (_G[k] for k in ("db", "hp_i", "kern_i", "pm", "pc", "pen"))
    Xi, yi, mi, Si = M._task_views(db, pm, pc, i)
    try:
        new, res = M._optim(hp_i[i],
                            lambda h: M.logL_GP_mod(h, Xi, yi, mi, kern_i, Si, pen),
                            lambda h: M.gr_GP_mod(h, Xi, yi, mi, kern_i, Si, pen))
    except (FloatingPointError, OverflowError, np.linalg.LinAlgError):
        # a line-search trial point without a finite objective: R's optim stops with "L-BFGS-B needs finite values of 'fn'"
        # and train_magma aborts.  The CPU baseline is a timing: the task keeps its old HPs and the step goes on (the
        # product ends such a task at its last accepted iterate and reports it).
        return i, dict(hp_i[i]), 0
    return i, new, res["counts"]


def _task_ll(i):
    db, hp_i, kern_i, pm, pc, pen = (_G[k] for k in ("db", "new_hp_i", "kern_i", "pm", "pc", "pen"))
    Xi, yi, mi, Si = M._task_views(db, pm, pc, i)
    return M.logL_GP_mod(hp_i[i], Xi, yi, mi, kern_i, Si, pen)


def _task_inv(i):
    db, hp_i, kern_i, pen = (_G[k] for k in ("db", "hp_i", "kern_i", "pen"))
    _, Xi, _ = db.tasks[i]
    return M.K.chol_inv_jitter(M.K.kern_to_cov(Xi, kern_i, hp_i[i]), pen)


def e_step(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag, pool, chunksize):
    """oracle.magma.e_step with the per-task chol_inv_jitter spread over the pool; the inverses are added in task order by
    the parent, exactly as the sequential loop does (em-magma.R:39-46).  Same operations in the same order; the last bits
    can differ from oracle.magma.e_step because OpenBLAS blocks dpotrf / dpotri differently with one thread (workers) than
    with several (a stand-alone call).  Timing baseline only -- the parity tests use oracle.magma."""
    U = db.grid_X.shape[0]
    m_0 = M._mean_vec(m_0, U)
    inv_0 = M.K.chol_inv_jitter(M.K.kern_to_cov(db.grid_X, kern_0, hp_0), pen_diag)
    post_inv = inv_0.copy()
    weighted_0 = inv_0 @ m_0
    list_inv = dict(zip(db.ids, pool.map(_task_inv, db.ids, chunksize=chunksize)))
    for i in db.ids:
        idx = db.index[i]
        post_inv[np.ix_(idx, idx)] += list_inv[i]
    post_cov = M.K.chol_inv_jitter(post_inv, pen_diag)
    for i in db.ids:
        weighted_0[db.index[i]] += list_inv[i] @ db.tasks[i][2]
    return post_cov @ weighted_0, post_cov


def em_step(db, m_0, kern_0, kern_i, hp_0, hp_i, pen_diag=1e-10, procs=None):
    """One iteration of train_magma's loop body (training.R:905-1030), shared_hp = FALSE.
    Returns (new_hp_0, new_hp_i, logL, n_task_evals)."""
    procs = procs or os.cpu_count()
    _limit_blas_threads()       # 64 x 64 matrices: BLAS threads only add overhead, and the pool below already uses every core
    U = db.grid_X.shape[0]
    m0 = M._mean_vec(m_0, U)
    ctx = mp.get_context("fork")
    chunk = max(1, len(db.ids) // (4 * procs))
    _G.update(db=db, hp_i=hp_i, kern_i=kern_i, pen=pen_diag)
    with ctx.Pool(procs, initializer=_limit_blas_threads) as pool:
        pm, pc = e_step(db, m0, kern_0, kern_i, hp_0, hp_i, pen_diag, pool, chunk)
    _G.update(pm=pm, pc=pc)
    with ctx.Pool(procs, initializer=_limit_blas_threads) as pool:
        fut = pool.map_async(_task_opt, db.ids, chunksize=max(1, len(db.ids) // (4 * procs)))
        new_hp_0, _ = M._optim(
            hp_0,
            lambda h: M.logL_GP_mod(h, db.grid_X, pm, m0, kern_0, pc, pen_diag),
            lambda h: M.gr_GP_mod(h, db.grid_X, pm, m0, kern_0, pc, pen_diag))
        out = fut.get()
    new_hp_i = {i: h for i, h, _ in out}
    evals = sum(c for _, _, c in out)
    _G.update(new_hp_i=new_hp_i)
    with ctx.Pool(procs, initializer=_limit_blas_threads) as pool:
        lls = pool.map(_task_ll, db.ids, chunksize=max(1, len(db.ids) // (4 * procs)))
    from scipy.linalg import lapack
    ll_0 = M.logL_GP_mod(new_hp_0, db.grid_X, pm, m0, kern_0, pc, pen_diag)
    c, info = lapack.dpotrf(np.asfortranarray(pc), lower=0, clean=1)
    det = float(np.sum(np.log(np.diag(c)))) if info == 0 else math.nan
    s = 0.0
    for v in lls:
        s += v
    return new_hp_0, new_hp_i, -ll_0 - s + det, evals
