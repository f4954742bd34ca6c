"""GPU parity of the new-task path against the oracle: the marginal forms logL_GP / gr_GP (likelihoods.R:73-86,
gradients-likelihoods.R:22-49), train_gp (training.R:78-287) and pred_magma / pred_magmaclust (prediction.R:1146-1197,
2319-2404) batched over several new tasks against a hyper-posterior that stays resident on the device."""
import numpy as np
import pytest

from helpers import make_case, upload, relerr
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
from oracle import magma as OM
from oracle import magmaclust as OC

pytestmark = pytest.mark.gpu


def _new_tasks(seed, sizes, kern="SE"):
    """A handful of new tasks with ragged sizes on the training pool, HPs drawn like hp() does."""
    df, _ = simu_db(M=len(sizes), N=max(sizes), seed=seed)
    rows = []
    for t, i in enumerate(sorted(df.ID.unique(), key=lambda s: s.encode())):
        rows.append(df[df.ID == i].iloc[:sizes[t]])
    import pandas as pd
    df = pd.concat(rows)
    df = df.assign(ID=["new" + s for s in df.ID.values])
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    odb = OM.Dataset(df.ID.values, df.Input.values, df.Output.values)
    hi = ini_hp(kern, prep.ids, False, True, seed + 9)
    names = [c for c in hi.columns if c != "ID"]
    hp = hi[names].to_numpy(dtype=np.float64)
    return prep, odb, hp, names


def _magma_hyperpost(ctx, seed=21, kern_i="SE"):
    """Train-side setup: a Magma hyper-posterior on (training inputs U new-task inputs U prediction grid), computed by
    the library and left resident; the oracle's copy for comparison."""
    case = make_case(10, 10, seed, kern_i=kern_i)
    upload(ctx, case)
    new_prep, new_odb, hp_new, names = _new_tasks(seed + 100, [3, 10, 7, 1, 6], kern_i)
    Xp = np.round(np.linspace(-1, 1, 41), 3)[:, None]
    extra = np.vstack([new_prep.X, Xp])
    ohyp = OM.hyperposterior(case["odb"], case["o_hp0"], case["o_hpi"], case["kern_0"], kern_i, None, extra)
    prep = case["prep"]
    prep.set_grid(extra)
    assert [k.decode() for k in prep.grid_keys] == ohyp["keys"]
    ctx.hyperpost_keep(True)
    pm, pc = ctx.hyperposterior(case["k0"], case["hp0"], case["ki"], case["hpi"], prep.grid_X, prep.grid_index, 0.0)
    ctx.hyperpost_keep(False)
    assert relerr(pm, ohyp["mean"]) < 1e-4 and relerr(pc, ohyp["cov"]) < 1e-4
    # the new tasks become the resident dataset, indexed into the hyper-posterior's grid
    gi_new = prep.index_of(new_prep.keys)
    ctx.upload_dataset(new_prep.offs, new_prep.X, new_prep.y, gi_new, prep.grid_X)
    return case, prep, new_prep, new_odb, hp_new, names, Xp, pm, pc, gi_new


def test_logL_GP_and_train_gp_tasks(ctx):
    """Marginal value + gradient of every new task in one launch, then train_gp of all of them in lock-step: same
    evaluation counts / convergence codes as the oracle's L-BFGS-B, HPs to 1e-6."""
    case, prep, new_prep, new_odb, hp_new, names, Xp, pm, pc, gi = _magma_hyperpost(ctx)
    ki = case["ki"]
    f, g, r = ctx.logL_GP_tasks(ki, hp_new)
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        idx = gi[new_prep.offs[t]:new_prep.offs[t + 1]]
        hpd = dict(zip(names, hp_new[t]))
        # the oracle is evaluated on the LIBRARY's hyper-posterior block so that only this step is compared
        of = OM.logL_GP(hpd, X, y, pm[idx], "SE", pc[np.ix_(idx, idx)], 1e-10)
        og = OM.gr_GP(hpd, X, y, pm[idx], "SE", pc[np.ix_(idx, idx)], 1e-10)
        assert abs(f[t] - of) <= 1e-9 * max(1.0, abs(of)), (t, f[t], of)
        np.testing.assert_allclose(g[t], og, rtol=1e-8, atol=1e-9)
        assert r[t] == 0
    hp_fit, stats = ctx.train_gp_tasks(ki, hp_new)
    conv, nev = ctx.last_task_status()
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        idx = gi[new_prep.offs[t]:new_prep.offs[t + 1]]
        ohp, res = OM.train_gp(X, y, pm[idx], "SE", dict(zip(names, hp_new[t])), pc[np.ix_(idx, idx)], 1e-10,
                               return_info=True)
        assert nev[t] == res["counts"] and conv[t] == res["convergence"], (t, nev[t], res["counts"], conv[t])
        np.testing.assert_allclose(hp_fit[t], [ohp[n] for n in names], rtol=1e-6, atol=1e-7)
    assert stats[2] == nev.sum()


def test_sum_logL_GP_terms(ctx):
    """post_cov = NULL and a constant mean: the per-task terms of sum_logL_GP (likelihoods.R:113-129)."""
    case, prep, new_prep, new_odb, hp_new, names, Xp, pm, pc, gi = _magma_hyperpost(ctx)
    ctx.hyperpost_set(np.full(ctx.U, 0.7))
    f, g, _ = ctx.logL_GP_tasks(case["ki"], hp_new[0], shared=True)
    hpd = dict(zip(names, hp_new[0]))
    total = 0.0
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        of = OM.logL_GP(hpd, X, y, 0.7, "SE", 0.0, 1e-10)
        assert abs(f[t] - of) <= 1e-9 * max(1.0, abs(of))
        total += of
    assert abs(f.sum() - total) <= 1e-9 * abs(total)


def test_pred_magma_tasks(ctx):
    """pred_magma for all new tasks in one call == the oracle's per-task pred_magma (Mean, Var)."""
    case, prep, new_prep, new_odb, hp_new, names, Xp, pm, pc, gi = _magma_hyperpost(ctx)
    from magmaclustr_b200.dataprep import reference_keys, r_signif
    Xp_r = r_signif(Xp)
    kp = reference_keys(Xp_r)
    op = np.argsort(kp, kind="stable")
    ip = prep.index_of(kp[op])
    mean, var, _, _, r = ctx.pred_magma_tasks(case["ki"], hp_new, Xp_r[op], ip)
    hyp = {"keys": [k.decode() for k in prep.grid_keys], "mean": pm, "cov": pc}
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        o = OM.pred_magma(X, y, Xp, "SE", dict(zip(names, hp_new[t])), hyp)
        np.testing.assert_allclose(mean[t], o["Mean"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(var[t], o["Var"], rtol=1e-7, atol=1e-9)
        assert r[t, 0] == 0
    # the same call cut into chunks of two tasks (the library bounds its scratch table that way at large sizes)
    import os
    os.environ["MB_PRED_CHUNK"] = "2"
    try:
        mean2, var2, _, _, _ = ctx.pred_magma_tasks(case["ki"], hp_new, Xp_r[op], ip)
    finally:
        del os.environ["MB_PRED_CHUNK"]
    assert np.array_equal(mean2, mean) and np.array_equal(var2, var)


def test_pred_magmaclust_tasks(ctx):
    """K = 3 resident cluster hyper-posteriors, mixture probabilities per new task: per-cluster and mixture outputs."""
    KC = 3
    case = make_case(12, 10, 4, shared=True, K=KC)
    upload(ctx, case)
    rng = np.random.default_rng(4)
    lab = rng.integers(0, KC, size=12)
    mix = np.zeros((12, KC))
    mix[np.arange(12), lab] = 1.0
    hp_k = np.array([[1.0 + 0.3 * k, -1.5 - 0.2 * k] for k in range(KC)])
    o_hp_k = [dict(zip(case["names_0"], hp_k[k])) for k in range(KC)]
    new_prep, new_odb, hp_new, names = _new_tasks(77, [4, 9, 2, 10])
    Xp = np.round(np.linspace(-1, 1, 150), 3)[:, None]        # more than one block of prediction points
    extra = np.vstack([new_prep.X, Xp])
    prep = case["prep"]
    ohyp = OC.hyperposterior_clust(case["odb"], o_hp_k, case["o_hpi"], "SE", "SE", [0.0] * KC, mix, extra)
    prep.set_grid(extra)
    ctx.hyperpost_keep(True)
    pm, pc = ctx.hyperposterior_clust(KC, case["k0"], hp_k, case["ki"], case["hpi"], prep.grid_X, prep.grid_index,
                                      np.zeros((KC, 1)), mix)
    ctx.hyperpost_keep(False)
    gi_new = prep.index_of(new_prep.keys)
    ctx.upload_dataset(new_prep.offs, new_prep.X, new_prep.y, gi_new, prep.grid_X)
    from magmaclustr_b200.dataprep import reference_keys, r_signif
    Xp_r = r_signif(Xp)
    kp = reference_keys(Xp_r)
    op = np.argsort(kp, kind="stable")
    ip = prep.index_of(kp[op])
    tau = rng.dirichlet(np.ones(KC), size=new_prep.n_tasks)
    mean, var, mk, vk, r = ctx.pred_magma_tasks(case["ki"], hp_new, Xp_r[op], ip, tau=tau, per_cluster=True, K=KC)
    hyp = {"keys": [k.decode() for k in prep.grid_keys], "mean": list(pm), "cov": list(pc)}
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        o = OC.pred_magmaclust(X, y, Xp, "SE", dict(zip(names, hp_new[t])), hyp, tau[t])
        for k in range(KC):
            np.testing.assert_allclose(mk[t, k], o["pred"][k]["Mean"], rtol=1e-8, atol=1e-8)
            np.testing.assert_allclose(vk[t, k], o["pred"][k]["Var"], rtol=1e-7, atol=1e-9)
        np.testing.assert_allclose(mean[t], o["Mean"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(var[t], o["Var"], rtol=1e-7, atol=1e-9)
    assert np.all(var >= 0) and np.all(r == 0)


def test_train_gp_clust_tasks(ctx):
    """train_gp_clust of several new tasks in one call: every task runs its own EM loop (M step L-BFGS-B on
    sum_logL_GP_clust, E step update_mixture, monitoring) and stops at its own iteration -- same iteration counts,
    convergence flags and mixtures as the oracle's stand-alone loop, HPs to 1e-6."""
    KC = 3
    case = make_case(12, 10, 4, shared=True, K=KC)
    upload(ctx, case)
    rng = np.random.default_rng(4)
    lab = rng.integers(0, KC, size=12)
    mix = np.zeros((12, KC))
    mix[np.arange(12), lab] = 1.0
    hp_k = np.array([[1.0 + 0.3 * k, -1.5 - 0.2 * k] for k in range(KC)])
    new_prep, new_odb, hp_new, names = _new_tasks(78, [5, 10, 3, 8, 10, 2])
    prep = case["prep"]
    prep.set_grid(new_prep.X)
    ctx.hyperpost_keep(True)
    pm, pc = ctx.hyperposterior_clust(KC, case["k0"], hp_k, case["ki"], case["hpi"], prep.grid_X, prep.grid_index,
                                      np.zeros((KC, 1)), mix)
    ctx.hyperpost_keep(False)
    gi = prep.index_of(new_prep.keys)
    ctx.upload_dataset(new_prep.offs, new_prep.X, new_prep.y, gi, prep.grid_X)
    prop = mix.mean(axis=0)
    hp_fit, mixture, nit, cv, stats = ctx.train_gp_clust_tasks(case["ki"], hp_new, prop)
    its = []
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        idx = gi[new_prep.offs[t]:new_prep.offs[t + 1]]
        o = OC.train_gp_clust(X, y, [pm[k][idx] for k in range(KC)], [pc[k][np.ix_(idx, idx)] for k in range(KC)], "SE",
                              dict(zip(names, hp_new[t])), prop)
        its.append(o["n_iter"])
        assert nit[t] == o["n_iter"] and bool(cv[t]) == o["converged"], (t, nit[t], o["n_iter"], cv[t], o["converged"])
        np.testing.assert_allclose(mixture[t], o["mixture"], atol=2e-5)
        np.testing.assert_allclose(hp_fit[t], [o["hp"][n] for n in names], rtol=1e-6, atol=1e-7)
    assert stats[2] > 0
    print("EM iterations per task", its)


def test_train_shared_gp(ctx):
    """train_shared_gp / precondition_hp (training.R:352-446): one shared HP vector, L-BFGS-B with optim's central
    differences on sum_logL_GP.  The finite-difference gradient divides the objective's rounding noise by 2e-3, so the
    trajectories may part by an evaluation; both optima must be equally good and the HPs agree to 1e-4."""
    case = make_case(8, 10, 31, shared=True)
    upload(ctx, case)
    ki, names = case["ki"], case["names_i"]
    hp0 = case["hpi"][0]
    hp_fit, stats = ctx.train_shared_gp(ki, hp0, mean=0.0, maxit=100)
    ohp, res = OM.train_shared_gp(case["odb"], 0.0, "SE", dict(zip(names, hp0)), 1e-10, 100, return_info=True)
    print("gpu", hp_fit, stats, "oracle", ohp, res["counts"], res["convergence"], res["value"])
    f_gpu = OM.sum_logL_GP(dict(zip(names, hp_fit)), case["odb"], 0.0, "SE", 1e-10)
    f_ini = OM.sum_logL_GP(dict(zip(names, hp0)), case["odb"], 0.0, "SE", 1e-10)
    assert f_gpu < f_ini and abs(f_gpu - res["value"]) <= 1e-6 * abs(res["value"])
    np.testing.assert_allclose(hp_fit, [ohp[n] for n in names], rtol=1e-4, atol=1e-5)
    assert abs(int(stats[0]) - res["counts"]) <= 2 and stats[1] == res["convergence"]
    assert stats[2] == stats[0] * (1 + 2 * len(names)) * 8


@pytest.mark.parametrize("kern", ["RQ + PERIO", "SE * LIN", "PERIO"])
def test_new_task_path_compound_kernels(ctx, kern):
    """The new-task entry points with the generic kernel grammar (kernel-to-matrices.R:163-340): marginal value and
    gradient per task, and the batched prediction, against the oracle on the library's own hyper-posterior."""
    case, prep, new_prep, new_odb, _, _, Xp, pm, pc, gi = _magma_hyperpost(ctx, seed=33)
    k = L.parse_kernel(kern, True)
    names = L.hp_names(k)
    rng = np.random.default_rng(3)
    hp = rng.uniform(0.2, 1.2, size=(new_prep.n_tasks, len(names)))
    hp[:, names.index("noise")] = rng.uniform(-3.0, -1.0, size=new_prep.n_tasks)
    f, g, r = ctx.logL_GP_tasks(k, hp)
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        idx = gi[new_prep.offs[t]:new_prep.offs[t + 1]]
        hpd = dict(zip(names, hp[t]))
        of = OM.logL_GP(hpd, X, y, pm[idx], kern, pc[np.ix_(idx, idx)], 1e-10)
        og = OM.gr_GP(hpd, X, y, pm[idx], kern, pc[np.ix_(idx, idx)], 1e-10)
        assert abs(f[t] - of) <= 1e-9 * max(1.0, abs(of)), (kern, t, f[t], of)
        np.testing.assert_allclose(g[t], og, rtol=1e-7, atol=1e-8, err_msg="%s task %d" % (kern, t))
    from magmaclustr_b200.dataprep import reference_keys, r_signif
    Xp_r = r_signif(Xp)
    kp = reference_keys(Xp_r)
    op = np.argsort(kp, kind="stable")
    mean, var, _, _, r = ctx.pred_magma_tasks(k, hp, Xp_r[op], prep.index_of(kp[op]))
    hyp = {"keys": [q.decode() for q in prep.grid_keys], "mean": pm, "cov": pc}
    for t, i in enumerate(new_prep.ids):
        _, X, y = new_odb.tasks[i]
        o = OM.pred_magma(X, y, Xp, kern, dict(zip(names, hp[t])), hyp)
        np.testing.assert_allclose(mean[t], o["Mean"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(var[t], o["Var"], rtol=1e-7, atol=1e-9)
