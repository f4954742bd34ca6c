"""GPU parity for the MagmaClust path (ve_step, update_mixture, ELBO + gradient, vm_step,
elbo_monitoring_VEM, full VEM loop) against the oracle."""
import numpy as np
import pytest

from helpers import make_case, upload, relerr, ulp_perturb
from magmaclustr_b200 import _lib as L
from oracle import magma as OM
from oracle import magmaclust as OC

pytestmark = pytest.mark.gpu
KC = 3


def _setup(ctx, M=12, N=10, seed=4, shared_i=True):
    case = make_case(M, N, seed, shared=shared_i, K=KC)
    upload(ctx, case)
    rng = np.random.default_rng(seed)
    # hard initial mixture (what ini_mixture's k-means gives), rows in prep.ids order
    lab = rng.integers(0, KC, size=M)
    mix = np.zeros((M, KC))
    mix[np.arange(M), lab] = 1.0
    hp_k = np.array([[1.0 + 0.3 * k, -1.5 - 0.2 * k] for k in range(KC)])
    o_hp_k = [dict(zip(case["names_0"], hp_k[k])) for k in range(KC)]
    return case, mix, hp_k, o_hp_k


def _oracle_ve(case, mix, o_hp_k, it=1, prop=None):
    U = case["odb"].grid_X.shape[0]
    return OC.ve_step(case["odb"], [np.zeros(U)] * KC, "SE", "SE", o_hp_k, case["o_hpi"], mix, it, 1e-10, prop)


def test_ve_step_and_update_mixture(ctx):
    case, mix, hp_k, o_hp_k = _setup(ctx)
    U = ctx.U
    pm, pc, tau = ctx.ve_step(KC, case["k0"], hp_k, case["ki"], case["hpi"], np.zeros((KC, U)), mix)
    omean, ocov, omix = _oracle_ve(case, mix, o_hp_k)
    assert np.array_equal(tau, mix)                         # mixture kept at iteration 1 (:130-131)
    rng = np.random.default_rng(0)
    for k in range(KC):
        sens = 0.0
        for _ in range(3):
            hk = [dict(zip(case["names_0"], ulp_perturb(hp_k[q], rng))) for q in range(KC)]
            pmean, pcov, _ = _oracle_ve(case, mix, hk)
            sens = max(sens, relerr(pcov[k], ocov[k]), relerr(pmean[k], omean[k]))
        assert relerr(pc[k], ocov[k]) < 30 * sens + 1e-12
        assert relerr(pm[k], omean[k]) < 30 * sens + 1e-12
    # responsibilities from the ORACLE's hyper-posterior: identical 5-decimal values and clusters
    prop = mix.mean(axis=0)
    otau = OC.update_mixture(case["odb"], omean, ocov, case["o_hpi"], "SE", prop, 1e-10)
    gtau = ctx.update_mixture(KC, case["ki"], case["hpi"], np.array(omean), np.array(ocov), prop)
    assert np.array_equal(np.argmax(gtau, axis=1), np.argmax(otau, axis=1))      # cluster assignment: exact
    assert np.max(np.abs(gtau - otau)) <= 1e-5 + 1e-12                            # at most one rounding unit
    assert np.mean(gtau == otau) > 0.95


@pytest.mark.parametrize("shared_i", [True, False])
def test_elbo_and_gradient(ctx, shared_i):
    case, mix, hp_k, o_hp_k = _setup(ctx, shared_i=shared_i)
    omean, ocov, _ = _oracle_ve(case, mix, o_hp_k)
    soft = np.round(0.8 * mix + 0.2 / KC, 5)
    f, g = ctx.elbo_clust_tasks(KC, case["ki"], case["hpi"], np.array(omean), np.array(ocov), soft)
    for t, i in enumerate(case["prep"].ids):
        of = OC.elbo_clust_multi_GP(case["o_hpi"][i], case["odb"], i, t, omean, ocov, soft, "SE", 1e-10)
        og = OC.gr_clust_multi_GP(case["o_hpi"][i], case["odb"], i, t, omean, ocov, soft, "SE", 1e-10)
        assert abs(f[t] - of) <= 1e-9 * abs(of)
        np.testing.assert_allclose(g[t], og, rtol=1e-8, atol=1e-9 * np.max(np.abs(og)))


@pytest.mark.parametrize("shared_k,shared_i", [(True, True), (False, False)])
def test_vm_step_and_monitoring(ctx, shared_k, shared_i):
    case, mix, hp_k, o_hp_k = _setup(ctx, shared_i=shared_i)
    U = ctx.U
    omean, ocov, _ = _oracle_ve(case, mix, o_hp_k)
    o_m, o_hk, o_hi, o_prop = OC.vm_step(case["odb"], o_hp_k, case["o_hpi"], omean, ocov, mix, "SE", "SE",
                                         [np.zeros(U)] * KC, shared_k, shared_i, 1e-10)
    m_k, h_k, h_i, prop, st = ctx.vm_step(KC, case["k0"], hp_k, case["ki"], case["hpi"], np.zeros((KC, U)),
                                          np.array(omean), np.array(ocov), mix, shared_k, shared_i)
    np.testing.assert_array_equal(prop, o_prop)                                   # colMeans(tau): exact
    np.testing.assert_allclose(m_k, np.array(o_m), rtol=1e-13, atol=1e-13)        # mean(post mean)
    for t, i in enumerate(case["prep"].ids):                                      # task HPs: well conditioned
        np.testing.assert_allclose(h_i[t], [o_hi[i][n] for n in case["names_i"]], rtol=1e-6, atol=1e-7)
    # cluster HPs ride on cond(K_k) ~ 1e12 like hp_0 (test_gpu_magma.py): optimiser paths may part,
    # so both optima must be equally good for the ORACLE's objective (within optim's factr rule)
    for k in range(KC):
        hg = dict(zip(case["names_0"], h_k[k]))
        if shared_k:
            fg = OC.elbo_GP_mod_shared_hp_k(hg, case["odb"], omean, o_m, ocov, "SE", 1e-10)
            fo = OC.elbo_GP_mod_shared_hp_k(o_hk[k], case["odb"], omean, o_m, ocov, "SE", 1e-10)
        else:
            fg = OM.logL_GP_mod(hg, case["odb"].grid_X, omean[k], o_m[k], "SE", ocov[k], 1e-10)
            fo = OM.logL_GP_mod(o_hk[k], case["odb"].grid_X, omean[k], o_m[k], "SE", ocov[k], 1e-10)
        assert abs(fg - fo) < 2 * 2.2e-3 * max(abs(fo), 1.0), (k, fg, fo)
    # monitoring at the SAME hyper-parameters
    hk_d = [dict(zip(case["names_0"], h_k[k])) for k in range(KC)]
    hi_d = {i: dict(zip(case["names_i"], h_i[t])) for t, i in enumerate(case["prep"].ids)}
    o_elbo = OC.elbo_monitoring_VEM(hk_d, hi_d, case["odb"], "SE", "SE", omean, ocov, mix, o_m, o_prop, 1e-10)
    elbo = ctx.elbo_monitoring(KC, case["k0"], h_k, case["ki"], h_i, m_k, np.array(omean), np.array(ocov), mix, prop)
    rng = np.random.default_rng(3)
    sens = 0.0
    for _ in range(3):
        hp = [dict(zip(case["names_0"], ulp_perturb(h_k[k], rng))) for k in range(KC)]
        sens = max(sens, abs(OC.elbo_monitoring_VEM(hp, hi_d, case["odb"], "SE", "SE", omean, ocov, mix, o_m, o_prop,
                                                    1e-10) - o_elbo))
    print("elbo gpu %.8f oracle %.8f diff %.2e sens %.2e" % (elbo, o_elbo, abs(elbo - o_elbo), sens))
    assert abs(elbo - o_elbo) < 30 * sens + 1e-9 * abs(o_elbo)


def test_train_magmaclust_small(ctx):
    """Shape of BASELINE config 3 at test size: K clusters, shared_hp_i = TRUE, VEM to convergence."""
    case, mix, hp_k, o_hp_k = _setup(ctx, M=12, N=10, seed=4, shared_i=True)
    res = ctx.train_magmaclust(KC, case["k0"], hp_k, case["ki"], case["hpi"], mix, True, True, n_iter_max=6)
    ores = OC.train_magmaclust(case["odb"], KC, o_hp_k, case["o_hpi"], "SE", "SE", mix, True, True, n_iter_max=6)
    print("gpu", res["seq_elbo"], "oracle", ores["seq_elbo"])
    assert len(res["seq_elbo"]) == len(ores["seq_elbo"]) and res["converged"] == ores["converged"]
    np.testing.assert_allclose(res["seq_elbo"], ores["seq_elbo"], rtol=2e-3)
    assert np.array_equal(np.argmax(res["mixture"], axis=1), np.argmax(ores["mixture"], axis=1))
    assert np.all(np.diff(res["seq_elbo"]) > -1e-6 * np.abs(res["seq_elbo"][1:]))
    # the resident loop (one library call) and the same steps driven from the host give the same bits
    step = ctx.train_magmaclust_stepwise(KC, case["k0"], hp_k, case["ki"], case["hpi"], mix, True, True, n_iter_max=6)
    for key in ("seq_elbo", "hp_k", "hp_i", "m_k", "mixture", "prop_mixture", "post_mean", "post_cov"):
        assert np.array_equal(np.asarray(res[key]), np.asarray(step[key])), key
    # ... also with per-task / per-cluster HPs.  The resident loop takes the ELBO's task and cluster terms from the VM step's
    # optimisers instead of evaluating them again; the stepwise path always re-evaluates: same bits.
    case2, mix2, hp_k2, _ = _setup(ctx, M=12, N=10, seed=4, shared_i=False)
    res2 = ctx.train_magmaclust(KC, case2["k0"], hp_k2, case2["ki"], case2["hpi"], mix2, False, False, n_iter_max=4)
    step2 = ctx.train_magmaclust_stepwise(KC, case2["k0"], hp_k2, case2["ki"], case2["hpi"], mix2, False, False, n_iter_max=4)
    for key in ("seq_elbo", "hp_k", "hp_i", "m_k", "mixture", "prop_mixture"):
        assert np.array_equal(np.asarray(res2[key]), np.asarray(step2[key])), key
    ctx.set_option("monitor_recompute", 1)
    try:
        res3 = ctx.train_magmaclust(KC, case2["k0"], hp_k2, case2["ki"], case2["hpi"], mix2, False, False, n_iter_max=4)
    finally:
        ctx.set_option("monitor_recompute", 0)
    assert np.array_equal(res3["seq_elbo"], res2["seq_elbo"])
    case, mix, hp_k, o_hp_k = _setup(ctx, M=12, N=10, seed=4, shared_i=True)
    # fast_approx: one VE step, ELBO of the initial hyper-parameters (training.R:1633-1647)
    fa = ctx.train_magmaclust(KC, case["k0"], hp_k, case["ki"], case["hpi"], mix, True, True, fast_approx=True)
    assert len(fa["seq_elbo"]) == 1 and not fa["converged"] and np.array_equal(fa["hp_k"], hp_k)
    omean, ocov, omix = _oracle_ve(case, mix, o_hp_k)
    o_prop = mix.mean(axis=0)
    o_fa = OC.elbo_monitoring_VEM(o_hp_k, case["o_hpi"], case["odb"], "SE", "SE", omean, ocov, omix,
                                  [np.zeros(ctx.U)] * KC, o_prop, 1e-10)
    assert abs(fa["seq_elbo"][0] - o_fa) < 2e-3 * abs(o_fa)


def test_hyperposterior_clust_and_pred_magmaclust(ctx):
    """hyperposterior_clust on data U grid (prediction.R:1685-1764), then pred_magmaclust for a new task
    (prediction.R:2298-2413): per-cluster conditionals and the mixture mean / variance."""
    from magmaclustr_b200.dataprep import reference_keys, r_signif
    case, mix, hp_k, o_hp_k = _setup(ctx, shared_i=True)
    prep, odb = case["prep"], case["odb"]
    rng = np.random.default_rng(5)
    Xp = np.round(np.linspace(-1, 1, 9), 2)[:, None]
    i0 = prep.ids[0]
    Xo, yo = odb.tasks[i0][1][:4], odb.tasks[i0][2][:4]
    ohyp = OC.hyperposterior_clust(odb, o_hp_k, case["o_hpi"], "SE", "SE", [0.0] * KC, mix, np.vstack([Xo, Xp]))
    prep.set_grid(np.vstack([Xo, Xp]))
    assert [k.decode() for k in prep.grid_keys] == ohyp["keys"]
    pm, pc = ctx.hyperposterior_clust(KC, case["k0"], hp_k, case["ki"], case["hpi"], prep.grid_X, prep.grid_index,
                                      np.zeros((KC, 1)), mix)
    for k in range(KC):
        assert relerr(pm[k], ohyp["mean"][k]) < 1e-4 and relerr(pc[k], ohyp["cov"][k]) < 1e-4
    proba = np.array([0.2, 0.5, 0.3])[:KC]
    proba = proba / proba.sum()
    hp_new = case["o_hpi"][i0]
    opred = OC.pred_magmaclust(Xo, yo, Xp, "SE", hp_new, ohyp, proba)
    Xo_r, Xp_r = r_signif(Xo), r_signif(Xp)
    ko, kp = reference_keys(Xo_r), reference_keys(Xp_r)
    oo, op = np.argsort(ko, kind="stable"), np.argsort(kp, kind="stable")
    io, ip = prep.index_of(ko[oo]), prep.index_of(kp[op])
    Hm, Hc = np.array(ohyp["mean"]), np.array(ohyp["cov"])
    mean, var, cov, mm, mv = ctx.pred_magmaclust(
        KC, case["ki"], np.array([hp_new[n] for n in case["names_i"]]), Xo_r[oo], yo[oo], Xp_r[op],
        Hm[:, io], Hm[:, ip], np.array([Hc[k][np.ix_(io, io)] for k in range(KC)]),
        np.array([Hc[k][np.ix_(ip, ip)] for k in range(KC)]), np.array([Hc[k][np.ix_(io, ip)] for k in range(KC)]), proba)
    for k in range(KC):
        np.testing.assert_allclose(mean[k], opred["pred"][k]["Mean"], rtol=1e-8, atol=1e-9)
        np.testing.assert_allclose(var[k], opred["pred"][k]["Var"], rtol=1e-7, atol=1e-10)
        assert relerr(cov[k], opred["pred"][k]["cov"]) < 1e-7
    np.testing.assert_allclose(mm, opred["Mean"], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(mv, opred["Var"], rtol=1e-7, atol=1e-10)
    assert np.all(mv >= 0)
    prep.set_grid(None)


def test_select_nb_cluster(ctx):
    """select_nb_cluster (training.R:2286-2454): one fast_approx training per K and the variational BIC; K = 1 is
    train_magma's E step + logL_monitoring, K > 1 the first VE step + elbo_monitoring_VEM.  Against the oracle's steps."""
    import math
    from magmaclustr_b200.selection import select_nb_cluster, vbic
    case, _, _, _ = _setup(ctx, M=12, N=10, seed=4, shared_i=True)
    prep, odb = case["prep"], case["odb"]
    rng = np.random.default_rng(3)
    mixes = {}
    for k in (2, 3):
        lab = rng.integers(0, k, size=12)
        lab[:k] = np.arange(k)                       # every cluster populated
        m = np.zeros((12, k))
        m[np.arange(12), lab] = 1.0
        mixes[k] = m
    hp_k1 = np.array([1.1, -1.4])
    res = select_nb_cluster(ctx, prep, case["k0"], case["ki"], hp_k1, case["hpi"], grid_nb_cluster=[1, 2, 3],
                            fast_approx=True, ini_mixtures=mixes)
    U = ctx.U
    o_hp = dict(zip(case["names_0"], hp_k1))
    # K = 1
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE", o_hp, case["o_hpi"], 1e-10)
    ll1 = OM.logL_monitoring(o_hp, case["o_hpi"], odb, 0.0, "SE", "SE", pm, pc, 1e-10)
    expect = {1: ll1}
    for k in (2, 3):
        hk = [dict(o_hp) for _ in range(k)]
        mean_k, cov_k, omix = OC.ve_step(odb, [np.zeros(U)] * k, "SE", "SE", hk, case["o_hpi"], mixes[k], 1, 1e-10, None)
        expect[k] = OC.elbo_monitoring_VEM(hk, case["o_hpi"], odb, "SE", "SE", mean_k, cov_k, omix, [np.zeros(U)] * k,
                                           mixes[k].mean(axis=0), 1e-10)
    for k in (1, 2, 3):
        assert abs(res["seq_elbo"][k] - expect[k]) < 2e-3 * abs(expect[k]), (k, res["seq_elbo"][k], expect[k])
        # shared hp_i (3 distinct values) + one hp_k row (2 distinct values) + K - 1 proportions
        pen = 0.5 * (2 + 3 + k - 1) * math.log(12)
        assert abs(res["seq_vbic"][k] - (res["seq_elbo"][k] - pen)) < 1e-9
    assert res["best_k"] == max((1, 2, 3), key=lambda k: expect[k] - 0.5 * (5 + k - 1) * math.log(12))
    assert res["model"] is not None and "V-BIC" in res["model"]
