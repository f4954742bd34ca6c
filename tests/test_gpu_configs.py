"""Oracle parity for the BASELINE.json configurations that round 1 only ran:
  C1  hyperposterior on data U seq(0, 10, 0.01) (U ~ 1050, the multi-CTA large-matrix path) and pred_magma, G = 1001;
  C2' e_step + hp_0 objective on a 2000-point union grid (the stress variant of config 2), exact jitter retry counts;
  C3  train_magmaclust shapes at M = 20 000, K = 5: VE step at full size, 12 sampled tasks against the oracle;
  C4  one task of N = 4096 on the common grid against LAPACK;
  C5  hyperposterior_clust on a 2000-point sub-grid of the 2-D prediction grid against the oracle.
Mean-process quantities (cond(K_0) ~ 1e11-1e12) are compared with a tolerance MEASURED on the oracle: the deviation of
its own result under a one-ulp perturbation of the hyper-parameters (helpers.ulp_perturb), times 30 -- no implementation,
LAPACK included, can reproduce them more closely."""
import math

import numpy as np
import pytest

from helpers import relerr, ulp_perturb
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.dataprep import Prepared, reference_keys, r_signif
from magmaclustr_b200.simulate import simu_db, ini_hp
from oracle import kernels as OK
from oracle import magma as OM
from oracle import magmaclust as OC

pytestmark = pytest.mark.gpu


def _hp(kern, ids, shared, noise, seed):
    tab = ini_hp(kern, ids, shared, noise, seed)
    names = [c for c in tab.columns if c != "ID"]
    return names, tab[names].to_numpy(dtype=np.float64)


def test_c1_hyperposterior_and_prediction_on_seq_0_10(ctx):
    """README example shape (BASELINE config 1): 10 training tasks x 10 points on [0, 10], one new task with 7 observations,
    prediction grid seq(0, 10, 0.01) -> G = 1001 (prediction.R:583-594: below the 2000 / 5000 guards)."""
    pool = np.round(np.linspace(0.0, 10.0, 401), 3)                     # multiples of 0.025: half of them off the 0.01 grid
    df, _ = simu_db(M=11, N=10, seed=17, pool=pool)
    ids = sorted(df.ID.unique(), key=lambda s: s.encode())
    new_id = ids[-1]
    train = df[df.ID != new_id]
    new = df[df.ID == new_id].iloc[:7]
    prep = Prepared(train.ID.values, train.Input.values, train.Output.values)
    odb = OM.Dataset(train.ID.values, train.Input.values, train.Output.values)
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    n0, hp0 = _hp("SE", ["0"], True, False, 18)
    ni, hpi = _hp("SE", prep.ids, False, True, 19)
    hp0 = hp0[0]
    o_hp0 = dict(zip(n0, hp0))
    o_hpi = {i: dict(zip(ni, hpi[t])) for t, i in enumerate(prep.ids)}
    grid = np.round(np.arange(0.0, 10.0 + 1e-9, 0.01), 2)[:, None]
    assert len(grid) == 1001
    extra = np.vstack([new.Input.values[:, None], grid])
    ohyp = OM.hyperposterior(odb, o_hp0, o_hpi, "SE", "SE", None, extra)
    prep.set_grid(extra)
    assert [k.decode() for k in prep.grid_keys] == ohyp["keys"]                     # union grid and its order: exact
    U = len(prep.grid_keys)
    assert 1001 < U <= 1001 + 107
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    pm, pc = ctx.hyperposterior(k0, hp0, ki, hpi, prep.grid_X, prep.grid_index, 0.0)
    assert np.array_equal(pc, pc.T)
    rng = np.random.default_rng(1)
    sens_m = sens_c = 0.0
    for _ in range(2):
        h = OM.hyperposterior(odb, dict(zip(n0, ulp_perturb(hp0, rng))), o_hpi, "SE", "SE", None, extra)
        sens_m, sens_c = max(sens_m, relerr(h["mean"], ohyp["mean"])), max(sens_c, relerr(h["cov"], ohyp["cov"]))
    print("C1 hyperposterior U=%d: mean %.2e (oracle sensitivity %.2e), cov %.2e (%.2e)"
          % (U, relerr(pm, ohyp["mean"]), sens_m, relerr(pc, ohyp["cov"]), sens_c))
    assert relerr(pm, ohyp["mean"]) < 30 * sens_m + 1e-9
    assert relerr(pc, ohyp["cov"]) < 30 * sens_c + 1e-9
    # pred_magma of the new task on the 1001-point grid, both fed the ORACLE's hyper-posterior: 1e-8 / 1e-7
    n_new, hp_new = _hp("SE", [new_id], False, True, 20)
    hp_new = hp_new[0]
    o = OM.pred_magma(new.Input.values, new.Output.values, grid, "SE", dict(zip(n_new, hp_new)), ohyp)
    Xo, Xp = r_signif(new.Input.values[:, None]), r_signif(grid)
    ko, kp = reference_keys(Xo), reference_keys(Xp)
    oo, op = np.argsort(ko, kind="stable"), np.argsort(kp, kind="stable")
    assert [k.decode() for k in kp[op]] == o["keys"]
    io, ip = prep.index_of(ko[oo]), prep.index_of(kp[op])
    Hm, Hc = ohyp["mean"], ohyp["cov"]
    mean, var, cov = ctx.pred_magma(ki, hp_new, Xo[oo], new.Output.values[oo], Xp[op], Hm[io], Hm[ip], Hc[np.ix_(io, io)],
                                    Hc[np.ix_(ip, ip)], Hc[np.ix_(io, ip)])
    np.testing.assert_allclose(mean, o["Mean"], rtol=1e-8, atol=1e-8)
    np.testing.assert_allclose(var, o["Var"], rtol=1e-7, atol=1e-9)
    assert relerr(cov, o["cov"]) < 1e-7
    # and the whole chain on the library's own hyper-posterior: within the measured sensitivity of the hyper-posterior
    mean2, var2, _ = ctx.pred_magma(ki, hp_new, Xo[oo], new.Output.values[oo], Xp[op], pm[io], pm[ip], pc[np.ix_(io, io)],
                                    pc[np.ix_(ip, ip)], pc[np.ix_(io, ip)])
    assert relerr(mean2, o["Mean"]) < 30 * sens_m + 1e-8
    assert relerr(var2, o["Var"]) < 30 * sens_c + 1e-7


def test_c2_stress_union_grid_2000(ctx):
    """Config 2's stress variant: a 2000-point pool (the largest simu_data allows without warning_grid = FALSE,
    simulate-data.R:746) -> U = 2000 mean-process matrices on the multi-CTA path: e_step and the hp_0 objective against the
    oracle, jitter retry counts exact."""
    pool = np.round(np.linspace(-1.0, 1.0, 2000), 6)
    M, N = 400, 30
    df, _ = simu_db(M=M, N=N, seed=7, pool=pool)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    odb = OM.Dataset(df.ID.values, df.Input.values, df.Output.values)
    U = len(prep.grid_keys)
    assert U > 1990
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    n0, hp0 = _hp("SE", ["0"], True, False, 8)
    ni, hpi = _hp("SE", prep.ids, False, True, 9)
    hp0 = hp0[0]
    o_hp0 = dict(zip(n0, hp0))
    o_hpi = {i: dict(zip(ni, hpi[t])) for t, i in enumerate(prep.ids)}
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    pm, pc, info = ctx.e_step(k0, hp0, ki, hpi, np.zeros(U))
    opm, opc, oinfo = OM.e_step(odb, 0.0, "SE", "SE", o_hp0, o_hpi, 1e-10, return_info=True)
    print("U=2000 retries gpu", info, "oracle", oinfo)
    assert int(info[0]) == int(oinfo["inv_0"]) and int(info[1]) == int(oinfo["post"])   # discrete decisions: exact
    assert int(info[2]) == max(oinfo["tasks"].values())
    rng = np.random.default_rng(2)
    h = OM.e_step(odb, 0.0, "SE", "SE", dict(zip(n0, ulp_perturb(hp0, rng))), o_hpi, 1e-10)
    sens_m, sens_c = relerr(h[0], opm), relerr(h[1], opc)
    print("U=%d e_step: mean %.2e (sens %.2e) cov %.2e (sens %.2e)" % (U, relerr(pm, opm), sens_m, relerr(pc, opc), sens_c))
    assert relerr(pm, opm) < 30 * sens_m + 1e-9 and relerr(pc, opc) < 30 * sens_c + 1e-9
    assert np.array_equal(pc, pc.T)
    # hp_0 objective + gradient on the oracle's hyper-posterior
    f0, g0 = ctx.logL_GP_mod(k0, hp0, prep.grid_X, opm, np.zeros(U), opc)
    of = OM.logL_GP_mod(o_hp0, odb.grid_X, opm, np.zeros(U), "SE", opc, 1e-10)
    og = OM.gr_GP_mod(o_hp0, odb.grid_X, opm, np.zeros(U), "SE", opc, 1e-10)
    sf = sg = 0.0
    for _ in range(2):
        hq = dict(zip(n0, ulp_perturb(hp0, rng)))
        sf = max(sf, abs(OM.logL_GP_mod(hq, odb.grid_X, opm, np.zeros(U), "SE", opc, 1e-10) - of))
        sg = max(sg, float(np.max(np.abs(OM.gr_GP_mod(hq, odb.grid_X, opm, np.zeros(U), "SE", opc, 1e-10) - og))))
    print("hp_0 objective U=%d: |df| %.3e (sens %.3e), |dg| %.3e (sens %.3e)" % (U, abs(f0 - of), sf, float(np.max(np.abs(g0 - og))), sg))
    assert abs(f0 - of) < 30 * sf + 1e-9 * abs(of)
    assert float(np.max(np.abs(g0 - og))) < 30 * sg + 1e-8 * float(np.max(np.abs(og)))


def test_c3_magmaclust_full_size_sampled_tasks(ctx):
    """Config 3 (M = 20 000 x N = 32, K = 5, shared_hp_i = TRUE): the VE step and the mixture update at full size; 12
    sampled tasks against the oracle -- their likelihood terms to 1e-9, responsibilities to one unit of round(5), cluster
    assignments exactly (vem-magmaclust.R:330-396)."""
    M, N, KC = 20000, 32, 5
    df, gen = simu_db(M=M, N=N, K=KC, seed=4)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    U = len(prep.grid_keys)
    assert U == 201
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    ni, hpi = _hp("SE", prep.ids, True, True, 5)                         # shared_hp_i: one row for all tasks
    hp_k = np.array([[1.0 + 0.2 * k, -1.5 - 0.1 * k] for k in range(KC)])
    rng = np.random.default_rng(4)
    lab = rng.integers(0, KC, size=M)
    mix = np.zeros((M, KC))
    mix[np.arange(M), lab] = 1.0
    prop = mix.mean(axis=0)
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    pm, pc, tau = ctx.ve_step(KC, k0, hp_k, ki, hpi, np.zeros((KC, U)), mix, prop_mixture=prop, update=True)
    assert np.all(np.isfinite(pm)) and np.all(np.isfinite(pc)) and tau.shape == (M, KC)
    for k in range(KC):
        assert np.array_equal(pc[k], pc[k].T)
    assert np.all(np.abs(tau.sum(axis=1) - 1.0) < 1e-4)
    # sampled tasks: -logL_GP_mod per cluster from the library's hyper-posterior, softmax, round(5)
    names = ni
    for t in sorted(rng.choice(M, size=12, replace=False)):
        sl = slice(prep.offs[t], prep.offs[t + 1])
        X, y, idx = prep.X[sl], prep.y[sl], prep.grid_index[sl]
        hpd = dict(zip(names, hpi[t]))
        Lk = np.array([-OM.logL_GP_mod(hpd, X, y, pm[k][idx], "SE", pc[k][np.ix_(idx, idx)], 1e-10) for k in range(KC)])
        w = prop * np.exp(Lk - Lk.max())
        ot = np.array([OC.r_round(v, 5) for v in w / w.sum()])
        assert int(np.argmax(tau[t])) == int(np.argmax(ot))                          # cluster assignment: exact
        assert np.max(np.abs(tau[t] - ot)) <= 1e-5 + 1e-12, (t, tau[t], ot)
    # ELBO value + gradient of the sampled tasks (elbos.R:18-55, gradients-elbos.R:17-90)
    f, g = ctx.elbo_clust_tasks(KC, ki, hpi, pm, pc, tau)
    odb = None
    for t in sorted(rng.choice(M, size=6, replace=False)):
        sl = slice(prep.offs[t], prep.offs[t + 1])
        X, y, idx = prep.X[sl], prep.y[sl], prep.grid_index[sl]
        hpd = dict(zip(names, hpi[t]))
        inv = OK.kern_to_inv(X, "SE", hpd, 1e-10)
        c1 = sum(tau[t, k] * pm[k][idx] for k in range(KC))
        c2 = sum(tau[t, k] * (np.outer(pm[k][idx], pm[k][idx]) + pc[k][np.ix_(idx, idx)]) for k in range(KC))
        of = OM.neg_dmnorm_log(y, np.zeros(len(y)), inv) - float(y @ inv @ c1) + 0.5 * float(np.sum(inv * c2))
        assert abs(f[t] - of) <= 1e-9 * max(1.0, abs(of)), (t, f[t], of)


def test_c4_one_task_of_4096_points(ctx):
    """Config 4 (common grid, N = U = 4096): chol_inv_jitter of ONE task's covariance on the large-matrix path against
    LAPACK (dpotrf / dpotri, what chol2inv(chol(.)) runs in R), and logL_GP_mod + gr_GP_mod of that task."""
    n = 4096
    x = np.linspace(-1.0, 1.0, n)
    hp = {"se_variance": 0.3, "se_lengthscale": -5.5, "noise": -2.0}
    K = OK.kern_to_cov(x[:, None], "SE", hp)
    inv, retries, jit = ctx.chol_inv_jitter(K, 1e-10)
    ref, r_ref = OK.chol_inv_jitter(K, 1e-10, return_info=True)
    assert retries == r_ref == 0 and jit == 1e-10
    assert np.array_equal(inv, inv.T)
    assert np.max(np.abs(inv - ref)) <= 1e-10 * np.max(np.abs(ref))
    ki = L.parse_kernel("SE", True)
    rng = np.random.default_rng(3)
    y = rng.normal(size=n)
    mean = 0.1 * x
    Sd = np.exp(-0.5 * (x[:, None] - x[None, :]) ** 2 / 0.05) * 0.2 + 1e-3 * np.eye(n)     # a smooth "post_cov"
    f, g = ctx.logL_GP_mod(ki, np.array([hp["se_variance"], hp["se_lengthscale"], hp["noise"]]), x[:, None], y, mean, Sd)
    of = OM.logL_GP_mod(hp, x[:, None], y, mean, "SE", Sd, 1e-10)
    og = OM.gr_GP_mod(hp, x[:, None], y, mean, "SE", Sd, 1e-10)
    assert abs(f - of) <= 1e-9 * abs(of), (f, of)
    np.testing.assert_allclose(g, og, rtol=1e-8, atol=1e-8 * float(np.max(np.abs(og))))


def test_c5_hyperposterior_clust_on_2d_subgrid(ctx):
    """Config 5's hyper-posterior (2-D covariate grid, kernel RQ + PERIO for the tasks, K = 3) re-evaluated on a 2000-point
    sub-grid of the 100 x 100 prediction grid (the oracle needs U^3 per cluster) -- prediction.R:1691-1747."""
    KC, T, NOBS = 3, 40, 20
    rng = np.random.default_rng(5)
    ax = np.linspace(0, 10, 100)
    grid = np.array([[a, b] for a in ax for b in ax])
    sub = grid[rng.choice(len(grid), size=2000, replace=False)]
    ids, Xs, ys = [], [], []
    for t in range(T):
        pts = grid[rng.choice(len(grid), size=NOBS, replace=False)]
        ids += ["t%03d" % t] * NOBS
        Xs.append(pts)
        ys.append(np.sin(pts[:, 0]) + 0.3 * pts[:, 1] + 0.1 * rng.normal(size=NOBS) + (t % KC))
    X, y = np.vstack(Xs), np.concatenate(ys)
    prep = Prepared(ids, X, y)
    odb = OM.Dataset(ids, X, y)
    kern_i = "RQ + PERIO"
    kk, ki = L.parse_kernel("SE", False), L.parse_kernel(kern_i, True)
    names_i = L.hp_names(ki)
    base = {"rq_variance": 0.3, "rq_lengthscale": 0.5, "rq_scale": 1.5, "perio_variance": -0.5, "perio_lengthscale": 0.4,
            "period": 1.2, "noise": -2.0}
    hpi = np.array([[base[nm] + 0.01 * ((t * 7 + q) % 5) for q, nm in enumerate(names_i)] for t in range(T)])
    o_hpi = {i: dict(zip(names_i, hpi[t])) for t, i in enumerate(prep.ids)}
    hp_k = np.array([[1.0 + 0.2 * k, 0.8 - 0.1 * k] for k in range(KC)])
    o_hp_k = [dict(zip(L.hp_names(kk), hp_k[k])) for k in range(KC)]
    mix = rng.dirichlet(np.ones(KC), size=T).round(5)
    ohyp = OC.hyperposterior_clust(odb, o_hp_k, o_hpi, "SE", kern_i, [0.0] * KC, mix, sub)
    prep.set_grid(sub)
    U = len(prep.grid_keys)
    assert [k.decode() for k in prep.grid_keys] == ohyp["keys"] and U >= 2000
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    if U > 2000:
        with pytest.warns(UserWarning, match="hyper-posterior grid"):      # R/prediction.R:590-593
            pm, pc = ctx.hyperposterior_clust(KC, kk, hp_k, ki, hpi, prep.grid_X, prep.grid_index, np.zeros((KC, U)), mix)
    else:
        pm, pc = ctx.hyperposterior_clust(KC, kk, hp_k, ki, hpi, prep.grid_X, prep.grid_index, np.zeros((KC, U)), mix)
    with pytest.raises(ValueError, match="5000 points"):                   # R/prediction.R:584-589, before any GPU work
        ctx.hyperposterior_clust(KC, kk, hp_k, ki, hpi, np.zeros((5001, 2)), np.zeros(1, dtype=np.int32), np.zeros((KC, 5001)), mix)
    rng2 = np.random.default_rng(6)
    pert = [dict(zip(L.hp_names(kk), ulp_perturb(hp_k[k], rng2))) for k in range(KC)]
    h2 = OC.hyperposterior_clust(odb, pert, o_hpi, "SE", kern_i, [0.0] * KC, mix, sub)
    for k in range(KC):
        sm, sc = relerr(h2["mean"][k], ohyp["mean"][k]), relerr(h2["cov"][k], ohyp["cov"][k])
        print("C5 sub-grid U=%d cluster %d: mean %.2e (sens %.2e) cov %.2e (sens %.2e)"
              % (U, k, relerr(pm[k], ohyp["mean"][k]), sm, relerr(pc[k], ohyp["cov"][k]), sc))
        assert relerr(pm[k], ohyp["mean"][k]) < 30 * sm + 1e-9
        assert relerr(pc[k], ohyp["cov"][k]) < 30 * sc + 1e-9
        assert np.array_equal(pc[k], pc[k].T)
