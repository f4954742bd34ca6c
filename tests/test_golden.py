"""Committed golden fixtures (tests/golden/, made by tests/tools/make_golden.py in the build container).

Inputs are the reference's own golden datasets (tests/testthat/fixtures/simu_data_reference.rds) and its
fixed 5-point test dataset; outputs were produced by the oracle (R cannot run here, so for e_step, m_step,
logL_monitoring, pred_magma and the MagmaClust steps parity with R itself is unpinned).
  * not-gpu tests: the oracle still reproduces the frozen numbers and the index mapping;
  * gpu tests: the CUDA library reproduces them through the C ABI, with no access to /root/reference.
Tolerances: the union-grid index mapping, task order and keys are exact; well-conditioned per-task
quantities 1e-9 relative; quantities that pass through the mean process (cond(K_0) ~ 1e11-1e12 at these
HPs) carry the tolerance measured in test_gpu_magma.py (1e-4 relative)."""
import glob
import json
import os

import numpy as np
import pytest

from magmaclustr_b200.dataprep import Prepared
from oracle import kernels as OK
from oracle import magma as OM
from oracle import magmaclust as OC

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MAGMA = sorted(glob.glob(os.path.join(GOLD, "magma_*.json")))
TIGHT, LOOSE = 1e-9, 1e-4


def _load(path):
    return json.load(open(path))


def _hp_dicts(c):
    hp_0 = dict(zip(c["hp_names_0"], c["hp_0"]))
    hp_i = {i: dict(zip(c["hp_names_i"], v)) for i, v in c["hp_i"].items()}
    return hp_0, hp_i


def _close(a, b, rtol, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = max(float(np.max(np.abs(b))), 1e-300)
    err = float(np.max(np.abs(a - b))) / scale
    assert err <= rtol, "%s: relative error %.3e > %.1e" % (what, err, rtol)


def test_fixture_set_is_complete():
    assert len(MAGMA) == 5
    assert os.path.exists(os.path.join(GOLD, "magmaclust_so_notshared_hp.json"))
    assert os.path.exists(os.path.join(GOLD, "kat_five_point.json"))


def test_five_point_check_values():
    k = _load(os.path.join(GOLD, "kat_five_point.json"))
    X, y, hp = np.array(k["X"]), np.array(k["y"]), {"se_variance": 1.0, "se_lengthscale": 0.5}
    cov = OK.kern_to_cov(X, "SE", hp)
    s = k["survey_check_values"]
    assert OM.logL_GP_mod(hp, X, y, 0.0, "SE", cov, 0) == pytest.approx(s["logL_GP_mod"], rel=1e-13)
    assert OM.logL_GP(hp, X, y, 0.0, "SE", cov, 0) == pytest.approx(s["logL_GP"], rel=1e-13)
    np.testing.assert_allclose(OM.gr_GP_mod(hp, X, y, 0.0, "SE", cov, 0), s["gr_GP_mod"], rtol=1e-8)
    np.testing.assert_allclose(OM.gr_GP(hp, X, y, 0.0, "SE", cov, 0), s["gr_GP"], rtol=1e-8)
    assert k["logL_GP_mod"] == pytest.approx(s["logL_GP_mod"], rel=1e-13)


@pytest.mark.parametrize("path", MAGMA, ids=[os.path.basename(p)[6:-5] for p in MAGMA])
def test_oracle_reproduces_golden(path):
    c = _load(path)
    db = OM.Dataset(c["ids"], np.array(c["X"]), np.array(c["y"]))
    # index mapping: exact (and identical between the oracle's and the product's host data path)
    assert db.ids == c["task_order"] and db.grid_keys == c["grid_keys"]
    prep = Prepared(c["ids"], np.array(c["X"]), np.array(c["y"]))
    assert prep.ids == c["task_order"]
    assert [k.decode() for k in prep.grid_keys] == c["grid_keys"]
    for t, i in enumerate(prep.ids):
        assert prep.grid_index[prep.offs[t]:prep.offs[t + 1]].tolist() == c["grid_index"][i]
        assert db.index[i].tolist() == c["grid_index"][i]
    np.testing.assert_array_equal(prep.grid_X, np.array(c["grid_X"]))
    hp_0, hp_i = _hp_dicts(c)
    pm, pc = OM.e_step(db, 0.0, c["kern_0"], c["kern_i"], hp_0, hp_i, c["pen_diag"])
    _close(pm, c["post_mean"], LOOSE, "post_mean")
    _close(pc, c["post_cov"], LOOSE, "post_cov")
    gpm, gpc = np.array(c["post_mean"]), np.array(c["post_cov"])
    for t, i in enumerate(db.ids):
        idx = db.index[i]
        f = OM.logL_GP_mod(hp_i[i], db.tasks[i][1], db.tasks[i][2], gpm[idx], c["kern_i"], gpc[np.ix_(idx, idx)], c["pen_diag"])
        g = OM.gr_GP_mod(hp_i[i], db.tasks[i][1], db.tasks[i][2], gpm[idx], c["kern_i"], gpc[np.ix_(idx, idx)], c["pen_diag"])
        assert f == pytest.approx(c["logL_GP_mod_tasks"][t], rel=TIGHT)
        _close(g, c["gr_GP_mod_tasks"][t], 1e-8, "gr_GP_mod task %s" % i)
        _close(OK.kern_to_inv(db.tasks[i][1], c["kern_i"], hp_i[i], c["pen_diag"]), c["inv_i"][i], TIGHT, "inv_i")


# ---- GPU: the CUDA library against the frozen numbers -------------------------------------------------------
def _gpu_case(ctx, c):
    from magmaclustr_b200 import _lib as L
    prep = Prepared(c["ids"], np.array(c["X"]), np.array(c["y"]))
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    k0 = L.parse_kernel(c.get("kern_0", c.get("kern_k")), False)
    ki = L.parse_kernel(c["kern_i"], True)
    assert L.hp_names(ki) == c["hp_names_i"]
    hpi = np.array([c["hp_i"][i] for i in prep.ids])
    return prep, k0, ki, hpi


@pytest.mark.gpu
@pytest.mark.parametrize("path", MAGMA, ids=[os.path.basename(p)[6:-5] for p in MAGMA])
def test_gpu_reproduces_golden_magma(ctx, path):
    c = _load(path)
    prep, k0, ki, hpi = _gpu_case(ctx, c)
    U = ctx.U
    hp0 = np.array(c["hp_0"])
    # list_kern_to_inv: per-task inverses (well conditioned)
    mats, retries = ctx.list_kern_to_mat(ki, prep.offs, prep.X, hpi, inverse=True, pen_diag=c["pen_diag"])
    for t, i in enumerate(prep.ids):
        _close(mats[t], c["inv_i"][i], TIGHT, "inv_i %s" % i)
    assert retries.max() == 0
    # e_step
    pm, pc, info = ctx.e_step(k0, hp0, ki, hpi, np.zeros(U), pen_diag=c["pen_diag"])
    _close(pm, c["post_mean"], LOOSE, "post_mean")
    _close(pc, c["post_cov"], LOOSE, "post_cov")
    # logL_GP_mod / gr_GP_mod of every task from the frozen hyper-posterior
    f, g, _ = ctx.logL_GP_mod_tasks(ki, hpi, np.array(c["post_mean"]), np.array(c["post_cov"]), pen_diag=c["pen_diag"])
    np.testing.assert_allclose(f, c["logL_GP_mod_tasks"], rtol=TIGHT)
    _close(g, c["gr_GP_mod_tasks"], 1e-8, "gr_GP_mod tasks")
    # mean-process objective (ill conditioned)
    f0, g0 = ctx.logL_GP_mod(k0, hp0, prep.grid_X, np.array(c["post_mean"]), np.zeros(U), np.array(c["post_cov"]),
                             pen_diag=c["pen_diag"])
    assert f0 == pytest.approx(c["logL_GP_mod_hp0"], rel=LOOSE, abs=LOOSE)
    _close(g0, c["gr_GP_mod_hp0"], 1e-3, "gr_GP_mod hp_0")
    # m_step: per-task trajectories are exact (evaluation counts), HPs to 1e-6
    if c["m_step_ok"]:
        h0, hi, st = ctx.m_step(k0, hp0, ki, hpi, np.zeros(U), np.array(c["post_mean"]), np.array(c["post_cov"]),
                                pen_diag=c["pen_diag"])
        conv, nev = ctx.last_task_status()
        assert [int(v) for v in nev] == [c["m_step_task_evals"][i] for i in prep.ids]
        assert [int(v) for v in conv] == [c["m_step_task_convergence"][i] for i in prep.ids]
        np.testing.assert_allclose(hi, [c["m_step_hp_i"][i] for i in prep.ids], rtol=1e-6, atol=1e-7)
        ll = ctx.logL_monitoring(k0, np.array(c["m_step_hp_0"]), ki, np.array([c["m_step_hp_i"][i] for i in prep.ids]),
                                 np.zeros(U), np.array(c["post_mean"]), np.array(c["post_cov"]), pen_diag=c["pen_diag"])
        assert ll == pytest.approx(c["logL_monitoring"], rel=LOOSE)
    # hyperposterior on data U grid, then pred_magma for the first task
    p = c["pred"]
    prep.set_grid(np.vstack([np.array(p["X_obs"]), np.array(p["X_pred"])]))
    assert [k.decode() for k in prep.grid_keys] == p["hyper_keys"]            # index mapping: exact
    hm, hc = ctx.hyperposterior(k0, hp0, ki, hpi, prep.grid_X, prep.grid_index, 0.0, pen_diag=c["pen_diag"])
    _close(hm, p["hyper_mean"], LOOSE, "hyperposterior mean")
    _close(hc, p["hyper_cov"], LOOSE, "hyperposterior cov")
    from magmaclustr_b200.dataprep import reference_keys, r_signif
    Xo, Xp = r_signif(np.array(p["X_obs"])), r_signif(np.array(p["X_pred"]))
    ko, kp = reference_keys(Xo), reference_keys(Xp)
    oo, op = np.argsort(ko, kind="stable"), np.argsort(kp, kind="stable")
    assert [k.decode() for k in kp[op]] == p["pred_keys"]
    io, ip = prep.index_of(ko[oo]), prep.index_of(kp[op])
    Hm, Hc = np.array(p["hyper_mean"]), np.array(p["hyper_cov"])
    mean, var, cov = ctx.pred_magma(ki, np.array(c["hp_i"][p["task"]]), Xo[oo], np.array(p["y_obs"])[oo], Xp[op],
                                    Hm[io], Hm[ip], Hc[np.ix_(io, io)], Hc[np.ix_(ip, ip)], Hc[np.ix_(io, ip)],
                                    pen_diag=c["pen_diag"])
    _close(mean, p["Mean"], 1e-8, "pred Mean")
    _close(var, p["Var"], 1e-7, "pred Var")
    _close(cov, p["cov"], 1e-7, "pred cov")
    prep.set_grid(None)


@pytest.mark.gpu
def test_gpu_reproduces_golden_magmaclust(ctx):
    c = _load(os.path.join(GOLD, "magmaclust_so_notshared_hp.json"))
    prep, kk, ki, hpi = _gpu_case(ctx, c)
    K, U = c["K"], ctx.U
    hp_k, mix, prop = np.array(c["hp_k"]), np.array(c["mixture"]), np.array(c["prop_mixture"])
    pm, pc, tau1 = ctx.ve_step(K, kk, hp_k, ki, hpi, np.zeros((K, U)), mix, prop, update=False, pen_diag=c["pen_diag"])
    _close(pm, c["ve_mean"], LOOSE, "ve_step mean")
    _close(pc, c["ve_cov"], LOOSE, "ve_step cov")
    np.testing.assert_array_equal(tau1, mix)                                  # iteration 1 keeps the mixture
    gm, gc = np.array(c["ve_mean"]), np.array(c["ve_cov"])
    tau2 = ctx.update_mixture(K, ki, hpi, gm, gc, prop, pen_diag=c["pen_diag"])
    assert np.max(np.abs(tau2 - np.array(c["update_mixture"]))) <= 1e-5 + 1e-12      # round(5): one last-digit flip at most
    assert np.array_equal(np.argmax(tau2, axis=1), np.argmax(np.array(c["update_mixture"]), axis=1))   # assignments: exact
    f, g = ctx.elbo_clust_tasks(K, ki, hpi, gm, gc, mix, pen_diag=c["pen_diag"])
    np.testing.assert_allclose(f, c["elbo_tasks"], rtol=TIGHT)
    _close(g, c["gr_elbo_tasks"], 1e-8, "gr_clust_multi_GP")
    elbo = ctx.elbo_monitoring(K, kk, hp_k, ki, hpi, np.zeros((K, U)), gm, gc, mix, prop, pen_diag=c["pen_diag"])
    assert elbo == pytest.approx(c["elbo_monitoring"], rel=LOOSE)
