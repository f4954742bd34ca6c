"""Parity at BASELINE.json's FULL sizes: the CUDA path runs the whole configuration, the oracle (which needs seconds per
task batch) checks a seeded sample of tasks / grid points of the very same run."""
import numpy as np
import pytest

from magmaclustr_b200 import _lib as L
from magmaclustr_b200.dataprep import Prepared, reference_keys, r_signif
from magmaclustr_b200.simulate import simu_db, ini_hp
from oracle import magma as OM
from oracle import magmaclust as OC

pytestmark = pytest.mark.gpu


def test_c2_full_size_sampled_tasks(ctx):
    """Config 2 (M = 10 000 x N = 64, SE, shared_hp = FALSE): E step, all-task objective / gradient and the lock-step M step
    at full size; 12 sampled tasks against the oracle -- inverse-based quantities to 1e-9 / 1e-8, optimiser trajectories
    (evaluation counts, convergence codes) exactly, HPs to 1e-6."""
    M, N = 10000, 64
    df, _ = simu_db(M=M, N=N, seed=2024)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    names = L.hp_names(ki)
    hp0 = ini_hp("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy(dtype=np.float64)
    hpi = ini_hp("SE", prep.ids, False, True, 107).drop(columns="ID").to_numpy(dtype=np.float64)   # one HP row per task
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    U = ctx.U
    assert U == 201 and ctx.n_tasks == M
    pm, pc, info = ctx.e_step(k0, hp0, ki, hpi, np.zeros(U))
    assert np.array_equal(pc, pc.T) and np.all(np.isfinite(pm))
    f, g, r = ctx.logL_GP_mod_tasks(ki, hpi, pm, pc)
    assert np.all(np.isfinite(f)) and np.all(np.isfinite(g)) and r.max() == 0
    h0, hi, st = ctx.m_step(k0, hp0, ki, hpi, np.zeros(U), pm, pc)
    conv, nev = ctx.last_task_status()
    assert st[2] == nev.sum() and np.all(np.isfinite(hi))
    assert st[3] == np.sum(conv == 99) and st[3] <= 5      # trial points with a non-finite objective (R would raise): reported
    rng = np.random.default_rng(0)
    for t in sorted(rng.choice(M, size=12, replace=False)):
        sl = slice(prep.offs[t], prep.offs[t + 1])
        X, y, idx = prep.X[sl], prep.y[sl], prep.grid_index[sl]
        hpd = dict(zip(names, hpi[t]))
        S = pc[np.ix_(idx, idx)]
        of = OM.logL_GP_mod(hpd, X, y, pm[idx], "SE", S, 1e-10)
        og = OM.gr_GP_mod(hpd, X, y, pm[idx], "SE", S, 1e-10)
        assert abs(f[t] - of) <= 1e-9 * max(1.0, abs(of)), (t, f[t], of)
        np.testing.assert_allclose(g[t], og, rtol=1e-8, atol=1e-8)
        ohp, res = OM._optim(hpd, lambda h: OM.logL_GP_mod(h, X, y, pm[idx], "SE", S, 1e-10),
                             lambda h: OM.gr_GP_mod(h, X, y, pm[idx], "SE", S, 1e-10))
        assert nev[t] == res["counts"] and conv[t] == res["convergence"], (t, nev[t], res["counts"])
        np.testing.assert_allclose(hi[t], [ohp[n] for n in names], rtol=1e-6, atol=1e-7)


def test_c5_full_size_sampled_predictions(ctx):
    """Config 5 (2-D 100 x 100 grid, kernel RQ + PERIO, K = 3, 1000 new tasks): hyperposterior_clust on the 10 000-point grid
    stays on the device, every new task is predicted on every grid point in one call; three tasks x 60 grid points against
    the oracle's pred_magmaclust fed with the library's own hyper-posterior."""
    KC, G1, T, NOBS = 3, 100, 1000, 20
    rng = np.random.default_rng(5)
    ax = np.linspace(0, 10, G1)
    grid = np.array([[a, b] for a in ax for b in ax])
    G = len(grid)
    kern = "RQ + PERIO"

    def make_tasks(m, n, prefix, clusters):
        ids, X, y = [], [], []
        for t in range(m):
            x = grid[rng.choice(G, size=n, replace=False)]
            c = clusters[t]
            ids += ["%s%05d" % (prefix, t)] * n
            X.append(x)
            y.append((2.0 + c) * np.sin(0.6 * x[:, 0] + c) + 0.5 * (1 + c) * np.cos(0.9 * x[:, 1]) + 3.0 * c
                     + 0.1 * rng.standard_normal(n))
        return Prepared(np.array(ids), np.vstack(X), np.concatenate(y))

    cl_tr = rng.integers(0, KC, size=100)
    train = make_tasks(100, 30, "tr", cl_tr)
    train.set_grid(grid)
    assert train.grid_X.shape[0] == G
    mix = np.zeros((train.n_tasks, KC))
    mix[np.arange(train.n_tasks), cl_tr] = 1.0
    kk, ki = L.parse_kernel(kern, False), L.parse_kernel(kern, True)
    names_k, names_i = L.hp_names(kk), L.hp_names(ki)
    hp_k = ini_hp(kern, [str(k) for k in range(KC)], False, False, 50)[names_k].to_numpy(dtype=np.float64).copy()
    hp_i_row = ini_hp(kern, ["x"], True, True, 51)[names_i].to_numpy(dtype=np.float64)[0]
    ctx.upload_dataset(train.offs, train.X, train.y, train.grid_index, train.grid_X)
    ctx.hyperpost_keep(True)
    pm, pc = ctx.hyperposterior_clust(KC, kk, hp_k, ki, np.tile(hp_i_row, (train.n_tasks, 1)), train.grid_X,
                                      train.grid_index, np.zeros((KC, 1)), mix, grid_guard=False)   # 10 000 points
    ctx.hyperpost_keep(False)
    assert np.all(np.isfinite(pm)) and all(np.array_equal(pc[k], pc[k].T) for k in range(KC))
    new = make_tasks(T, NOBS, "nw", rng.integers(0, KC, size=T))
    ctx.upload_dataset(new.offs, new.X, new.y, train.index_of(new.keys), train.grid_X)
    hp_new = np.tile(hp_i_row, (T, 1))
    tau = rng.dirichlet(np.ones(KC), size=T)
    gp = r_signif(grid)
    kp = reference_keys(gp)
    op = np.argsort(kp, kind="stable")
    mean, var, _, _, r = ctx.pred_magma_tasks(ki, hp_new, gp[op], train.index_of(kp[op]), tau=tau, K=KC)
    ctx.hyperpost_keep(-1)
    assert mean.shape == (T, G) and np.all(np.isfinite(mean)) and np.all(var > 0) and r.max() == 0
    hyp = {"keys": [k.decode() for k in train.grid_keys], "mean": list(pm), "cov": list(pc)}
    sub = np.sort(rng.choice(G, size=60, replace=False))
    for t in (0, T // 2, T - 1):
        sl = slice(new.offs[t], new.offs[t + 1])
        o = OC.pred_magmaclust(new.X[sl], new.y[sl], gp[op][sub], kern, dict(zip(names_i, hp_new[t])), hyp, tau[t])
        oo = np.argsort(reference_keys(gp[op][sub]), kind="stable")          # the oracle answers in Reference order
        np.testing.assert_allclose(mean[t][sub][oo], o["Mean"], rtol=1e-8, atol=1e-8)
        np.testing.assert_allclose(var[t][sub][oo], o["Var"], rtol=1e-7, atol=1e-9)
