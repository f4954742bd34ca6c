"""GPU parity: CUDA library (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerances (FP64, relative unless stated) are written next to each assertion.  Integer /
index results (union-grid mapping, jitter retry counts, EM step counts) must be identical.
"""
import numpy as np
import pytest

from helpers import make_case, upload, relerr
from magmaclustr_b200 import _lib as L
from oracle import kernels as OK
from oracle import magma as OM

pytestmark = pytest.mark.gpu


def _hp_dict(k, hp):
    return dict(zip(L.hp_names(k), [float(v) for v in hp]))


# ---- the five native builders (src/code.cpp) ----------------------------------------------
@pytest.mark.parametrize("d", [1, 2, 3])
def test_legacy_builders(ctx, d):
    rng = np.random.default_rng(3 + d)
    m1 = rng.normal(size=(7, d))
    m2 = rng.normal(size=(5, d))
    m2[2] = m1[3]  # one coincident pair for cpp_noise
    np.testing.assert_allclose(ctx.cpp_dist(m1, m2), OK.cpp_dist(m1, m2), rtol=1e-15, atol=0)
    np.testing.assert_allclose(ctx.cpp_prod(m1, m2), OK.cpp_prod(m1, m2), rtol=1e-15, atol=1e-16)
    np.testing.assert_allclose(ctx.cpp_perio(m1, m2, 0.3), OK.cpp_perio(m1, m2, 0.3), rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(ctx.cpp_perio_deriv(m1, m2, 0.3), OK.cpp_perio_deriv(m1, m2, 0.3),
                               rtol=1e-12, atol=1e-14)
    got = ctx.cpp_noise(m1, m2, -1.5)
    np.testing.assert_allclose(got, OK.cpp_noise(m1, m2, -1.5), rtol=1e-15)
    assert got[3, 2] > 0 and np.count_nonzero(got) == 1


def test_legacy_builder_dimension_mismatch(ctx):
    with pytest.raises(RuntimeError):
        ctx.cpp_dist(np.zeros((3, 2)), np.zeros((3, 1)))


# ---- kernels at zero distance: test-kernels.R:1-23 -------------------------------------------
@pytest.mark.parametrize("kern,hp", [("SE", [1, 1]), ("PERIO", [1, 1, 1]), ("RQ", [1, 1, 1])])
def test_kernel_zero_distance(ctx, kern, hp):
    k = L.parse_kernel(kern, False)
    got = ctx.kern_to_cov(k, hp, np.array([2.0]))
    assert abs(got[0, 0] - np.e) < 1e-15 * 4


# ---- kern_to_cov incl. compound kernels and derivatives: test-kern_to_cov.R:118-185 -----------
@pytest.mark.parametrize("kern", ["SE", "PERIO", "RQ", "LIN", "SE + RQ + PERIO", "SE * RQ * PERIO",
                                  "SE * RQ + PERIO", "SE * LIN + RQ", "RQ + PERIO"])
@pytest.mark.parametrize("noise", [False, True])
def test_kern_to_cov_and_derivatives(ctx, kern, noise):
    rng = np.random.default_rng(11)
    x = np.round(rng.uniform(-1, 1, size=(9, 2)), 2)
    x2 = np.round(rng.uniform(-1, 1, size=(4, 2)), 2)
    k = L.parse_kernel(kern, noise)
    names = L.hp_names(k)
    hp = rng.uniform(0.2, 1.5, size=len(names))
    hpd = dict(zip(names, hp))
    np.testing.assert_allclose(ctx.kern_to_cov(k, hp, x), OK.kern_to_cov(x, kern, hpd), rtol=2e-14, atol=1e-15)
    np.testing.assert_allclose(ctx.kern_to_cov(k, hp, x, x2), OK.kern_to_cov(x, kern, hpd, input_2=x2),
                               rtol=2e-14, atol=1e-15)
    for p, name in enumerate(names):
        np.testing.assert_allclose(ctx.kern_to_cov(k, hp, x, deriv=p), OK.kern_to_cov(x, kern, hpd, deriv=name),
                                   rtol=5e-13, atol=1e-14, err_msg=name)


# ---- kern_to_inv = solve(cov): test-kern_to_inv.R:1-36 ------------------------------------------
def test_kern_to_inv_matches_solve(ctx):
    k = L.parse_kernel("SE", False)
    x = np.array([[2.0], [3.0], [4.0]])
    mats, r = ctx.list_kern_to_mat(k, [0, 3], x, [[2.0, 1.0]], inverse=True, pen_diag=1e-10)
    cov = OK.kern_to_cov(x, "SE", {"se_variance": 2.0, "se_lengthscale": 1.0})
    np.testing.assert_allclose(mats[0], np.linalg.inv(cov), atol=1.5e-8)   # testthat tolerance
    np.testing.assert_allclose(mats[0], OK.kern_to_inv(x, "SE", {"se_variance": 2.0, "se_lengthscale": 1.0}),
                               rtol=1e-12)
    assert r[0] == 0


def test_chol_inv_jitter_cumulative(ctx):
    # matrix(1,3,3), pen 1e-20 succeeds after 5 retries, shift 1.11111e-15 (Q1)
    inv, retries, jitter = ctx.chol_inv_jitter(np.ones((3, 3)), 1e-20)
    oinv, oretries = OK.chol_inv_jitter(np.ones((3, 3)), 1e-20, return_info=True)
    assert retries == oretries == 5
    assert abs(jitter - 1.11111e-15) < 1e-20
    # (the inverse itself is noise at cond ~ 3e15: only the discrete decisions are compared)
    assert np.all(np.isfinite(inv)) and oinv.shape == inv.shape
    rng = np.random.default_rng(0)
    a = rng.normal(size=(40, 40))
    spd = a @ a.T + 40 * np.eye(40)
    inv, retries, _ = ctx.chol_inv_jitter(spd, 1e-10)
    np.testing.assert_allclose(inv, OK.chol_inv_jitter(spd, 1e-10), rtol=1e-11)
    assert retries == 0 and np.array_equal(inv, inv.T)


def test_list_kern_to_inv_ragged(ctx):
    rng = np.random.default_rng(5)
    sizes = [1, 2, 7, 16, 33, 64, 5]
    offs = np.r_[0, np.cumsum(sizes)]
    X = np.concatenate([np.sort(rng.choice(np.round(np.arange(-1, 1.001, 0.01), 2), n, replace=False)) for n in sizes])[:, None]
    k = L.parse_kernel("SE", True)
    hp = np.column_stack([rng.uniform(-0.7, 2.3, len(sizes)), rng.uniform(-3, 0, len(sizes)), rng.uniform(-5, -2, len(sizes))])
    mats, r = ctx.list_kern_to_mat(k, offs, X, hp, inverse=True)
    covs, _ = ctx.list_kern_to_mat(k, offs, X, hp, inverse=False)
    for t in range(len(sizes)):
        xt = X[offs[t]:offs[t + 1]]
        hpd = _hp_dict(k, hp[t])
        oinv, oretry = OK.chol_inv_jitter(OK.kern_to_cov(xt, "SE", hpd), 1e-10, return_info=True)
        assert r[t] == oretry
        np.testing.assert_allclose(covs[t], OK.kern_to_cov(xt, "SE", hpd), rtol=1e-14)
        # conditioning of these matrices is <= 1e5: 1e-9 leaves four digits of margin
        assert relerr(mats[t], oinv) < 1e-9


# ---- logL_GP_mod / gr_GP_mod check values (test-gradients-logL.R:32-58) ------------
def test_logL_GP_mod_check_values(ctx):
    X = np.column_stack([np.arange(1, 6), np.arange(3, 8)]).astype(float)
    y = np.arange(2, 7).astype(float)
    k = L.parse_kernel("SE", False)
    hp = [1.0, 0.5]
    cov = OK.kern_to_cov(X, "SE", {"se_variance": 1.0, "se_lengthscale": 0.5})
    f, g = ctx.logL_GP_mod(k, hp, X, y, np.zeros(5), cov, pen_diag=0.0)
    assert abs(f - 18.13284374363541) < 1e-11
    np.testing.assert_allclose(g, [-9.4131031, -2.30758734], rtol=1e-8)
    # the reference's own assertion: analytic gradient == finite difference to 3 decimals
    eps = 1e-6
    for p in range(2):
        hp2 = list(hp)
        hp2[p] += eps
        f2, _ = ctx.logL_GP_mod(k, hp2, X, y, np.zeros(5), cov, pen_diag=0.0, grad=False)
        assert round((f2 - f) / eps, 3) == round(g[p], 3)


@pytest.mark.parametrize("M,N,seed", [(10, 10, 17), (6, 33, 3), (5, 64, 9)])
def test_task_objective_and_gradient(ctx, M, N, seed):
    case = make_case(M, N, seed)
    upload(ctx, case)
    odb = case["odb"]
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
    f, g, r = ctx.logL_GP_mod_tasks(case["ki"], case["hpi"], pm, pc)
    for t, i in enumerate(case["prep"].ids):
        Xi, yi, mi, Si = OM._task_views(odb, pm, pc, i)
        of = OM.logL_GP_mod(case["o_hpi"][i], Xi, yi, mi, "SE", Si, 1e-10)
        og = OM.gr_GP_mod(case["o_hpi"][i], Xi, yi, mi, "SE", Si, 1e-10)
        assert abs(f[t] - of) <= 1e-9 * abs(of)
        np.testing.assert_allclose(g[t], og, rtol=1e-8, atol=1e-9 * np.max(np.abs(og)))
        assert r[t] == 0


# ---- e_step --------------------------------------------------------------------------------------
@pytest.mark.parametrize("M,N,seed", [(10, 10, 17), (40, 20, 5)])
def test_e_step(ctx, M, N, seed):
    case = make_case(M, N, seed)
    upload(ctx, case)
    pm, pc, info = ctx.e_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(ctx.U))
    opm, opc, oinfo = OM.e_step(case["odb"], 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10,
                                return_info=True)
    assert info[0] == oinfo["inv_0"] and info[1] == oinfo["post"] and info[2] == max(oinfo["tasks"].values())
    # K_0 + 1e-10 I has cond ~1e11-1e13: two correct implementations agree on the
    # hyper-posterior to ~cond*eps, not to 1e-9.  Tolerances are relative to the largest entry.
    assert relerr(pc, opc) < 1e-6
    assert relerr(pm, opm) < 1e-6
    assert np.array_equal(pc, pc.T)


# ---- m_step, logL_monitoring, full EM --------------------------------------------------------------
def test_m_step_matches_oracle(ctx):
    case = make_case(10, 10, 17)
    upload(ctx, case)
    odb = case["odb"]
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
    stats = {}
    o0, oi = OM.m_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], pm, pc, False, 1e-10, stats)
    h0, hi, st = ctx.m_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(ctx.U), pm, pc)
    # identical optimiser trajectories: same number of objective evaluations
    assert st[0] == stats["hp_0"]["counts"]
    assert st[2] == sum(s["counts"] for s in stats["tasks"].values())
    np.testing.assert_allclose(h0, [o0[n] for n in case["names_0"]], rtol=1e-6, atol=1e-7)
    for t, i in enumerate(case["prep"].ids):
        np.testing.assert_allclose(hi[t], [oi[i][n] for n in case["names_i"]], rtol=1e-6, atol=1e-7)
    ll = ctx.logL_monitoring(case["k0"], h0, case["ki"], hi, np.zeros(ctx.U), pm, pc)
    oll = OM.logL_monitoring(o0, oi, odb, 0.0, "SE", "SE", pm, pc, 1e-10)
    assert abs(ll - oll) < 1e-7 * abs(oll)


def test_m_step_shared_hp(ctx):
    case = make_case(8, 12, 23, shared=True)
    upload(ctx, case)
    odb = case["odb"]
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
    o0, oi = OM.m_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], pm, pc, True, 1e-10)
    h0, hi, st = ctx.m_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(ctx.U), pm, pc,
                            shared_hp=True)
    np.testing.assert_allclose(h0, [o0[n] for n in case["names_0"]], rtol=1e-6, atol=1e-7)
    first = case["prep"].ids[0]
    for t in range(len(case["prep"].ids)):
        np.testing.assert_allclose(hi[t], [oi[first][n] for n in case["names_i"]], rtol=1e-6, atol=1e-7)


def test_train_magma_config1(ctx):
    """BASELINE config 1 (README example shape): M=10, N=10, SE, shared_hp=FALSE."""
    case = make_case(10, 10, 17)
    upload(ctx, case)
    res = ctx.train_magma(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(ctx.U))
    ores = OM.train_magma(case["odb"], None, case["o_hp0"], case["o_hpi"], "SE", "SE", shared_hp=False)
    assert len(res["seq_loglikelihood"]) == len(ores["seq_loglikelihood"])      # EM step count: exact
    assert res["converged"] == ores["converged"]
    np.testing.assert_allclose(res["seq_loglikelihood"], ores["seq_loglikelihood"], rtol=1e-6)
    np.testing.assert_allclose(res["hp_0"], [ores["hp_0"][n] for n in case["names_0"]], rtol=1e-4, atol=1e-5)
    assert relerr(res["post_mean"], ores["post_mean"]) < 1e-5
