"""GPU parity of the large-matrix path (dense_big.cu): DMMA GEMM, blocked Cholesky / inverse with the
reference's jitter escalation, on matrices above the shared-memory-resident limit (n > 224).

Oracle: numpy / LAPACK (what R's `%*%`, chol(), chol2inv() call).  Tolerances are written at each assertion.
"""
import numpy as np
import pytest

from magmaclustr_b200 import _lib as L
from oracle import kernels as OK

pytestmark = pytest.mark.gpu


def _total_jitter(pen, retries):
    tot, p = 0.0, pen
    for _ in range(retries + 1):
        tot += p
        p = 10 * p
    return tot


def _se_cov(n, v, l, noise, lo=-1.0, hi=1.0):
    x = np.linspace(lo, hi, n)
    d2 = (x[:, None] - x[None, :]) ** 2
    return np.exp(v - np.exp(-l) * 0.5 * d2) + noise * np.eye(n)


# ---- R's %*% / crossprod / tcrossprod ----------------------------------------------------------------
@pytest.mark.parametrize("m,n,k", [(200, 200, 64), (257, 131, 515), (128, 384, 1000), (513, 300, 77)])
@pytest.mark.parametrize("ta,tb", [(False, False), (True, False), (False, True), (True, True)])
def test_matmul_matches_dgemm(ctx, m, n, k, ta, tb):
    rng = np.random.default_rng(m + 3 * n + 7 * k)
    a = rng.normal(size=(k, m) if ta else (m, k))
    b = rng.normal(size=(n, k) if tb else (k, n))
    ref = (a.T if ta else a) @ (b.T if tb else b)
    got = ctx.matmul(a, b, ta, tb)
    # both are FP64 dot products of length k in different summation orders: |err| <= k * eps * sum|a||b|
    bound = k * 2.3e-16 * (np.abs(a.T if ta else a) @ np.abs(b.T if tb else b))
    assert np.all(np.abs(got - ref) <= bound + 1e-300)


def test_matmul_linearity_large(ctx):
    # size-independent property at a BASELINE-sized operand (4096): (A + B) C == A C + B C to rounding
    rng = np.random.default_rng(5)
    n = 1024
    a, b, c = rng.normal(size=(n, n)), rng.normal(size=(n, n)), rng.normal(size=(n, 96))
    lhs = ctx.matmul(a + b, c)
    rhs = ctx.matmul(a, c) + ctx.matmul(b, c)
    assert np.max(np.abs(lhs - rhs)) < 1e-10


# ---- chol_inv_jitter above the shared-memory limit ---------------------------------------------------
@pytest.mark.parametrize("n", [225, 256, 300, 513, 1000])
def test_big_chol_inv_jitter_well_conditioned(ctx, n):
    a = _se_cov(n, 0.3, -5.0, 0.05)
    inv, retries, jit = ctx.chol_inv_jitter(a, 1e-10)
    assert retries == 0 and jit == 1e-10
    ref = np.linalg.inv(a + 1e-10 * np.eye(n))
    # cond ~ 30: LAPACK-level agreement; 1e-11 relative to the largest element
    assert np.max(np.abs(inv - ref)) <= 1e-11 * np.max(np.abs(ref))
    assert np.array_equal(inv, inv.T)   # chol2inv + symmetrise: exactly symmetric


@pytest.mark.parametrize("n", [300, 640])
def test_big_chol_inv_jitter_ill_conditioned(ctx, n):
    # the mean-process matrix of the EM loop: SE kernel without noise, only the 1e-10 jitter
    a = _se_cov(n, 0.5, -4.0, 0.0)
    inv, retries, jit = ctx.chol_inv_jitter(a, 1e-10)
    ref, r_ref = OK.chol_inv_jitter(a, 1e-10, return_info=True)
    assert retries == r_ref and jit == _total_jitter(1e-10, r_ref)       # same escalation as the reference's recursion
    # residual check that does not depend on the conditioning of the comparison: (A + jI) inv ~ I
    aj = a + jit * np.eye(n)
    res = aj @ inv - np.eye(n)
    res_ref = aj @ ref - np.eye(n)
    assert np.max(np.abs(res)) <= 10 * max(np.max(np.abs(res_ref)), 1e-12)


def test_big_chol_inv_jitter_retry(ctx):
    # rank-deficient matrix: needs escalation; the count and the total shift must equal the oracle's
    n = 320
    rng = np.random.default_rng(9)
    b = rng.normal(size=(n, 40))
    a = b @ b.T
    inv, retries, jit = ctx.chol_inv_jitter(a, 1e-20)
    ref, r_ref = OK.chol_inv_jitter(a, 1e-20, return_info=True)
    print("rank-deficient 320 x 320: retries gpu %d oracle %d" % (retries, r_ref))
    assert retries > 0
    assert retries == r_ref               # the escalation count is a discrete decision of the reference: exact
    assert jit == _total_jitter(1e-20, retries)


def test_big_cholesky_roundtrip_4096(ctx):
    # BASELINE.json config 4 size: inverse of a 4096 x 4096 covariance; property: inv(inv(A)) == A
    n = 4096
    a = _se_cov(n, 0.0, -5.5, 0.1)
    inv, retries, _ = ctx.chol_inv_jitter(a, 1e-10)
    assert retries == 0
    back, r2, _ = ctx.chol_inv_jitter(inv, 0.0)
    assert r2 == 0
    assert np.max(np.abs(back - a)) < 1e-8


def test_bench_dense_runs(ctx):
    for what in (0, 1, 2):
        assert ctx.bench_dense(what, 512, reps=1) > 0


# ======================================================================================================
# tasks with more than 96 points (bigtasks.cu): same assertions as test_gpu_magma.py / test_gpu_magmaclust.py
# ======================================================================================================
from helpers import make_case, upload, relerr, ulp_perturb  # noqa: E402
from oracle import magma as OM  # noqa: E402
from oracle import magmaclust as OC  # noqa: E402


def _hp_dict(k, hp):
    return dict(zip(L.hp_names(k), [float(v) for v in hp]))


def test_big_list_kern_to_inv_ragged(ctx):
    rng = np.random.default_rng(6)
    sizes = [130, 97, 12, 190]
    offs = np.r_[0, np.cumsum(sizes)]
    pool = np.round(np.arange(-1, 1.001, 0.01), 2)
    X = np.concatenate([np.sort(rng.choice(pool, n, replace=False)) for n in sizes])[:, None]
    k = L.parse_kernel("SE", True)
    hp = np.column_stack([rng.uniform(-0.7, 2.3, len(sizes)), rng.uniform(-3, 0, len(sizes)), rng.uniform(-5, -2, len(sizes))])
    mats, r = ctx.list_kern_to_mat(k, offs, X, hp, inverse=True)
    for t in range(len(sizes)):
        xt = X[offs[t]:offs[t + 1]]
        oinv, oretry = OK.chol_inv_jitter(OK.kern_to_cov(xt, "SE", _hp_dict(k, hp[t])), 1e-10, return_info=True)
        assert r[t] == oretry
        assert relerr(mats[t], oinv) < 1e-9          # cond <= 1e5 (noise floor exp(-5))
        assert np.array_equal(mats[t], mats[t].T)


@pytest.mark.parametrize("M,N,seed,common", [(4, 128, 3, False), (3, 150, 11, True), (3, 97, 5, False)])
def test_big_task_objective_and_gradient(ctx, M, N, seed, common):
    case = make_case(M, N, seed, common_input=common)
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


def test_big_task_compound_kernel_gradient(ctx):
    case = make_case(3, 110, 8, kern_i="SE + PERIO")
    upload(ctx, case)
    odb = case["odb"]
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE + PERIO", case["o_hp0"], case["o_hpi"], 1e-10)
    f, g, r = ctx.logL_GP_mod_tasks(case["ki"], case["hpi"], pm, pc)
    for t, i in enumerate(case["prep"].ids):
        Xi, yi, mi, Si = OM._task_views(odb, pm, pc, i)
        of = OM.logL_GP_mod(case["o_hpi"][i], Xi, yi, mi, "SE + PERIO", Si, 1e-10)
        og = OM.gr_GP_mod(case["o_hpi"][i], Xi, yi, mi, "SE + PERIO", Si, 1e-10)
        assert abs(f[t] - of) <= 1e-9 * abs(of)
        np.testing.assert_allclose(g[t], og, rtol=1e-8, atol=1e-9 * np.max(np.abs(og)))


def test_big_e_step(ctx):
    case = make_case(5, 120, 7)
    upload(ctx, case)
    pm, pc, info = ctx.e_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(ctx.U))
    opm, opc, oinfo = OM.e_step(case["odb"], 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10, return_info=True)
    assert info[0] == oinfo["inv_0"] and info[1] == oinfo["post"] and info[2] == max(oinfo["tasks"].values())
    rng = np.random.default_rng(0)
    sens_c = sens_m = 0.0
    for _ in range(3):   # measured sensitivity of the oracle to a one-ulp change of the HPs (cond(K_0) ~ 1e12)
        h0 = dict(zip(case["names_0"], ulp_perturb(case["hp0"], rng)))
        hi = {i: dict(zip(case["names_i"], ulp_perturb(case["hpi"][t], rng))) for t, i in enumerate(case["prep"].ids)}
        ppm, ppc = OM.e_step(case["odb"], 0.0, "SE", "SE", h0, hi, 1e-10)
        sens_c, sens_m = max(sens_c, relerr(ppc, opc)), max(sens_m, relerr(ppm, opm))
    assert relerr(pc, opc) < 30 * sens_c + 1e-12
    assert relerr(pm, opm) < 30 * sens_m + 1e-12
    assert np.array_equal(pc, pc.T)


def test_big_m_step_matches_oracle(ctx):
    case = make_case(3, 100, 21)
    upload(ctx, case)
    odb = case["odb"]
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
    stats = {}
    o0, oi = OM.m_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], pm, pc, False, 1e-10, stats)
    h0, hi, st = ctx.m_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(ctx.U), pm, pc)
    conv, nev = ctx.last_task_status()
    # identical optimiser trajectories task by task (evaluation counts and convergence codes), HPs to 1e-6
    assert [int(v) for v in nev] == [stats["tasks"][i]["counts"] for i in case["prep"].ids]
    assert [int(v) for v in conv] == [stats["tasks"][i]["convergence"] for i in case["prep"].ids]
    for t, i in enumerate(case["prep"].ids):
        np.testing.assert_allclose(hi[t], [oi[i][n] for n in case["names_i"]], rtol=1e-6, atol=1e-7)
    ll = ctx.logL_monitoring(case["k0"], h0, case["ki"], hi, np.zeros(ctx.U), pm, pc)
    oll = OM.logL_monitoring(dict(zip(case["names_0"], h0)), {i: dict(zip(case["names_i"], hi[t])) for t, i in enumerate(case["prep"].ids)},
                             odb, 0.0, "SE", "SE", pm, pc, 1e-10)
    rng = np.random.default_rng(2)
    f0 = OM.logL_GP_mod(dict(zip(case["names_0"], h0)), odb.grid_X, pm, 0.0, "SE", pc, 1e-10)
    sf = max(abs(OM.logL_GP_mod(dict(zip(case["names_0"], ulp_perturb(h0, rng))), odb.grid_X, pm, 0.0, "SE", pc, 1e-10) - f0)
             for _ in range(4))
    assert abs(ll - oll) < 30 * sf + 1e-9 * abs(oll)


def test_big_elbo_and_mixture(ctx):
    KC = 3
    case = make_case(4, 105, 4, shared=True, K=KC)
    upload(ctx, case)
    rng = np.random.default_rng(4)
    M = 4
    lab = rng.integers(0, KC, size=M)
    mix = np.zeros((M, KC))
    mix[np.arange(M), lab] = 1.0
    hp_k = np.array([[1.0 + 0.3 * k, -1.5 - 0.2 * k] for k in range(KC)])
    o_hp_k = [dict(zip(case["names_0"], hp_k[k])) for k in range(KC)]
    U = case["odb"].grid_X.shape[0]
    omean, ocov, _ = OC.ve_step(case["odb"], [np.zeros(U)] * KC, "SE", "SE", o_hp_k, case["o_hpi"], mix, 1, 1e-10, None)
    soft = np.round(0.8 * mix + 0.2 / KC, 5)
    f, g = ctx.elbo_clust_tasks(KC, case["ki"], case["hpi"], np.array(omean), np.array(ocov), soft)
    for t, i in enumerate(case["prep"].ids):
        of = OC.elbo_clust_multi_GP(case["o_hpi"][i], case["odb"], i, t, omean, ocov, soft, "SE", 1e-10)
        og = OC.gr_clust_multi_GP(case["o_hpi"][i], case["odb"], i, t, omean, ocov, soft, "SE", 1e-10)
        assert abs(f[t] - of) <= 1e-9 * abs(of)
        np.testing.assert_allclose(g[t], og, rtol=1e-8, atol=1e-9 * np.max(np.abs(og)))
    prop = mix.mean(axis=0)
    otau = OC.update_mixture(case["odb"], omean, ocov, case["o_hpi"], "SE", prop, 1e-10)
    gtau = ctx.update_mixture(KC, case["ki"], case["hpi"], np.array(omean), np.array(ocov), prop)
    assert np.array_equal(np.argmax(gtau, axis=1), np.argmax(otau, axis=1))
    assert np.max(np.abs(gtau - otau)) <= 1e-5 + 1e-12
    # VE step through the large-task inverse path
    pm, pc, tau = ctx.ve_step(KC, case["k0"], hp_k, case["ki"], case["hpi"], np.zeros((KC, U)), mix)
    for k in range(KC):
        sens = 0.0
        for _ in range(3):
            hk = [dict(zip(case["names_0"], ulp_perturb(hp_k[q], rng))) for q in range(KC)]
            pmean, pcov, _ = OC.ve_step(case["odb"], [np.zeros(U)] * KC, "SE", "SE", hk, case["o_hpi"], mix, 1, 1e-10, None)
            sens = max(sens, relerr(pcov[k], ocov[k]), relerr(pmean[k], omean[k]))
        assert relerr(pc[k], ocov[k]) < 30 * sens + 1e-12
        assert relerr(pm[k], omean[k]) < 30 * sens + 1e-12
