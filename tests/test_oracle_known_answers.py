"""The oracle against every known-answer / self-consistency test the reference holds for the
hot path: /root/reference/tests/testthat/test-kernels.R, test-kern_to_cov.R,
test-kern_to_inv.R, test-gradients-logL.R, plus the check values derived in SURVEY.md 8c."""
import math

import numpy as np
import pytest

from oracle import kernels as K
from oracle import magma as M
from oracle.lbfgsb import optim_lbfgsb

UNIT = {"SE": {"se_variance": 1.0, "se_lengthscale": 1.0},
        "PERIO": {"perio_variance": 1.0, "perio_lengthscale": 1.0, "period": 1.0},
        "RQ": {"rq_variance": 1.0, "rq_lengthscale": 1.0, "rq_scale": 1.0},
        "LIN": {"lin_slope": 1.0, "lin_offset": 1.0}}


@pytest.mark.parametrize("kern", ["SE", "PERIO", "RQ"])
def test_kernels_zero_distance_equal_e(kern):           # test-kernels.R:1-23
    assert K.kern_to_cov(np.array([2.0]), kern, UNIT[kern])[0, 0] == pytest.approx(math.e, rel=1e-15)


def test_lin_kernel_value():                              # test-kernels.R (LIN): exp(1)*4 + exp(1)
    assert K.kern_to_cov(np.array([2.0]), "LIN", UNIT["LIN"])[0, 0] == pytest.approx(5 * math.e)


@pytest.mark.parametrize("kern", ["SE", "PERIO", "RQ", "LIN"])
def test_kernel_derivatives_match_finite_differences(kern):   # test-kernels.R:25-147
    x = np.array([[2.0, 1.0], [3.5, 0.5], [4.0, 2.0]])
    hp = {k: 0.7 + 0.1 * i for i, k in enumerate(UNIT[kern])}
    for name in hp:
        h2 = dict(hp)
        h2[name] += 1e-7
        fd = (K.kern_to_cov(x, kern, h2) - K.kern_to_cov(x, kern, hp)) / 1e-7
        np.testing.assert_allclose(K.kern_to_cov(x, kern, hp, deriv=name), fd, rtol=1e-4, atol=1e-5)


def test_kern_to_cov_equals_elementwise_kernel():         # test-kern_to_cov.R:1-35
    x = np.array([[1.0, 4.0], [2.0, 5.0], [3.0, 6.0]])
    hp = {"se_variance": 2.0, "se_lengthscale": 1.0}
    ref = np.array([[math.exp(2.0 - 0.5 * math.exp(-1.0) * float(np.sum((a - b) ** 2))) for b in x] for a in x])
    np.testing.assert_allclose(K.kern_to_cov(x, "SE", hp), ref, rtol=1e-15)


def test_compound_kernels():                              # test-kern_to_cov.R:118-151
    x = np.array([[1.0], [2.5], [4.0]])
    hp = {**UNIT["SE"], **UNIT["RQ"], **UNIT["PERIO"]}
    se, rq, pe = (K.kern_to_cov(x, k, hp) for k in ("SE", "RQ", "PERIO"))
    np.testing.assert_allclose(K.kern_to_cov(x, "SE + RQ + PERIO", hp), se + rq + pe, rtol=1e-15)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE * RQ * PERIO", hp), se * rq * pe, rtol=1e-15)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE * RQ + PERIO", hp), se * rq + pe, rtol=1e-15)


def test_compound_kernel_derivatives():                   # test-kern_to_cov.R:153-185
    x = np.array([[1.0], [2.5], [4.0]])
    hp = {**UNIT["SE"], **UNIT["RQ"], **UNIT["PERIO"]}
    se, rq, pe = (K.kern_to_cov(x, k, hp) for k in ("SE", "RQ", "PERIO"))
    dse = K.kern_to_cov(x, "SE", hp, deriv="se_lengthscale")
    drq = K.kern_to_cov(x, "RQ", hp, deriv="rq_scale")
    dpe = K.kern_to_cov(x, "PERIO", hp, deriv="period")
    np.testing.assert_allclose(K.kern_to_cov(x, "SE + RQ + PERIO", hp, "se_lengthscale"), dse)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE + RQ + PERIO", hp, "rq_scale"), drq)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE + RQ + PERIO", hp, "period"), dpe)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE * RQ * PERIO", hp, "rq_scale"), se * drq * pe)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE * RQ + PERIO", hp, "se_lengthscale"), dse * rq)
    np.testing.assert_allclose(K.kern_to_cov(x, "SE * RQ + PERIO", hp, "period"), dpe)


def test_star_after_plus_is_an_error():                   # kernel-to-matrices.R:273-277
    with pytest.raises(ValueError):
        K.kern_to_cov(np.array([1.0, 2.0]), "SE + RQ * PERIO", {**UNIT["SE"], **UNIT["RQ"], **UNIT["PERIO"]},
                      deriv="se_variance")


def test_kern_to_inv_equals_solve():                      # test-kern_to_inv.R:1-36
    x = np.array([2.0, 3.0, 4.0])
    hp = {"se_variance": 2.0, "se_lengthscale": 1.0}
    inv = K.kern_to_inv(x, "SE", hp)
    d = np.max(np.abs(inv - np.linalg.inv(K.kern_to_cov(x, "SE", hp))))
    assert d < 1.5e-8                 # testthat tolerance
    assert d == pytest.approx(7.3e-10, rel=0.05)   # SURVEY.md 8c check value (the 1e-10 jitter)


def test_cumulative_jitter():
    inv, retries = K.chol_inv_jitter(np.ones((3, 3)), 1e-20, return_info=True)
    assert retries == 5


def test_noise_on_coincident_pairs_only():                # code.cpp:105-131, Q4
    x = np.array([[0.0], [1.0], [1.0 + 5e-7]])
    n = K.cpp_noise(x, x, -1.0)
    assert n[1, 2] == n[2, 1] == n[0, 0] == math.exp(-1.0) and n[0, 1] == 0.0


def _five_point():
    X = np.column_stack([np.arange(1, 6), np.arange(3, 8)]).astype(float)
    return X, np.arange(2, 7).astype(float), {"se_variance": 1.0, "se_lengthscale": 0.5}


def test_check_values_logL_and_gradients():               # SURVEY.md 8c on test-gradients-logL.R:2-58
    X, y, hp = _five_point()
    cov = K.kern_to_cov(X, "SE", hp)
    assert M.logL_GP_mod(hp, X, y, 0.0, "SE", cov, 0) == pytest.approx(18.13284374363541, rel=1e-13)
    np.testing.assert_allclose(M.gr_GP_mod(hp, X, y, 0.0, "SE", cov, 0), [-9.4131031, -2.30758734], rtol=1e-8)
    assert M.logL_GP(hp, X, y, 0.0, "SE", cov, 0) == pytest.approx(12.659160142830544, rel=1e-13)
    np.testing.assert_allclose(M.gr_GP(hp, X, y, 0.0, "SE", cov, 0), [-1.10327578, -1.32513776], rtol=1e-8)


@pytest.mark.parametrize("fn,gr", [(M.logL_GP, M.gr_GP), (M.logL_GP_mod, M.gr_GP_mod)])
def test_gradients_match_finite_differences(fn, gr):      # test-gradients-logL.R:1-58 (3 decimals)
    X, y, hp = _five_point()
    cov = K.kern_to_cov(X, "SE", hp)
    g = gr(hp, X, y, 0.0, "SE", cov, 0)
    for p, name in enumerate(hp):
        h2 = dict(hp)
        h2[name] += 1e-8
        emp = (fn(h2, X, y, 0.0, "SE", cov, 0) - fn(hp, X, y, 0.0, "SE", cov, 0)) / 1e-8
        assert round(g[p], 3) == round(emp, 3)


def test_shared_hp_gradient_matches_finite_differences():  # test-gradients-logL.R:62-110
    ids = np.repeat(np.arange(1, 6), 4)
    inp = np.arange(2, 22).astype(float)
    cov_col = np.array(list(range(1, 11)) + [23, 77] + list(range(1, 9)), dtype=float)
    out = np.arange(1, 21).astype(float)
    db = M.Dataset(ids, np.column_stack([inp, cov_col]), out)
    hp = {"se_variance": 1.0, "se_lengthscale": 0.5}
    U = db.grid_X.shape[0]
    pc = K.kern_to_cov(db.grid_X, "SE", hp)
    pm = np.zeros(U)
    g = M.gr_GP_mod_shared_hp(hp, db, pm, "SE", pc, 0)
    for p, name in enumerate(hp):
        h2 = dict(hp)
        h2[name] += 1e-7
        emp = (M.logL_GP_mod_shared_hp(h2, db, pm, "SE", pc, 0) - M.logL_GP_mod_shared_hp(hp, db, pm, "SE", pc, 0)) / 1e-7
        assert g[p] == pytest.approx(emp, rel=1e-4, abs=1e-3)


def test_lbfgsb_rosenbrock_like_r_optim():
    """optim(c(-1.2, 1), fr, grr, method = "L-BFGS-B") in R's documentation reports
    $value 2.2676e-13 with 47 function / gradient evaluations; the restated optimiser follows the
    same trajectory (recalled from R's ?optim example output; R cannot run here)."""
    def fg(x):
        return (100 * (x[1] - x[0] ** 2) ** 2 + (1 - x[0]) ** 2,
                [-400 * x[0] * (x[1] - x[0] ** 2) - 2 * (1 - x[0]), 200 * (x[1] - x[0] ** 2)])
    res = optim_lbfgsb([-1.2, 1.0], fg, factr=1e7, maxit=100)
    assert res["counts"] == 47 and res["convergence"] == 0
    assert res["value"] == pytest.approx(2.2676e-13, rel=1e-3)
    np.testing.assert_allclose(res["par"], [0.9999997, 0.9999995], atol=1e-7)


def test_oracle_em_increases_likelihood():                # test-training-mo.R:35 (single-output analogue)
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import make_case
    case = make_case(5, 8, 1)
    res = M.train_magma(case["odb"], None, case["o_hp0"], case["o_hpi"], "SE", "SE", shared_hp=False,
                        n_iter_max=3)
    seq = res["seq_loglikelihood"]
    assert all(np.isfinite(seq)) and seq[-1] > seq[0]
    assert np.all(np.diag(res["post_cov"]) >= 0)


def test_convolution_kernel_restatement():
    """convolution_kernel (kernels.R:16-140): the vectorised branch equals the single-pair formula of the non-vectorised
    branch (:31-47) element by element; derivatives agree with central differences; with ONE output and l_u -> the kernel is
    symmetric positive definite.  The same function (value, noise placement) reproduces the reference's two multi-output
    simu_data fixtures through magmaclustr_b200/rsim.py (tests/test_rsim.py)."""
    import math
    rng = np.random.default_rng(11)
    O, n = 2, 9
    X = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(0, 3, n), rng.integers(0, O, n)])
    hp = {"l_t_1": -1.2, "l_t_2": -0.7, "S_t_1": 0.4, "S_t_2": -0.3, "l_u_t": -2.0, "noise_1": -1.5, "noise_2": -2.5}
    Kmat = K.kern_to_cov(X, "CONV2", hp)
    for i in range(n):
        for j in range(n):
            oi, oj = int(X[i, 2]), int(X[j, 2])
            den = math.exp(hp["l_t_%d" % (oi + 1)]) + math.exp(hp["l_t_%d" % (oj + 1)]) + math.exp(hp["l_u_t"])
            d2 = float(np.sum((X[i, :2] - X[j, :2]) ** 2))
            ref = math.exp(hp["S_t_%d" % (oi + 1)]) * math.exp(hp["S_t_%d" % (oj + 1)]) / (2 * math.pi * den) ** (2 / 2) \
                * math.exp(-0.5 * d2 / den)
            if i == j:
                ref += math.exp(hp["noise_%d" % (oi + 1)])
            assert Kmat[i, j] == pytest.approx(ref, rel=1e-14)
    assert np.all(np.linalg.eigvalsh(Kmat) > 0)
    names = K.hp_names("CONV2", True)
    assert names == ["l_t_1", "l_t_2", "S_t_1", "S_t_2", "l_u_t", "noise_1", "noise_2"]
    for nm in names:
        d = K.kern_to_cov(X, "CONV2", hp, deriv=nm)
        hp1, hp2 = dict(hp), dict(hp)
        hp1[nm] += 1e-6
        hp2[nm] -= 1e-6
        fd = (K.kern_to_cov(X, "CONV2", hp1) - K.kern_to_cov(X, "CONV2", hp2)) / 2e-6
        np.testing.assert_allclose(d, fd, rtol=1e-6, atol=1e-9)
    # cross-covariance: no noise (kernel-to-matrices.R:146-157)
    C = K.kern_to_cov(X[:4], "CONV2", hp, input_2=X[4:])
    assert C.shape == (4, 5) and np.allclose(C, Kmat[:4, 4:])
