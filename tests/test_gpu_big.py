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
    assert retries > 0
    assert abs(retries - r_ref) <= 1      # the failing pivot sits at rounding level: allow one step
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
