"""Synthetic datasets with the distribution of the reference's ``simu_data``.

Input generation only (numpy on the host) -- not part of the hot path.  R's RNG is
not reproducible outside R, so draws come from ``numpy.random.default_rng(seed)``
but follow /root/reference/R/simulate-data.R:722-743,810-905:
pool ``seq(-1, 1, 0.01)``; prior mean ``a + b x`` with ``a~U(-50,50)``,
``b~U(-5,5)``; mean process ``mu_0 ~ N(a + b x, K_SE(v0, l0))``; per task N pool
points without replacement (sorted), ``v_i~U(0, log 1.5)``,
``l_i~U(log 1/300, log 1/200)``, ``sigma_i~U(0, 0.2)`` (added to the diagonal as a
variance, simulate-data.R:613-617), all rounded to 2 decimals (``draw``, :13-19);
Cholesky draws with the escalating jitter of ``.rmvnorm_chol`` (:27-44).
Initial hyper-parameters follow ``hp()``'s ranges (hyperparameters.R:70-77).
"""
import math

import numpy as np
import pandas as pd


def _se_cov(x, v, l):
    d = x[:, None] - x[None, :]
    return np.exp(v - 0.5 * math.exp(-l) * d * d)


def _rmvnorm_chol(rng, mu, sigma):
    base = max(float(np.mean(np.diag(sigma))), 1.0)
    for k in range(9):
        jit = 0.0 if k == 0 else base * 1e-10 * 10 ** (k - 1)
        try:
            L = np.linalg.cholesky(sigma + jit * np.eye(len(mu)))
            return mu + L @ rng.standard_normal(len(mu))
        except np.linalg.LinAlgError:
            continue
    raise FloatingPointError("Covariance matrix is not positive definite")


def _draw(rng, lo, hi, size=None):
    return np.round(rng.uniform(min(lo, hi), max(lo, hi), size), 2)


def simu_db(M=10, N=10, K=1, common_input=False, seed=0, pool=None):
    """Returns a legacy-format DataFrame (ID, Input, Output) with 1-D inputs, plus a
    dict of the generating quantities.

    The draws go through LAPACK / BLAS (Cholesky factor of an ill-conditioned covariance, L z), whose rounding depends on
    the number of BLAS threads -- and torchrun sets OMP_NUM_THREADS = 1 for world > 1 only.  The generator therefore always
    runs single-threaded, so that every rank of every world size simulates the SAME bits for a seed."""
    try:
        from threadpoolctl import threadpool_limits
    except ImportError:                      # pragma: no cover
        return _simu_db(M, N, K, common_input, seed, pool)
    with threadpool_limits(limits=1):
        return _simu_db(M, N, K, common_input, seed, pool)


def _simu_db(M, N, K, common_input, seed, pool):
    rng = np.random.default_rng(seed)
    if pool is None:
        pool = np.round(np.arange(-1.0, 1.0 + 1e-9, 0.01), 2)
    pool = np.asarray(pool, dtype=np.float64)
    means = []
    for _ in range(K):
        a, b = _draw(rng, -50, 50), _draw(rng, -5, 5)
        v0, l0 = _draw(rng, 0, math.log(3)), _draw(rng, math.log(1 / 300), math.log(1 / 200))
        means.append(_rmvnorm_chol(rng, a + b * pool, _se_cov(pool, v0, l0)))
    clusters = rng.integers(0, K, size=M) if K > 1 else np.zeros(M, dtype=np.int64)
    v = _draw(rng, 0, math.log(1.5), M)
    l = _draw(rng, math.log(1 / 300), math.log(1 / 200), M)
    s = _draw(rng, 0, 0.2, M)
    shared_idx = np.sort(rng.choice(len(pool), size=N, replace=False)) if common_input else None
    ids, xs, ys = [], [], []
    for i in range(M):
        idx = shared_idx if common_input else np.sort(rng.choice(len(pool), size=N, replace=False))
        x = pool[idx]
        cov = _se_cov(x, v[i], l[i]) + s[i] * np.eye(N)
        y = _rmvnorm_chol(rng, means[clusters[i]][idx], cov)
        ids += [str(i + 1)] * N
        xs.append(x)
        ys.append(y)
    df = pd.DataFrame({"ID": ids, "Input": np.concatenate(xs), "Output": np.concatenate(ys)})
    return df, {"clusters": clusters, "means": means, "pool": pool}


def simu_db_spectral(M=10, N=10, common_input=False, seed=0, pool=None, n_features=256):
    """Same table, same hyper-parameter ranges and the same model as `simu_db` (one mean process, SE kernels), but every
    Gaussian-process draw is a random-Fourier-feature sum  sqrt(2 e^v / F) sum_j cos(w_j x + b_j),  w_j ~ N(0, e^{-l}),
    b_j ~ U(0, 2 pi)  -- the spectral representation of the SE kernel -- instead of a Cholesky factor times a normal vector.
    O(N F) per task instead of O(N^3): what makes config 4 (256 tasks on a 4096-point grid) generable in seconds on the
    benchmark box.  Element-wise numpy only, so every rank generates the same bits."""
    rng = np.random.default_rng(seed)
    if pool is None:
        pool = np.round(np.arange(-1.0, 1.0 + 1e-9, 0.01), 2)
    pool = np.asarray(pool, dtype=np.float64)

    def gp_draw(x, v, l):
        w = rng.standard_normal(n_features) * math.exp(-0.5 * l)
        b = rng.uniform(0.0, 2.0 * math.pi, n_features)
        out = np.zeros(len(x))
        for j in range(n_features):          # fixed summation order
            out += np.cos(w[j] * x + b[j])
        return math.sqrt(2.0 * math.exp(v) / n_features) * out

    a, b = _draw(rng, -50, 50), _draw(rng, -5, 5)
    v0, l0 = _draw(rng, 0, math.log(3)), _draw(rng, math.log(1 / 300), math.log(1 / 200))
    mean = a + b * pool + gp_draw(pool, v0, l0)
    v = _draw(rng, 0, math.log(1.5), M)
    l = _draw(rng, math.log(1 / 300), math.log(1 / 200), M)
    s = _draw(rng, 0, 0.2, M)
    shared_idx = np.sort(rng.choice(len(pool), size=N, replace=False)) if common_input else None
    ids, xs, ys = [], [], []
    for i in range(M):
        idx = shared_idx if common_input else np.sort(rng.choice(len(pool), size=N, replace=False))
        x = pool[idx]
        y = mean[idx] + gp_draw(x, v[i], l[i]) + math.sqrt(s[i]) * rng.standard_normal(N)
        ids += [str(i + 1)] * N
        xs.append(x)
        ys.append(y)
    df = pd.DataFrame({"ID": ids, "Input": np.concatenate(xs), "Output": np.concatenate(ys)})
    return df, {"means": [mean], "pool": pool}


def ini_hp(kern, ids, shared, noise, seed):
    """Random initial hyper-parameters in hp()'s ranges, as a DataFrame with an ID
    column (one row per id; identical rows when `shared`)."""
    rng = np.random.default_rng(seed)
    n = 1 if shared else len(ids)
    cols = {}
    for tok in kern.split(" "):
        if tok == "SE":
            cols["se_variance"] = rng.uniform(math.log(0.5), math.log(10), n)
            cols["se_lengthscale"] = rng.uniform(-3, 0, n)
        elif tok == "PERIO":
            cols["perio_variance"] = rng.uniform(math.log(0.5), math.log(10), n)
            cols["perio_lengthscale"] = rng.uniform(-3, 0, n)
            cols["period"] = rng.uniform(0, 2 * math.pi, n)
        elif tok == "RQ":
            cols["rq_variance"] = rng.uniform(math.log(0.5), math.log(10), n)
            cols["rq_lengthscale"] = rng.uniform(-3, 0, n)
            cols["rq_scale"] = rng.uniform(0.1, 10, n)
        elif tok == "LIN":
            cols["lin_slope"] = rng.uniform(-3, 10, n)
            cols["lin_offset"] = rng.uniform(-3, 10, n)
    if noise:
        cols["noise"] = rng.uniform(-5, -2, n)
    df = pd.DataFrame({k: (np.repeat(v, len(ids)) if shared else v) for k, v in cols.items()})
    df.insert(0, "ID", [str(i) for i in ids])
    return df
