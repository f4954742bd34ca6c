"""``simu_data`` with R's own random number stream.

The reference simulates its datasets with ``set.seed()`` + ``stats::runif / rnorm / sample``
(/root/reference/R/simulate-data.R:13-44,116-124,386-528,569-656,722-924) and ships the exact outputs of six configurations
as its only numeric fixture (tests/testthat/fixtures/simu_data_reference.rds, compared with tolerance 1e-6 by
tests/testthat/test-simu_data-characterization.R:43-52).  This module restates

* R's default generators (third-party to the reference: R's src/main/RNG.c, src/nmath/snorm.c, qnorm.c, unif.c, and
  src/main/unique.c / random.c for ``sample``): Mersenne-Twister seeded by ``set.seed`` (LCG scrambling, 69069),
  ``unif_rand`` with the open-interval fix-up, ``norm_rand`` by INVERSION (two uniforms, 2^27 grid, Wichura's AS 241
  quantile function) and ``sample`` by REJECTION sampling of random bits (R >= 3.6.0), drawing without replacement with the
  swap-with-last scheme;
* ``draw``, ``.rmvnorm_chol`` and the single-output branch of ``simu_data`` with R's order of evaluation -- the promises
  ``v = draw(int_mu_v), l = draw(int_mu_l)`` of ``generate_mean_process`` are forced only when the covariance is built,
  i.e. AFTER the grid permutation and the prior-mean draws (simulate-data.R:404-447,826-834);

so that the Python side reproduces the reference's simulated datasets from a seed (tests/test_rsim.py checks all six
configurations of the fixture, the two multi-output ones through the convolution kernel of kernels.R:16-98).  Input generation only: host numpy, not part of the hot path.
"""
import math

import numpy as np

_M32 = 0xFFFFFFFF


class RRng:
    """R's RNGkind("Mersenne-Twister", "Inversion", "Rejection") -- the defaults since R 3.6.0."""

    N, M = 624, 397

    def __init__(self, seed):
        self.set_seed(seed)

    def set_seed(self, seed):
        s = int(seed) & _M32
        for _ in range(50):                       # RNG.c:RNG_Init initial scrambling
            s = (69069 * s + 1) & _M32
        st = []
        for _ in range(self.N + 1):
            s = (69069 * s + 1) & _M32
            st.append(s)
        self.mt = st[1:]                          # dummy[0] is mti; FixupSeeds sets it to N: regenerate at first use
        self.mti = self.N

    def _genrand(self):
        mt, N, M = self.mt, self.N, self.M
        if self.mti >= N:
            for kk in range(N):
                y = (mt[kk] & 0x80000000) | (mt[(kk + 1) % N] & 0x7FFFFFFF)
                mt[kk] = mt[(kk + M) % N] ^ (y >> 1) ^ (0x9908B0DF if (y & 1) else 0)
            self.mti = 0
        y = mt[self.mti]
        self.mti += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return (y & _M32) * 2.3283064365386963e-10          # [0, 1)

    def unif_rand(self):
        v = self._genrand()
        i2 = 2.328306437080797e-10
        if v <= 0.0:
            return 0.5 * i2
        if 1.0 - v <= 0.0:
            return 1.0 - 0.5 * i2
        return v

    def runif(self, a, b):
        if a == b:
            return a
        while True:
            u = self.unif_rand()
            if 0.0 < u < 1.0:
                return a + (b - a) * u

    def norm_rand(self):
        big = 134217728.0
        u = self.unif_rand()
        u = int(big * u) + self.unif_rand()
        return qnorm(u / big)

    def rnorm(self, n):
        return np.array([self.norm_rand() for _ in range(n)])

    def _rbits(self, bits):
        v = 0
        n = 0
        while n <= bits:
            v = 65536 * v + int(math.floor(self.unif_rand() * 65536))
            n += 16
        if bits < 64:
            v &= (1 << bits) - 1
        return v

    def unif_index(self, dn):
        if dn <= 0:
            return 0
        bits = int(math.ceil(math.log2(dn)))
        while True:
            dv = self._rbits(bits)
            if dv < dn:
                return dv

    def sample_int(self, n, size):
        """sample.int(n, size, replace = FALSE): 0-based indices (R returns them + 1)."""
        x = list(range(n))
        out = []
        for _ in range(size):
            j = self.unif_index(n)
            out.append(x[j])
            n -= 1
            x[j] = x[n]
        return out


def qnorm(p):
    """R's qnorm5(p, 0, 1, lower = TRUE, log = FALSE): Wichura (1988), algorithm AS 241 (PPND16)."""
    if p <= 0.0:
        return -math.inf
    if p >= 1.0:
        return math.inf
    q = p - 0.5
    if abs(q) <= 0.425:
        r = 0.180625 - q * q
        return q * (((((((r * 2509.0809287301226727 + 33430.575583588128105) * r + 67265.770927008700853) * r
                        + 45921.953931549871457) * r + 13731.693765509461125) * r + 1971.5909503065514427) * r
                      + 133.14166789178437745) * r + 3.387132872796366608) / \
            (((((((r * 5226.495278852545925 + 28729.085735721942674) * r + 39307.89580009271061) * r
                 + 21213.794301586595867) * r + 5394.1960214247511077) * r + 687.1870074920579083) * r
               + 42.313330701600911252) * r + 1.0)
    lp = p if q < 0 else 1.0 - p
    r = math.sqrt(-math.log(lp))
    if r <= 5.0:
        r -= 1.6
        val = (((((((r * 7.7454501427834140764e-4 + .0227238449892691845833) * r + .24178072517745061177) * r
                   + 1.27045825245236838258) * r + 3.64784832476320460504) * r + 5.7694972214606914055) * r
                + 4.6303378461565452959) * r + 1.42343711074968357734) / \
            (((((((r * 1.05075007164441684324e-9 + 5.475938084995344946e-4) * r + .0151986665636164571966) * r
                 + .14810397642748007459) * r + .68976733498510000455) * r + 1.6763848301838038494) * r
               + 2.05319162663775882187) * r + 1.0)
    else:
        r -= 5.0
        val = (((((((r * 2.01033439929228813265e-7 + 2.71155556874348757815e-5) * r + .0012426609473880784386) * r
                   + .026532189526576123093) * r + .29656057182850489123) * r + 1.7848265399172913358) * r
                + 5.4637849111641143699) * r + 6.6579046435011037772) / \
            (((((((r * 2.04426310338993978564e-15 + 1.4215117583164458887e-7) * r + 1.8463183175100546818e-5) * r
                 + 7.868691311456132591e-4) * r + .0148753612908506148525) * r + .13692988092273580531) * r
               + .59983224884798220045) * r + 1.0)
    return -val if q < 0.0 else val


def r_round(x, digits):
    """R's round(x, digits) for digits > 0 (nmath/fround.c, R >= 4.0.0): the nearer of the two neighbouring decimals
    (evaluated in long double), ties to the even one."""
    if not math.isfinite(x):
        return x
    sgn = 1.0
    if x < 0:
        sgn, x = -1.0, -x
    ld = np.longdouble
    p10 = ld(10.0) ** digits
    x10 = p10 * ld(x)
    i10 = np.floor(x10)
    xd, xu = float(i10 / p10), float(np.ceil(x10) / p10)
    du, dd = xu - x, x - xd
    even = float(i10) % 2.0 == 0.0
    return sgn * (xd if (dd < du or (dd == du and even)) else xu)


def _draw(rng, interval):
    lo, hi = sorted(interval)
    return r_round(rng.runif(lo, hi), 2)


def _se_cov(X, v, l):
    """kern_to_cov(X, "SE", (v, l)): exp(v - 0.5 exp(-l) cpp_dist) (kernels.R:164-189, src/code.cpp:6-27)."""
    X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
    d2 = np.zeros((len(X), len(X)))
    for c in range(X.shape[1]):
        diff = X[:, c][:, None] - X[:, c][None, :]
        d2 += diff ** 2
    return np.exp(v - 0.5 * math.exp(-l) * d2)


def _conv_cov(X, out_ids, l_t, S_t, l_u_t, noise=None):
    """kern_to_cov(input, convolution_kernel, hp) with ONE latent process (kernels.R:16-98, kernel-to-matrices.R:88-160):
    K_ij = S_i S_j / (2 pi (l_i + l_j + l_u))^(D/2) exp(-d_ij^2 / (2 (l_i + l_j + l_u))) with l = exp(l_t), S = exp(S_t) of
    each point's output, + diag(exp(noise)) of each point's output when the HP table carries a noise column."""
    X = np.asarray(X, dtype=np.float64).reshape(len(X), -1)
    D = X.shape[1]
    d2 = np.zeros((len(X), len(X)))
    for c in range(D):
        diff = X[:, c][:, None] - X[:, c][None, :]
        d2 += diff ** 2
    l = np.exp(np.asarray(l_t, dtype=np.float64))[out_ids]
    S = np.exp(np.asarray(S_t, dtype=np.float64))[out_ids]
    denom = (l[:, None] + l[None, :]) + math.exp(l_u_t)
    K = (S[:, None] * S[None, :]) / ((2 * math.pi * denom) ** (D / 2)) * np.exp(-0.5 * d2 / denom)
    if noise is not None:
        K = K + np.diag(np.exp(np.asarray(noise, dtype=np.float64)[out_ids]))
    return K


def _rmvnorm_chol(rng, mu, sigma):
    """.rmvnorm_chol (simulate-data.R:27-44): mu + R' z with R = chol(Sigma + jitter I), escalating jitter."""
    from scipy.linalg import cholesky
    p = len(mu)
    base = max(float(np.mean(np.diag(sigma))), 1.0)
    R = None
    for k in range(9):
        jitter = 0.0 if k == 0 else base * 1e-10 * 10.0 ** (k - 1)
        try:
            R = cholesky(sigma + jitter * np.eye(p), lower=False)
            break
        except np.linalg.LinAlgError:
            continue
    if R is None:
        raise FloatingPointError("Covariance matrix is not positive definite (Cholesky failed).")
    z = rng.rnorm(p)
    return np.asarray(mu) + R.T @ z


def _seq(lo, hi, by):
    """R's seq(lo, hi, by) for doubles: lo + (0:n) * by with n = floor((hi - lo) / by + 1e-10)."""
    n = int(math.floor((hi - lo) / by + 1e-10))
    return lo + np.arange(n + 1) * by


def _sample_sorted(rng, grid, size):
    """sample_grid_points (simulate-data.R:116-124): `size` rows without replacement, sorted (all columns, in order)."""
    grid = np.asarray(grid, dtype=np.float64)
    idx = rng.sample_int(len(grid), size)
    pts = grid[idx]
    if pts.ndim == 1:
        return np.sort(pts)
    order = np.lexsort(tuple(pts[:, c] for c in range(pts.shape[1] - 1, -1, -1)))
    return pts[order]


def simu_data(seed, n_tasks=10, n_points=10, n_clusters=1, n_outputs=1, input2=False, prior_means=None,
              grid_input=None, grid_input2=None, shared_input=False, shared_hp=True, add_hp=False, add_clust=False,
              int_mu_v=(0.0, math.log(3)), int_mu_l=(math.log(1 / 300), math.log(1 / 200)),
              int_i_v=(0.0, math.log(1.5)), int_i_l=(math.log(1 / 300), math.log(1 / 200)), int_i_sigma=(0.0, 0.2),
              m0_slope=(-5.0, 5.0), m0_intercept=(-50.0, 50.0), warning_grid=True):
    """``set.seed(seed); simu_data(...)`` of the reference for ONE output (simulate-data.R:722-924).

    Returns a dict of columns in the reference's long format: Task_ID, Input_ID, Input, Output_ID, Output (+ se_variance,
    se_lengthscale, noise with add_hp; Cluster with add_clust), one row per (observation, input dimension)."""
    if n_outputs != 1:
        return _simu_data_mo(seed, n_tasks, n_points, n_clusters, n_outputs, input2, prior_means, grid_input, grid_input2,
                             shared_input, shared_hp, add_hp, add_clust, int_mu_v, int_mu_l, int_i_v, int_i_l, int_i_sigma,
                             m0_slope, m0_intercept, warning_grid)
    if prior_means is not None:
        raise NotImplementedError("explicit prior_means")
    rng = RRng(seed)
    grid_input = _seq(-1.0, 1.0, 0.01) if grid_input is None else np.asarray(grid_input, dtype=np.float64)
    grid_input2 = _seq(-1.0, 1.0, 0.5) if grid_input2 is None else np.asarray(grid_input2, dtype=np.float64)
    if len(grid_input) > 2000 and warning_grid:
        raise ValueError("The defined input grid is > 2000 points.")
    if input2:                                    # tidyr::expand_grid(Input, Input2): Input varies slowest
        domain = np.array([[a, b] for a in grid_input for b in grid_input2])
        if len(domain) > 2000 and warning_grid:
            raise ValueError("The defined input grid is > 2000 points.")
    else:
        domain = grid_input
    rows = {k: [] for k in ("Task_ID", "Input_ID", "Input", "Output_ID", "Output")}
    extra = {"se_variance": [], "se_lengthscale": [], "noise": [], "Cluster": []}
    D = 2 if input2 else 1
    for k in range(1, n_clusters + 1):
        # ---- generate_mean_process: the whole domain, permuted then sorted; prior mean; THEN the promises v, l
        grid = _sample_sorted(rng, domain, len(domain))
        first = grid[:, 0] if grid.ndim == 2 else grid
        m0 = _draw(rng, m0_intercept) + _draw(rng, m0_slope) * first
        v0 = _draw(rng, int_mu_v)
        l0 = _draw(rng, int_mu_l)
        mean_proc = _rmvnorm_chol(rng, m0, _se_cov(grid, v0, l0))
        key = {tuple(np.atleast_1d(g)): i for i, g in enumerate(grid)}
        if shared_hp:
            hp_shared = (_draw(rng, int_i_v), _draw(rng, int_i_l), _draw(rng, int_i_sigma))
        shared_grid = _sample_sorted(rng, grid, n_points) if shared_input else None
        for i in range(1, n_tasks + 1):
            task_id = ("ID%d-Clust%d" % (i, k)) if n_clusters > 1 else str(i)
            v, l, sigma = hp_shared if shared_hp else (_draw(rng, int_i_v), _draw(rng, int_i_l), _draw(rng, int_i_sigma))
            pts = shared_grid if shared_input else _sample_sorted(rng, grid, n_points)
            cov = _se_cov(pts, v, l) + sigma * np.eye(len(pts))
            mu = np.array([mean_proc[key[tuple(np.atleast_1d(p))]] for p in pts])
            y = _rmvnorm_chol(rng, mu, cov)
            for r in range(len(pts)):             # pivot_inputs_longer: one row per input dimension
                for q in range(D):
                    rows["Task_ID"].append(task_id)
                    rows["Input_ID"].append(str(q + 1))
                    rows["Input"].append(float(pts[r, q]) if D == 2 else float(pts[r]))
                    rows["Output_ID"].append("1")
                    rows["Output"].append(float(y[r]))
                    extra["se_variance"].append(v)
                    extra["se_lengthscale"].append(l)
                    extra["noise"].append(sigma)
                    extra["Cluster"].append(k)
    if add_hp:
        for c in ("se_variance", "se_lengthscale", "noise"):
            rows[c] = extra[c]
    if add_clust:
        rows["Cluster"] = extra["Cluster"]
    return rows


def _simu_data_mo(seed, n_tasks, n_points, n_clusters, n_outputs, input2, prior_means, grid_input, grid_input2,
                  shared_input, shared_hp, add_hp, add_clust, int_mu_v, int_mu_l, int_i_v, int_i_l, int_i_sigma, m0_slope,
                  m0_intercept, warning_grid):
    """The multi-output branch of simu_data (simulate-data.R:404-520,569-656,775-905): convolution kernel with one latent
    process, hyper-parameters drawn by hp(kern = convolution_kernel, ..., hp_config) (hyperparameters.R:90-210)."""
    if prior_means is not None or input2:
        raise NotImplementedError("multi-output simu_data with explicit prior_means / a second input")
    rng = RRng(seed)
    pool = _seq(-1.0, 1.0, 0.01) if grid_input is None else np.asarray(grid_input, dtype=np.float64)
    if len(pool) > 2000 and warning_grid:
        raise ValueError("The defined input grid is > 2000 points.")
    O = n_outputs
    rows = {k: [] for k in ("Task_ID", "Input_ID", "Input", "Output_ID", "Output")}
    extra = {k: [] for k in ("l_t", "S_t", "noise", "l_u_t", "Cluster")}

    def draw_task_hp():
        # hp(): l_t, S_t, noise per output (vectorised runif, in that order), then ONE l_u_t
        l_t = [rng.runif(*sorted_pair(int_i_l, False)) for _ in range(O)]
        S_t = [rng.runif(*sorted_pair(int_i_v, False)) for _ in range(O)]
        nz = [rng.runif(*sorted_pair(int_i_sigma, False)) for _ in range(O)]
        l_u = rng.runif(*sorted_pair(int_i_l, False))
        return l_t, S_t, nz, l_u

    def sorted_pair(iv, _):
        return float(iv[0]), float(iv[1])          # hp_config keeps the interval as given (no sort(), unlike draw())

    for k in range(1, n_clusters + 1):
        grids = [_sample_sorted(rng, pool, len(pool)) for _ in range(O)]
        means = [_draw(rng, m0_intercept) + _draw(rng, m0_slope) * g for g in grids]
        lu0 = rng.runif(float(int_mu_l[0]), float(int_mu_l[1]))
        l0 = [rng.runif(float(int_mu_l[0]), float(int_mu_l[1])) for _ in range(O)]
        S0 = [rng.runif(float(int_mu_v[0]), float(int_mu_v[1])) for _ in range(O)]
        X = np.concatenate(grids)
        oid = np.concatenate([np.full(len(g), o) for o, g in enumerate(grids)])
        mean_proc = _rmvnorm_chol(rng, np.concatenate(means), _conv_cov(X, oid, l0, S0, lu0))
        key = {(int(o), float(x)): i for i, (o, x) in enumerate(zip(oid, X))}
        if shared_hp:
            hp_shared = draw_task_hp()
        shared_grids = [_sample_sorted(rng, g, n_points) for g in grids] if shared_input else None
        for i in range(1, n_tasks + 1):
            task_id = ("ID%d-Clust%d" % (i, k)) if n_clusters > 1 else str(i)
            l_t, S_t, nz, l_u = hp_shared if shared_hp else draw_task_hp()
            tg = shared_grids if shared_input else [_sample_sorted(rng, g, n_points) for g in grids]
            Xt = np.concatenate(tg)
            ot = np.concatenate([np.full(len(g), o) for o, g in enumerate(tg)])
            mu = np.array([mean_proc[key[(int(o), float(x))]] for o, x in zip(ot, Xt)])
            y = _rmvnorm_chol(rng, mu, _conv_cov(Xt, ot, l_t, S_t, l_u, nz))
            for r in range(len(Xt)):
                rows["Task_ID"].append(task_id)
                rows["Input_ID"].append("1")
                rows["Input"].append(float(Xt[r]))
                rows["Output_ID"].append(str(int(ot[r]) + 1))
                rows["Output"].append(float(y[r]))
                extra["l_t"].append(l_t[ot[r]])
                extra["S_t"].append(S_t[ot[r]])
                extra["noise"].append(nz[ot[r]])
                extra["l_u_t"].append(l_u)
                extra["Cluster"].append(k)
    if add_hp:
        for c in ("l_t", "S_t", "noise", "l_u_t"):
            rows[c] = extra[c]
    if add_clust:
        rows["Cluster"] = extra["Cluster"]
    return rows
