"""Kernel -> matrix layer of the reference, restated (oracle; test infrastructure).

Follows /root/reference/src/code.cpp (five native builders),
/root/reference/R/kernels.R:164-398 (se/perio/rq/lin kernels and derivatives) and
/root/reference/R/kernel-to-matrices.R:71-503,553-562,670-680
(kern_to_cov string parser, kern_to_inv, chol_inv_jitter).

Inputs are ``n x D`` float64 arrays (the tibble minus ID/Output/Reference),
hyper-parameters are dicts in the tibble's column order.
"""
import math

import numpy as np
from scipy.linalg import lapack


def _rexp(x):
    """exp() with R's (IEEE) overflow behaviour: Inf instead of Python's OverflowError.  An optimiser trial point far outside
    the sensible range then yields a non-finite objective, which optim reports as "L-BFGS-B needs finite values of 'fn'"
    (oracle/lbfgsb.py raises FloatingPointError there) -- not a Python-specific exception from inside the kernel."""
    try:
        return math.exp(x)
    except OverflowError:
        return math.inf


MAX_JITTER_RETRY = 40  # R has no cap (recursion until the C stack ends); see DESIGN.md


def _mat(x):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x[:, None]
    return x


# ---- src/code.cpp -----------------------------------------------------------
def cpp_dist(m1, m2):
    """code.cpp:6-27: sum_c (m1[r1,c]-m2[r2,c])^2, accumulated over c in order."""
    m1, m2 = _mat(m1), _mat(m2)
    if m1.shape[1] != m2.shape[1]:
        raise RuntimeError("Incompatible number of dimensions")
    out = np.zeros((m1.shape[0], m2.shape[0]))
    for c in range(m1.shape[1]):
        d = m1[:, c][:, None] - m2[:, c][None, :]
        out += d * d
    return out


def cpp_perio(m1, m2, period):
    """code.cpp:31-52."""
    m1, m2 = _mat(m1), _mat(m2)
    out = np.zeros((m1.shape[0], m2.shape[0]))
    with np.errstate(all="ignore"):
        w = float(np.float64(math.pi) / np.exp(np.float64(period)))   # IEEE like C++: pi / 0 = Inf
    for c in range(m1.shape[1]):
        with np.errstate(all="ignore"):
            s = np.sin(w * np.abs(m1[:, c][:, None] - m2[:, c][None, :]))
        out += s * s
    return out


def cpp_perio_deriv(m1, m2, period):
    """code.cpp:56-76."""
    m1, m2 = _mat(m1), _mat(m2)
    out = np.zeros((m1.shape[0], m2.shape[0]))
    with np.errstate(all="ignore"):
        w = float(np.float64(math.pi) / np.exp(np.float64(period)))
    for c in range(m1.shape[1]):
        with np.errstate(all="ignore"):
            a = w * np.abs(m1[:, c][:, None] - m2[:, c][None, :])
            out += 2 * np.sin(a) * np.cos(a) * a
    return out


def cpp_prod(m1, m2):
    """code.cpp:80-101."""
    m1, m2 = _mat(m1), _mat(m2)
    if m1.shape[1] != m2.shape[1]:
        raise RuntimeError("Incompatible number of dimensions")
    out = np.zeros((m1.shape[0], m2.shape[0]))
    for c in range(m1.shape[1]):
        out += m1[:, c][:, None] * m2[:, c][None, :]
    return out


def cpp_noise(m1, m2, noise):
    """code.cpp:105-131: exp(noise) where the L1 distance < 1e-6, else 0."""
    m1, m2 = _mat(m1), _mat(m2)
    diff = np.zeros((m1.shape[0], m2.shape[0]))
    for c in range(m1.shape[1]):
        diff += np.abs(m1[:, c][:, None] - m2[:, c][None, :])
    return np.where(diff < 0.000001, _rexp(noise), 0.0)


# ---- R/kernels.R ------------------------------------------------------------
def se_kernel(x, y, hp, deriv=None):
    """kernels.R:164-196."""
    distance = cpp_dist(x, y)
    top_term = _rexp(-hp["se_lengthscale"]) * 0.5 * distance
    if deriv is None or deriv == "se_variance":
        return np.exp(hp["se_variance"] - top_term)
    if deriv == "se_lengthscale":
        return _rexp(hp["se_variance"]) * top_term * np.exp(-top_term)
    raise ValueError("invalid deriv")


def perio_kernel(x, y, hp, deriv=None):
    """kernels.R:221-270."""
    perio_term = cpp_perio(x, y, hp["period"])
    sum_deriv = cpp_perio_deriv(x, y, hp["period"])
    v, l = hp["perio_variance"], hp["perio_lengthscale"]
    if deriv is None or deriv == "perio_variance":
        return _rexp(v) * np.exp(-2 * _rexp(-l) * perio_term)
    if deriv == "period":
        return (_rexp(v) * np.exp(-2 * _rexp(-l) * perio_term)
                * 2 * _rexp(-l) * sum_deriv)
    if deriv == "perio_lengthscale":
        return (_rexp(v) * 2 * perio_term * _rexp(-l)
                * np.exp(-2 * perio_term * _rexp(-l)))
    raise ValueError("invalid deriv")


def rq_kernel(x, y, hp, deriv=None):
    """kernels.R:294-341 (rq_scale is used raw, not exponentiated)."""
    distance = cpp_dist(x, y)
    v, l, a = hp["rq_variance"], hp["rq_lengthscale"], hp["rq_scale"]
    term = 1 + distance * _rexp(-l) / (2 * a)
    if deriv is None or deriv == "rq_variance":
        return _rexp(v) * term ** (-a)
    if deriv == "rq_scale":
        return (_rexp(v) * term ** (-a)
                * (distance * _rexp(-l) / (2 * a * term) - np.log(term)))
    if deriv == "rq_lengthscale":
        return _rexp(v) * distance * 0.5 * _rexp(-l) * term ** (-a - 1)
    raise ValueError("invalid deriv")


def lin_kernel(x, y, hp, deriv=None):
    """kernels.R:364-398."""
    prod = cpp_prod(x, y)
    if deriv is None:
        return _rexp(hp["lin_slope"]) * prod + _rexp(hp["lin_offset"])
    if deriv == "lin_offset":
        return np.exp(hp["lin_offset"] * np.ones_like(prod))
    if deriv == "lin_slope":
        return _rexp(hp["lin_slope"]) * prod
    raise ValueError("invalid deriv")


_KERNELS = {
    "SE": (se_kernel, ("se_variance", "se_lengthscale")),
    "PERIO": (perio_kernel, ("perio_variance", "perio_lengthscale", "period")),
    "RQ": (rq_kernel, ("rq_variance", "rq_lengthscale", "rq_scale")),
    "LIN": (lin_kernel, ("lin_slope", "lin_offset")),
}


def _compound(kern, x, y, hp, deriv):
    """The string parser / derivative state machine, kernel-to-matrices.R:163-340.

    Restated token by token, including the `break`s and the `out <- 0` reset.
    """
    str_kern = [t for t in kern.split(" ") if t != ""]
    if str_kern[0] in ("+", "*"):
        raise ValueError("The string in 'kern' should start with a kernel not an operator.")
    s_all = ["start"] + str_kern + ["stop"]
    out = 0.0
    op = "+"

    def apply(o, a, b):
        return a + b if o == "+" else a * b

    true_deriv = False
    past_true_deriv = False
    for i in range(1, len(s_all) - 1):
        s_m, s, s_p = s_all[i - 1], s_all[i], s_all[i + 1]
        if deriv is None:
            if s in _KERNELS:
                out = apply(op, out, _KERNELS[s][0](x, y, hp, None))
            elif s in ("+", "*"):
                op = s
            else:
                raise ValueError("Incorrect character string specified in the 'kern' argument.")
        else:
            if s in _KERNELS:
                fn, names = _KERNELS[s]
                if deriv in names:
                    true_deriv = True
                    past_true_deriv = True
                if s_m == "start":
                    if s_p == "stop":
                        out = fn(x, y, hp, deriv)
                    elif s_p == "+":
                        if true_deriv:
                            out = apply(op, out, fn(x, y, hp, deriv))
                    elif s_p == "*":
                        out = apply(op, out, fn(x, y, hp, deriv if true_deriv else None))
                    else:
                        raise ValueError("Incorrect character string specified in 'kern'.")
                elif s_m == "+":
                    if s_p == "stop":
                        if true_deriv:
                            out = fn(x, y, hp, deriv)
                    elif s_p == "+":
                        if true_deriv:
                            out = fn(x, y, hp, deriv)
                            break
                    elif s_p == "*":
                        raise ValueError("The '*' operators should come before '+' operators.")
                    else:
                        raise ValueError("Incorrect character string specified in 'kern'.")
                elif s_m == "*":
                    if s_p == "stop":
                        out = apply(op, out, fn(x, y, hp, deriv if true_deriv else None))
                    elif s_p == "+":
                        if true_deriv:
                            out = apply(op, out, fn(x, y, hp, deriv))
                            break
                        elif past_true_deriv:
                            out = apply(op, out, fn(x, y, hp, None))
                            break
                        else:
                            out = 0.0
                    elif s_p == "*":
                        out = apply(op, out, fn(x, y, hp, deriv if true_deriv else None))
                    else:
                        raise ValueError("Incorrect character string specified in 'kern'.")
                else:
                    raise ValueError("Incorrect character string specified in 'kern'.")
            elif s in ("+", "*"):
                op = s
            else:
                raise ValueError("Incorrect character string specified in the 'kern' argument.")
        true_deriv = False
    return out


def convolution_kernel(x, y, hp, n_outputs, deriv=None):
    """The vectorised branch of convolution_kernel (kernels.R:50-140) for ONE latent process plus kern_to_cov's multi-output
    wrapper (kernel-to-matrices.R:88-160): x, y are matrices whose LAST column is the 0-based output index (the reference's
    Output_ID column); hp is flatten_hp_mo()'s named vector with output labels 1..O: l_t_<o>, S_t_<o>, l_u_t[, noise_<o>]
    (hyperparameters.R:348-396).  Noise (value and derivative) only on the square matrix of a set of points against itself."""
    x, y = _mat(x), _mat(y)
    O = int(n_outputs)
    same_inputs = x.shape == y.shape and np.array_equal(x, y)
    ox, oy = x[:, -1].astype(int), y[:, -1].astype(int)
    xc, yc = x[:, :-1], y[:, :-1]
    input_dim = xc.shape[1]
    if deriv is not None and deriv.startswith("noise"):           # :121-141
        mat = np.zeros((len(x), len(y)))
        if same_inputs:
            out_lab = int(deriv.split("_")[-1]) - 1
            mat[np.arange(len(x)), np.arange(len(x))] = np.where(ox == out_lab, _rexp(hp[deriv]), 0.0)
        return mat
    distance_sq = cpp_dist(xc, yc)
    l_t = np.array([hp["l_t_%d" % (o + 1)] for o in range(O)])
    S_t = np.array([hp["S_t_%d" % (o + 1)] for o in range(O)])
    l1, l2 = np.exp(l_t[ox]), np.exp(l_t[oy])
    S1, S2 = np.exp(S_t[ox]), np.exp(S_t[oy])
    l_u = _rexp(hp["l_u_t"])
    denom = np.add.outer(l1, l2) + l_u
    K = np.multiply.outer(S1, S2) / ((2 * math.pi * denom) ** (input_dim / 2)) * np.exp(-0.5 * distance_sq / denom)
    if deriv is None:
        if any(k.startswith("noise") for k in hp) and same_inputs:        # :146-157
            noise = np.array([hp["noise_%d" % (o + 1)] for o in range(O)])
            K = K + np.diag(np.exp(noise[ox]))
        return K
    common = K * ((-0.5 * input_dim / denom) + (0.5 * distance_sq / (denom ** 2)))
    if deriv == "l_u_t":
        return common * l_u
    out_lab = int(deriv.split("_")[-1]) - 1
    mask_i = np.multiply.outer(ox == out_lab, np.ones(len(y), dtype=bool))
    mask_j = np.multiply.outer(np.ones(len(x), dtype=bool), oy == out_lab)
    if deriv.startswith("l_t"):
        l_i_mat = np.repeat(l1[:, None], len(y), axis=1)
        l_j_mat = np.repeat(l2[None, :], len(x), axis=0)
        return common * (l_i_mat * mask_i + l_j_mat * mask_j)
    if deriv.startswith("S_t"):
        return K * (mask_i.astype(float) + mask_j.astype(float))
    raise ValueError("Invalid 'deriv' argument: " + deriv)


def kern_to_cov(inputs, kern, hp, deriv=None, input_2=None):
    """kernel-to-matrices.R:71-503 for character kernels (se/perio/rq/lin); "CONV<O>" stands for
    kern = convolution_kernel on inputs whose last column is the output index."""
    x = _mat(inputs)
    y = x if input_2 is None else _mat(input_2)
    if kern.startswith("CONV"):
        return convolution_kernel(x, y, hp, int(kern[4:]), deriv)
    if deriv is not None and deriv == "noise":
        return cpp_noise(x, y, hp["noise"])                       # :440-447
    mat = _compound(kern, x, y, hp, deriv)
    if isinstance(mat, float):                                    # deriv matched nothing
        mat = np.full((x.shape[0], y.shape[0]), mat)
    if "noise" in hp and deriv is None:                           # :481-487
        mat = mat + cpp_noise(x, y, hp["noise"])
    return mat


def chol_inv_jitter(mat, pen_diag, return_info=False):
    """kernel-to-matrices.R:670-680.

    diag += pen; chol2inv(chol(.)) via LAPACK dpotrf('U')/dpotri('U'); on failure
    recurse with 10*pen on the ALREADY jittered matrix (cumulative jitter, Q1).
    Returns the symmetric inverse (chol2inv fills both triangles).
    """
    mat = np.array(mat, dtype=np.float64, order="F", copy=True)
    n = mat.shape[0]
    idx = np.arange(n)
    pen = float(pen_diag)
    retries = 0
    while True:
        mat[idx, idx] += pen
        c, info = lapack.dpotrf(mat, lower=0, clean=1, overwrite_a=0)
        if info == 0:
            break
        retries += 1
        if retries > MAX_JITTER_RETRY:
            raise FloatingPointError("chol_inv_jitter: no positive definite jitter found")
        pen = 10 * pen
    inv, info2 = lapack.dpotri(c, lower=0, overwrite_c=0)
    if info2 != 0:
        raise FloatingPointError("dpotri failed")
    inv = np.triu(inv) + np.triu(inv, 1).T
    if return_info:
        return inv, retries
    return inv


def kern_to_inv(inputs, kern, hp, pen_diag=1e-10, deriv=None):
    """kernel-to-matrices.R:553-562."""
    return chol_inv_jitter(kern_to_cov(inputs, kern, hp, deriv), pen_diag)


def logdet_lu(mat):
    """determinant(mat, logarithm=TRUE)$modulus = sum log|U_ii| of dgetrf
    (likelihoods.R:37-41)."""
    lu, piv, info = lapack.dgetrf(np.asfortranarray(mat))
    return float(np.sum(np.log(np.abs(np.diag(lu)))))


def hp_names(kern, noise):
    """Column order of hp() for a character kernel (hyperparameters.R)."""
    if kern.startswith("CONV"):
        O = int(kern[4:])
        return (["l_t_%d" % (o + 1) for o in range(O)] + ["S_t_%d" % (o + 1) for o in range(O)] + ["l_u_t"] +
                (["noise_%d" % (o + 1) for o in range(O)] if noise else []))
    names = []
    for t in kern.split(" "):
        if t in _KERNELS:
            names += list(_KERNELS[t][1])
    if noise:
        names.append("noise")
    return names
