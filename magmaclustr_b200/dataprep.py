"""Host data path above the C ABI: the reference's table preparation, vectorised.

Mirrors /root/reference/R/training.R:690-720 (and prediction.R:534-567, 937-976):
``signif`` to 6 digits, ``Reference`` = input columns pasted with ":" through
``as.character(<double>)`` (15 significant digits, fixed unless scientific is shorter),
rows ordered by ID then Reference and the union grid by Reference in byte-wise (C-locale)
order (dplyr::arrange).  The result is the CSR layout of include/magma_b200.h with
``grid_index = match(Reference, all_input) - 1``.  SURVEY.md Appendix B.
"""
import math

import numpy as np


def r_signif(x, digits=6):
    """R's signif() (nmath/fprec.c) on an array."""
    x = np.asarray(x, dtype=np.float64)
    out = x.copy()
    ok = np.isfinite(x) & (x != 0)
    if not ok.any():
        return out
    ax = np.abs(x[ok])
    e10 = (digits - 1 - np.floor(np.log10(ax))).astype(np.int64)
    pos = e10 > 0
    res = np.empty_like(ax)
    extra = np.power(10.0, np.maximum(e10 - 308, 0).astype(np.float64))   # max10e split
    e10 = np.minimum(e10, 308)
    p = np.power(10.0, np.where(pos, e10, 0).astype(np.float64))
    q = np.power(10.0, np.where(pos, 0, -e10).astype(np.float64))
    res[pos] = np.rint((ax[pos] * p[pos]) * extra[pos]) / p[pos] / extra[pos]
    res[~pos] = np.rint(ax[~pos] / q[~pos]) * q[~pos]
    out[ok] = np.sign(x[ok]) * res
    return out


def r_as_character(v):
    """as.character(<double>) for one value."""
    v = float(v)
    if math.isnan(v):
        return "NA"
    if math.isinf(v):
        return "Inf" if v > 0 else "-Inf"
    if v == 0.0:
        return "0"
    neg = v < 0
    ref = float("%.14e" % v)
    nsig = 15
    for n in range(1, 16):
        if float("%.*e" % (n - 1, v)) == ref:
            nsig = n
            break
    mant, exp = ("%.*e" % (nsig - 1, abs(v))).split("e")
    e = int(exp)
    digits = mant.replace(".", "")
    wexp = 2 if abs(e) < 100 else 3
    w_sci = neg + (nsig + 1 if nsig > 1 else 1) + 2 + wexp
    left = e + 1 if e >= 0 else 1
    rgt = max(0, nsig - e - 1)
    w_fix = neg + left + (rgt + 1 if rgt > 0 else 0)
    if w_fix <= w_sci:
        if e >= 0:
            s = digits[: e + 1].ljust(e + 1, "0") + ("." + digits[e + 1:] if rgt > 0 else "")
        else:
            s = "0." + "0" * (-e - 1) + digits
    else:
        s = mant + "e" + ("-" if e < 0 else "+") + ("%0*d" % (wexp, abs(e)))
    return ("-" if neg else "") + s


def reference_keys(X):
    """Reference strings (bytes dtype, so comparisons are byte-wise) of the rows of X."""
    X = np.asarray(X, dtype=np.float64)
    if X.ndim == 1:
        X = X[:, None]
    cols = []
    for c in range(X.shape[1]):
        u, inv = np.unique(X[:, c], return_inverse=True)
        s = np.array([r_as_character(v).encode() for v in u], dtype="S")
        cols.append(s[inv])
    key = cols[0]
    for c in cols[1:]:
        key = np.char.add(np.char.add(key, b":"), c)
    return key


class Prepared:
    """CSR view of a training table."""

    def __init__(self, ids, X, y, digits=6):
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X[:, None]
        y = np.asarray(y, dtype=np.float64)
        keep = np.isfinite(X).all(axis=1) & np.isfinite(y)            # tidyr::drop_na
        ids = np.asarray([str(i) for i in ids])[keep]
        X, y = r_signif(X[keep], digits), y[keep]
        keys = reference_keys(X)
        idb = np.char.encode(ids, "utf-8")
        order = np.lexsort((keys, idb))                                # by ID, then Reference
        self.X, self.y, self.keys = np.ascontiguousarray(X[order]), y[order].copy(), keys[order]
        idb = idb[order]
        change = np.flatnonzero(np.r_[True, idb[1:] != idb[:-1]])
        self.ids = [b.decode() for b in idb[change]]
        self.offs = np.r_[change, len(idb)].astype(np.int64)
        # duplicates inside a task (training.R:703-714)
        same = (idb[1:] == idb[:-1]) & (self.keys[1:] == self.keys[:-1])
        if same.any():
            raise ValueError("At least one individual have several Outputs on the same grid point.")
        self.set_grid(None)

    def set_grid(self, grid_X, digits=6):
        """Union grid = observed keys (U grid_X), sorted by Reference."""
        allX, allk = self.X, self.keys
        if grid_X is not None:
            g = np.asarray(grid_X, dtype=np.float64)
            if g.ndim == 1:
                g = g[:, None]
            g = r_signif(g, digits)
            allX = np.vstack([allX, g])
            allk = np.concatenate([allk, reference_keys(g)])
        gk, first = np.unique(allk, return_index=True)                 # bytes: sorted byte-wise
        self.grid_keys = gk
        self.grid_X = np.ascontiguousarray(allX[first])
        self.grid_index = np.searchsorted(gk, self.keys).astype(np.int32)
        return self

    @property
    def n_tasks(self):
        return len(self.ids)

    def index_of(self, keys):
        """0-based grid positions of Reference keys (all must be present)."""
        pos = np.searchsorted(self.grid_keys, keys)
        pos = np.clip(pos, 0, len(self.grid_keys) - 1)
        if not np.all(self.grid_keys[pos] == keys):
            raise KeyError("some inputs are not on the grid")
        return pos.astype(np.int32)


def prepare_rows_native(ids, X, y=None, grid_extra=None, digits=6):
    """The same preparation done by the library's compiled host routine (mb_prepare_rows, csrc/hostprep.cu): what an R
    shim would call instead of the tidyverse pipeline of training.R:690-720.  Returns a dict with the fields of `Prepared`
    (X, y, offs, grid_index, grid_X, perm).  Rows with missing values must already be dropped (tidyr::drop_na)."""
    import ctypes as C
    from . import _lib as L
    X = np.ascontiguousarray(np.atleast_2d(np.asarray(X, dtype=np.float64).T).T)
    n, d = X.shape
    yv = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
    idb = [str(i).encode("utf-8") for i in ids]
    arr = (C.c_char_p * n)(*idb)
    ge = None if grid_extra is None else np.ascontiguousarray(np.atleast_2d(np.asarray(grid_extra, dtype=np.float64).T).T)
    ne = 0 if ge is None else ge.shape[0]
    perm = np.zeros(n, dtype=np.int64)
    Xo = np.zeros((n, d))
    yo = np.zeros(n)
    nt = C.c_int64(0)
    offs = np.zeros(n + 1, dtype=np.int64)
    gi = np.zeros(n, dtype=np.int32)
    ng = C.c_int32(0)
    gX = np.zeros((n + ne, d))
    L.check(L.lib().mb_prepare_rows(n, d, L.ptr(X), L.ptr(yv) if yv is not None else None, arr, int(digits),
                                    L.ptr(ge) if ge is not None else None, ne, L.ptr(perm), L.ptr(Xo),
                                    L.ptr(yo) if yv is not None else None, C.byref(nt), L.ptr(offs), L.ptr(gi), C.byref(ng),
                                    L.ptr(gX)))
    return {"X": Xo, "y": yo if yv is not None else None, "offs": offs[:nt.value + 1].copy(), "grid_index": gi,
            "grid_X": gX[:ng.value].copy(), "perm": perm}
