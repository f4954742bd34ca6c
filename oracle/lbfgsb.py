"""R's ``optim(method = "L-BFGS-B")`` restated for the unbounded case (oracle).

TEST INFRASTRUCTURE.  The reference calls ``stats::optim(par, fn, gr,
method="L-BFGS-B", control=list(factr=1e13, maxit=25))`` with no bounds
(/root/reference/R/em-magma.R:157-168,178-189,215-226;
vem-magmaclust.R:239-245,253-259,273-279,288-294).  R's optimiser is its C
translation of the Netlib L-BFGS-B 2.x code (Zhu, Byrd, Lu, Nocedal),
``src/appl/lbfgsb.c`` + the driver in ``src/appl/optim.c`` -- a third-party,
un-vendored dependency of the reference (not in /root/reference; the R sources
are not available in this image): **parity unpinned**.  This file restates the
published algorithm with the subroutine structure of that code (mainlb, cauchy,
formk, cmprlb, subsm, lnsrlb, matupd, formt, dcsrch, dcstep, dpofa, dtrsl),
specialised to ``nbd == 0`` for every variable:

  * ``active`` marks every variable free, ``cnstnd = boxed = FALSE``;
  * when the memory is empty (``col == 0``) the generalised Cauchy point is
    ``x - g/theta`` with ``theta = 1`` (one segment, ``dtm = -f1/f2``), and the
    subspace step is skipped; the first line search starts at ``1/||d||``;
  * when ``col > 0`` the Cauchy search is skipped (``!cnstnd && col > 0``),
    ``r = -g`` (cmprlb) and subsm solves with the 2col x 2col matrix formed by
    formk;
  * ``lmm = 5``, ``pgtol = 0``, ``stpmx = 1e10``, line search
    ``ftol = 1e-3, gtol = 0.9, xtol = 0.1``;
  * optim's driver: ``iter`` counts NEW_X returns and stops with
    ``fail = 1`` once ``iter > maxit``; a non-finite ``fn`` raises.

All dot products run sequentially in index order so that the C++ / CUDA port
in the product (csrc/lbfgsb.h) can be compared step for step.
"""
import math

DBL_EPSILON = 2.220446049250313e-16


def _ddot(a, b):
    s = 0.0
    for u, v in zip(a, b):
        s += u * v
    return s


def _dpofa(a, n):
    """LINPACK dpofa on the upper triangle of a (list of lists, a[i][j]); in place."""
    for j in range(n):
        s = 0.0
        for k in range(j):
            t = a[k][j]
            for i in range(k):
                t -= a[i][k] * a[i][j]
            t = t / a[k][k]
            a[k][j] = t
            s += t * t
        s = a[j][j] - s
        if s <= 0.0:
            return j + 1
        a[j][j] = math.sqrt(s)
    return 0


def _dtrsl_upper(t, n, b, trans):
    """LINPACK dtrsl, t upper triangular; job 01 (t x = b) or 11 (t' x = b)."""
    for j in range(n):
        if t[j][j] == 0.0:
            return j + 1
    if not trans:            # job 01: solve t*x = b
        b[n - 1] = b[n - 1] / t[n - 1][n - 1]
        for jj in range(2, n + 1):
            j = n - jj
            temp = -b[j + 1]
            for i in range(j + 1):                # daxpy(j, temp, t(1,j+1), b)
                b[i] += temp * t[i][j + 1]
            b[j] = b[j] / t[j][j]
    else:                    # job 11: solve trans(t)*x = b
        b[0] = b[0] / t[0][0]
        for j in range(1, n):
            s = 0.0
            for i in range(j):
                s += t[i][j] * b[i]
            b[j] = b[j] - s
            b[j] = b[j] / t[j][j]
    return 0


class _Dcsrch:
    """MINPACK-2 dcsrch / dcstep (More'-Thuente), reverse communication."""

    def __init__(self, ftol=1e-3, gtol=0.9, xtol=0.1, stpmin=0.0, stpmax=1e10):
        self.ftol, self.gtol, self.xtol = ftol, gtol, xtol
        self.stpmin, self.stpmax = stpmin, stpmax
        self.task = "START"

    def step(self, f, g, stp):
        xtrapl, xtrapu, p5, p66 = 1.1, 4.0, 0.5, 0.66
        if self.task == "START":
            if stp < self.stpmin or stp > self.stpmax or g >= 0.0:
                self.task = "ERROR"
                return stp
            self.brackt = False
            self.stage = 1
            self.finit = f
            self.ginit = g
            self.gtest = self.ftol * self.ginit
            self.width = self.stpmax - self.stpmin
            self.width1 = self.width / p5
            self.stx, self.fx, self.gx = 0.0, self.finit, self.ginit
            self.sty, self.fy, self.gy = 0.0, self.finit, self.ginit
            self.stmin = 0.0
            self.stmax = stp + xtrapu * stp
            self.task = "FG"
            return stp
        ftest = self.finit + stp * self.gtest
        if self.stage == 1 and f <= ftest and g >= 0.0:
            self.stage = 2
        task = "FG"
        if self.brackt and (stp <= self.stmin or stp >= self.stmax):
            task = "WARNING"
        if self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax:
            task = "WARNING"
        if stp == self.stpmax and f <= ftest and g <= self.gtest:
            task = "WARNING"
        if stp == self.stpmin and (f > ftest or g >= self.gtest):
            task = "WARNING"
        if f <= ftest and abs(g) <= self.gtol * (-self.ginit):
            task = "CONVERGENCE"
        if task != "FG":
            self.task = task
            return stp
        if self.stage == 1 and f <= self.fx and f > ftest:
            fm = f - stp * self.gtest
            fxm = self.fx - self.stx * self.gtest
            fym = self.fy - self.sty * self.gtest
            gm = g - self.gtest
            gxm = self.gx - self.gtest
            gym = self.gy - self.gtest
            (self.stx, fxm, gxm, self.sty, fym, gym, stp, self.brackt) = _dcstep(
                self.stx, fxm, gxm, self.sty, fym, gym, stp, fm, gm,
                self.brackt, self.stmin, self.stmax)
            self.fx = fxm + self.stx * self.gtest
            self.fy = fym + self.sty * self.gtest
            self.gx = gxm + self.gtest
            self.gy = gym + self.gtest
        else:
            (self.stx, self.fx, self.gx, self.sty, self.fy, self.gy, stp,
             self.brackt) = _dcstep(self.stx, self.fx, self.gx, self.sty, self.fy,
                                    self.gy, stp, f, g, self.brackt, self.stmin,
                                    self.stmax)
        if self.brackt:
            if abs(self.sty - self.stx) >= p66 * self.width1:
                stp = self.stx + p5 * (self.sty - self.stx)
            self.width1 = self.width
            self.width = abs(self.sty - self.stx)
        if self.brackt:
            self.stmin = min(self.stx, self.sty)
            self.stmax = max(self.stx, self.sty)
        else:
            self.stmin = stp + xtrapl * (stp - self.stx)
            self.stmax = stp + xtrapu * (stp - self.stx)
        stp = max(stp, self.stpmin)
        stp = min(stp, self.stpmax)
        if (self.brackt and (stp <= self.stmin or stp >= self.stmax)) or \
           (self.brackt and self.stmax - self.stmin <= self.xtol * self.stmax):
            stp = self.stx
        self.task = "FG"
        return stp


def _dcstep(stx, fx, dx, sty, fy, dy, stp, fp, dp, brackt, stpmin, stpmax):
    sgnd = dp * (dx / abs(dx))
    if fp > fx:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s))
        if stp < stx:
            gamma = -gamma
        p = (gamma - dx) + theta
        q = ((gamma - dx) + gamma) + dp
        r = p / q
        stpc = stx + r * (stp - stx)
        stpq = stx + ((dx / ((fx - fp) / (stp - stx) + dx)) / 2.0) * (stp - stx)
        if abs(stpc - stx) < abs(stpq - stx):
            stpf = stpc
        else:
            stpf = stpc + (stpq - stpc) / 2.0
        brackt = True
    elif sgnd < 0.0:
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt((theta / s) * (theta / s) - (dx / s) * (dp / s))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = ((gamma - dp) + gamma) + dx
        r = p / q
        stpc = stp + r * (stx - stp)
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if abs(stpc - stp) > abs(stpq - stp):
            stpf = stpc
        else:
            stpf = stpq
        brackt = True
    elif abs(dp) < abs(dx):
        theta = 3.0 * (fx - fp) / (stp - stx) + dx + dp
        s = max(abs(theta), abs(dx), abs(dp))
        gamma = s * math.sqrt(max(0.0, (theta / s) * (theta / s) - (dx / s) * (dp / s)))
        if stp > stx:
            gamma = -gamma
        p = (gamma - dp) + theta
        q = (gamma + (dx - dp)) + gamma
        r = p / q
        if r < 0.0 and gamma != 0.0:
            stpc = stp + r * (stx - stp)
        elif stp > stx:
            stpc = stpmax
        else:
            stpc = stpmin
        stpq = stp + (dp / (dp - dx)) * (stx - stp)
        if brackt:
            if abs(stpc - stp) < abs(stpq - stp):
                stpf = stpc
            else:
                stpf = stpq
            if stp > stx:
                stpf = min(stp + 0.66 * (sty - stp), stpf)
            else:
                stpf = max(stp + 0.66 * (sty - stp), stpf)
        else:
            if abs(stpc - stp) > abs(stpq - stp):
                stpf = stpc
            else:
                stpf = stpq
            stpf = min(stpmax, stpf)
            stpf = max(stpmin, stpf)
    else:
        if brackt:
            theta = 3.0 * (fp - fy) / (sty - stp) + dy + dp
            s = max(abs(theta), abs(dy), abs(dp))
            gamma = s * math.sqrt((theta / s) * (theta / s) - (dy / s) * (dp / s))
            if stp > sty:
                gamma = -gamma
            p = (gamma - dp) + theta
            q = ((gamma - dp) + gamma) + dy
            r = p / q
            stpc = stp + r * (sty - stp)
            stpf = stpc
        elif stp > stx:
            stpf = stpmax
        else:
            stpf = stpmin
    if fp > fx:
        sty, fy, dy = stp, fp, dp
    else:
        if sgnd < 0.0:
            sty, fy, dy = stx, fx, dx
        stx, fx, dx = stp, fp, dp
    return stx, fx, dx, sty, fy, dy, stpf, brackt


class LbfgsbState:
    """Reverse-communication state machine: ``x`` is where (f, g) is wanted next.

    Usage: ``st = LbfgsbState(x0, ...)``; loop ``while st.active: st.advance(f, g)``
    with (f, g) evaluated at ``st.x``.  ``st.x`` is optim's ``$par`` at the end.
    ``fail``: 0 converged, 1 maxit reached, 51 warning (abnormal line search),
    52 error.
    """

    def __init__(self, x0, m=5, factr=1e13, pgtol=0.0, maxit=25):
        self.n = len(x0)
        self.m = m
        self.x = [float(v) for v in x0]
        self.tol = factr * DBL_EPSILON
        self.pgtol = pgtol
        self.maxit = maxit
        self.active = self.n > 0 and maxit > 0
        self.fail = 0
        self.iter = 0          # mainlb iteration counter
        self.niter = 0         # optim.c NEW_X counter
        self.nfgv = 0
        self.f = float("nan")
        self.g = [0.0] * self.n
        self._reset_memory()
        self.stage = "FG_START"
        self.msg = ""

    def _reset_memory(self):
        self.col = 0
        self.theta = 1.0
        self.ws = []   # stored s vectors (oldest first)
        self.wy = []
        self.sy = []   # sy[i][j] = s_i'y_j (lower triangle used), rows/cols over memory
        self.ss = []
        self.updatd = False

    # -- pieces of mainlb -------------------------------------------------
    def _search_direction(self):
        """L222..L555: returns False when the memory had to be refreshed."""
        n, x, g = self.n, self.x, self.g
        while True:
            if self.col == 0:
                # cauchy(): d = -g, one unbounded segment, dtm = -f1/f2, theta = 1
                f1 = 0.0
                d = [0.0] * n
                for i in range(n):
                    neggi = -g[i]
                    d[i] = neggi
                    f1 -= neggi * neggi
                f2 = -self.theta * f1
                dtm = -f1 / f2
                if dtm <= 0.0:
                    dtm = 0.0
                z = [x[i] + dtm * d[i] for i in range(n)]
                self.z = z
            else:
                col, theta = self.col, self.theta
                z = list(x)
                # formk(): WN for all variables free (L_a = 0, S'AA'S = 0)
                m2 = 2 * col
                wn = [[0.0] * m2 for _ in range(m2)]
                for iy in range(col):
                    for jy in range(iy + 1):
                        wn[jy][iy] = _ddot(self.wy[iy], self.wy[jy]) / theta
                    for jy in range(iy, col):
                        # R_z: s_iy' y_jy for iy <= jy, stored at wn[jy][col+iy]
                        wn[jy][col + iy] = _ddot(self.ws[iy], self.wy[jy])
                    wn[iy][iy] += self.sy[iy][iy]
                info = _dpofa(wn, col)
                if info == 0:
                    for js in range(col, m2):
                        colv = [wn[i][js] for i in range(col)]
                        _dtrsl_upper(wn, col, colv, trans=True)
                        for i in range(col):
                            wn[i][js] = colv[i]
                    for is_ in range(col, m2):
                        for js in range(is_, m2):
                            s = 0.0
                            for k in range(col):
                                s += wn[k][is_] * wn[k][js]
                            wn[is_][js] += s
                    sub = [[wn[col + i][col + j] for j in range(col)] for i in range(col)]
                    info = _dpofa(sub, col)
                    for i in range(col):
                        for j in range(col):
                            wn[col + i][col + j] = sub[i][j]
                if info != 0:
                    self._reset_memory()
                    continue
                # cmprlb(): r = -g ; subsm()
                r = [-g[i] for i in range(n)]
                wv = [0.0] * m2
                for i in range(col):
                    t1 = 0.0
                    t2 = 0.0
                    for k in range(n):
                        t1 += self.wy[i][k] * r[k]
                        t2 += self.ws[i][k] * r[k]
                    wv[i] = t1
                    wv[col + i] = theta * t2
                info = _dtrsl_upper(wn, m2, wv, trans=True)
                if info == 0:
                    for i in range(col):
                        wv[i] = -wv[i]
                    info = _dtrsl_upper(wn, m2, wv, trans=False)
                if info != 0:
                    self._reset_memory()
                    continue
                for jy in range(col):
                    js = col + jy
                    for i in range(n):
                        r[i] += self.wy[jy][i] * wv[jy] / theta + self.ws[jy][i] * wv[js]
                for i in range(n):
                    r[i] /= theta
                for i in range(n):
                    z[i] += 1.0 * r[i]
                self.z = z
            self.d = [self.z[i] - x[i] for i in range(n)]
            return

    def _lnsrlb_start(self):
        d = self.d
        self.dtd = _ddot(d, d)
        self.dnorm = math.sqrt(self.dtd)
        self.stpmx = 1e10
        if self.iter == 0:
            self.stp = min(1.0 / self.dnorm, self.stpmx) if self.dnorm > 0 else self.stpmx
        else:
            self.stp = 1.0
        self.t = list(self.x)
        self.r = list(self.g)
        self.fold = self.f
        self.ifun = 0
        self.iback = 0
        self.ls = _Dcsrch(stpmax=self.stpmx)

    def _lnsrlb(self):
        """Returns 'FG' (x updated, evaluate), 'NEW_X', or 'FAIL' (info != 0)."""
        self.gd = _ddot(self.g, self.d)
        if self.ifun == 0:
            self.gdold = self.gd
            if self.gd >= 0.0:
                return "FAIL"
        self.stp = self.ls.step(self.f, self.gd, self.stp)
        if self.ls.task == "ERROR":
            return "FAIL"
        if self.ls.task not in ("CONVERGENCE", "WARNING"):
            self.ifun += 1
            self.nfgv += 1
            self.iback = self.ifun - 1
            if self.stp == 1.0:
                self.x = list(self.z)
            else:
                self.x = [self.stp * self.d[i] + self.t[i] for i in range(self.n)]
            return "FG"
        return "NEW_X"

    def _begin_iteration(self):
        """L222 -> L666 first entry; may loop on restarts.  Sets self.stage."""
        while True:
            self._search_direction()
            self._lnsrlb_start()
            res = self._lnsrlb()
            if res == "FG":
                self.stage = "FG_LN"
                return
            # res == FAIL at ifun == 0 (gd >= 0) -- NEW_X cannot happen here
            if not self._line_search_failed(info_nonzero=True):
                return

    def _line_search_failed(self, info_nonzero):
        """Restore the previous iterate; abnormal end or restart (mainlb after L666).
        Returns True when the iteration must be restarted from L222."""
        self.x = list(self.t)
        self.g = list(self.r)
        self.f = self.fold
        if self.col == 0:
            self.msg = "ABNORMAL_TERMINATION_IN_LNSRCH"
            self.fail = 52
            self.active = False
            return False
        self._reset_memory()
        return True

    def advance(self, f, g):
        """Feed (f, g) evaluated at self.x."""
        if not self.active:
            return
        if not math.isfinite(f):
            raise FloatingPointError("L-BFGS-B needs finite values of 'fn'")
        self.f = float(f)
        self.g = [float(v) for v in g]
        if self.stage == "FG_START":
            self.nfgv = 1
            sbgnrm = max(abs(v) for v in self.g)
            if sbgnrm <= self.pgtol:
                self.msg = "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL"
                self.active = False
                return
            self._begin_iteration()
            return
        # stage FG_LN: continue the line search
        res = self._lnsrlb()
        if res == "FG" and self.iback < 20:
            return
        if res == "FAIL" or (res == "FG" and self.iback >= 20):
            if self._line_search_failed(info_nonzero=(res == "FAIL")):
                self._begin_iteration()
            return
        # NEW_X
        self.iter += 1
        sbgnrm = max(abs(v) for v in self.g)
        self.niter += 1
        if self.niter > self.maxit:
            self.fail = 1
            self.msg = "NEW_X"
            self.active = False
            return
        if sbgnrm <= self.pgtol:
            self.msg = "CONVERGENCE: NORM OF PROJECTED GRADIENT <= PGTOL"
            self.active = False
            return
        ddum = max(abs(self.fold), abs(self.f), 1.0)
        if self.fold - self.f <= self.tol * ddum:
            self.msg = "CONVERGENCE: REL_REDUCTION_OF_F <= FACTR*EPSMCH"
            self.active = False
            return
        n = self.n
        r = [self.g[i] - self.r[i] for i in range(n)]
        rr = _ddot(r, r)
        d = self.d
        if self.stp == 1.0:
            dr = self.gd - self.gdold
            ddum = -self.gdold
        else:
            dr = (self.gd - self.gdold) * self.stp
            d = [self.stp * v for v in d]
            ddum = -self.gdold * self.stp
        if dr <= DBL_EPSILON * ddum:
            self.updatd = False
        else:
            self.updatd = True
            self._matupd(d, r, rr, dr)
            if not self._formt():
                self._reset_memory()
        self._begin_iteration()

    def _matupd(self, d, r, rr, dr):
        m = self.m
        if self.col == m:
            self.ws.pop(0)
            self.wy.pop(0)
            self.sy = [row[1:] for row in self.sy[1:]]
            self.ss = [row[1:] for row in self.ss[1:]]
            self.col -= 1
        col = self.col
        # last row of SY (s_new'y_j) and last column of SS (s_j's_new)
        new_sy_row = [_ddot(d, self.wy[j]) for j in range(col)]
        new_ss_col = [_ddot(self.ws[j], d) for j in range(col)]
        self.ws.append(list(d))
        self.wy.append(list(r))
        for i in range(col):
            self.sy[i].append(0.0)          # upper part of SY is never read
            self.ss[i].append(new_ss_col[i])
        ss_diag = self.dtd if self.stp == 1.0 else self.stp * self.stp * self.dtd
        self.sy.append(new_sy_row + [dr])
        self.ss.append([0.0] * col + [ss_diag])
        self.col = col + 1
        self.theta = rr / dr

    def _formt(self):
        """T = theta*SS + L D^-1 L' (upper), Cholesky by dpofa; False if it fails."""
        col, theta = self.col, self.theta
        wt = [[0.0] * col for _ in range(col)]
        for j in range(col):
            wt[0][j] = theta * self.ss[0][j]
        for i in range(1, col):
            for j in range(i, col):
                k1 = min(i, j)
                ddum = 0.0
                for k in range(k1):
                    ddum += self.sy[i][k] * self.sy[j][k] / self.sy[k][k]
                wt[i][j] = ddum + theta * self.ss[i][j]
        return _dpofa(wt, col) == 0


def optim_lbfgsb(par, fn_gr, m=5, factr=1e13, pgtol=0.0, maxit=25):
    """``optim(par, fn, gr, method="L-BFGS-B")``; fn_gr(x) -> (f, g).
    Returns dict(par, value, counts, convergence, message, iterations)."""
    st = LbfgsbState(par, m=m, factr=factr, pgtol=pgtol, maxit=maxit)
    f = None
    while st.active:
        f, g = fn_gr(list(st.x))
        st.advance(f, g)
    return {"par": list(st.x), "value": st.f, "counts": st.nfgv,
            "convergence": st.fail, "message": st.msg, "iterations": st.niter}
