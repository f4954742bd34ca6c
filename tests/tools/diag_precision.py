"""How accurate are the GPU and the oracle (LAPACK) hyper-posteriors against an extended-precision
(mpmath, 40 digits) evaluation of the same E step?  cond(K_0) ~ 1e12 makes both inexact; this tells
whether the CUDA path is as accurate as LAPACK or noisier."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import mpmath as mp
from helpers import make_case, upload, relerr
from magmaclustr_b200.engine import Context
from oracle import magma as OM, kernels as OK
mp.mp.dps = 40
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 17
case = make_case(10, 10, seed)
ctx = Context(0)
upload(ctx, case)
odb = case["odb"]
U = odb.grid_X.shape[0]
def mpinv(A):
    return mp.inverse(mp.matrix(A.tolist()))
def tonp(Mx):
    return np.array([[float(Mx[i, j]) for j in range(Mx.cols)] for i in range(Mx.rows)])
K0 = OK.kern_to_cov(odb.grid_X, "SE", case["o_hp0"])
print("U", U, "cond(K0 + 1e-10 I) = %.3e" % np.linalg.cond(K0 + 1e-10 * np.eye(U)))
inv0 = mpinv(K0 + 1e-10 * np.eye(U))
post_inv = inv0.copy()
w = mp.matrix(U, 1)
for i in odb.ids:
    Ki = OK.kern_to_cov(odb.tasks[i][1], "SE", case["o_hpi"][i]) + 1e-10 * np.eye(len(odb.tasks[i][2]))
    inv_i = mpinv(Ki)
    idx = odb.index[i]
    yi = mp.matrix(odb.tasks[i][2].tolist())
    wi = inv_i * yi
    for a, ia in enumerate(idx):
        w[int(ia)] += wi[a]
        for b, ib in enumerate(idx):
            post_inv[int(ia), int(ib)] += inv_i[a, b]
for j in range(U):
    post_inv[j, j] += mp.mpf(1e-10)
pc_x = mp.inverse(post_inv)
pm_x = pc_x * w
pc_x, pm_x = tonp(pc_x), np.array([float(v) for v in pm_x])
pm_o, pc_o = OM.e_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
pm_g, pc_g, info = ctx.e_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(U))
print("post_cov : oracle vs exact %.3e | gpu vs exact %.3e | gpu vs oracle %.3e" % (relerr(pc_o, pc_x), relerr(pc_g, pc_x), relerr(pc_g, pc_o)))
print("post_mean: oracle vs exact %.3e | gpu vs exact %.3e | gpu vs oracle %.3e" % (relerr(pm_o, pm_x), relerr(pm_g, pm_x), relerr(pm_g, pm_o)))
# inverse of K0 alone
A_o = OK.chol_inv_jitter(K0, 1e-10)
A_g, nr, jt = ctx.chol_inv_jitter(K0, 1e-10)
A_x = tonp(inv0)
print("inv_0    : oracle vs exact %.3e | gpu vs exact %.3e | gpu vs oracle %.3e  (retries gpu %d)" % (relerr(A_o, A_x), relerr(A_g, A_x), relerr(A_g, A_o), nr))
# mean-process objective and gradient at hp_0 with the exact hyper-posterior
f_o = OM.logL_GP_mod(case["o_hp0"], odb.grid_X, pm_x, 0.0, "SE", pc_x, 1e-10)
g_o = OM.gr_GP_mod(case["o_hp0"], odb.grid_X, pm_x, 0.0, "SE", pc_x, 1e-10)
f_g, g_g = ctx.logL_GP_mod(case["k0"], case["hp0"], odb.grid_X, pm_x, np.zeros(U), pc_x)
# extended precision objective: 0.5 (n log 2pi - log det A + r'Ar) + 0.5 tr(A S)
r = mp.matrix(pm_x.tolist())
Ax = inv0
quad = (r.T * Ax * r)[0]
ldA = -mp.log(mp.det(mp.matrix((K0 + 1e-10 * np.eye(U)).tolist())))
tr = sum(Ax[i, j] * mp.mpf(pc_x[i, j]) for i in range(U) for j in range(U))
f_x = float(0.5 * (U * mp.log(2 * mp.pi) - ldA + quad) + 0.5 * tr)
print("f(hp_0)  : exact %.8f | oracle %.8f (err %.2e) | gpu %.8f (err %.2e)" % (f_x, f_o, abs(f_o - f_x), f_g, abs(f_g - f_x)))
print("grad     : oracle %s | gpu %s" % (g_o, g_g))
S = mp.matrix(pc_x.tolist())
al = Ax * r
Cm = al * al.T + Ax * (S * Ax - mp.eye(U))
v, l = case["o_hp0"]["se_variance"], case["o_hp0"]["se_lengthscale"]
X1 = odb.grid_X
d2 = ((X1[:, None, :] - X1[None, :, :]) ** 2).sum(axis=2)
top = np.exp(-l) * 0.5 * d2
Kse = np.exp(v - top)
dKv, dKl = mp.matrix(Kse.tolist()), mp.matrix((Kse * top).tolist())
gx = [float(-0.5 * sum(Cm[i, j] * dK[j, i] for i in range(U) for j in range(U))) for dK in (dKv, dKl)]
print("grad     : exact %s | oracle err %s | gpu err %s" % (gx, np.abs(np.array(g_o) - gx), np.abs(np.array(g_g) - gx)))
