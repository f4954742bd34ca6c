import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from helpers import make_case, upload, relerr
from magmaclustr_b200.engine import Context
from oracle import kernels as OK, magma as OM
ctx = Context(0)
for (M, N, seed) in [(10, 10, 17), (40, 20, 5)]:
    case = make_case(M, N, seed); upload(ctx, case)
    odb = case["odb"]
    U = ctx.U
    K0 = OK.kern_to_cov(odb.grid_X, "SE", case["o_hp0"])
    oinv0 = OK.chol_inv_jitter(K0, 1e-10)
    ginv0, r, j = ctx.chol_inv_jitter(K0, 1e-10)
    rng = np.random.default_rng(0)
    pert = K0 * (1 + 1.1e-16 * np.triu(rng.uniform(-1, 1, K0.shape))); pert = np.triu(pert) + np.triu(pert, 1).T
    pinv0 = OK.chol_inv_jitter(pert, 1e-10)
    print("U", U, "cond", np.linalg.cond(K0 + 1e-10*np.eye(U)), "hp0", case["o_hp0"])
    print(" inv0 gpu-vs-oracle", relerr(ginv0, oinv0), " oracle sens(1ulp)", relerr(pinv0, oinv0), "retries", r)
    pm, pc, info = ctx.e_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(U))
    opm, opc = OM.e_step(odb, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
    print(" post_cov", relerr(pc, opc), "post_mean", relerr(pm, opm), "info", info)
    # dense objective on the mean process
    f, g = ctx.logL_GP_mod(case["k0"], case["hp0"], odb.grid_X, opm, np.zeros(U), opc)
    of = OM.logL_GP_mod(case["o_hp0"], odb.grid_X, opm, 0.0, "SE", opc, 1e-10)
    og = OM.gr_GP_mod(case["o_hp0"], odb.grid_X, opm, 0.0, "SE", opc, 1e-10)
    print(" f0 gpu", f, "oracle", of, "rel", abs(f-of)/abs(of)); print(" g0 gpu", g, "oracle", og)
    # pieces
    inv = oinv0
    print("  oracle pieces: logdet_lu", OK.logdet_lu(inv), " -2sumlogdiag(chol)", -2*np.sum(np.log(np.diag(np.linalg.cholesky(K0+1e-10*np.eye(U))))),
          "quad", opm@inv@opm, "trace", np.sum(inv*opc))
    inv2 = pinv0
    print("  perturbed   : logdet_lu", OK.logdet_lu(inv2), "quad", opm@inv2@opm, "trace", np.sum(inv2*opc))
    print("  gpu inv     : logdet_lu", OK.logdet_lu(ginv0), "quad", opm@ginv0@opm, "trace", np.sum(ginv0*opc))
