"""Where do the GPU and oracle EM trajectories of config 1 (seed 17) part?  Per EM iteration, from the
ORACLE's current HPs: e_step both ways, then m_step both ways from the oracle's hyper-posterior."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from helpers import make_case, upload, relerr
from magmaclustr_b200.engine import Context
from oracle import magma as OM
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 17
case = make_case(10, 10, seed)
ctx = Context(0)
upload(ctx, case)
odb = case["odb"]
o0, oi = dict(case["o_hp0"]), {k: dict(v) for k, v in case["o_hpi"].items()}
n0, ni, ids = case["names_0"], case["names_i"], case["prep"].ids
m0 = np.zeros(ctx.U)
for it in range(8):
    h0 = np.array([o0[n] for n in n0]); hi = np.array([[oi[i][n] for n in ni] for i in ids])
    pm, pc = OM.e_step(odb, 0.0, "SE", "SE", o0, oi, 1e-10)
    gpm, gpc, info = ctx.e_step(case["k0"], h0, case["ki"], hi, m0)
    st = {}
    no0, noi = OM.m_step(odb, 0.0, "SE", "SE", o0, oi, pm, pc, False, 1e-10, st)
    g0, gi, gst = ctx.m_step(case["k0"], h0, case["ki"], hi, m0, pm, pc)
    fo = OM.logL_GP_mod(no0, odb.grid_X, pm, 0.0, "SE", pc, 1e-10)
    fg = OM.logL_GP_mod(dict(zip(n0, g0)), odb.grid_X, pm, 0.0, "SE", pc, 1e-10)
    f_start = OM.logL_GP_mod(o0, odb.grid_X, pm, 0.0, "SE", pc, 1e-10)
    gf, gg = ctx.logL_GP_mod(case["k0"], h0, odb.grid_X, pm, m0, pc)
    og = OM.gr_GP_mod(o0, odb.grid_X, pm, 0.0, "SE", pc, 1e-10)
    ll = OM.logL_monitoring(no0, noi, odb, 0.0, "SE", "SE", pm, pc, 1e-10)
    print("it %d  e_step relerr mean %.2e cov %.2e info %s | hp0 oracle %s (evals %d, msg %s) gpu %s (evals %d) | f_start %.6f gpu-f_start %.6f f_oracle_opt %.6f f_gpu_opt %.6f | grad oracle %s gpu %s | LL %.4f"
          % (it, relerr(gpm, pm), relerr(gpc, pc), info.tolist(), [round(no0[n], 5) for n in n0], st["hp_0"]["counts"], st["hp_0"].get("message", "?"),
             np.round(g0, 5).tolist(), gst[0], f_start, gf, fo, fg, np.round(og, 5).tolist(), np.round(gg, 5).tolist(), ll), flush=True)
    o0, oi = no0, noi
