import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
t = time.time(); df, _ = simu_db(M=M, N=N, seed=2024); print("simu", time.time() - t)
t = time.time(); prep = Prepared(df.ID.values, df.Input.values, df.Output.values); print("prep", time.time() - t, "U", len(prep.grid_keys))
k0 = L.parse_kernel("SE", False); ki = L.parse_kernel("SE", True)
hp0 = ini_hp("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy()
hpi = ini_hp("SE", prep.ids, True, True, 106).drop(columns="ID").to_numpy()
ctx = Context(0)
t = time.time(); ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X); print("upload", time.time() - t)
m0 = np.zeros(ctx.U)
ctx.set_state(k0, hp0, ki, hpi, m0)
ctx.profile(True)
for it in range(4):
    ctx.flush_l2()
    w = time.time(); ctx.timer_start()
    ll, st = ctx.em_step_resident(k0, ki, restart=(it < 2))
    ms = ctx.timer_stop_ms(); w = time.time() - w
    pr = ctx.profile_read()
    flops = pr["evals"] * 5.67 * N ** 3
    print("step", it, "logL", ll, "ms", ms, "wall", w * 1e3, "stats", st, "prof", pr, "obj TFLOP/s", flops / (pr["ms"] * 1e-3) / 1e12 if pr["ms"] else None)
# phases
import ctypes as C
t = time.time(); pm, pc, info = ctx.e_step(k0, hp0, ki, hpi, m0); print("e_step wall ms", (time.time() - t) * 1e3, info)
t = time.time(); h0, hi, st = ctx.m_step(k0, hp0, ki, hpi, m0, pm, pc); print("m_step wall ms", (time.time() - t) * 1e3, st)
t = time.time(); ll = ctx.logL_monitoring(k0, h0, ki, hi, m0, pm, pc); print("monitor wall ms", (time.time() - t) * 1e3, ll)
t = time.time(); f, g, r = ctx.logL_GP_mod_tasks(ki, hpi, pm, pc); print("obj+grad all tasks wall ms", (time.time() - t) * 1e3)
t = time.time(); f, g, r = ctx.logL_GP_mod_tasks(ki, hpi, pm, pc, grad=False); print("obj only all tasks wall ms", (time.time() - t) * 1e3)
