import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
from oracle import magma as OM, kernels as OK
from oracle.lbfgsb import optim_lbfgsb
M, N = 10000, 64
df, _ = simu_db(M=M, N=N, seed=2024)
prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
k0 = L.parse_kernel("SE", False); ki = L.parse_kernel("SE", True)
hp0 = ini_hp("SE", ["0"], True, False, 1).drop(columns="ID").iloc[0].to_numpy()
hpi = ini_hp("SE", prep.ids, False, True, 2).drop(columns="ID").to_numpy()
ctx = Context(0)
ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
m0 = np.zeros(ctx.U)
pm, pc, info = ctx.e_step(k0, hp0, ki, hpi, m0); print("e_step info", info, np.isfinite(pm).all(), np.isfinite(pc).all())
h0, hi, st = ctx.m_step(k0, hp0, ki, hpi, m0, pm, pc); print("m_step", st, "hp0", h0)
conv, nev = ctx.last_task_status()
print("conv codes", np.unique(conv, return_counts=True), "nev max", nev.max(), "hist", np.bincount(nev)[:45])
bad = np.flatnonzero(conv != 0)[:6]
names = L.hp_names(ki)
for t in bad:
    sl = slice(prep.offs[t], prep.offs[t+1]); idx = prep.grid_index[sl]
    X, y = prep.X[sl], prep.y[sl]; mi = pm[idx]; S = pc[np.ix_(idx, idx)]
    def fg(x):
        h = dict(zip(names, x))
        return OM.logL_GP_mod(h, X, y, mi, "SE", S, 1e-10), OM.gr_GP_mod(h, X, y, mi, "SE", S, 1e-10)
    try:
        res = optim_lbfgsb(list(hpi[t]), fg)
        print("task", t, "gpu conv", conv[t], "nev", nev[t], "hp", hi[t], "| oracle", res["convergence"], res["counts"], res["par"], res["message"])
    except Exception as e:
        print("task", t, "gpu conv", conv[t], "nev", nev[t], "hp", hi[t], "| oracle raised", repr(e))
print("nan hp rows", np.flatnonzero(~np.isfinite(hi).all(axis=1))[:10])
try:
    ll = ctx.logL_monitoring(k0, h0, ki, hi, m0, pm, pc); print("ll", ll)
except Exception as e: print("monitor failed", e)
f, g, r = ctx.logL_GP_mod_tasks(ki, hi, pm, pc, grad=False); print("nonfinite f", np.flatnonzero(~np.isfinite(f))[:10], "retries", np.unique(r, return_counts=True))
