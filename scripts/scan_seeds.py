import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
M, N = 10000, 64
df, _ = simu_db(M=M, N=N, seed=2024)
prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
k0 = L.parse_kernel("SE", False); ki = L.parse_kernel("SE", True)
ctx = Context(0)
ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
m0 = np.zeros(ctx.U)
for seed in range(1, 13):
    hp0 = ini_hp("SE", ["0"], True, False, seed).drop(columns="ID").iloc[0].to_numpy().copy()
    hpi = np.ascontiguousarray(ini_hp("SE", prep.ids, True, True, seed + 100).drop(columns="ID").to_numpy())
    ll_prev = -np.inf; out = []
    for it in range(25):
        try:
            ll, st = ctx.em_step(k0, hp0, ki, hpi, m0)
        except Exception as e:
            out.append("ERR " + str(e)[-60:]); break
        conv, nev = ctx.last_task_status()
        out.append((round(ll, 3), int(st[0]), int(st[1]), int(st[2]), int((conv == 99).sum()), int((conv == 1).sum())))
        eps = (ll - ll_prev) / abs(ll)
        ll_prev = ll
        if not np.isfinite(ll) or abs(eps) < 1e-3: break
    print("seed", seed, out, flush=True)
