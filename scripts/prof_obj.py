import sys, os
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
hp0 = ini_hp("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy()
hpi = ini_hp("SE", prep.ids, True, True, 106).drop(columns="ID").to_numpy()
ctx = Context(0)
ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
m0 = np.zeros(ctx.U)
pm, pc, info = ctx.e_step(k0, hp0, ki, hpi, m0)
for _ in range(3):
    f, g, r = ctx.logL_GP_mod_tasks(ki, hpi, pm, pc)
print("ok", np.isfinite(f).all())
