"""One EM step at C2 (M = 10 000 x N = 64) through the resident-state entry point: the process to put under ncu for the
kernels other than the objective (k_task_inverse3, k_lbfgsb_advance, k_dense_cholinv_sm, k_reduce_partials ...)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
df, _ = simu_db(M=10000, N=64, seed=2024)
prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
hp0 = ini_hp("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy()
hpi = ini_hp("SE", prep.ids, True, True, 106).drop(columns="ID").to_numpy()
ctx = Context(0)
ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
ctx.set_state(k0, hp0, ki, hpi, np.zeros(ctx.U))
ll, st = ctx.em_step_resident(k0, ki, restart=True)
print("ok", ll, st)
