"""precondition_hp / train_shared_gp (training.R:352-446) at C2 size: one shared HP vector for M = 10 000 x N = 64 tasks.
usage: bench_precondition.py [M] [N]"""
import sys, os, json, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
M = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
N = int(sys.argv[2]) if len(sys.argv) > 2 else 64
df, _ = simu_db(M=M, N=N, seed=2024)
prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
ki = L.parse_kernel("SE", True)
hp = ini_hp("SE", ["x"], True, True, 106).drop(columns="ID").to_numpy(dtype=np.float64)[0]
ctx = Context(0)
ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
for rep in range(2):
    t0 = time.perf_counter()
    hp_fit, st = ctx.train_shared_gp(ki, hp, mean=0.0)
    dt = time.perf_counter() - t0
print(json.dumps({"config": "train_shared_gp M=%d N=%d SE" % (M, N), "seconds": dt, "optimiser_evals": int(st[0]),
                  "convergence": int(st[1]), "task_objective_evals": int(st[2]), "hp": list(hp_fit),
                  "task_evals_per_s": int(st[2]) / dt}))
