"""HBM-bound kernels of the path at C2 (M = 10 000 x N = 64): device-only timings (mb_bench_hbm), and the process to
put under ncu for k_task_cov_sym / k_reduce_partials.  usage: prof_hbm.py [reps]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
M, N = 10000, 64
df, _ = simu_db(M=M, N=N, seed=2024)
prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
k0 = L.parse_kernel("SE", False); ki = L.parse_kernel("SE", True)
hp0 = ini_hp("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy()
hpi = ini_hp("SE", prep.ids, True, True, 106).drop(columns="ID").to_numpy()
ctx = Context(0)
ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
ctx.set_state(k0, hp0, ki, hpi, np.zeros(ctx.U))
peak = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
for what, name in ((0, "list_kern_to_cov"), (1, "e_step_reduce_partials")):
    ms, by = ctx.bench_hbm(what, ki, reps=reps)
    print("%-24s %.4f ms  %.1f MB  %.0f GB/s  %.1f %% of %.0f" % (name, ms, by / 1e6, by / ms / 1e6, 100 * by / ms / 1e6 / peak, peak))
