#!/usr/bin/env python
"""BASELINE.json config 4 shape: train_magma EM step with common_input = TRUE, M tasks on one N-point grid
(default N = U = 4096), SE kernels, shared_hp = FALSE.  Device-resident state, CUDA-event timing.
usage: bench_c4.py [M] [N] [steps]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclustr_b200 import _lib as L  # noqa: E402
from magmaclustr_b200.dataprep import Prepared  # noqa: E402
from magmaclustr_b200.engine import Context  # noqa: E402
from magmaclustr_b200.simulate import simu_db, ini_hp  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 1
    pool = np.linspace(-1.0, 1.0, N)
    t0 = time.perf_counter()
    df, _ = simu_db(M=M, N=N, common_input=True, seed=4096, pool=pool)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    hp0 = ini_hp("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy(dtype=np.float64)
    hpi = ini_hp("SE", prep.ids, False, True, 106).drop(columns="ID").to_numpy(dtype=np.float64)
    t_gen = time.perf_counter() - t0
    ctx = Context(0)
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    assert ctx.U == N
    ctx.set_state(k0, hp0, ki, hpi, np.zeros(ctx.U))
    out = []
    for s in range(steps + 1):          # first pass = warm-up (allocations)
        ctx.timer_start()
        ll, stats = ctx.em_step_resident(k0, ki, restart=True)
        ms = ctx.timer_stop_ms()
        evals = int(stats[2]) + 2 * M   # M-step evaluations + E-step inverse + monitoring per task
        out.append({"M": M, "N": N, "ms_per_em_step": ms, "task_evals": int(stats[2]), "rounds": int(stats[1]),
                    "logL": ll, "tflops_5n3": int(stats[2]) * 5.0 * N ** 3 / ms / 1e9, "warmup": s == 0,
                    "gen_s": t_gen})
        print(json.dumps(out[-1]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
