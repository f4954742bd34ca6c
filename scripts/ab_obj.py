#!/usr/bin/env python
"""A/B timing of the per-task kernels on the C2 workload with whatever library MB_LIB points at:
one all-task objective launch (10 000 evaluations, host copies of the HP table and of f / g included in both arms),
the E step, and one resident EM step.  usage: [MB_LIB=...] python scripts/ab_obj.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from magmaclustr_b200 import _lib as L  # noqa: E402
from magmaclustr_b200.engine import Context  # noqa: E402


def main():
    df, prep, hp0, hpi = bench.make_inputs()
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    ctx = Context(0)
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    m0 = np.zeros(ctx.U)
    pm, pc, _ = ctx.e_step(k0, hp0, ki, hpi, m0)
    t_obj, t_e = [], []
    for r in range(8):
        ctx.timer_start()
        f, g, _ = ctx.logL_GP_mod_tasks(ki, hpi, pm, pc)
        t_obj.append(ctx.timer_stop_ms())
        ctx.timer_start()
        ctx.e_step(k0, hp0, ki, hpi, m0)
        t_e.append(ctx.timer_stop_ms())
    ctx.set_state(k0, hp0, ki, hpi, m0)
    t_s = []
    for r in range(14):
        ctx.flush_l2()
        ctx.timer_start()
        ll, st = ctx.em_step_resident(k0, ki, restart=True)
        t_s.append(ctx.timer_stop_ms())
    ms_inv, _ = ctx.bench_task_inverse(ki, reps=5)
    print("lib %s" % os.path.basename(L.LIB_PATH))
    print("objective launch (10000 evals + copies) min %.4f ms median %.4f ms   sum f %.10e  sum g %.10e"
          % (min(t_obj[2:]), float(np.median(t_obj[2:])), f.sum(), g.sum()))
    print("e_step (host call) min %.4f ms median %.4f" % (min(t_e[2:]), float(np.median(t_e[2:]))))
    print("inverse only %.4f ms" % ms_inv)
    print("em step min %.3f ms median %.3f   logL %.6f stats %s" % (min(t_s[2:]), float(np.median(t_s[2:])), ll, list(st)))
    ctx.close()


if __name__ == "__main__":
    main()
