#!/usr/bin/env python
"""Per-phase SM cycles of k_task_objective3 on the C2 workload (diagnostic build, -DT3_TIMING).
usage: MB_LIB=magmaclustr_b200/lib/libmagma_b200_timing.so python scripts/phase_timing.py"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from magmaclustr_b200 import _lib as L  # noqa: E402
from magmaclustr_b200.engine import Context  # noqa: E402

NAMES = ["load+hp", "cov build", "chol+logdet", "trtri", "lauum", "vectors+gemv", "S gather + B=SA", "f reduce",
         "gradient (AB o dK)", "  chol: first panel", "  chol: far panel tiles + sync", "  chol: look-ahead update (warp 0)",
         "  chol: panel factor (warp 0)", "  chol: diag inverse (warp 0)", "  chol: wait for trailing update"]


def main():
    df, prep, hp0, hpi = bench.make_inputs()
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    ctx = Context(0)
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    pm, pc, _ = ctx.e_step(k0, hp0, ki, hpi, np.zeros(ctx.U))
    lib = L.lib()
    lib.mb_debug_phase_cycles.argtypes = [C.POINTER(C.c_ulonglong), C.c_int]
    buf = (C.c_ulonglong * 16)()
    ctx.logL_GP_mod_tasks(ki, hpi, pm, pc)          # warm-up
    lib.mb_debug_phase_cycles(buf, 1)
    ctx.logL_GP_mod_tasks(ki, hpi, pm, pc)
    lib.mb_debug_phase_cycles(buf, 1)
    tot = float(sum(buf[:9]))
    for i, nme in enumerate(NAMES):
        print("%-38s %12d cycles/task  %5.1f%%" % (nme, buf[i] / ctx.n_tasks, 100.0 * buf[i] / tot))
    print("total %.0f cycles per task (CTA latency)" % (tot / ctx.n_tasks))
    ctx.close()


if __name__ == "__main__":
    main()
