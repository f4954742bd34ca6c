#!/usr/bin/env python
"""Host-clock timing of mb_upload_dataset on the C2 workload from pinned host buffers (the e2e leg of bench.py pays it once
per step): total, and with the pair tables switched off.  usage: [MB_LIB=...] python scripts/ab_upload.py"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from magmaclustr_b200.engine import Context  # noqa: E402


def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()


def main():
    df, prep, hp0, hpi = bench.make_inputs()
    ctx = Context(0)
    offs, X, y, gi, gX = prep.offs, pin(prep.X), pin(prep.y), pin(prep.grid_index), pin(prep.grid_X)
    for on in (1, 0, 1):
        ctx.set_option("pair_tables", on)
        ts = []
        for r in range(12):
            t0 = time.perf_counter()
            ctx.upload_dataset(offs, X, y, gi, gX)
            ts.append((time.perf_counter() - t0) * 1e3)
        print("pair_tables=%d  upload_dataset min %.3f ms median %.3f ms  (%.1f MB)" %
              (on, min(ts[2:]), float(np.median(ts[2:])), (X.nbytes + y.nbytes + gi.nbytes + offs.nbytes) / 1e6))
    ctx.close()


if __name__ == "__main__":
    main()
