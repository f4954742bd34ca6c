#!/usr/bin/env python
"""One call of each large-matrix operation at size argv[1] (for ncu)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclustr_b200.engine import Context  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
what = [int(a) for a in sys.argv[2:]] or [0, 1, 2]
ctx = Context(0)
for w in what:
    print(w, ctx.bench_dense(w, n, reps=1))
ctx.close()
