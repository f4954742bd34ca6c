#!/usr/bin/env python
"""Device-only throughput of the large-matrix kernels (dense_big.cu): product, Cholesky, Cholesky + inverse."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclustr_b200.engine import Context  # noqa: E402

FLOPS = {0: lambda n: 2.0 * n ** 3, 1: lambda n: n ** 3 / 3.0, 2: lambda n: float(n) ** 3}
NAMES = {0: "gemm", 1: "chol", 2: "chol+inverse"}


def main():
    sizes = [int(a) for a in sys.argv[1:]] or [512, 1024, 2048, 4096, 8192]
    ctx = Context(0)
    out = []
    for n in sizes:
        for what in (0, 1, 2):
            ms = ctx.bench_dense(what, n, reps=3 if n >= 4096 else 10)
            out.append({"n": n, "op": NAMES[what], "ms": ms, "tflops": FLOPS[what](n) / ms / 1e9})
            print(json.dumps(out[-1]), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
