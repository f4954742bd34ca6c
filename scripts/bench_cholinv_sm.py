"""Phase timing of the single-CTA union-grid chol_inv_jitter kernel (k_dense_cholinv_sm) at n points (default 201)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclustr_b200.engine import Context
n = int(sys.argv[1]) if len(sys.argv) > 1 else 201
ctx = Context(0)
ms, ph = ctx.bench_dense(3, n, reps=20)
names = ["load", "cholesky", "logdet", "trtri", "lauum", "store"]
print("n=%d  %.1f us per call; cycles: %s  (sum %.0f)" % (n, ms * 1e3, ", ".join("%s %.0f" % (a, b) for a, b in zip(names, ph)), ph.sum()))
