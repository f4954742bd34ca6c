set -x
for w in 1 2 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node=$w --master-addr 127.0.0.1 --master-port 2964$w scripts/debug/multirank_dbg.py /tmp/dbg 2>&1 | grep -v Warning | tail -6
done
python - <<'PY'
import numpy as np
a = np.load("/tmp/dbg_w1_r0.npz")
def tree(x):
    x = list(x)
    while len(x) > 1:
        x = [x[i] + x[i + 1] for i in range(0, len(x), 2)]
    return x[0]
U = len(a["pm"]); UU = U * U
tw = tree(a["leaf"][:, UU:])
print("world1: tree(leaf weighted) vs dumped weighted: max diff", np.max(np.abs(tw - a["pinv"][UU:])), "argmax", int(np.argmax(np.abs(tw - a["pinv"][UU:]))))
b0, b1 = np.load("/tmp/dbg_w2_r0.npz"), np.load("/tmp/dbg_w2_r1.npz")
tw2 = tree(b0["leaf"][:, UU:]) + tree(b1["leaf"][:, UU:])
print("world2: tree vs dumped weighted: max diff", np.max(np.abs(tw2 - b0["pinv"][UU:])), "tree w1 == tree w2:", np.array_equal(tw, tw2))
d = np.abs(a["pinv"][UU:] - b0["pinv"][UU:]); print("positions where w1 != w2:", np.nonzero(d)[0][:20], d[np.nonzero(d)[0][:5]])
for w in (2, 4):
    for r in range(w):
        b = np.load("/tmp/dbg_w%d_r%d.npz" % (w, r))
        U = len(a["pm"]); UU = U * U
        print(w, r, "post_inv equal", np.array_equal(a["pinv"][:UU], b["pinv"][:UU]), "weighted equal", np.array_equal(a["pinv"][UU:], b["pinv"][UU:]),
              "leaf sums equal", np.array_equal(a["leaf"][r * len(b["leaf"]):(r + 1) * len(b["leaf"])], b["leaf"]),
              "max |dw|", np.max(np.abs(a["pinv"][UU:] - b["pinv"][UU:])), "max|w|", np.max(np.abs(a["pinv"][UU:])))
        print(w, r, "pm equal", np.array_equal(a["pm"], b["pm"]), "pc equal", np.array_equal(a["pc"], b["pc"]), np.max(np.abs(a["pc"] - b["pc"])) / np.max(np.abs(a["pc"])))
PY
