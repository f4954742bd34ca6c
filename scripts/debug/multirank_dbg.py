import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import make_case
from magmaclustr_b200.engine import Context
from magmaclustr_b200.parallel import shard_tasks
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()

class View:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}
calls = []
def hook(ptr, count):
    t = torch.as_tensor(View(ptr, count), device="cuda:0")
    h = t.cpu()
    before = h.clone()
    dist.all_reduce(h)
    calls.append((count, float(before.abs().sum()), float(h.abs().sum())))
    t.copy_(h)
    torch.cuda.synchronize()

case = make_case(70, 12, 11)
p = case["prep"]
M = len(p.offs) - 1
lo, hi = shard_tasks(p.offs, rank, world, len(p.grid_X))
r0, r1 = int(p.offs[lo]), int(p.offs[hi])
ctx = Context(0)
ctx.upload_dataset((p.offs[lo:hi + 1] - p.offs[lo]).astype(np.int64), p.X[r0:r1], p.y[r0:r1], p.grid_index[r0:r1], p.grid_X)
if world > 1:
    ctx.set_allreduce(hook, rank, world)
    ctx.set_shard(lo, M)
hpi = np.ascontiguousarray(case["hpi"][lo:hi])
U = ctx.U
outs = []
for rep in range(2):
    pm, pc, info = ctx.e_step(case["k0"], case["hp0"], case["ki"], hpi, np.zeros(U))
    outs.append((pm.copy(), pc.copy()))
print(rank, "shard", lo, hi, "U", U, "repeat identical:", np.array_equal(outs[0][0], outs[1][0]), np.array_equal(outs[0][1], outs[1][1]), "hook calls", calls[:4])
import ctypes as C
from magmaclustr_b200 import _lib as L
lib = L.lib()
lib.mb_debug_estep_buffers.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.c_int64]
cnt = U * U + U
pinv = np.zeros(cnt)
assert lib.mb_debug_estep_buffers(ctx._h, 0, L.ptr(pinv), cnt) == 0
nl = lib.mb_estep_leaves(M, U) // world
leaf = np.zeros(nl * cnt)
assert lib.mb_debug_estep_buffers(ctx._h, 1, L.ptr(leaf), nl * cnt) == 0
# host reference of the per-leaf weighted sums
from oracle import kernels as OK
o = case["odb"]
NLg = lib.mb_estep_leaves(M, U)
starts = [(l * M + NLg - 1) // NLg for l in range(NLg + 1)]
l0 = starts.index(lo)
href = np.zeros((nl, U))
for li in range(nl):
    for t in range(starts[l0 + li], starts[l0 + li + 1]):
        i = p.ids[t]
        inv = OK.kern_to_inv(o.tasks[i][1], "SE", case["o_hpi"][i], 1e-10)
        href[li, o.index[i]] += inv @ o.tasks[i][2]
lw = leaf.reshape(nl, cnt)[:, U * U:]
bad = np.argwhere(np.abs(lw - href) > 1e-9 * np.max(np.abs(href)))
print(rank, "leaf weighted vs host: max err", np.max(np.abs(lw - href)), "n bad", len(bad), bad[:6].tolist(),
      [(starts[l0 + b[0]], starts[l0 + b[0] + 1]) for b in bad[:3]])
np.savez(sys.argv[1] + "_w%d_r%d.npz" % (world, rank), pm=outs[0][0], pc=outs[0][1], pinv=pinv, leaf=leaf.reshape(nl, cnt))
ctx.close()
dist.destroy_process_group()
