"""Results must not depend on the number of GPUs (north star: convergence step count bit-exact).

The library is run with world = 1, 2 and 4 ranks as separate processes that share cuda:0, their all-reduces carried by the
caller-supplied hook (mb_set_allreduce) over torch.distributed's gloo backend.  Every rank holds the task range
mb_shard_range gives it.  Rank 0 of each world stores what it computed; the files must be BIT-identical:
e_step (post_mean, post_cov), the monitoring log-likelihood, a full train_magma fit (EM step count, log-likelihood
sequence, hp_0) and the MagmaClust steps (ve_step, vm_step incl. pi_k = colMeans(tau), elbo_monitoring_VEM).
/root/reference/R/em-magma.R:39-46 (sequential sum over tasks), R/vem-magmaclust.R:265-266, R/training.R:905-1030.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, "tests"))
import numpy as np, torch, torch.distributed as dist
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.parallel import shard_tasks

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
out = sys.argv[1]
KC = 3

class View:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 2}

def hook(ptr, count):                     # sum all-reduce of `count` device doubles through the CPU (gloo)
    t = torch.as_tensor(View(ptr, count), device="cuda:0")
    h = t.cpu()
    dist.all_reduce(h)
    t.copy_(h)
    torch.cuda.synchronize()

# the inputs were simulated ONCE by the parent (the generator draws through LAPACK, whose rounding depends on the thread
# count torchrun sets per world size): every world reads the same bits
inputs = np.load(sys.argv[2])
k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)

class Case(dict):
    pass

def load_case(tag):
    c = Case({k: inputs[tag + "_" + k] for k in ("offs", "X", "y", "grid_index", "grid_X", "hp0", "hpi")})
    c["k0"], c["ki"] = k0, ki
    return c

def shard(case, ctx):
    M = len(case["offs"]) - 1
    offs = case["offs"]
    lo, hi = shard_tasks(offs, rank, world, len(case["grid_X"]))
    r0, r1 = int(offs[lo]), int(offs[hi])
    ctx.upload_dataset((offs[lo:hi + 1] - offs[lo]).astype(np.int64), case["X"][r0:r1], case["y"][r0:r1],
                       case["grid_index"][r0:r1], case["grid_X"])
    if world > 1:
        ctx.set_allreduce(hook, rank, world)
        ctx.set_shard(lo, M)
    return lo, hi

res = {}
ctx = Context(0)
# ---- Magma: 70 ragged-grid tasks (64 leaves: some hold two tasks) ------------------------------------
case = load_case("magma")
lo, hi = shard(case, ctx)
hpi = np.ascontiguousarray(case["hpi"][lo:hi])
U = ctx.U
pm, pc, info = ctx.e_step(case["k0"], case["hp0"], case["ki"], hpi, np.zeros(U))
res["pm"], res["pc"] = pm, pc
res["ll"] = ctx.logL_monitoring(case["k0"], case["hp0"], case["ki"], hpi, np.zeros(U), pm, pc)
fit = ctx.train_magma(case["k0"], case["hp0"].copy(), case["ki"], hpi.copy(), np.zeros(U), n_iter_max=12)
res["seq"] = np.asarray(fit["seq_loglikelihood"])
res["hp0"] = np.asarray(fit["hp_0"])
res["fit_pc"] = np.asarray(fit["post_cov"])
hp_all = [None] * world
dist.all_gather_object(hp_all, np.asarray(fit["hp_i"]))
res["hpi"] = np.concatenate(hp_all, axis=0)
# shared HPs: one optimisation whose every evaluation sums over the tasks of all ranks
h0s, his = case["hp0"].copy(), np.repeat(case["hpi"][:1], hi - lo, axis=0)
h0s, his, st = ctx.m_step(case["k0"], h0s, case["ki"], his, np.zeros(U), pm, pc, shared_hp=True)
res["shared_hpi"] = his[0]
res["shared_evals"] = np.array([st[1]])
# ---- MagmaClust: ve_step / vm_step / elbo on 40 tasks, 3 clusters -------------------------------------
case = load_case("clust")
lo, hi = shard(case, ctx)
U = ctx.U
M = len(case["offs"]) - 1
rng = np.random.default_rng(4)
mix = rng.dirichlet(np.ones(KC), size=M).round(5)
hp_k = np.array([[1.0 + 0.3 * k, -1.5 - 0.2 * k] for k in range(KC)])
hpi = np.repeat(case["hpi"][:1], hi - lo, axis=0)       # shared_hp_i: the first task's block, the same on every rank
tau = np.ascontiguousarray(mix[lo:hi])
prop = mix.mean(axis=0)
vpm, vpc, tau2 = ctx.ve_step(KC, case["k0"], hp_k, case["ki"], hpi, np.zeros((KC, U)), tau, prop_mixture=prop, update=True)
res["vpm"], res["vpc"] = vpm, vpc
m_k, hk, hi_new, pi_k, st = ctx.vm_step(KC, case["k0"], hp_k.copy(), case["ki"], hpi.copy(), np.zeros((KC, U)), vpm, vpc, tau2,
                                        shared_hp_k=False, shared_hp_i=True)
res["pi_k"], res["hk"], res["hi_shared"] = np.asarray(pi_k), np.asarray(hk), np.asarray(hi_new)[0]
res["elbo"] = np.array([ctx.elbo_monitoring(KC, case["k0"], hk, case["ki"], hi_new, m_k, vpm, vpc, tau2, pi_k)])
tau_all = [None] * world
dist.all_gather_object(tau_all, np.asarray(tau2))
res["tau2"] = np.concatenate(tau_all, axis=0)
if rank == 0:
    np.savez(out, **{k: np.asarray(v) for k, v in res.items()})
ctx.close()
dist.destroy_process_group()
print("rank", rank, "done")
'''


def _inputs(tmp_path):
    from helpers import make_case
    arrs = {}
    for tag, case in (("magma", make_case(70, 12, 11)), ("clust", make_case(40, 10, 4, shared=False, K=3))):
        p = case["prep"]
        for k, v in (("offs", p.offs), ("X", p.X), ("y", p.y), ("grid_index", p.grid_index), ("grid_X", p.grid_X),
                     ("hp0", case["hp0"]), ("hpi", case["hpi"])):
            arrs[tag + "_" + k] = np.asarray(v)
    path = tmp_path / "inputs.npz"
    np.savez(path, **arrs)
    return path


def _run(world, tmp_path, port, inputs):
    script = tmp_path / "multirank.py"
    script.write_text(_SCRIPT % {"root": ROOT})
    out = tmp_path / ("world%d.npz" % world)
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(script), str(out), str(inputs)],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return np.load(out)


def test_results_bit_identical_for_world_1_2_4(tmp_path):
    inputs = _inputs(tmp_path)
    ref = _run(1, tmp_path, 29631, inputs)
    assert len(ref["seq"]) >= 2
    for world, port in ((2, 29632), (4, 29633)):
        got = _run(world, tmp_path, port, inputs)
        for key in ref.files:
            assert got[key].shape == ref[key].shape, (world, key)
            assert np.array_equal(got[key], ref[key], equal_nan=True), (world, key, float(np.max(np.abs(got[key] - ref[key]))))
