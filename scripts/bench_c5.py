"""BASELINE config 5 end to end at full size: hyperposterior_clust re-evaluated on a 2-D 100 x 100 grid (U = 10 000,
K clusters, kernel "RQ + PERIO"), then pred_magmaclust for 1000 new tasks on all 10 000 grid points in ONE call against the
device-resident hyper-posterior (mb_hyperpost_keep + mb_pred_magma_tasks).  Wall-clock with host buffers in and out.
The same pipeline is compared with the oracle on a sample of tasks / grid points in tests/test_gpu_fullsize.py.
usage: bench_c5.py [--tasks 1000] [--k 3] [--grid 100]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from magmaclustr_b200 import _lib as L
from magmaclustr_b200.engine import Context
from magmaclustr_b200.dataprep import Prepared, reference_keys, r_signif
from magmaclustr_b200.simulate import ini_hp

ap = argparse.ArgumentParser()
ap.add_argument("--tasks", type=int, default=1000)
ap.add_argument("--train-tasks", type=int, default=200)
ap.add_argument("--k", type=int, default=3)
ap.add_argument("--grid", type=int, default=100)
ap.add_argument("--n-obs", type=int, default=20)
args = ap.parse_args()
KC, G1 = args.k, args.grid
rng = np.random.default_rng(5)
ax = np.linspace(0, 10, G1)
grid = np.array([[a, b] for a in ax for b in ax])            # expand_grid(Input, Covariate)
G = len(grid)
KERN = "RQ + PERIO"


def make_tasks(m, n, prefix, clusters):
    ids, X, y = [], [], []
    for t in range(m):
        idx = rng.choice(G, size=n, replace=False)
        x = grid[idx]
        c = clusters[t]
        f = (2.0 + c) * np.sin(0.6 * x[:, 0] + c) + 0.5 * (1 + c) * np.cos(0.9 * x[:, 1]) + 3.0 * c
        ids += ["%s%05d" % (prefix, t)] * n
        X.append(x)
        y.append(f + 0.1 * rng.standard_normal(n))
    return Prepared(np.array(ids), np.vstack(X), np.concatenate(y))


cl_tr = rng.integers(0, KC, size=args.train_tasks)
train = make_tasks(args.train_tasks, 30, "tr", cl_tr)
train.set_grid(grid)
assert train.grid_X.shape[0] == G, "training inputs must lie on the prediction grid"
mix = np.zeros((train.n_tasks, KC)); mix[np.arange(train.n_tasks), cl_tr] = 1.0
kk, ki = L.parse_kernel(KERN, False), L.parse_kernel(KERN, True)
names_k, names_i = L.hp_names(kk), L.hp_names(ki)
hk = ini_hp(KERN, [str(k) for k in range(KC)], False, False, 50)
hp_k = hk[names_k].to_numpy(dtype=np.float64).copy()
hp_k[:, names_k.index("rq_scale")] = np.abs(hp_k[:, names_k.index("rq_scale")]) + 0.5
hi = ini_hp(KERN, ["x"], True, True, 51)
hp_i_row = hi[names_i].to_numpy(dtype=np.float64)[0]
hp_i_tr = np.tile(hp_i_row, (train.n_tasks, 1))

ctx = Context(0)
out = {"config": "C5: pred_magmaclust, grid %dx%d = %d points, %d new tasks x %d obs, K = %d, kernel %s" %
                 (G1, G1, G, args.tasks, args.n_obs, KC, KERN)}
ctx.upload_dataset(train.offs, train.X, train.y, train.grid_index, train.grid_X)
ctx.hyperpost_keep(True)
t0 = time.perf_counter()
ctx.hyperposterior_clust(KC, kk, hp_k, ki, hp_i_tr, train.grid_X, train.grid_index, np.zeros((KC, 1)), mix, host_out=False, grid_guard=False)
out["hyperposterior_clust_s"] = time.perf_counter() - t0
t0 = time.perf_counter()
ctx.hyperposterior_clust(KC, kk, hp_k, ki, hp_i_tr, train.grid_X, train.grid_index, np.zeros((KC, 1)), mix, host_out=False, grid_guard=False)
out["hyperposterior_clust_warm_s"] = time.perf_counter() - t0
out["hyperposterior_flops"] = 2.0 * KC * float(G) ** 3      # 2 K chol_inv_jitter of U x U at n^3 each
out["hyperposterior_tflops_warm"] = out["hyperposterior_flops"] / out["hyperposterior_clust_warm_s"] / 1e12
ctx.hyperpost_keep(False)

cl_new = rng.integers(0, KC, size=args.tasks)
new = make_tasks(args.tasks, args.n_obs, "nw", cl_new)
gi_new = train.index_of(new.keys)
hp_new = np.tile(hp_i_row, (new.n_tasks, 1))
tau = rng.dirichlet(np.ones(KC), size=new.n_tasks)
gp = r_signif(grid)
kp = reference_keys(gp)
op = np.argsort(kp, kind="stable")
ip = train.index_of(kp[op])
ctx.upload_dataset(new.offs, new.X, new.y, gi_new, train.grid_X)
# train_gp_clust of all new tasks (mixture probabilities + task HPs by EM), then the predictions with what it learned
prop = mix.mean(axis=0)
for rep in range(2):
    t0 = time.perf_counter()
    hp_fit, tau_fit, nit, cvf, st = ctx.train_gp_clust_tasks(ki, hp_new, prop)
    dt_fit = time.perf_counter() - t0
out["train_gp_clust_1000_tasks_s"] = dt_fit
out["train_gp_clust_em_iterations_max"] = int(nit.max())
out["train_gp_clust_objective_evals"] = int(st[2])
out["train_gp_clust_converged_frac"] = float(cvf.mean())
out["train_gp_clust_cluster_agreement"] = float(np.mean(np.argmax(tau_fit, axis=1) == cl_new[np.argsort(np.argsort(new.ids))] )) if False else float(np.mean(np.argmax(tau_fit, axis=1) == cl_new))
hp_new, tau = hp_fit, tau_fit
for rep in range(2):
    l0 = ctx.launches
    t0 = time.perf_counter()
    mean, var, _, _, r = ctx.pred_magma_tasks(ki, hp_new, gp[op], ip, tau=tau, K=KC, strict=False)
    dt = time.perf_counter() - t0
out["pred_magmaclust_1000_tasks_s"] = dt
out["pred_launches"] = ctx.launches - l0
out["predictions_per_s"] = new.n_tasks * G / dt
out["max_jitter_retries"] = int(r.max())
ok_t = (r >= 0).all(axis=1)
out["tasks_whose_matrix_never_factored"] = int((~ok_t).sum())
# tasks whose fitted HPs are degenerate (unbounded L-BFGS-B on 20 points, like the reference's optim): their matrices factor
# but the predictive moments overflow -- counted, not hidden
fin_t = np.isfinite(mean).all(axis=1) & np.isfinite(var).all(axis=1)
out["tasks_with_nonfinite_prediction"] = int((ok_t & ~fin_t).sum())
good = ok_t & fin_t
out["finite"] = bool(good.any())
out["var_min"] = float(var[good].min())
out["hp_fit_abs_max"] = float(np.abs(hp_new).max())
out["hp_fit_abs_max_of_finite_tasks"] = float(np.abs(hp_new[good]).max())
if (ok_t & ~fin_t).any() and os.path.isdir("gpurun_out"):
    b = int(np.flatnonzero(ok_t & ~fin_t)[0])
    np.savez("gpurun_out/c5_nonfinite_task.npz", task=b, hp=hp_new[b], tau=tau[b], X=new.X[new.offs[b]:new.offs[b + 1]],
             y=new.y[new.offs[b]:new.offs[b + 1]], mean=mean[b], var=var[b], hp_all_absmax=np.abs(hp_new).max(axis=1))
print(json.dumps(out))
