#!/usr/bin/env python
"""BASELINE.json config 3 at full size: train_magmaclust, K = 5 clusters, M = 20 000 tasks x N = 32 points, SE kernels,
shared_hp_i = TRUE (one shared-HP optimisation whose every evaluation reduces over all tasks), shared_hp_k = TRUE.
Times the VEM iterations (ve_step + vm_step + elbo_monitoring_VEM, training.R:1622-1758) through the C ABI with host
buffers (wall clock; the K x U x U hyper-posteriors cross PCIe every step like the reference's lists do).
usage: bench_c3.py [M] [N] [K] [iters]"""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from magmaclustr_b200 import _lib as L  # noqa: E402
from magmaclustr_b200.dataprep import Prepared  # noqa: E402
from magmaclustr_b200.engine import Context  # noqa: E402
from magmaclustr_b200.simulate import simu_db, ini_hp  # noqa: E402


def main():
    M = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    K = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
    df, gen = simu_db(M=M, N=N, K=K, seed=4)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    kk, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    hp_k = ini_hp("SE", [str(k) for k in range(K)], True, False, 7).drop(columns="ID").to_numpy(dtype=np.float64)
    hp_i = ini_hp("SE", prep.ids, True, True, 107).drop(columns="ID").to_numpy(dtype=np.float64)
    # ini_mixture: hard assignment (what the k-means initialisation returns); here the generating clusters of 80 % of the
    # tasks and a random cluster for the rest, rows in ID-string order
    rng = np.random.default_rng(4)
    order = np.argsort(np.char.encode(np.array([str(i + 1) for i in range(M)]), "utf-8"), kind="stable")
    lab = gen["clusters"][order].copy()
    flip = rng.random(M) < 0.2
    lab[flip] = rng.integers(0, K, size=int(flip.sum()))
    mixture = np.zeros((M, K)); mixture[np.arange(M), lab] = 1.0
    ctx = Context(0)
    ctx.upload_dataset(prep.offs, prep.X, prep.y, prep.grid_index, prep.grid_X)
    U = ctx.U
    m_k = np.zeros((K, U))
    prop = mixture.mean(axis=0)
    elbo_prev = -np.inf
    for it in range(1, iters + 1):
        l0 = ctx.launches
        t0 = time.perf_counter()
        pm, pc, post_mix = ctx.ve_step(K, kk, hp_k, ki, hp_i, m_k, mixture, prop, update=(it > 1))
        t1 = time.perf_counter()
        m_k, hp_k, hp_i, prop, stats = ctx.vm_step(K, kk, hp_k, ki, hp_i, m_k, pm, pc, post_mix, True, True)
        t2 = time.perf_counter()
        elbo = ctx.elbo_monitoring(K, kk, hp_k, ki, hp_i, m_k, pm, pc, post_mix, prop)
        t3 = time.perf_counter()
        mixture = post_mix
        agree = float(np.mean(np.argmax(mixture, axis=1) == gen["clusters"][order]))
        print(json.dumps({"config": "C3: train_magmaclust M=%d N=%d K=%d U=%d shared_hp_i=TRUE" % (M, N, K, U), "iter": it,
                          "ve_step_s": t1 - t0, "vm_step_s": t2 - t1, "elbo_monitoring_s": t3 - t2,
                          "vem_iteration_s": t3 - t0, "hp_k_evals": int(stats[0]), "shared_hp_i_evals": int(stats[1]),
                          "task_objective_evals": int(stats[2]), "launches": ctx.launches - l0, "elbo": elbo,
                          "elbo_increased": bool(elbo > elbo_prev), "cluster_agreement_with_generator": agree}), flush=True)
        elbo_prev = elbo
    ctx.close()


if __name__ == "__main__":
    main()
