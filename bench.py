#!/usr/bin/env python
"""bench.py -- train_magma EM steps/s at BASELINE config 2 (M = 10 000 tasks x N = 64 points,
SE kernel, shared_hp = FALSE, U = 201), measured through the C-ABI library.

A step = ONE EM iteration (e_step + m_step + logL_monitoring, /root/reference/R/training.R:905-1030)
started from the pristine initial hyper-parameters, so every timed step does identical work.

  value   : device-resident inputs (dataset + HP tables already in HBM), CUDA events on the
            library's stream, max over ranks.
  e2e     : the same step through mb_upload_dataset + mb_em_step with HOST buffers (pinned):
            host->device copy of the table and HP arrays and device->host read of the new
            HPs / log-likelihood inside the timed region.
  roofline: dominant kernel k_task_objective3, algorithmic flops = objective evaluations x
            5 N^3, divided by its summed event time; peak = cuBLAS DGEMM
            8192^3 measured in the same run (MEASURED_PEAKS.json has no FP64 figure).
  cpu_baseline / --impl reference: the oracle (CPU restatement of the reference's algorithm,
            all host cores) on a bounded task sample, scaled linearly in M.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

M_TASKS, N_POINTS = 10000, 64
DATA_SEED, HP0_SEED, HPI_SEED = 2024, 6, 106
# dram__bytes_read.sum + dram__bytes_write.sum of one full-grid launch of the dominant kernel (10 000 evaluations), from the
# ncu --set full capture summarised in profiles/r2_final_k_task_objective3_ncu.txt: 100.65 MB read + 3.71 MB written.  Of the
# reads 86 MB are the tasks' pair tables (8.6 KB per task: index words + distinct squared distances, pairs3.cu) -- they
# replace 4/5 of the exponentials; 19.0 MB without them, 145.5 MB in round 1 (write-back of the kernel-value cache).
NCU_TRAFFIC_BYTES_PER_LAUNCH = 104357120
METRIC = "train_magma EM steps/sec at M=10k,N=64"
FLOPS_PER_EVAL = 5.0  # x N^3: chol+inverse N^3, S A and A (S A) 4 N^3; log-det read off the Cholesky diagonal


def simu_db_c4(m, n):
    from magmaclustr_b200.simulate import simu_db
    return simu_db(M=m, N=n, common_input=True, seed=4096, pool=np.linspace(-1.0, 1.0, n))


def Prepared4(*a):
    from magmaclustr_b200.dataprep import Prepared
    return Prepared(*a)


def ini_hp4(*a):
    from magmaclustr_b200.simulate import ini_hp
    return ini_hp(*a)


def make_inputs(broadcast=False):
    """Dataset + initial HPs of config 2.  ini_hp_i: one draw PER TASK from hp()'s ranges (shared_hp = FALSE,
    hyperparameters.R:70-77); broadcast = True repeats ONE draw for every task (an ini_hp_i without an ID
    column, training.R:831-833) -- the variant round 1 measured, reported beside the headline."""
    from magmaclustr_b200.dataprep import Prepared
    from magmaclustr_b200.simulate import simu_db, ini_hp
    df, _ = simu_db(M=M_TASKS, N=N_POINTS, seed=DATA_SEED)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    hp0 = ini_hp("SE", ["0"], True, False, HP0_SEED).drop(columns="ID").iloc[0].to_numpy(dtype=np.float64)
    hpi = ini_hp("SE", prep.ids, broadcast, True, HPI_SEED).drop(columns="ID").to_numpy(dtype=np.float64)
    return df, prep, np.ascontiguousarray(hp0), np.ascontiguousarray(hpi)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, index):
        self.rows = []
        self.proc = None
        self.index = index

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                for n, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": float(max(mx)) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def dgemm_peak_tflops():
    import torch
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device="cuda")
    b = torch.randn(n, n, dtype=torch.float64, device="cuda")
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    del a, b
    torch.cuda.empty_cache()
    return 2.0 * n ** 3 / (best * 1e-3) / 1e12


def oracle_sample_step(df, m_sample, procs):
    """One EM step of the oracle on the first m_sample tasks (ID order), from the bench's own initial HPs; returns
    (seconds, evals, logL)."""
    from oracle import magma as OM
    from oracle import parallel as OP
    from magmaclustr_b200.simulate import ini_hp
    all_ids = sorted(df.ID.unique(), key=lambda s: s.encode())
    ids = all_ids[:m_sample]
    sub = df[df.ID.isin(ids)] if m_sample < len(all_ids) else df
    db = OM.Dataset(sub.ID.values, sub.Input.values, sub.Output.values)
    h0 = ini_hp("SE", ["0"], True, False, HP0_SEED).drop(columns="ID").iloc[0].to_dict()
    tab = ini_hp("SE", all_ids, False, True, HPI_SEED).set_index("ID")      # the same per-task draws as the GPU arm
    hi = {i: tab.loc[i].to_dict() for i in db.ids}
    t = time.perf_counter()
    _, _, ll, evals = OP.em_step(db, 0.0, "SE", "SE", h0, hi, 1e-10, procs)
    return time.perf_counter() - t, evals, ll


def run_reference(args, rank):
    """--impl reference: the reference's CPU algorithm (oracle port; R is not installable in this
    image) on all host cores, bounded sample per step, scaled linearly in M."""
    if rank != 0:
        return
    cores = os.cpu_count()
    df, _, _, _ = make_inputs()
    n_runs = args.warmup + args.steps
    budget_s = 200.0                        # the whole arm has to end within a few minutes
    if args.cpu_sample:
        m_sample = args.cpu_sample
    else:                                   # ALL tasks per step when that fits the budget, else the largest sample that does
        dt0, _, _ = oracle_sample_step(df, 1024, cores)
        est_full = dt0 * M_TASKS / 1024.0
        m_sample = M_TASKS if est_full * n_runs <= budget_s else int(max(1024, min(M_TASKS, M_TASKS * budget_s / (est_full * n_runs)))) // 256 * 256
    secs = []
    for s in range(n_runs):
        dt, evals, ll = oracle_sample_step(df, m_sample, cores)
        if s >= args.warmup:
            secs.append(dt)
    per_step_full = float(np.mean(secs)) * (M_TASKS / m_sample)
    value = 1.0 / per_step_full
    sample = ("%s of %d tasks per step (one EM step = e_step + m_step + logL_monitoring), %.2f s measured per step%s"
              % ("ALL" if m_sample == M_TASKS else "first %d" % m_sample, M_TASKS, float(np.mean(secs)),
                 "" if m_sample == M_TASKS else ", scaled linearly in M (DESCRIPTION:23-24)"))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "EM steps/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": per_step_full * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "C2: train_magma EM step (e_step + m_step + logL_monitoring), M=10000 tasks x N=64, SE kernel, "
                                   "shared_hp=FALSE, U=201, seeds data=2024 hp_0=6 hp_i=106 (one hp_i draw per task)",
                       "tasks_per_step": int(m_sample), "seconds_per_step_measured": float(np.mean(secs))},
            "cpu_baseline": {"value": value, "unit": "EM steps/s", "cores": cores, "kind": "port",
                             "sample": sample},
            "e2e": {"value": value, "unit": "EM steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def run_c4(args, rank, world, local_rank):
    """--config c4: BASELINE config 4 (common_input = TRUE, M = 256 tasks on one N = U = 4096-point grid, SE kernels,
    shared_hp = FALSE), the tasks sharded over the ranks.  Every task is a pipeline of 4096 x 4096 factorisations and
    products on HBM-resident matrices; a step is one EM iteration from the initial hyper-parameters."""
    import torch
    from magmaclustr_b200 import _lib as L
    from magmaclustr_b200.engine import Context
    from magmaclustr_b200.parallel import shard_tasks, init_library_nccl
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    m4, n4 = args.c4_tasks, args.c4_points
    from magmaclustr_b200.simulate import simu_db_spectral
    # random-Fourier-feature draws of simu_db's model: Cholesky draws of 256 covariances of 4096 points take minutes of CPU
    df4, _ = simu_db_spectral(M=m4, N=n4, common_input=True, seed=4096, pool=np.linspace(-1.0, 1.0, n4))
    prep = Prepared4(df4.ID.values, df4.Input.values, df4.Output.values)
    hp0 = ini_hp4("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy(dtype=np.float64)
    hpi = ini_hp4("SE", prep.ids, False, True, 106).drop(columns="ID").to_numpy(dtype=np.float64)
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    ctx = Context(local_rank)
    lo, hi_ = shard_tasks(prep.offs, rank, world, len(prep.grid_X))
    offs = (prep.offs[lo:hi_ + 1] - prep.offs[lo]).astype(np.int64)
    r0, r1 = int(prep.offs[lo]), int(prep.offs[hi_])
    X_h, y_h, gi_h = (np.ascontiguousarray(prep.X[r0:r1]), np.ascontiguousarray(prep.y[r0:r1]),
                      np.ascontiguousarray(prep.grid_index[r0:r1]))
    hpi_loc = np.ascontiguousarray(hpi[lo:hi_])
    if world > 1:
        init_library_nccl(ctx, dist, rank, world)
        ctx.set_shard(lo, len(prep.offs) - 1)
    ctx.upload_dataset(offs, X_h, y_h, gi_h, prep.grid_X)
    m0 = np.zeros(ctx.U)
    ctx.set_state(k0, hp0, ki, hpi_loc, m0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    for _ in range(args.warmup):
        ctx.em_step_resident(k0, ki, restart=True)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ms, stats, logL = [], None, None
    for _ in range(args.steps):
        barrier()                           # operands (256 x 134 MB of factors and products per pass) exceed L2 many times over
        ctx.timer_start()
        logL, stats = ctx.em_step_resident(k0, ki, restart=True)
        ms.append(ctx.timer_stop_ms())
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - launches0
    total_ms = float(np.sum(ms))
    evals = float(stats[2])
    if world > 1:
        t = torch.tensor([total_ms, 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t[0].item())
        t = torch.tensor([evals], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        evals = float(t[0].item())
    # end to end: host buffers in, host buffers out
    e2e_ms = []
    h0_h, hi_h = hp0.copy(), hpi_loc.copy()
    for s in range(1 + args.steps):
        h0_h[:] = hp0
        hi_h[:] = hpi_loc
        barrier()
        t0 = time.perf_counter()
        ctx.upload_dataset(offs, X_h, y_h, gi_h, prep.grid_X)
        ctx.em_step(k0, h0_h, ki, hi_h, m0)
        dt = (time.perf_counter() - t0) * 1e3
        barrier()
        if s >= 1:
            e2e_ms.append(dt)
    e2e_total = float(np.sum(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak = dgemm_peak_tflops()
    ms_g = ctx.bench_dense(0, n4, reps=3)
    gemm_tf = 2.0 * float(n4) ** 3 / ms_g / 1e9
    step_tf = evals * FLOPS_PER_EVAL * float(n4) ** 3 / (total_ms / args.steps) / 1e9
    line = {
        "metric": "train_magma EM steps/sec at M=%d,N=%d (config 4)" % (m4, n4), "value": args.steps / (total_ms * 1e-3),
        "unit": "EM steps/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic (spectral draws of simu_db's model)",
        "config": {"workload": "C4: train_magma EM step, common_input=TRUE, M=%d tasks on one N=U=%d grid, SE kernels, "
                               "shared_hp=FALSE" % (m4, n4),
                   "l2": "inputs larger than L2 (every task streams 134 MB matrices)", "tasks_per_rank": int(hi_ - lo),
                   "task_evals_per_step": int(evals), "hp0_evals": int(stats[0]), "logL": logL},
        "clocks": clocks,
        "e2e": {"value": args.steps / (e2e_total * 1e-3), "unit": "EM steps/s",
                "h2d_bytes_per_step": int(offs.nbytes + X_h.nbytes + y_h.nbytes + gi_h.nbytes + prep.grid_X.nbytes + hi_h.nbytes + m0.nbytes),
                "d2h_bytes_per_step": int(hi_h.nbytes + 8)},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "k_big_gemm (128 x 128 DMMA tiles)", "achieved": gemm_tf, "peak": peak,
                     "unit": "TFLOP/s", "frac": gemm_tf / peak if peak else None, "traffic": None,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run",
                     "whole_step_tflops_at_5n3_per_eval": step_tf,
                     "whole_step_frac_of_dgemm": step_tf / peak / world if peak else None},
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=None)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-fit", action="store_true")
    ap.add_argument("--no-c4", action="store_true")
    ap.add_argument("--no-variants", action="store_true")
    ap.add_argument("--config", default="c2", choices=["c2", "c4"])
    ap.add_argument("--c4-tasks", type=int, default=256)
    ap.add_argument("--c4-points", type=int, default=4096)
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # one BLAS thread per process from the very start (the pool already uses every core): OpenBLAS sizes its thread pool
        # when numpy is imported, so the variables have to be in the environment before this interpreter starts -- torchrun
        # sets OMP_NUM_THREADS=1 itself, a plain `python bench.py --impl reference` did not, and the two differed
        if rank == 0 and os.environ.get("OPENBLAS_NUM_THREADS") != "1" and not os.environ.get("MB_BENCH_REEXEC"):
            env = dict(os.environ, OPENBLAS_NUM_THREADS="1", OMP_NUM_THREADS="1", MKL_NUM_THREADS="1", MB_BENCH_REEXEC="1")
            sys.stdout.flush()
            os.execve(sys.executable, [sys.executable] + sys.argv, env)
        try:
            run_reference(args, rank)
        except Exception as e:               # never leave the driver without a line
            if rank == 0:
                print(json.dumps({"impl": "reference", "unavailable": "oracle arm failed: %s" % str(e)[:200]}))
        return
    if args.config == "c4":
        run_c4(args, rank, world, local_rank)
        return

    import torch
    from magmaclustr_b200 import _lib as L
    from magmaclustr_b200.engine import Context
    from magmaclustr_b200.parallel import shard_tasks, make_allreduce, init_library_nccl

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    df, prep, hp0, hpi = make_inputs()
    k0, ki = L.parse_kernel("SE", False), L.parse_kernel("SE", True)
    ctx = Context(local_rank)
    lo, hi_ = shard_tasks(prep.offs, rank, world, len(prep.grid_X))
    offs = (prep.offs[lo:hi_ + 1] - prep.offs[lo]).astype(np.int64)
    r0, r1 = int(prep.offs[lo]), int(prep.offs[hi_])
    # pinned host copies of the step's inputs (e2e path)
    def pinned(a):
        t = torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
        return t.numpy(), t
    X_h, _k1 = pinned(prep.X[r0:r1])
    y_h, _k2 = pinned(prep.y[r0:r1])
    gi_h, _k3 = pinned(prep.grid_index[r0:r1])
    gX_h, _k4 = pinned(prep.grid_X)
    hpi_loc = np.ascontiguousarray(hpi[lo:hi_])
    if world > 1:
        if os.environ.get("MB_ALLREDUCE", "nccl") == "hook":
            ctx.set_allreduce(make_allreduce(dist), rank, world)      # torch.distributed does the all-reduce
        else:
            init_library_nccl(ctx, dist, rank, world)                 # the library's own NCCL communicator
        ctx.set_shard(lo, len(prep.offs) - 1)                         # leaf-aligned shard: results independent of the GPU count
    ctx.upload_dataset(offs, X_h, y_h, gi_h, gX_h)
    m0 = np.zeros(ctx.U)
    ctx.set_state(k0, hp0, ki, hpi_loc, m0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident steps ------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()                     # started before the warm-up: the timed region itself is < 100 ms
    for _ in range(args.warmup):
        ctx.em_step_resident(k0, ki, restart=True)
    ctx.profile(True)
    ctx.profile_read()
    launches0 = ctx.launches
    ms = []
    stats = None
    logL = None
    for _ in range(args.steps):
        ctx.flush_l2()                      # 256 MiB write between timed iterations
        barrier()
        ctx.timer_start()
        logL, stats = ctx.em_step_resident(k0, ki, restart=True)
        ms.append(ctx.timer_stop_ms())
        barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - launches0
    prof = ctx.profile_read()
    cyc = ctx.mstep_cycles()
    ctx.profile(False)
    total_ms = float(np.sum(ms))
    evals_local = prof["evals"]
    if world > 1:
        t = torch.tensor([total_ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    value = args.steps / (total_ms * 1e-3)

    # ---- end-to-end steps: host buffers in, host buffers out -------------------------------
    e2e_ms = []
    h0_h, hi_h = hp0.copy(), hpi_loc.copy()
    for s in range(args.warmup + args.steps):
        h0_h[:] = hp0
        hi_h[:] = hpi_loc
        ctx.flush_l2()
        barrier()
        t0 = time.perf_counter()
        ctx.upload_dataset(offs, X_h, y_h, gi_h, gX_h)
        ll_e2e, _ = ctx.em_step(k0, h0_h, ki, hi_h, m0)
        dt = (time.perf_counter() - t0) * 1e3
        barrier()
        if s >= args.warmup:
            e2e_ms.append(dt)
    e2e_total = float(np.sum(e2e_ms))
    if world > 1:
        t = torch.tensor([e2e_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_total = float(t.item())
    h2d = int(offs.nbytes + X_h.nbytes + y_h.nbytes + gi_h.nbytes + gX_h.nbytes + hi_h.nbytes + m0.nbytes)
    d2h = int(hi_h.nbytes + 8)

    # ---- full fit (context for the step metric; outside the timed regions) --------------------
    fit = None
    if not args.no_fit:
        t0 = time.perf_counter()
        res = ctx.train_magma(k0, hp0, ki, hpi_loc, m0)
        fit = {"em_steps": int(len(res["seq_loglikelihood"])), "converged": bool(res["converged"]),
               "seconds": time.perf_counter() - t0, "final_logL": float(res["seq_loglikelihood"][-1])}

    # ---- the variant round 1 measured: ONE hp_i draw broadcast to every task (reported beside the headline) ----
    variants = None
    if world == 1 and not args.no_variants:
        _, _, _, hpi_b = make_inputs(broadcast=True)
        ctx.set_state(k0, hp0, ki, hpi_b, m0)
        for _ in range(2):
            ctx.em_step_resident(k0, ki, restart=True)
        msb = []
        for _ in range(max(3, args.steps // 2)):
            ctx.flush_l2()
            barrier()
            ctx.timer_start()
            llb, stb = ctx.em_step_resident(k0, ki, restart=True)
            msb.append(ctx.timer_stop_ms())
        variants = {"hp_i_broadcast_to_all_tasks": {"value": 1e3 / float(np.mean(msb)), "unit": "EM steps/s",
                                                    "ms_per_step": float(np.mean(msb)), "task_evals_per_step": int(stb[2]),
                                                    "lockstep_rounds": int(stb[1]), "logL": llb}}
        ctx.set_state(k0, hp0, ki, hpi_loc, m0)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    peak = dgemm_peak_tflops()
    achieved = evals_local * FLOPS_PER_EVAL * N_POINTS ** 3 / (prof["ms"] * 1e-3) / 1e12 if prof["ms"] > 0 else 0.0
    line = {
        "metric": METRIC, "value": value, "unit": "EM steps/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C2: train_magma EM step (e_step + m_step + logL_monitoring), "
                               "M=10000 tasks x N=64, SE kernel, shared_hp=FALSE, U=201, seeds data=2024 "
                               "hp_0=6 hp_i=106 (one hp_i draw per task)",
                   "step": "first EM iteration from the initial hyper-parameters, repeated",
                   "l2": "256 MiB flush between timed iterations",
                   "tasks_per_rank": int(hi_ - lo), "task_evals_per_step": int(stats[2]),
                   "lockstep_rounds": int(stats[1]), "hp0_evals": int(stats[0]),
                   "tasks_stopped_on_nonfinite_objective": int(stats[3]), "logL": logL,
                   "m_step": "lock-step rounds while > MB_FUSED_BELOW (default 3 x resident CTAs) tasks are active, then the "
                             "persistent per-task optimiser kernel (objective + L-BFGS-B step fused, atomic task queue)",
                   "m_step_eval_share_of_cta_cycles": (cyc["eval_cycles"] / (cyc["eval_cycles"] + cyc["optimiser_cycles"])
                                                        if cyc["eval_cycles"] > 0 else None)},
        "clocks": clocks,
        "e2e": {"value": args.steps / (e2e_total * 1e-3), "unit": "EM steps/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
        "gpu_launches": int(launches),
        "roofline": {"bound": "tensor", "kernel": "k_task_objective3<SE,8,D1,MOD>", "achieved": achieved, "peak": peak,
                     "unit": "TFLOP/s", "frac": achieved / peak if peak else None, "traffic": NCU_TRAFFIC_BYTES_PER_LAUNCH,
                     "algorithmic_flops_per_eval": FLOPS_PER_EVAL * N_POINTS ** 3,
                     "peak_source": "cuBLAS DGEMM 8192^3 measured in this run (no FP64 figure in MEASURED_PEAKS.json)",
                     "kernel_ms_per_step": prof["ms"] / args.steps,
                     "kernel_share_of_step": prof["ms"] / float(np.sum(ms)),
                     "persistent_m_step_kernel_ms_per_step": prof["persistent_kernel_ms"] / args.steps,
                     "between_launches_ms_per_step": prof["round_gap_ms"] / args.steps,
                     "phases_ms_per_step": {"e_step": prof["e_step_ms"] / args.steps, "m_step": prof["m_step_ms"] / args.steps,
                                            "logL_monitoring": prof["monitoring_ms"] / args.steps,
                                            "e_step_segments": {k: v / args.steps for k, v in prof["e_step_segments_ms"].items()}}},
    }
    if fit:
        line["fit"] = fit
    if variants:
        line["variants"] = variants
    if world == 1:
        # second half of BASELINE.json's metric: batched Cholesky FP64 TFLOP/s, device-only
        ms_inv, fl = ctx.bench_task_inverse(ki, reps=5)
        bc = {"tasks": {"kernel": "k_task_inverse3", "M": M_TASKS, "N": N_POINTS, "ms": ms_inv,
                        "tflops_chol_plus_inverse_n3": fl / ms_inv / 1e9,
                        "frac_of_dgemm": fl / ms_inv / 1e9 / peak if peak else None}}
        n_big = 4096
        for what, key, flops in ((1, "chol_n3_over_3", n_big ** 3 / 3.0), (2, "chol_plus_inverse_n3", float(n_big) ** 3),
                                 (0, "gemm_2n3", 2.0 * n_big ** 3)):
            ms_d = ctx.bench_dense(what, n_big, reps=3)
            bc.setdefault("dense_4096", {})[key] = {"ms": ms_d, "tflops": flops / ms_d / 1e9,
                                                    "frac_of_dgemm": flops / ms_d / 1e9 / peak if peak else None}
        if not args.no_c4:
            # BASELINE config 4 shape (common grid, N = U = 4096): one EM step over 8 of the 256 tasks -- the large
            # factorisations batched over the library's streams, their trailing updates on FP64 DMMA
            try:
                m4, n4 = 8, 4096
                df4, _ = simu_db_c4(m4, n4)
                prep4 = Prepared4(df4.ID.values, df4.Input.values, df4.Output.values)
                hp04 = ini_hp4("SE", ["0"], True, False, 6).drop(columns="ID").iloc[0].to_numpy(dtype=np.float64)
                hpi4 = ini_hp4("SE", prep4.ids, False, True, 106).drop(columns="ID").to_numpy(dtype=np.float64)
                ctx4 = Context(local_rank)
                ctx4.upload_dataset(prep4.offs, prep4.X, prep4.y, prep4.grid_index, prep4.grid_X)
                ctx4.set_state(k0, hp04, ki, hpi4, np.zeros(ctx4.U))
                ctx4.em_step_resident(k0, ki, restart=True)          # warm-up: allocations
                ctx4.timer_start()
                ll4, st4 = ctx4.em_step_resident(k0, ki, restart=True)
                ms4 = ctx4.timer_stop_ms()
                tf4 = int(st4[2]) * FLOPS_PER_EVAL * float(n4) ** 3 / ms4 / 1e9
                bc["c4_em_step_8_tasks_x_4096"] = {"ms": ms4, "task_evals": int(st4[2]), "tflops_5n3": tf4,
                                                   "frac_of_dgemm": tf4 / peak if peak else None, "logL": ll4}
                ctx4.close()
            except Exception as e:      # the headline line must survive a failure of this side measurement
                bc["c4_em_step_8_tasks_x_4096"] = {"error": str(e)[:200]}
        line["batched_chol"] = bc
        # HBM-bound kernels of the path: stand-alone covariance build and the E step's reduction,
        # device-only, against the measured copy bandwidth of MEASURED_PEAKS.json
        hbm_peak = None
        try:
            pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
            hbm_peak = float(pk.get("hbm_gbps_burst") or pk.get("hbm_gbps") or pk.get("hbm_gbs") or 0) or None
        except Exception:
            pass
        hk = {"peak_gbps": hbm_peak}
        for what, key in ((0, "list_kern_to_cov"), (1, "e_step_reduce_partials")):
            ms_h, by = ctx.bench_hbm(what, ki, reps=10)
            hk[key] = {"ms": ms_h, "algorithmic_bytes": by, "gbps": by / ms_h / 1e6,
                       "frac_of_hbm": (by / ms_h / 1e6 / hbm_peak) if hbm_peak else None}
        line["hbm_kernels"] = hk
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        try:                                 # a side measurement: the headline line must survive whatever happens here
            m_sample = args.cpu_sample
            if not m_sample:                 # size the sample for ~15 s of CPU work on this box's cores
                dt0, _, _ = oracle_sample_step(df, 512, cores)
                m_sample = int(min(M_TASKS, max(512, 512 * 15.0 / max(dt0, 1e-3)))) // 256 * 256
            dt, evals, _ = oracle_sample_step(df, m_sample, cores)
            line["cpu_baseline"] = {
                "value": 1.0 / (dt * M_TASKS / m_sample), "unit": "EM steps/s", "cores": cores,
                "kind": "port",
                "sample": "one EM step of the oracle (numpy/LAPACK restatement of the reference, one process per core) "
                          "on the first %d of %d tasks (%.1f s), scaled linearly in M" % (m_sample, M_TASKS, dt)}
        except Exception as e:
            line["cpu_baseline"] = {"value": None, "unit": "EM steps/s", "cores": cores, "kind": "port",
                                    "sample": "failed: %s" % str(e)[:200]}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
