"""CPU-side checks: data preparation (Reference keys, C-locale ordering, CSR), the C-ABI
library loads and exports every symbol of include/magma_b200.h, kernel-string parsing, the
host build of the L-BFGS-B state machine, and the sharding helpers (gloo, world size 2)."""
import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

from magmaclustr_b200 import _lib as L
from magmaclustr_b200 import dataprep as DP
from magmaclustr_b200.parallel import shard_tasks
from oracle import rformat as RF
from oracle import magma as OM
from oracle.lbfgsb import optim_lbfgsb


def test_r_number_formatting():
    cases = {1e-4: "1e-04", 0.001: "0.001", 1e5: "1e+05", 123457.0: "123457", 100000.0: "1e+05",
             -0.22: "-0.22", 0.3: "0.3", 10.0: "10", -1.0: "-1", 0.1 + 0.2: "0.3", 2.5e-7: "2.5e-07",
             0.0: "0", 1234567.891: "1234567.891", 1 / 3: "0.333333333333333"}
    for v, s in cases.items():
        assert DP.r_as_character(v) == s == RF.r_as_character(v)
    xs = np.array([1234567.891, 1 / 3, -0.0123456789, 123456.7, 0.999999499, 5e-324, 0.0])
    np.testing.assert_array_equal(DP.r_signif(xs), [RF.r_signif(v) for v in xs])
    assert DP.r_signif(np.array([123456.7]))[0] == 123457.0


def test_union_grid_is_string_sorted():
    # "-0.01" < "-0.02" < "-0.1" < "-1" < "0" < "0.01" < "1" ; "10" < "2"   
    x = np.array([2.0, 10.0, 0.0, -0.1, -1.0, -0.02, -0.01, 0.01, 1.0])
    p = DP.Prepared(["a"] * len(x), x, np.zeros(len(x)))
    assert [k.decode() for k in p.grid_keys] == ["-0.01", "-0.02", "-0.1", "-1", "0", "0.01", "1", "10", "2"]
    assert [k.decode() for k in p.keys] == [k.decode() for k in p.grid_keys]


def test_prepared_matches_oracle_dataset_bit_exact():
    from helpers import make_case
    for (m, n, seed) in [(10, 10, 17), (23, 7, 2)]:
        case = make_case(m, n, seed)
        p, o = case["prep"], case["odb"]
        assert p.ids == o.ids                                   # ID strings in C-locale order ("1","10","2",...)
        assert [k.decode() for k in p.grid_keys] == o.grid_keys
        np.testing.assert_array_equal(p.grid_X, o.grid_X)
        for t, i in enumerate(p.ids):
            sl = slice(p.offs[t], p.offs[t + 1])
            np.testing.assert_array_equal(p.grid_index[sl], o.index[i])     # index mapping: exact
            np.testing.assert_array_equal(p.y[sl], o.tasks[i][2])
            np.testing.assert_array_equal(p.X[sl], o.tasks[i][1])


def test_prepared_ragged_2d_and_duplicates():
    ids = ["b", "b", "a", "a", "a"]
    X = np.array([[1.0, 2.0], [0.5, 2.0], [1.0, 2.0], [10.0, 1.0], [2.0, 1.0]])
    p = DP.Prepared(ids, X, np.arange(5.0))
    assert p.ids == ["a", "b"] and list(p.offs) == [0, 3, 5]
    assert [k.decode() for k in p.keys[:3]] == ["10:1", "1:2", "2:1"]   # byte order: "0" (0x30) < ":" (0x3a)
    with pytest.raises(ValueError):
        DP.Prepared(["a", "a"], np.array([1.0, 1.0]), np.zeros(2))
    q = DP.Prepared(["a", "a", "b"], np.array([1.0, np.nan, 2.0]), np.array([1.0, 2.0, np.nan]))
    assert q.ids == ["a"] and len(q.y) == 1                       # drop_na


def test_extended_grid_for_prediction():
    p = DP.Prepared(["a", "a", "b"], np.array([0.1, 0.3, 0.2]), np.zeros(3))
    p.set_grid(np.array([0.0, 0.2, 0.5]))
    assert [k.decode() for k in p.grid_keys] == ["0", "0.1", "0.2", "0.3", "0.5"]
    np.testing.assert_array_equal(p.grid_index, [1, 3, 2])


def _declared_symbols():
    hdr = open(os.path.join(ROOT, "include", "magma_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    return sorted(set(re.findall(r"\b(mb_[a-zA-Z0-9_]+)\s*\(", hdr)) - {"mb_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(L.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), "missing export: " + n
    for n in names:                                              # the Python binding covers the ABI
        assert n in L._SIGS or n in ("mb_last_error", "mb_kernel_hp_name"), n


def test_no_gpu_means_loud_failure():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from magmaclustr_b200.engine import Context
    with pytest.raises(L.MagmaError):
        Context(0)


def test_kernel_parse():
    k = L.parse_kernel("SE * RQ + PERIO", True)
    assert (k.n_terms, k.n_prod, k.has_noise, k.n_hp) == (3, 2, 1, 9)
    assert L.hp_names(k) == ["se_variance", "se_lengthscale", "rq_variance", "rq_lengthscale", "rq_scale",
                             "perio_variance", "perio_lengthscale", "period", "noise"]
    assert L.parse_kernel("SE", False).n_hp == 2
    for bad in ["+ SE", "SE + RQ * PERIO", "SE *", "FOO", "SE SE", "SE + SE"]:
        with pytest.raises(L.MagmaError):
            L.parse_kernel(bad, False)


def test_host_lbfgsb_matches_oracle_bitwise():
    x = np.array([-1.2, 1.0])
    fval = ctypes.c_double(0)
    nev = ctypes.c_int32(0)
    fail = ctypes.c_int32(0)
    for factr, maxit in [(1e7, 100), (1e13, 25)]:
        xx = x.copy()
        L.check(L.lib().mb_lbfgsb_selftest(0, L.ptr(xx), 2, factr, maxit, ctypes.byref(fval), ctypes.byref(nev),
                                           ctypes.byref(fail)))

        def fg(v):
            return (100 * (v[1] - v[0] * v[0]) * (v[1] - v[0] * v[0]) + (1 - v[0]) * (1 - v[0]),
                    [-400 * v[0] * (v[1] - v[0] * v[0]) - 2 * (1 - v[0]), 200 * (v[1] - v[0] * v[0])])
        res = optim_lbfgsb(list(x), fg, factr=factr, maxit=maxit)
        assert list(xx) == res["par"] and fval.value == res["value"]
        assert nev.value == res["counts"] and fail.value == res["convergence"]


def test_shard_tasks_partition():
    """Shards are unions of whole E-step leaves (mb_shard_range): contiguous, covering, and nested -- the range of a rank
    at world = 2 is the union of two ranges at world = 4 -- which is what makes the sums independent of the GPU count."""
    import ctypes as C
    offs = np.r_[0, np.cumsum([64] * 10)]
    parts = [shard_tasks(offs, r, 4, 201) for r in range(4)]
    assert parts[0][0] == 0 and parts[-1][1] == 10
    assert all(parts[i][1] == parts[i + 1][0] for i in range(3))
    assert shard_tasks(offs, 0, 1, 201) == (0, 10)
    for m, u in ((10000, 201), (10000, 2000), (256, 4096), (70, 100), (11, 60)):
        offs = np.arange(m + 1) * 8
        nl = L.lib().mb_estep_leaves(m, u)
        assert nl >= 1 and (nl & (nl - 1)) == 0 and nl <= min(128, m)
        assert nl * (u * u + u) * 8 <= (1 << 31) or nl == 1
        if (m, u) == (256, 4096):
            assert nl == 8          # config 4 has to shard over 8 GPUs
        for world in (2, 4, 8):
            if nl < world:
                continue
            pr = [shard_tasks(offs, r, world, u) for r in range(world)]
            assert pr[0][0] == 0 and pr[-1][1] == m and all(pr[i][1] == pr[i + 1][0] for i in range(world - 1))
            assert all(hi > lo for lo, hi in pr)
            if world > 2:
                half = [shard_tasks(offs, r, world // 2, u) for r in range(world // 2)]
                assert all(half[r] == (pr[2 * r][0], pr[2 * r + 1][1]) for r in range(world // 2))
    with pytest.raises(L.MagmaError):
        shard_tasks(np.arange(4) * 8, 0, 8, 10)      # 3 tasks cannot feed 8 ranks


_GLOO = r'''
import os, sys
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "tests"))
import numpy as np, torch, torch.distributed as dist
from helpers import make_case
from magmaclustr_b200.parallel import shard_tasks
from oracle import kernels as OK, magma as OM
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
case = make_case(12, 9, 4)
p, o = case["prep"], case["odb"]
U = len(p.grid_keys)
lo, hi = shard_tasks(p.offs, rank, world, U)
part = np.zeros((U, U)); w = np.zeros(U)
for t in range(lo, hi):                      # this rank's partial precision / weighted sums
    i = p.ids[t]
    inv = OK.kern_to_inv(o.tasks[i][1], "SE", case["o_hpi"][i], 1e-10)
    idx = o.index[i]
    part[np.ix_(idx, idx)] += inv; w[idx] += inv @ o.tasks[i][2]
tp, tw = torch.from_numpy(part), torch.from_numpy(w)
dist.all_reduce(tp); dist.all_reduce(tw)
inv0 = OK.kern_to_inv(o.grid_X, "SE", case["o_hp0"], 1e-10)
pc = OK.chol_inv_jitter(inv0 + tp.numpy(), 1e-10); pm = pc @ tw.numpy()
opm, opc = OM.e_step(o, 0.0, "SE", "SE", case["o_hp0"], case["o_hpi"], 1e-10)
err = max(np.max(np.abs(pc - opc)) / np.max(np.abs(opc)), np.max(np.abs(pm - opm)) / np.max(np.abs(opm)))
assert err < 1e-6, err
print("rank", rank, "ok", err)
dist.destroy_process_group()
'''


def test_sharded_e_step_sums_with_gloo_world2(tmp_path):
    """The multi-GPU exchange: per-rank partial precision sums + all-reduce(sum) give the
    single-process hyper-posterior.  Host logic only (gloo, 2 processes, oracle arithmetic)."""
    script = tmp_path / "gloo_estep.py"
    script.write_text(_GLOO % (ROOT, ROOT))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert r.stdout.count("ok") == 2


# ---- compiled host data path (csrc/hostprep.cu): bit-exact with the Python preparation and the oracle ------------------
def test_native_number_formatting_matches_r_rules():
    import ctypes as C
    lib = L.lib()
    cases = {1e-4: "1e-04", 0.001: "0.001", 1e5: "1e+05", 123457.0: "123457", 100000.0: "1e+05",
             -0.22: "-0.22", 0.3: "0.3", 10.0: "10", -1.0: "-1", 0.1 + 0.2: "0.3", 2.5e-7: "2.5e-07",
             0.0: "0", 1234567.891: "1234567.891", 1 / 3: "0.333333333333333"}
    buf = C.create_string_buffer(64)
    for v, s in cases.items():
        L.check(lib.mb_r_as_character(v, buf, 64))
        assert buf.value.decode() == s
    rng = np.random.default_rng(12)
    xs = np.concatenate([rng.normal(size=300) * 10.0 ** rng.integers(-8, 9, size=300), np.round(rng.uniform(-5, 5, 200), 2),
                         [1234567.891, 1 / 3, -0.0123456789, 123456.7, 0.999999499, 5e-324, 0.0, 1e22, -1e-15]])
    for v in xs:
        r = lib.mb_r_signif(float(v), 6)
        assert r == RF.r_signif(float(v))
        L.check(lib.mb_r_as_character(r, buf, 64))
        assert buf.value.decode() == RF.r_as_character(r), v
    row = np.array([0.1 + 0.2, -1234567.0, 2.5e-7])
    L.check(lib.mb_reference_key(L.ptr(row), 3, 6, buf, 64))
    assert buf.value.decode() == RF.reference_key(row) == "0.3:-1234570:2.5e-07"


def test_native_prepare_rows_matches_python_and_oracle():
    rng = np.random.default_rng(8)
    for d in (1, 2):
        sizes = rng.integers(1, 9, size=13)
        ids, X = [], []
        pool = np.round(np.arange(-1, 1.001, 0.01), 2)
        for t, n in enumerate(sizes):
            ids += [str(t + 1)] * int(n)                      # "1", "10", "11", ... sort as strings
            pts = rng.choice(len(pool), size=(int(n),), replace=False)
            cols = [pool[pts]] + [np.round(rng.uniform(0, 10, int(n)), 3) for _ in range(d - 1)]
            X.append(np.column_stack(cols))
        X = np.vstack(X)
        y = rng.normal(size=len(X))
        shuffle = rng.permutation(len(X))                     # the table need not arrive sorted
        ids = [ids[i] for i in shuffle]
        X, y = X[shuffle], y[shuffle]
        extra = np.column_stack([np.round(np.linspace(-1, 1, 7), 3)] + [np.full(7, 1.5)] * (d - 1))
        nat = DP.prepare_rows_native(ids, X, y, grid_extra=extra)
        py = DP.Prepared(ids, X, y)
        py.set_grid(extra)
        np.testing.assert_array_equal(nat["offs"], py.offs)
        np.testing.assert_array_equal(nat["X"], py.X)
        np.testing.assert_array_equal(nat["y"], py.y)
        np.testing.assert_array_equal(nat["grid_index"], py.grid_index)       # index mapping: exact
        np.testing.assert_array_equal(nat["grid_X"], py.grid_X)
        # and against the oracle's Dataset (training.R:690-720 restated independently)
        odb = OM.Dataset(ids, X, y)
        keys, Xg, index = odb.extend_grid(extra)
        np.testing.assert_array_equal(nat["grid_X"], Xg)
        for t, i in enumerate(odb.ids):
            np.testing.assert_array_equal(nat["grid_index"][nat["offs"][t]:nat["offs"][t + 1]], index[i])


def test_native_prepare_rows_rejects_duplicates():
    ids = ["a", "a", "b"]
    X = np.array([[0.5], [0.5000001], [0.5]])       # equal after signif(6): two Outputs of "a" on one grid point
    with pytest.raises(L.MagmaError) as e:
        DP.prepare_rows_native(ids, X, np.zeros(3))
    assert e.value.args[0] == 3 or "several Outputs" in str(e.value)


def test_cpu_baseline_pool_matches_sequential_oracle():
    """oracle/parallel.py (the all-cores driver behind bench.py's cpu_baseline / --impl reference) runs the SAME EM step as the
    sequential oracle: hyper-posterior to rounding (OpenBLAS blocks differently with one thread), identical optimiser
    evaluation counts, log-likelihood to 1e-9."""
    import multiprocessing as mp
    from helpers import make_case
    from oracle import magma as OM, parallel as OP
    case = make_case(12, 10, 23)
    db, h0, hi = case["odb"], case["o_hp0"], case["o_hpi"]
    pm, pc = OM.e_step(db, 0.0, "SE", "SE", h0, hi, 1e-10)
    OP._G.update(db=db, hp_i=hi, kern_i="SE", pen=1e-10)
    with mp.get_context("fork").Pool(2, initializer=OP._limit_blas_threads) as pool:
        pm2, pc2 = OP.e_step(db, 0.0, "SE", "SE", h0, hi, 1e-10, pool, 2)
    assert np.max(np.abs(pc2 - pc)) <= 1e-9 * np.max(np.abs(pc)) and np.max(np.abs(pm2 - pm)) <= 1e-7 * np.max(np.abs(pm))
    stats = {}
    n0, ni = OM.m_step(db, 0.0, "SE", "SE", h0, hi, pm, pc, False, 1e-10, stats)
    ll = OM.logL_monitoring(n0, ni, db, 0.0, "SE", "SE", pm, pc, 1e-10)
    _, ni2, ll2, evals = OP.em_step(db, 0.0, "SE", "SE", h0, hi, 1e-10, procs=2)
    ref_evals = sum(r["counts"] for r in stats["tasks"].values())
    assert abs(evals - ref_evals) <= max(2, ref_evals // 50)     # equal here; a BLAS thread-count rounding difference may move one line search
    assert abs(ll2 - ll) <= 1e-6 * abs(ll)
    for i in db.ids:
        for k in ni[i]:
            assert abs(ni2[i][k] - ni[i][k]) <= 1e-4 * max(1.0, abs(ni[i][k]))    # observed 4e-6 (factr = 1e13 stops early)
