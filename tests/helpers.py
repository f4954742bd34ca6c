"""Shared builders for the parity tests: the same seeded inputs go to the oracle (CPU
restatement of the reference) and to the CUDA library."""
import numpy as np

from magmaclustr_b200 import _lib as L
from magmaclustr_b200.dataprep import Prepared
from magmaclustr_b200.simulate import simu_db, ini_hp
from oracle import magma as OM


def make_case(M, N, seed, kern_i="SE", kern_0="SE", shared=False, common_input=False, K=1):
    df, gen = simu_db(M=M, N=N, K=K, common_input=common_input, seed=seed)
    prep = Prepared(df.ID.values, df.Input.values, df.Output.values)
    odb = OM.Dataset(df.ID.values, df.Input.values, df.Output.values)
    h0 = ini_hp(kern_0, ["0"], True, False, seed + 1).drop(columns="ID")
    hi = ini_hp(kern_i, prep.ids, shared, True, seed + 2)
    names_i = [c for c in hi.columns if c != "ID"]
    names_0 = list(h0.columns)
    hp0 = h0.iloc[0].to_numpy(dtype=np.float64)
    hpi = hi[names_i].to_numpy(dtype=np.float64)
    o_hp0 = dict(zip(names_0, hp0))
    o_hpi = {i: dict(zip(names_i, hpi[t])) for t, i in enumerate(prep.ids)}
    k0 = L.parse_kernel(kern_0, False)
    ki = L.parse_kernel(kern_i, True)
    assert L.hp_names(k0) == names_0 and L.hp_names(ki) == names_i
    return dict(df=df, prep=prep, odb=odb, hp0=hp0, hpi=hpi, o_hp0=o_hp0, o_hpi=o_hpi, k0=k0, ki=ki,
                names_0=names_0, names_i=names_i, kern_0=kern_0, kern_i=kern_i, gen=gen)


def upload(ctx, case):
    p = case["prep"]
    ctx.upload_dataset(p.offs, p.X, p.y, p.grid_index, p.grid_X)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def ulp_perturb(x, rng, k=1.0):
    """Multiply every entry by (1 +/- k * 2^-52): the smallest change an input can undergo.
    Quantities whose oracle value moves by d under this perturbation cannot be reproduced by ANY
    other correct implementation to better than ~d."""
    x = np.asarray(x, dtype=np.float64)
    return x * (1.0 + k * 2.0 ** -52 * rng.choice([-1.0, 1.0], size=x.shape))


def oracle_sensitivity(fn, args_perturbed_list, base):
    """max over perturbed evaluations of the relative deviation (w.r.t. max |base|) of fn."""
    worst = 0.0
    for a in args_perturbed_list:
        worst = max(worst, relerr(fn(*a), base))
    return worst
