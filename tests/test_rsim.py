"""R-RNG-compatible simu_data (magmaclustr_b200/rsim.py) against the reference's own golden fixture
(tests/testthat/fixtures/simu_data_reference.rds decoded into tests/golden/simu_data_reference.json by
tests/tools/make_simu_golden.py) with the criterion of tests/testthat/test-simu_data-characterization.R:43-52:
expect_equal(new, old, tolerance = 1e-6), i.e. mean relative difference per numeric column, identical ID columns."""
import json
import math
import os

import numpy as np
import pytest

from magmaclustr_b200 import rsim

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "simu_data_reference.json")))
BOOL = ("shared_input", "shared_hp", "input2", "add_hp", "add_clust")


def _kwargs(cfg):
    kw = {}
    for k, v in cfg.items():
        if k in ("grid_input", "grid_input2"):
            kw[k] = np.array(v, dtype=np.float64)
        elif k in BOOL:
            kw[k] = bool(v[0])
        else:
            kw[k] = int(v[0])
    return kw


def test_r_generators_known_answers():
    r = rsim.RRng(101)                      # set.seed(101); runif(3)
    np.testing.assert_allclose([r.unif_rand() for _ in range(3)], [0.37219838, 0.04382482, 0.70968402], atol=5e-9)
    # qnorm: R's documented values
    assert abs(rsim.qnorm(0.975) - 1.959963984540054) < 1e-15
    assert abs(rsim.qnorm(0.5)) == 0.0 and rsim.qnorm(1e-10) == pytest.approx(-6.361340902404056, abs=1e-13)
    assert rsim.qnorm(0.0) == -math.inf
    # round half: R >= 4.0.0 picks the representable neighbour that is nearer to x
    assert rsim.r_round(0.15, 1) == 0.1 and rsim.r_round(0.25, 1) == 0.2 and rsim.r_round(2.675, 2) == 2.67
    assert rsim.r_round(-1.005, 2) == -1.0 and rsim.r_round(0.285, 2) == 0.28
    # sample.int without replacement is a permutation, and uses one uniform per accepted 8-bit draw for n = 201
    r = rsim.RRng(1)
    perm = r.sample_int(201, 201)
    assert sorted(perm) == list(range(201))


@pytest.mark.parametrize("name", GOLD["names"])
def test_simu_data_reproduces_reference_fixture(name):
    got = rsim.simu_data(int(GOLD["seeds"][name]), **_kwargs(GOLD["configs"][name]))
    ref = GOLD["results"][name]
    assert list(got.keys()) == GOLD["columns"][name]                        # expect_named
    n = len(ref["Output"])
    for col in GOLD["columns"][name]:
        assert len(got[col]) == n, col
        if col in ("Task_ID", "Input_ID", "Output_ID"):
            assert [str(v) for v in got[col]] == [str(v) for v in ref[col]], col      # character columns: identical
        elif col == "Input":
            assert np.array_equal(np.asarray(got[col]), np.asarray(ref[col], dtype=np.float64))   # sampled grid points: exact
        else:
            g, r = np.asarray(got[col], dtype=np.float64), np.asarray(ref[col], dtype=np.float64)
            assert not np.any(np.isnan(g))
            # all.equal(): mean absolute difference relative to the mean absolute target value
            assert np.mean(np.abs(g - r)) / np.mean(np.abs(r)) < 1e-6, (col, float(np.max(np.abs(g - r))))
            if col != "Output":            # hyper-parameter columns come straight from runif: exact
                assert np.array_equal(g, r), col
