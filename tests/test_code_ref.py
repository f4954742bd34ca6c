"""The reference's own object code for the five native builders (oracle/_ref/libcode_ref.so = /root/reference/src/code.cpp
compiled unchanged, oracle/Makefile) against (i) the numpy restatement in oracle/kernels.py -- this PINS the oracle for
rows a1-a3 of SURVEY.md 8 -- and (ii) the CUDA builders behind the C ABI (mb_cpp_*)."""
import numpy as np
import pytest

from oracle import code_ref as REF
from oracle import kernels as OK

needs_ref = pytest.mark.skipif(not REF.available(), reason="oracle/_ref not built and /root/reference absent")


def _ulps(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.spacing(np.maximum(np.abs(a), np.abs(b))), 5e-324)
    return float(np.max(np.abs(a - b) / scale)) if a.size else 0.0


def _cases():
    rng = np.random.default_rng(7)
    out = []
    for d in (1, 2, 3):
        for n1, n2 in ((1, 1), (5, 5), (17, 9), (64, 64)):
            m1 = rng.normal(size=(n1, d)) * 3
            m2 = rng.normal(size=(n2, d)) * 3
            if n1 == n2:
                m2[: n1 // 2] = m1[: n1 // 2]               # coincident pairs (noise placement, zero distances)
                if n1 > 2:
                    m2[-1] = m1[-1] + 4e-7 / d              # L1 distance just under the 1e-6 threshold of cpp_noise
            out.append((m1, m2))
    pool = np.round(np.arange(-1, 1.001, 0.01), 2)[:, None]    # the simulated pool: many exactly equal inputs
    out.append((pool[::3], pool[::5]))
    return out


@needs_ref
def test_oracle_restatement_matches_reference_object_code():
    worst = {}
    for m1, m2 in _cases():
        for name, args in (("cpp_dist", ()), ("cpp_prod", ()), ("cpp_perio", (0.7,)), ("cpp_perio_deriv", (0.7,)),
                           ("cpp_perio", (-1.3,)), ("cpp_noise", (-2.5,))):
            r = getattr(REF, name)(m1, m2, *args)
            o = getattr(OK, name)(m1, m2, *args)
            assert r.shape == o.shape == (len(m1), len(m2))
            worst[name] = max(worst.get(name, 0.0), _ulps(o, r))
    # sums of squares / products / indicator: the same IEEE operations in the same order -> identical bits;
    # the periodic kernels go through libm's sin / cos / exp in both, where numpy may differ from glibc by an ulp
    assert worst["cpp_dist"] == 0 and worst["cpp_prod"] == 0 and worst["cpp_noise"] == 0, worst
    assert worst["cpp_perio"] <= 4 and worst["cpp_perio_deriv"] <= 64, worst


@needs_ref
def test_reference_object_code_rejects_mismatched_dimensions():
    with pytest.raises(ValueError):
        REF.cpp_dist(np.zeros((3, 2)), np.zeros((3, 1)))      # code.cpp:11-13
    # only cpp_dist checks: the other four read m2 with m1's column count (code.cpp:31-131), nothing to pin there


@needs_ref
@pytest.mark.gpu
def test_cuda_builders_match_reference_object_code(ctx):
    for m1, m2 in _cases():
        np.testing.assert_allclose(ctx.cpp_dist(m1, m2), REF.cpp_dist(m1, m2), rtol=1e-15, atol=0)
        scale = float(np.max(np.abs(m1)) * np.max(np.abs(m2)))     # products may cancel: absolute floor of a few ulps of the terms
        np.testing.assert_allclose(ctx.cpp_prod(m1, m2), REF.cpp_prod(m1, m2), rtol=1e-15, atol=4e-16 * scale)
        assert np.array_equal(ctx.cpp_noise(m1, m2, -2.5) != 0, REF.cpp_noise(m1, m2, -2.5) != 0)   # placement: exact
        np.testing.assert_allclose(ctx.cpp_noise(m1, m2, -2.5), REF.cpp_noise(m1, m2, -2.5), rtol=1e-15)
        for p in (0.7, -1.3):
            np.testing.assert_allclose(ctx.cpp_perio(m1, m2, p), REF.cpp_perio(m1, m2, p), rtol=1e-13, atol=1e-15)
            np.testing.assert_allclose(ctx.cpp_perio_deriv(m1, m2, p), REF.cpp_perio_deriv(m1, m2, p), rtol=1e-12, atol=1e-13)
