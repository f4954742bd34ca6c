"""The `.Call` shim (shim/magma_b200_shim.c) compiled against the stand-in R runtime of shim/mock and linked with the
product library: (CPU) it builds warning-free, registers the five native builders under the reference's own names and
arities (/root/reference/src/init.c:17-24) plus one routine per accelerated R function, every registered routine exists;
(GPU) routines called through the mock runtime return what the C ABI returns through ctypes, and a library error becomes
an R error (Rf_error) after the call returned."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "shim", "_build", "libmagma_b200_shim_mock.so")
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def _lib():
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "shim")])
    l = C.CDLL(SO)
    l.mock_routine_name.restype = C.c_char_p
    l.mock_last_error.restype = C.c_char_p
    for f in ("mock_real", "mock_int", "mock_lgl", "mock_str", "mock_nil", "mock_call", "mock_list_get"):
        getattr(l, f).restype = C.c_void_p
    l.mock_real.argtypes = [_dp, C.c_ssize_t, C.c_int, C.c_int]
    l.mock_int.argtypes = [_ip, C.c_ssize_t]
    l.mock_str.argtypes = [C.c_char_p]
    l.mock_call.argtypes = [C.c_char_p, C.c_int, C.POINTER(C.c_void_p)]
    l.mock_list_get.argtypes = [C.c_void_p, C.c_char_p]
    l.mock_real_ptr.restype = _dp
    l.mock_real_ptr.argtypes = [C.c_void_p]
    l.mock_length.restype = C.c_ssize_t
    for f in ("mock_length", "mock_nrow", "mock_ncol", "mock_type"):
        getattr(l, f).argtypes = [C.c_void_p]
    assert l.mock_init() == 1
    return l


def _real(l, a, matrix=False):
    a = np.asarray(a, dtype=np.float64)
    if matrix:                       # R matrix: column-major with a dim attribute
        f = np.asfortranarray(a)
        return l.mock_real(f.ctypes.data_as(_dp), f.size, f.shape[0], f.shape[1])
    f = np.ascontiguousarray(a).ravel()
    return l.mock_real(f.ctypes.data_as(_dp), f.size, 0, 0)


def _call(l, name, *args):
    arr = (C.c_void_p * len(args))(*args)
    r = l.mock_call(name.encode(), len(args), arr)
    if not r:
        raise RuntimeError(l.mock_last_error().decode())
    return r


def _values(l, s, shape=None):
    n = l.mock_length(s)
    v = np.ctypeslib.as_array(l.mock_real_ptr(s), shape=(n,)).copy()
    return v if shape is None else v.reshape(shape, order="F")


def test_shim_builds_and_registers_every_routine():
    l = _lib()
    reg = {l.mock_routine_name(i).decode(): l.mock_routine_nargs(i) for i in range(l.mock_n_routines())}
    # the reference's own registration table (src/init.c:17-24)
    legacy = {"_MagmaClustR_cpp_dist": 2, "_MagmaClustR_cpp_noise": 3, "_MagmaClustR_cpp_perio": 3,
              "_MagmaClustR_cpp_perio_deriv": 3, "_MagmaClustR_cpp_prod": 2}
    init_c = "/root/reference/src/init.c"
    if os.path.exists(init_c):
        import re
        found = dict((m.group(1), int(m.group(2))) for m in
                     re.finditer(r'\{"(_MagmaClustR_\w+)",\s*\(DL_FUNC\)\s*&\w+,\s*(\d+)\}', open(init_c).read()))
        assert found == legacy
    for k, v in legacy.items():
        assert reg.get(k) == v
    for r in ("kern_to_cov", "list_kern_to_mat", "chol_inv_jitter", "upload_dataset", "e_step", "m_step", "logL_monitoring",
              "logL_GP_mod", "logL_GP_mod_tasks", "train_magma", "hyperposterior", "pred_magma", "hyperposterior_clust",
              "pred_magmaclust", "hyperpost_keep", "hyperpost_set", "logL_GP_tasks", "train_gp_tasks", "train_shared_gp",
              "train_gp_clust_tasks", "pred_magma_tasks", "ve_step", "update_mixture", "elbo_clust_tasks", "vm_step",
              "elbo_monitoring", "train_magmaclust", "set_option"):
        assert "mbR_" + r in reg, r
    # every routine of the table is a real symbol of the object, with the arity its definition has
    src = open(os.path.join(ROOT, "shim", "magma_b200_shim.c")).read()
    import re
    for name, nargs in reg.items():
        assert hasattr(l, name)
        m = re.search(r"SEXP %s\(([^)]*)\)" % re.escape(name), src)
        if m:                                    # (the five builders come from a macro)
            assert m.group(1).count("SEXP") == nargs, name


@pytest.mark.gpu
def test_shim_routines_match_the_c_abi(ctx):
    from helpers import make_case
    l = _lib()
    rng = np.random.default_rng(0)
    m1, m2 = rng.normal(size=(6, 2)), rng.normal(size=(4, 2))
    out = _call(l, "_MagmaClustR_cpp_dist", _real(l, m1, True), _real(l, m2, True))
    assert (l.mock_nrow(out), l.mock_ncol(out)) == (6, 4)
    assert np.array_equal(_values(l, out, (6, 4)), ctx.cpp_dist(m1, m2))
    out = _call(l, "_MagmaClustR_cpp_noise", _real(l, m1, True), _real(l, m1, True), _real(l, [-1.5]))
    assert np.array_equal(_values(l, out, (6, 6)), ctx.cpp_noise(m1, m1, -1.5))
    with pytest.raises(RuntimeError, match="Incompatible number of dimensions"):
        _call(l, "_MagmaClustR_cpp_dist", _real(l, m1, True), _real(l, rng.normal(size=(4, 3)), True))
    # kern_to_cov with a derivative
    hp = np.array([0.3, -1.0, -2.0])
    out = _call(l, "mbR_kern_to_cov", l.mock_str(b"SE"), l.mock_lgl(1), _real(l, hp), _real(l, m1, True), l.mock_nil(),
                _real(l, [1.0]))
    from magmaclustr_b200 import _lib as L
    assert np.array_equal(_values(l, out, (6, 6)), ctx.kern_to_cov(L.parse_kernel("SE", True), hp, m1, deriv=1))
    # resident dataset + e_step + logL_monitoring through the shim == through ctypes
    case = make_case(8, 9, 3)
    p = case["prep"]
    gi = np.ascontiguousarray(p.grid_index, dtype=np.int32)
    _call(l, "mbR_upload_dataset", _real(l, p.offs.astype(np.float64)), _real(l, p.X), _real(l, p.y),
          l.mock_int(gi.ctypes.data_as(_ip), gi.size), _real(l, p.grid_X), _real(l, [1.0]))
    U = len(p.grid_X)
    res = _call(l, "mbR_e_step", l.mock_str(b"SE"), _real(l, case["hp0"]), l.mock_str(b"SE"), _real(l, case["hpi"]),
                l.mock_lgl(0), _real(l, np.zeros(U)), _real(l, [1e-10]))
    mean, cov = _values(l, l.mock_list_get(res, b"mean")), _values(l, l.mock_list_get(res, b"cov"), (U, U))
    ctx.upload_dataset(p.offs, p.X, p.y, p.grid_index, p.grid_X)
    pm, pc, _ = ctx.e_step(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(U))
    assert np.array_equal(mean, pm) and np.array_equal(cov, pc)
    ll = _call(l, "mbR_logL_monitoring", l.mock_str(b"SE"), _real(l, case["hp0"]), l.mock_str(b"SE"), _real(l, case["hpi"]),
               l.mock_lgl(0), _real(l, np.zeros(U)), _real(l, pm), _real(l, pc), _real(l, [1e-10]))
    assert _values(l, ll)[0] == ctx.logL_monitoring(case["k0"], case["hp0"], case["ki"], case["hpi"], np.zeros(U), pm, pc)
    # library switches through the shim: the element-by-element SE kernel gives the bits of the pair-table path
    _call(l, "mbR_set_option", l.mock_str(b"pair_tables"), _real(l, [0.0]))
    res0 = _call(l, "mbR_e_step", l.mock_str(b"SE"), _real(l, case["hp0"]), l.mock_str(b"SE"), _real(l, case["hpi"]),
                 l.mock_lgl(0), _real(l, np.zeros(U)), _real(l, [1e-10]))
    assert np.array_equal(_values(l, l.mock_list_get(res0, b"cov"), (U, U)), pc)
    _call(l, "mbR_set_option", l.mock_str(b"pair_tables"), _real(l, [1.0]))
    with pytest.raises(RuntimeError, match="unknown option"):
        _call(l, "mbR_set_option", l.mock_str(b"no_such_switch"), _real(l, [1.0]))
    # a library failure surfaces as an R error with the library's message
    with pytest.raises(RuntimeError, match="magma_b200"):
        _call(l, "mbR_kern_to_cov", l.mock_str(b"SE * * RQ"), l.mock_lgl(1), _real(l, hp), _real(l, m1, True), l.mock_nil(),
              _real(l, [-1.0]))
    l.mock_unload()
