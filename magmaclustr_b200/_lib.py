"""ctypes binding of libmagma_b200.so (include/magma_b200.h).

The library is the product: there is no Python or CPU fallback.  If the shared object is
missing or no CUDA device is present, calls fail loudly.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MB_LIB") or os.path.join(_HERE, "lib", "libmagma_b200.so")

MB_MAX_TERMS = 4
MB_MAX_HP = 12


class MbKernel(C.Structure):
    _fields_ = [("n_terms", C.c_int32), ("n_prod", C.c_int32),
                ("types", C.c_int32 * MB_MAX_TERMS), ("has_noise", C.c_int32),
                ("n_hp", C.c_int32)]


class MagmaError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("magma_b200 error %d: %s" % (code, msg))
        self.code = code


_lib = None
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_lp = C.POINTER(C.c_int64)
_kp = C.POINTER(MbKernel)
_vp = C.c_void_p

ALLREDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_int64)

_SIGS = {
    "mb_create": [C.POINTER(_vp), C.c_int],
    "mb_destroy": [_vp],
    "mb_kernel_parse": [C.c_char_p, C.c_int, _kp],
    "mb_launch_count": [_vp],
    "mb_set_option": [_vp, C.c_char_p, C.c_int64],
    "mb_pair_table_counts": [_vp, _ip],
    "mb_set_allreduce": [_vp, ALLREDUCE_FN, _vp, C.c_int, C.c_int],
    "mb_cpp_dist": [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp],
    "mb_cpp_prod": [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, _dp],
    "mb_cpp_perio": [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, _dp],
    "mb_cpp_perio_deriv": [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, _dp],
    "mb_cpp_noise": [_vp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_double, _dp],
    "mb_kern_to_cov": [_vp, _kp, _dp, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, _dp],
    "mb_list_kern_to_mat": [_vp, _kp, C.c_int64, _lp, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int,
                            C.c_double, _lp, _dp, _ip],
    "mb_chol_inv_jitter": [_vp, _dp, C.c_int, C.c_int, C.c_double, _dp, _ip, _dp],
    "mb_upload_dataset": [_vp, C.c_int64, _lp, C.c_int, _dp, _dp, _ip, C.c_int32, _dp],
    "mb_e_step": [_vp, _kp, _dp, _kp, _dp, C.c_int, _dp, C.c_double, _dp, _dp, _ip],
    "mb_logL_GP_mod_tasks": [_vp, _kp, _dp, C.c_int, _dp, _dp, C.c_double, _dp, _dp, _ip],
    "mb_logL_GP_mod": [_vp, _kp, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp],
    "mb_m_step": [_vp, _kp, _dp, _kp, _dp, C.c_int, _dp, _dp, _dp, C.c_double, _lp],
    "mb_last_task_status": [_vp, _ip, _ip],
    "mb_logL_monitoring": [_vp, _kp, _dp, _kp, _dp, C.c_int, _dp, _dp, _dp, C.c_double, _dp],
    "mb_em_step": [_vp, _kp, _dp, _kp, _dp, C.c_int, _dp, C.c_double, _dp, _dp, _dp, _lp],
    "mb_train_magma": [_vp, _kp, _dp, _kp, _dp, C.c_int, _dp, C.c_double, C.c_int, C.c_double, _dp,
                       _ip, _ip, _dp, _dp],
    "mb_hyperposterior": [_vp, _kp, _dp, _kp, _dp, C.c_int, C.c_int32, _dp, _ip, _dp, C.c_double,
                          _dp, _dp],
    "mb_pred_magma": [_vp, _kp, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp,
                      _dp, C.c_double, _dp, _dp, _dp],
    "mb_hyperposterior_clust": [_vp, C.c_int, _kp, _dp, _kp, _dp, C.c_int, C.c_int32, _dp, _ip, _dp, _dp,
                                C.c_double, _dp, _dp],
    "mb_pred_magmaclust": [_vp, C.c_int, _kp, _dp, _dp, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp,
                           _dp, _dp, C.c_double, _dp, _dp, _dp, _dp, _dp],
    "mb_ve_step": [_vp, C.c_int, _kp, _dp, _kp, _dp, C.c_int, _dp, _dp, _dp, C.c_int, C.c_double,
                   _dp, _dp, _dp],
    "mb_update_mixture": [_vp, C.c_int, _kp, _dp, C.c_int, _dp, _dp, _dp, C.c_double, _dp],
    "mb_elbo_clust_tasks": [_vp, C.c_int, _kp, _dp, C.c_int, _dp, _dp, _dp, C.c_double, _dp, _dp],
    "mb_vm_step": [_vp, C.c_int, _kp, _dp, _kp, _dp, C.c_int, C.c_int, _dp, _dp, _dp, _dp,
                   C.c_double, _dp, _lp],
    "mb_elbo_monitoring": [_vp, C.c_int, _kp, _dp, _kp, _dp, C.c_int, _dp, _dp, _dp, _dp, _dp,
                           C.c_double, _dp],
    "mb_train_magmaclust": [_vp, C.c_int, _kp, _dp, _kp, _dp, C.c_int, C.c_int, _dp, _dp, C.c_double, C.c_int, C.c_double,
                            C.c_int, _dp, _ip, _ip, _dp, _dp, _dp, _dp],
    "mb_set_state": [_vp, _kp, _dp, _kp, _dp, _dp],
    "mb_em_step_resident": [_vp, _kp, _kp, C.c_int, C.c_double, C.c_int, _dp, _lp],
    "mb_get_state": [_vp, _kp, _dp, _kp, _dp],
    "mb_timer": [_vp, C.c_int],
    "mb_timer_ms": [_vp, _dp],
    "mb_profile": [_vp, C.c_int],
    "mb_profile_read": [_vp, _dp],
    "mb_profile_read_gaps": [_vp, _dp],
    "mb_profile_read_phases": [_vp, _dp],
    "mb_mstep_cycles": [_vp, _dp],
    "mb_flush_l2": [_vp, C.c_int64],
    "mb_matmul": [_vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.c_int, _dp, C.c_int, _dp, C.c_int],
    "mb_bench_dense": [_vp, C.c_int, C.c_int, C.c_int, _dp],
    "mb_bench_task_inverse": [_vp, _kp, C.c_double, C.c_int, _dp, _dp],
    "mb_nccl_unique_id": [C.c_char_p, C.c_int],
    "mb_nccl_init": [_vp, C.c_char_p, C.c_int, C.c_int, C.c_int],
    "mb_estep_leaves": [C.c_int64, C.c_int32],
    "mb_shard_range": [C.c_int64, C.c_int32, C.c_int, C.c_int, _lp, _lp],
    "mb_set_shard": [_vp, C.c_int64, C.c_int64],
    "mb_hyperpost_keep": [_vp, C.c_int],
    "mb_hyperpost_set": [_vp, C.c_int, C.c_int32, _dp, _dp],
    "mb_logL_GP_tasks": [_vp, _kp, _dp, C.c_int, C.c_int, C.c_double, _dp, _dp, _ip],
    "mb_train_gp_tasks": [_vp, _kp, _dp, C.c_int, C.c_double, _lp],
    "mb_train_shared_gp": [_vp, _kp, _dp, C.c_double, C.c_double, C.c_int, _lp],
    "mb_train_gp_clust_tasks": [_vp, _kp, _dp, _dp, C.c_double, C.c_int, C.c_double, _dp, _ip, _ip, _lp],
    "mb_pred_magma_tasks": [_vp, _kp, _dp, _dp, C.c_int32, _dp, _ip, C.c_double, _dp, _dp, _dp, _dp, _ip],
    "mb_r_signif": [C.c_double, C.c_int],
    "mb_r_as_character": [C.c_double, C.c_char_p, C.c_int],
    "mb_reference_key": [_dp, C.c_int, C.c_int, C.c_char_p, C.c_int],
    "mb_prepare_rows": [C.c_int64, C.c_int, _dp, _dp, C.POINTER(C.c_char_p), C.c_int, _dp, C.c_int64, _lp, _dp, _dp, _lp, _lp,
                        _ip, _ip, _dp],
    "mb_bench_hbm": [_vp, C.c_int, _kp, C.c_int, _dp, _dp],
    "mb_lbfgsb_selftest": [C.c_int, _dp, C.c_int, C.c_double, C.c_int, _dp, _ip, _ip],
}


def lib():
    """The loaded shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "libmagma_b200.so is not built (%s). Run `python -m magmaclustr_b200.build` "
                "or __graft_entry__.build(); there is no CPU fallback." % LIB_PATH)
        l = C.CDLL(LIB_PATH)
        for name, sig in _SIGS.items():
            fn = getattr(l, name)
            fn.argtypes = sig
            fn.restype = C.c_int64 if name == "mb_launch_count" else (C.c_double if name == "mb_r_signif" else C.c_int)
        l.mb_last_error.restype = C.c_char_p
        l.mb_kernel_hp_name.argtypes = [_kp, C.c_int]
        l.mb_kernel_hp_name.restype = C.c_char_p
        _lib = l
    return _lib


def check(rc):
    if rc != 0:
        raise MagmaError(rc, lib().mb_last_error().decode())


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    if a is None:
        return None
    if a.dtype == np.float64:
        return a.ctypes.data_as(_dp)
    if a.dtype == np.int32:
        return a.ctypes.data_as(_ip)
    if a.dtype == np.int64:
        return a.ctypes.data_as(_lp)
    raise TypeError(a.dtype)


def parse_kernel(kern, noise):
    k = MbKernel()
    check(lib().mb_kernel_parse(kern.encode(), 1 if noise else 0, C.byref(k)))
    return k


def hp_names(k):
    return [lib().mb_kernel_hp_name(C.byref(k), p).decode() for p in range(k.n_hp)]
