"""ctypes handle on oracle/_ref/libcode_ref.so -- the REFERENCE'S OWN object code for the five native builders
(/root/reference/src/code.cpp compiled unchanged against the stand-in Rcpp.h of oracle/ref_shim, recipe oracle/Makefile).

Test infrastructure only (imported by tests/ and __graft_entry__.build()): it pins the numpy restatement in
oracle/kernels.py and the CUDA builders (mb_cpp_*) to what the reference's compiled code returns.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libcode_ref.so")
REFERENCE = os.environ.get("MB_REFERENCE_DIR", "/root/reference")
_lib = None


def build():
    """Compile oracle/_ref/libcode_ref.so when the reference tree is present (this container); a no-op on the GPU box,
    which receives the prebuilt file.  Returns the path or None."""
    if os.path.exists(os.path.join(REFERENCE, "src", "code.cpp")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "REFERENCE=" + REFERENCE])
    return SO if os.path.exists(SO) else None


def available():
    return os.path.exists(SO) or build() is not None


def _l():
    global _lib
    if _lib is None:
        if not available():
            raise ImportError("oracle/_ref/libcode_ref.so is not built and %s is absent" % REFERENCE)
        _lib = C.CDLL(SO)
        dp = C.POINTER(C.c_double)
        for name in ("ref_cpp_dist", "ref_cpp_prod"):
            getattr(_lib, name).argtypes = [dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, dp]
        for name in ("ref_cpp_perio", "ref_cpp_perio_deriv", "ref_cpp_noise"):
            getattr(_lib, name).argtypes = [dp, C.c_int, dp, C.c_int, C.c_int, C.c_int, C.c_double, dp]
    return _lib


def _call(name, m1, m2, par=None):
    # R numeric matrices are column-major
    a = np.asfortranarray(np.atleast_2d(np.asarray(m1, dtype=np.float64)))
    b = np.asfortranarray(np.atleast_2d(np.asarray(m2, dtype=np.float64)))
    out = np.zeros((a.shape[0], b.shape[0]), order="F")
    dp = C.POINTER(C.c_double)
    args = [a.ctypes.data_as(dp), a.shape[0], b.ctypes.data_as(dp), b.shape[0], a.shape[1], b.shape[1]]
    if par is not None:
        args.append(float(par))
    args.append(out.ctypes.data_as(dp))
    rc = getattr(_l(), name)(*args)
    if rc != 0:
        raise ValueError("Incompatible number of dimensions")     # the std::runtime_error of code.cpp:12
    return np.ascontiguousarray(out)


def cpp_dist(m1, m2):
    return _call("ref_cpp_dist", m1, m2)


def cpp_prod(m1, m2):
    return _call("ref_cpp_prod", m1, m2)


def cpp_perio(m1, m2, period):
    return _call("ref_cpp_perio", m1, m2, period)


def cpp_perio_deriv(m1, m2, period):
    return _call("ref_cpp_perio_deriv", m1, m2, period)


def cpp_noise(m1, m2, noise):
    return _call("ref_cpp_noise", m1, m2, noise)
