"""ctypes binding of oracle/c/knot_ref.cpp (CPU port of the reference algorithm).

TEST INFRASTRUCTURE / CPU BASELINE ONLY -- see oracle/__init__.py.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libknot_ref.so")
_lib = None


def build(force=False):
    src = os.path.join(_HERE, "c", "knot_ref.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(_HERE, "c")])
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        dp = ctypes.POINTER(ctypes.c_double)
        ints = [ctypes.c_int] * 8
        _lib.pbo_residual.argtypes = ints + [dp, dp, dp, dp, ctypes.c_int]
        _lib.pbo_jacobian.argtypes = ints + [dp, dp, dp, dp, ctypes.c_int]
        _lib.pbo_hessian.argtypes = ints + [dp, dp, dp, dp, dp, ctypes.c_int]
    return _lib


def _p(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _args(prob):
    G0 = np.asfortranarray(prob.G0)
    Gj = np.ascontiguousarray(np.stack([np.asfortranarray(g).reshape(-1, order="F")
                                        for g in prob.Gj])) if prob.m else np.zeros(1)
    return (prob.b, prob.n_b, prob.m, prob.K, prob.D, prob.x_off, prob.dt_off, prob.u_off), G0, Gj


def max_threads():
    return lib().pbo_max_threads()


def residual(prob, Z, threads=0):
    ints, G0, Gj = _args(prob)
    Zf = np.asfortranarray(Z, dtype=float)
    out = np.empty(prob.dim)
    rc = lib().pbo_residual(*ints, _p(G0), _p(Gj), _p(Zf), _p(out), threads)
    assert rc == 0
    return out


def jacobian_values(prob, Z, threads=0):
    ints, G0, Gj = _args(prob)
    Zf = np.asfortranarray(Z, dtype=float)
    out = np.empty((prob.K - 1) * prob.nnz_jac_knot)
    rc = lib().pbo_jacobian(*ints, _p(G0), _p(Gj), _p(Zf), _p(out), threads)
    assert rc == 0
    return out


def hessian_values(prob, Z, mu, threads=0):
    ints, G0, Gj = _args(prob)
    Zf = np.asfortranarray(Z, dtype=float)
    mu = np.ascontiguousarray(mu, dtype=float)
    out = np.empty((prob.K - 1) * prob.nnz_hess_knot)
    rc = lib().pbo_hessian(*ints, _p(G0), _p(Gj), _p(Zf), _p(mu), _p(out), threads)
    assert rc == 0
    return out
