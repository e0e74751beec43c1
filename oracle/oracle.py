"""ctypes bindings + numpy helpers for oracle/eikonal_oracle.c (test infrastructure)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    src = os.path.join(_HERE, "eikonal_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "liboracle.so"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        L.oracle_eikonal2d_forward.restype = ctypes.c_int
        L.oracle_eikonal2d_forward.argtypes = [_dp, _dp, ctypes.c_int, ctypes.c_int, ctypes.c_double,
                                               ctypes.c_int, ctypes.c_int, _ip]
        L.oracle_eikonal2d_backward.restype = ctypes.c_int
        L.oracle_eikonal2d_backward.argtypes = [_dp, _dp, _dp, _dp, ctypes.c_int, ctypes.c_int,
                                                ctypes.c_double, ctypes.c_int, ctypes.c_int]
        L.oracle_eikonal3d_sweep.restype = None
        L.oracle_eikonal3d_sweep.argtypes = [_dp, _dp, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                             ctypes.c_int, ctypes.c_int]
        L.oracle_eikonal3d_forward.restype = ctypes.c_int
        L.oracle_eikonal3d_forward.argtypes = [_dp, _dp, _dp, ctypes.c_double, ctypes.c_int, ctypes.c_int,
                                               ctypes.c_int, ctypes.c_double, ctypes.c_int, ctypes.c_int, _dp]
        L.oracle_eikonal3d_backward.restype = ctypes.c_int
        L.oracle_eikonal3d_backward.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, ctypes.c_double,
                                                ctypes.c_int, ctypes.c_int, ctypes.c_int]
        _lib = L
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def eikonal2d_forward(f, h, ix, jx):
    """f: (n+1, m+1) array [row j, col i]; ix, jx 0-based.  Returns (u, rounds, converged)."""
    f, fp = _c(f)
    n1, m1 = f.shape
    u = np.empty_like(f)
    conv = ctypes.c_int(0)
    it = lib().oracle_eikonal2d_forward(u.ctypes.data_as(_dp), fp, m1 - 1, n1 - 1, float(h), int(ix), int(jx),
                                        ctypes.byref(conv))
    return u, it, bool(conv.value)


def eikonal2d_backward(grad_u, u, f, h, ix, jx):
    f, fp = _c(f)
    u, up = _c(u)
    g, gp = _c(grad_u)
    n1, m1 = f.shape
    gf = np.empty_like(f)
    rc = lib().oracle_eikonal2d_backward(gf.ctypes.data_as(_dp), gp, up, fp, m1 - 1, n1 - 1, float(h), int(ix),
                                         int(jx))
    return gf, rc


def eikonal3d_sweep(u, f, h, sweep_id):
    """One directional sweep in place on a C-contiguous (m,n,l) array."""
    assert u.flags.c_contiguous and u.dtype == np.float64
    f, fp = _c(f)
    m, n, l = u.shape
    lib().oracle_eikonal3d_sweep(u.ctypes.data_as(_dp), fp, float(h), m, n, l, int(sweep_id))
    return u


def eikonal3d_forward(u0, f, h, tol=1e-6, verbose=False, max_rounds=20):
    """Returns (u, rounds, last_err)."""
    u0, u0p = _c(u0)
    f, fp = _c(f)
    m, n, l = u0.shape
    u = np.empty_like(u0)
    err = ctypes.c_double(0.0)
    it = lib().oracle_eikonal3d_forward(u.ctypes.data_as(_dp), u0p, fp, float(h), m, n, l, float(tol),
                                        int(bool(verbose)), int(max_rounds), ctypes.byref(err))
    return u, it, err.value


def eikonal3d_backward(grad_u, u, u0, f, h):
    """Returns (grad_u0, grad_f, n_pinned)."""
    g, gp = _c(grad_u)
    u, up = _c(u)
    u0, u0p = _c(u0)
    f, fp = _c(f)
    m, n, l = u.shape
    gu0 = np.empty_like(u)
    gf = np.empty_like(u)
    npin = lib().oracle_eikonal3d_backward(gu0.ctypes.data_as(_dp), gf.ctypes.data_as(_dp), gp, up, u0p, fp,
                                           float(h), m, n, l)
    return gu0, gf, npin
