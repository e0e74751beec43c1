"""ctypes bindings for oracle/_ref/libref_eikonal.so: the REFERENCE's own Eikonal solvers
(deps/CustomOps/Eikonal/Eikonal.h, deps/CustomOps/Eikonal3D/Eikonal3D.cpp), compiled unmodified from
/root/reference against oracle/eigen_stub (see oracle/Makefile).  TEST INFRASTRUCTURE: used to pin the oracle
and to generate the committed golden vectors (tests/golden/make_golden_ref.py).  The library can only be BUILT
where /root/reference is mounted; a prebuilt copy travels with the repo snapshot (oracle/_ref/ is git-ignored)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libref_eikonal.so")
REF_SRC = "/root/reference/deps/CustomOps"
_lib = None
_dp = ctypes.POINTER(ctypes.c_double)


def can_build():
    return os.path.exists(os.path.join(REF_SRC, "Eikonal3D", "Eikonal3D.cpp"))


def available():
    """True when the compiled reference exists or can be built here."""
    return os.path.exists(_SO) or can_build()


def build(force=False):
    if can_build():
        subprocess.check_call(["make", "-s", "-C", _HERE] + (["-B"] if force else []) + ["_ref/libref_eikonal.so"])
    if not os.path.exists(_SO):
        raise RuntimeError("oracle/_ref/libref_eikonal.so is missing and /root/reference is not mounted")
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_SO)
        i, d = ctypes.c_int, ctypes.c_double
        L.ref_eikonal3d_forward.restype = None
        L.ref_eikonal3d_forward.argtypes = [_dp, _dp, _dp, d, i, i, i, d, i]
        L.ref_eikonal3d_backward.restype = None
        L.ref_eikonal3d_backward.argtypes = [_dp, _dp, _dp, _dp, _dp, _dp, d, i, i, i]
        L.ref_eikonal2d_forward.restype = None
        L.ref_eikonal2d_forward.argtypes = [_dp, _dp, i, i, d, i, i]
        L.ref_eikonal2d_backward.restype = None
        L.ref_eikonal2d_backward.argtypes = [_dp, _dp, _dp, _dp, i, i, d, i, i]
        _lib = L
    return _lib


def _c(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


def eikonal3d_forward(u0, f, h, tol=1e-6, verbose=False):
    """Eikonal3D::forward (Eikonal3D.cpp:90-94).  u0, f: (m, n, l)."""
    u0, u0p = _c(u0)
    f, fp = _c(f)
    m, n, l = u0.shape
    u = np.empty_like(u0)
    lib().ref_eikonal3d_forward(u.ctypes.data_as(_dp), u0p, fp, float(h), m, n, l, float(tol), int(bool(verbose)))
    return u


def eikonal3d_backward(grad_u, u, u0, f, h):
    """Eikonal3D::backward (Eikonal3D.cpp:96-198) -> (grad_u0, grad_f).  Dense-LU stub: keep m*n*l <= ~4000."""
    g, gp = _c(grad_u)
    u, up = _c(u)
    u0, u0p = _c(u0)
    f, fp = _c(f)
    m, n, l = u.shape
    gu0 = np.empty_like(u)
    gf = np.empty_like(u)
    lib().ref_eikonal3d_backward(gu0.ctypes.data_as(_dp), gf.ctypes.data_as(_dp), gp, up, u0p, fp, float(h), m, n, l)
    return gu0, gf


def eikonal2d_forward(f, h, ix, jx):
    """forward (Eikonal.h:54-93).  f: (n+1, m+1) [row j, col i]; ix, jx 0-based."""
    f, fp = _c(f)
    n1, m1 = f.shape
    u = np.empty_like(f)
    lib().ref_eikonal2d_forward(u.ctypes.data_as(_dp), fp, m1 - 1, n1 - 1, float(h), int(ix), int(jx))
    return u


def eikonal2d_backward(grad_u, u, f, h, ix, jx):
    """backward (Eikonal.h:95-200) -> grad_f."""
    f, fp = _c(f)
    u, up = _c(u)
    g, gp = _c(grad_u)
    n1, m1 = f.shape
    gf = np.empty_like(f)
    lib().ref_eikonal2d_backward(gf.ctypes.data_as(_dp), gp, up, fp, m1 - 1, n1 - 1, float(h), int(ix), int(jx))
    return gf
