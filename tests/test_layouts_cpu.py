"""CPU: the level-major layout machinery of the CUDA forward kernel (csrc/layouts.h) and a serial
host emulation of the kernel's two-phase level loop (tests/emul/emulate_v1.cpp) against the oracle.
This pins the index maps / buffer rotation of kernels_fwd_v1.cuh bit for bit without a GPU."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
_dp = ctypes.POINTER(ctypes.c_double)


@pytest.fixture(scope="module")
def emul():
    so = os.path.join(HERE, "emul", "libemul_v1.so")
    src = os.path.join(HERE, "emul", "emulate_v1.cpp")
    deps = [src] + [os.path.join(HERE, "..", "adtomo.jl_b200", "csrc", n) for n in ("layouts.h", "eik_core.h")]
    if not os.path.exists(so) or any(os.path.getmtime(so) < os.path.getmtime(d) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-std=c++17", "-shared", "-fPIC", "-o", so, src])
    L = ctypes.CDLL(so)
    L.emul_fwd3d_v1.restype = ctypes.c_int
    L.emul_fwd3d_v1.argtypes = [_dp, _dp, _dp, ctypes.c_double, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                ctypes.c_double, ctypes.c_int, _dp]
    L.emul_layout_offsets.restype = ctypes.c_int
    L.emul_layout_offsets.argtypes = [ctypes.c_int] * 4 + [ctypes.POINTER(ctypes.c_int)]
    return L


@pytest.mark.parametrize("dims", [(2, 2, 2), (5, 4, 3), (3, 9, 4), (8, 8, 8), (7, 3, 11), (16, 12, 6)])
def test_layouts_are_injective(emul, dims):
    m, n, l = dims
    N = m * n * l
    for q in range(5):
        out = np.empty(N, dtype=np.int32)
        assert emul.emul_layout_offsets(m, n, l, q, out.ctypes.data_as(ctypes.POINTER(ctypes.c_int))) == 0
        assert len(set(out.tolist())) == N and out.min() >= 0, f"layout {q} is not injective"


@pytest.mark.parametrize("dims,tol", [((2, 2, 2), 1e-9), ((5, 4, 3), 1e-9), ((3, 9, 4), 1e-6), ((9, 7, 6), 1e-6),
                                      ((12, 12, 12), 1e-3), ((7, 3, 11), 0.0), ((16, 12, 6), 1e-6),
                                      ((6, 16, 12), 1e-6), ((12, 6, 16), 1e-4), ((24, 19, 15), 1e-3)])
def test_emulated_kernel_bitexact(emul, oracle, dims, tol):
    rng = np.random.default_rng(sum(dims))
    f = 0.5 + rng.random(dims)
    u0 = np.full(dims, 1000.0)
    for _ in range(2):
        u0[tuple(rng.integers(0, d) for d in dims)] = float(rng.random() * 0.1)
    h = 0.3
    u_ref, r_ref, e_ref = oracle.eikonal3d_forward(u0, f, h, tol)
    u = np.empty_like(u0)
    err = ctypes.c_double(0)
    r = emul.emul_fwd3d_v1(u.ctypes.data_as(_dp), u0.ctypes.data_as(_dp), f.ctypes.data_as(_dp), h, *dims, tol, 20,
                           ctypes.byref(err))
    assert abs(r) == r_ref and (r > 0) == (tol > 0)
    assert err.value == e_ref
    np.testing.assert_array_equal(u, u_ref)
