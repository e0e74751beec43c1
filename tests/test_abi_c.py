"""The C header against the binary: tests/abi_check.c includes ONLY include/adtomo_b200.h and links
libadtomo_b200.so -- what a Julia ccall / cgo / JNI binding of the reference's maintainers would see.
CPU: it compiles, links and finds every symbol.  GPU (-m gpu): it runs forward + backward + the fused step on a
9 x 7 x 6 problem and the results are compared with the oracle (forward bit for bit, gradients to 1e-10)."""
import os
import subprocess

import numpy as np
import pytest

import ref_misfit

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def abi_exe(lib, tmp_path_factory):
    d = tmp_path_factory.mktemp("abi")
    exe = str(d / "abi_check")
    libdir = os.path.dirname(lib.LIB_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "abi_check.c"), "-o", exe, "-L", libdir, "-ladtomo_b200", "-lm",
                           "-Wl,-rpath," + libdir])
    return exe


def test_header_links_against_library(abi_exe):
    out = subprocess.run([abi_exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert "symbols linked" in out.stdout


def _lcg(n, seed):
    out = np.empty(n)
    s = seed
    for q in range(n):
        s = (s * 1664525 + 1013904223) & 0xFFFFFFFF
        out[q] = (s >> 8) / 16777216.0
    return out, s


@pytest.mark.gpu
def test_c_caller_matches_oracle(abi_exe, oracle, tmp_path):
    m, n, l = 9, 7, 6
    N = m * n * l
    out = subprocess.run([abi_exe, str(tmp_path / "out.bin")], capture_output=True, text=True)
    assert out.returncode == 0, out.stdout + out.stderr
    raw = np.fromfile(tmp_path / "out.bin", dtype=np.float64)
    assert raw.size == 5 * N + 1
    u, gu0, gf, ub, packed = raw[:N], raw[N:2 * N], raw[2 * N:3 * N], raw[3 * N:4 * N], raw[4 * N:]
    r, s = _lcg(N, 12345)
    f = (0.5 + r).reshape(m, n, l)
    gu = (_lcg(N, s)[0] - 0.5).reshape(m, n, l)
    u0 = np.full((m, n, l), 1000.0)
    u0[4, 3, 2] = 0.0
    h, tol = 0.3, 1e-9
    ur, _, _ = oracle.eikonal3d_forward(u0, f, h, tol)
    assert np.array_equal(u.reshape(m, n, l), ur)
    assert np.array_equal(ub.reshape(m, n, l), ur)
    g0r, gfr = oracle.eikonal3d_backward(gu, ur, u0, f, h)[:2]
    assert np.array_equal(gu0.reshape(m, n, l), g0r)
    assert np.abs(gf.reshape(m, n, l) - gfr).max() <= 1e-10 * np.abs(gfr).max()
    rcv = np.array([[1.0, 1.0, 1.0], [7.25, 5.5, 4.75], [2.0, 6.0, 0.5]])
    mis, gur = ref_misfit.misfit_and_grad_u(ur, rcv, np.array([1.0, -1.0, 2.5]), np.array([1.0, 0.7, 0.4]))
    gfm = oracle.eikonal3d_backward(gur, ur, u0, f, h)[1]
    assert abs(packed[N] - mis) <= 1e-12 * abs(mis)
    assert np.abs(packed[:N].reshape(m, n, l) - gfm).max() <= 1e-10 * np.abs(gfm).max()
