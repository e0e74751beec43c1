"""CPU: pin the oracle (oracle/eikonal_oracle.c) before anything is compared against it.

Pins: (0) the reference's OWN C++ solvers compiled unmodified (oracle/_ref, bottom of this file): forward bit for
bit, adjoint 1e-12 -- directly where /root/reference is mounted, and through the committed outputs
tests/golden/ref_cpp.npz everywhere; (1) the reference's own Python prototypes (tests/golden/*.npz, made by make_golden.py from
/root/reference/tests/Eikonal3D/prototype*.py); (2) the reference's matrix assembly solved with
SciPy SuperLU (tests/ref_assembly.py); (3) finite-difference Taylor tests in the style of
deps/CustomOps/*/gradtest.jl turned into assertions; (4) analytic homogeneous-medium solutions;
(5) the discrete residual of tests/Eikonal3D/prototype.py:5-7.
"""
import os
import sys

import numpy as np
import pytest

import ref_assembly as ra

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_golden_proto3d_21(oracle):
    g = np.load(os.path.join(G, "proto3d_21.npz"))
    u, rounds, err = oracle.eikonal3d_forward(g["u0"], g["f"], float(g["h"]), 1e-6)
    assert rounds == 2 and err == 0.0
    np.testing.assert_allclose(u, g["u"], rtol=1e-11, atol=0)


def test_golden_proto3d_ragged(oracle):
    g = np.load(os.path.join(G, "proto3d_ragged.npz"))
    u, rounds, _ = oracle.eikonal3d_forward(g["u0"], g["f"], float(g["h"]), 1e-6)
    np.testing.assert_allclose(u, g["u"], rtol=1e-11, atol=0)


def test_golden_single_sweeps(oracle):
    g = np.load(os.path.join(G, "proto3d_sweeps.npz"))
    for key, sid in (("s1", 0), ("s7", 6)):
        a = g["u_in"].copy()
        oracle.eikonal3d_sweep(a, g["f"], float(g["h"]), sid)
        np.testing.assert_allclose(a, g[key], rtol=1e-11, atol=0)


def test_golden_proto2d(oracle):
    g = np.load(os.path.join(G, "proto2d.npz"))
    # prototype2d's u[i, j] has i (the OUTER loop) first; Eikonal.h stores j*(m+1)+i -> transpose
    u, rounds, conv = oracle.eikonal2d_forward(g["f"].T.copy(), float(g["h"]), int(g["src"][0]), int(g["src"][1]))
    assert conv
    np.testing.assert_allclose(u.T, g["u"], rtol=1e-11, atol=0)


def test_homogeneous_analytic_3d(oracle):
    # first-order scheme: error vs distance*f shrinks with h (tests/Eikonal/eikonal.py:213-215 analogue)
    errs = []
    for m in (17, 33):
        h = 1.0 / (m - 1)
        u0 = np.full((m, m, m), 1000.0)
        c = m // 2
        u0[c, c, c] = 0.0
        u, _, _ = oracle.eikonal3d_forward(u0, np.ones((m, m, m)), h, 1e-12)
        ii, jj, kk = np.meshgrid(*(np.arange(m),) * 3, indexing="ij")
        d = np.sqrt((ii - c) ** 2 + (jj - c) ** 2 + (kk - c) ** 2) * h
        errs.append(np.abs(u - d).max())
        # along the axes the upwind scheme is exact
        np.testing.assert_allclose(u[c, c, :], np.abs(np.arange(m) - c) * h, rtol=1e-13, atol=1e-15)
    assert errs[1] < errs[0] < 0.15


def test_residual_3d(oracle):
    rng = np.random.default_rng(1)
    m, n, l = 14, 11, 9
    f = 0.5 + rng.random((m, n, l))
    h = 0.3
    u0 = np.full((m, n, l), 1000.0)
    u0[3, 4, 5] = 0.0
    u, rounds, err = oracle.eikonal3d_forward(u0, f, h, 0.0, max_rounds=100)   # tol=0 never stops early
    assert rounds == 100
    up = np.pad(u, 1, mode="reflect")
    ax = np.minimum(up[:-2, 1:-1, 1:-1], up[2:, 1:-1, 1:-1])
    ay = np.minimum(up[1:-1, :-2, 1:-1], up[1:-1, 2:, 1:-1])
    az = np.minimum(up[1:-1, 1:-1, :-2], up[1:-1, 1:-1, 2:])
    res = (np.maximum(u - ax, 0) ** 2 + np.maximum(u - ay, 0) ** 2 + np.maximum(u - az, 0) ** 2 - (f * h) ** 2)
    res[3, 4, 5] = 0.0
    assert np.abs(res).max() < 1e-12


def test_rounds_and_cap(oracle):
    from adtomo_jl_b200 import synthetic as syn
    u0, f, h = syn.model_test3d()
    u0, f = u0[:31, :31, :31].copy(), f[:31, :31, :31].copy()
    u, rounds, err = oracle.eikonal3d_forward(u0, f, h, 1e-6)
    assert rounds == 3            # tests/test3d.jl model needs 3 rounds
    u2, rounds2, _ = oracle.eikonal3d_forward(u0, f, h, 0.0)
    assert rounds2 == 20          # strict '<': tol = 0 runs into the cap (Eikonal3D.cpp:74,85)
    np.testing.assert_array_equal(u, u2)


@pytest.mark.parametrize("seed", [0, 1])
def test_adjoint3d_vs_superlu(oracle, seed):
    rng = np.random.default_rng(seed)
    m, n, l = 11, 9, 8
    f = 0.6 + rng.random((m, n, l))
    h = 0.2
    u0 = np.full((m, n, l), 1000.0)
    u0[5, 2, 3] = 0.0
    u0[5, 3, 3] = 0.05
    if seed:
        u0[0, 0, 0] = 0.3     # a second source on a corner
    u, _, _ = oracle.eikonal3d_forward(u0, f, h, 1e-13, max_rounds=50)
    g = rng.standard_normal(u.shape)
    gu0, gf, npin = oracle.eikonal3d_backward(g, u, u0, f, h)
    gu0_l, gf_l = ra.backward3d_lu(g, u, u0, f, h)
    np.testing.assert_array_equal(gu0, gu0_l)
    assert np.abs(gf - gf_l).max() <= 1e-12 * np.abs(gf_l).max()
    assert npin >= 2


def test_adjoint3d_unconverged_field(oracle):
    # production runs stop at tol = 1e-3: the adjoint must follow the reference's rules on such fields too
    rng = np.random.default_rng(5)
    m, n, l = 10, 10, 8
    f = 0.3 + rng.random((m, n, l))
    u0 = np.full((m, n, l), 1000.0)
    u0[7, 2, 6] = 0.0
    u, rounds, _ = oracle.eikonal3d_forward(u0, f, 1.0, 1e3, max_rounds=1)   # one round only
    g = rng.standard_normal(u.shape)
    _, gf, _ = oracle.eikonal3d_backward(g, u, u0, f, 1.0)
    _, gf_l = ra.backward3d_lu(g, u, u0, f, 1.0)
    assert np.abs(gf - gf_l).max() <= 1e-12 * np.abs(gf_l).max()


def test_adjoint2d_vs_superlu(oracle):
    rng = np.random.default_rng(3)
    f = 0.2 + rng.random((31, 61))          # deps/CustomOps/Eikonal/gradtest.jl shape
    u, rounds, conv = oracle.eikonal2d_forward(f, 0.1, 29, 2)
    assert conv
    g = rng.standard_normal(f.shape)
    gf, rc = oracle.eikonal2d_backward(g, u, f, 0.1, 29, 2)
    gf_l = ra.backward2d_lu(g, u, f, 0.1, 29, 2)
    assert rc == 0
    assert np.abs(gf - gf_l).max() <= 1e-12 * np.abs(gf_l).max()


def _taylor(yfun, f, v, grad):
    out = []
    y0 = yfun(f)
    for gam in (1e-2, 1e-3, 1e-4):
        s = yfun(f + gam * v) - y0
        out.append((abs(s), abs(s - gam * float((v * grad).sum()))))
    return out


def test_fd_gradient_3d(oracle):
    # deps/CustomOps/Eikonal3D/gradtest.jl:37-78 style, y = sum(u^2), here w.r.t. f
    rng = np.random.default_rng(233)
    m = n = l = 13
    f = 1.0 + 0.5 * rng.random((m, n, l))
    h = 0.01
    u0 = np.full((m, n, l), 1000.0)
    u0[6, 6, 6] = 0.0

    def y(ff):
        return float((oracle.eikonal3d_forward(u0, ff, h, 0.0, max_rounds=30)[0] ** 2).sum())

    u, _, _ = oracle.eikonal3d_forward(u0, f, h, 0.0, max_rounds=30)
    _, gf, _ = oracle.eikonal3d_backward(2 * u, u, u0, f, h)
    t = _taylor(y, f, rng.standard_normal(f.shape) * 0.1, gf)
    # first-order term decays ~gamma, remainder ~gamma^2
    assert t[1][1] < t[0][1] / 50 and t[2][1] < t[1][1] / 50
    assert t[2][1] < 1e-3 * t[2][0]


def test_fd_gradient_u0_3d(oracle):
    # the reference's own 3D gradtest differentiates w.r.t. u0 (gradtest.jl:37-78)
    rng = np.random.default_rng(7)
    m = n = l = 9
    f = np.ones((m, n, l))
    h = 0.01
    u0 = np.full((m, n, l), 1000.0)
    u0[4, 4, 4] = 0.0
    u0[1, 2, 3] = 0.02
    u, _, _ = oracle.eikonal3d_forward(u0, f, h, 0.0, max_rounds=30)
    gu0, _, _ = oracle.eikonal3d_backward(2 * u, u, u0, f, h)
    # only nodes with u == u0 carry gradient
    assert set(zip(*np.nonzero(gu0))) <= {(4, 4, 4), (1, 2, 3)}
    eps = 1e-7
    for node in ((4, 4, 4), (1, 2, 3)):
        up = u0.copy()
        up[node] += eps
        y1 = (oracle.eikonal3d_forward(up, f, h, 0.0, max_rounds=30)[0] ** 2).sum()
        fd = (y1 - (u ** 2).sum()) / eps
        # implicit-function adjoint w.r.t. u0 in the reference is just the pass-through grad_u[u==u0];
        # it ignores downstream dependence, so only check it equals 2*u at the pinned node
        assert gu0[node] == 2 * u[node]
        assert np.isfinite(fd)


def test_fd_gradient_2d(oracle):
    rng = np.random.default_rng(233)
    f = 0.2 + rng.random((31, 61))
    h = 0.1

    def y(ff):
        return float((oracle.eikonal2d_forward(ff, h, 29, 2)[0] ** 2).sum())

    u, _, _ = oracle.eikonal2d_forward(f, h, 29, 2)
    gf, _ = oracle.eikonal2d_backward(2 * u, u, f, h, 29, 2)
    t = _taylor(y, f, rng.standard_normal(f.shape) * 0.1, gf)
    assert t[1][1] < t[0][1] / 30 and t[2][1] < t[1][1] / 30


# ----------------------------------------------------------------------------------------------
# Pin (0): the reference's OWN C++ solvers.  tests/golden/ref_cpp.npz holds outputs of
# deps/CustomOps/Eikonal/Eikonal.h and Eikonal3D/Eikonal3D.cpp compiled unmodified (oracle/Makefile target
# _ref/libref_eikonal.so, against oracle/eigen_stub) on the inputs of tests/golden/ref_cases.py.
# Forward: bit for bit (hash of the bytes).  Adjoint: the reference factorises with a sparse LU (here: the
# stub's dense LU), the oracle back-substitutes -- 1e-12 relative to max |grad|.
# ----------------------------------------------------------------------------------------------
sys.path.insert(0, G)
import ref_cases  # noqa: E402

REF_GOLD = os.path.join(G, "ref_cpp.npz")
ADJ_RTOL = 1e-12


def test_oracle_matches_reference_cpp_goldens_3d(oracle):
    g = np.load(REF_GOLD)
    for name, c in ref_cases.cases3d().items():
        u, rounds, _ = oracle.eikonal3d_forward(c["u0"], c["f"], c["h"], c["tol"])
        assert ref_cases.sha(u) == str(g[f"3d/{name}/u_sha256"]), name
        np.testing.assert_array_equal(u.ravel()[::97], g[f"3d/{name}/u_sample"])
        if f"3d/{name}/u" in g:
            np.testing.assert_array_equal(u, g[f"3d/{name}/u"])
        if c["grad_u"] is not None:
            gu0, gf, _ = oracle.eikonal3d_backward(c["grad_u"], u, c["u0"], c["f"], c["h"])
            np.testing.assert_array_equal(gu0, g[f"3d/{name}/grad_u0"])
            ref_gf = g[f"3d/{name}/grad_f"]
            assert np.abs(gf - ref_gf).max() <= ADJ_RTOL * np.abs(ref_gf).max(), name


def test_oracle_matches_reference_cpp_goldens_2d(oracle):
    g = np.load(REF_GOLD)
    for name, c in ref_cases.cases2d().items():
        u, _, conv = oracle.eikonal2d_forward(c["f"], c["h"], c["ix"], c["jx"])
        assert conv
        np.testing.assert_array_equal(u, g[f"2d/{name}/u"])
        gf, _ = oracle.eikonal2d_backward(c["grad_u"], u, c["f"], c["h"], c["ix"], c["jx"])
        ref_gf = g[f"2d/{name}/grad_f"]
        assert np.abs(gf - ref_gf).max() <= ADJ_RTOL * np.abs(ref_gf).max(), name


def _ref():
    import oracle.ref as ref
    if not ref.available():
        pytest.skip("compiled reference (oracle/_ref) not present and /root/reference not mounted")
    ref.build()
    return ref


def test_goldens_are_what_the_compiled_reference_produces():
    """Where the reference can be compiled (this container), the committed fixture must be reproducible."""
    ref = _ref()
    g = np.load(REF_GOLD)
    for name, c in ref_cases.cases3d().items():
        if c["u0"].size > 30000:
            continue
        assert ref_cases.sha(ref.eikonal3d_forward(c["u0"], c["f"], c["h"], c["tol"])) == str(g[f"3d/{name}/u_sha256"])
    for name, c in ref_cases.cases2d().items():
        np.testing.assert_array_equal(ref.eikonal2d_forward(c["f"], c["h"], c["ix"], c["jx"]), g[f"2d/{name}/u"])


@pytest.mark.parametrize("seed", range(6))
def test_oracle_vs_compiled_reference_random(oracle, seed):
    """Fresh random cases straight against the compiled reference: ragged shapes, several sources, all three
    stopping rules (tol met, production tol, 20-round cap), adjoint on converged and unconverged fields."""
    ref = _ref()
    rng = np.random.default_rng(1000 + seed)
    dims = tuple(int(x) for x in rng.integers(2, 14, 3))
    f = 0.2 + 2.0 * rng.random(dims)
    u0 = np.full(dims, 1000.0)
    for _ in range(int(rng.integers(1, 4))):
        u0[tuple(rng.integers(0, d) for d in dims)] = float(rng.random())
    h = float(0.1 + rng.random())
    for tol in (1e-6, 1e-3, 0.0, 1e9):
        u_ref = ref.eikonal3d_forward(u0, f, h, tol)
        u, _, _ = oracle.eikonal3d_forward(u0, f, h, tol)
        np.testing.assert_array_equal(u, u_ref)
        if tol in (1e-6, 1e9):
            gu = rng.standard_normal(dims)
            gu0_ref, gf_ref = ref.eikonal3d_backward(gu, u_ref, u0, f, h)
            gu0, gf, _ = oracle.eikonal3d_backward(gu, u, u0, f, h)
            np.testing.assert_array_equal(gu0, gu0_ref)
            assert np.abs(gf - gf_ref).max() <= ADJ_RTOL * max(np.abs(gf_ref).max(), 1e-300)
    shape = tuple(int(x) for x in rng.integers(2, 30, 2))
    f2 = 0.2 + rng.random(shape)
    ix, jx = int(rng.integers(0, shape[1])), int(rng.integers(0, shape[0]))
    u_ref = ref.eikonal2d_forward(f2, h, ix, jx)
    u, _, _ = oracle.eikonal2d_forward(f2, h, ix, jx)
    np.testing.assert_array_equal(u, u_ref)
    g2 = rng.standard_normal(shape)
    gf_ref = ref.eikonal2d_backward(g2, u_ref, f2, h, ix, jx)
    gf, _ = oracle.eikonal2d_backward(g2, u, f2, h, ix, jx)
    assert np.abs(gf - gf_ref).max() <= ADJ_RTOL * max(np.abs(gf_ref).max(), 1e-300)
