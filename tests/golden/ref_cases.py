"""The inputs of the golden cases computed by the compiled reference (make_golden_ref.py): regenerated from
seeds / recipes, so that the fixture file only has to carry OUTPUTS (and hashes of the large ones).
Shared by the generator and by the tests."""
import hashlib

import numpy as np


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def cases3d():
    """name -> dict(u0, f, h, tol, grad_u or None).  Sizes: small ones store full outputs, large ones hashes."""
    out = {}
    rng = np.random.default_rng(20261017)

    def rand(dims, nsrc, h, lo=0.5):
        f = lo + rng.random(dims)
        u0 = np.full(dims, 1000.0)
        for _ in range(nsrc):
            u0[tuple(rng.integers(0, d) for d in dims)] = float(rng.random() * 0.1)
        return u0, f, h

    for name, dims, nsrc, tol, adj in (("r_9x7x6", (9, 7, 6), 1, 1e-6, True), ("r_12x11x10", (12, 11, 10), 2, 1e-6, True),
                                      ("r_2x2x2", (2, 2, 2), 1, 1e-6, True), ("r_5x4x3_cap", (5, 4, 3), 1, 0.0, True),
                                      ("r_17x2x33", (17, 2, 33), 2, 1e-9, False), ("r_24x19x15_cap", (24, 19, 15), 3, 0.0, False),
                                      ("r_40x33x18_tol1e-3", (40, 33, 18), 2, 1e-3, False), ("r_64cubed", (64, 64, 64), 1, 1e-6, False)):
        u0, f, h = rand(dims, nsrc, 0.3)
        out[name] = dict(u0=u0, f=f, h=h, tol=tol, grad_u=rng.standard_normal(dims) if adj else None)
    # deps/CustomOps/Eikonal3D/gradtest.jl:14-23: 21^3, f = 1, h = 0.01, source (10,10,10) 0-based
    u0 = np.full((21, 21, 21), 1000.0)
    u0[10, 10, 10] = 0.0
    out["gradtest_21cubed"] = dict(u0=u0, f=np.ones((21, 21, 21)), h=0.01, tol=1e-6, grad_u=None)
    # tests/test3d.jl:13-26: 51^3, f = 1 with f[5:8,5:8,5:8] = 2 (1-based), source (10,10,10) 1-based, h = 5, tol 1e-6
    f = np.ones((51, 51, 51))
    f[4:8, 4:8, 4:8] = 2.0
    u0 = np.full((51, 51, 51), 1000.0)
    u0[9, 9, 9] = 0.0
    out["test3d_jl_51cubed"] = dict(u0=u0, f=f, h=5.0, tol=1e-6, grad_u=None)
    # an adjoint on an UNCONVERGED field (one round): the assembly rules must hold for any u
    u0, f, h = rand((10, 9, 8), 2, 0.3)
    out["r_10x9x8_tol_huge"] = dict(u0=u0, f=f, h=h, tol=1e9, grad_u=rng.standard_normal((10, 9, 8)))
    return out


def cases2d():
    """name -> dict(f (rows, cols), h, ix, jx (0-based), grad_u)."""
    out = {}
    rng = np.random.default_rng(233)
    # tests/2D_test.jl:17-24 model: 30 x 40, 1/6 background, two blocks
    f = np.ones((30, 40)) / 6.0
    f[15:20, 19:24] = 1.0 / 5.0
    f[7:14, 9:18] = 1.0 / 7.0
    for k, (ix, jx) in enumerate(((4, 7), (39, 29), (0, 0), (20, 15))):
        out[f"test2d_jl_src{k}"] = dict(f=f, h=1.0, ix=ix, jx=jx, grad_u=rng.standard_normal(f.shape))
    # deps/CustomOps/Eikonal/gradtest.jl:41-51: 31 rows x 61 cols, f = 1 with rows 12-18 (1-based) = 10, srcx = 30, srcy = 3, h = 0.1
    f = np.ones((31, 61))
    f[11:18, :] = 10.0
    out["gradtest_31x61"] = dict(f=f, h=0.1, ix=29, jx=2, grad_u=rng.standard_normal(f.shape))
    f = 0.5 + rng.random((8, 5))
    out["r_8x5"] = dict(f=f, h=0.7, ix=4, jx=0, grad_u=rng.standard_normal(f.shape))
    return out
