"""Synthetic inputs of the named benchmark/test configurations, numpy only.

Recipes restated from the reference's data-preparation scripts:
  GIL7 layered model         scripts/gene_vel0.jl:13-34
  checkerboard perturbation  scripts/gene_check.jl:26-43 (len = 10, +-0.8 km/s: set_config.jl:100-103)
  2D test model              tests/2D_test.jl:17-24
  3D single-source test      tests/test3d.jl:13-26
RNG: numpy default_rng(233) (tests/test3d.jl:7 seeds Julia with 233; its stream is not reproducible).
"""
import numpy as np

GIL7_DEPTH = [0, 1, 3, 4, 5, 17, 25]
GIL7_VP = [3.20, 4.50, 4.80, 5.51, 6.21, 6.89, 7.83]


def gil7_velocity(m, n, l, h, dz=2):
    """Layered P velocity: layer index advances when (k1 - dz)*h >= next depth (k1 is 1-based)."""
    vel = np.ones((m, n, l))
    nl, nvel = 0, GIL7_VP[0]
    for k1 in range(1, l + 1):
        if nl < 6 and (k1 - dz) * h >= GIL7_DEPTH[nl + 1]:
            nl += 1
            nvel = GIL7_VP[nl]
        vel[:, :, k1 - 1] = nvel
    return vel


def checkerboard(vel, length=10, change=0.8):
    m, n, l = vel.shape
    i, j, k = np.meshgrid(np.arange(m) // length, np.arange(n) // length, np.arange(l) // length, indexing="ij")
    sign = np.where((i + j + k) % 2 == 0, 1.0, -1.0)
    return vel + change * sign


def stations_events(m, n, l, S, E, h=1.0, dz=2, seed=233):
    """Fractional 0-based station (near-surface) and event (interior) coordinates."""
    rng = np.random.default_rng(seed)
    sta = np.stack([rng.uniform(2, m - 3, S), rng.uniform(2, n - 3, S), rng.uniform(dz - 1, dz, S)], axis=1)
    zmax = min(l - 3, 15.0 / h + dz)
    eve = np.stack([rng.uniform(2, m - 3, E), rng.uniform(2, n - 3, E), rng.uniform(dz, zmax, E)], axis=1)
    return sta, eve


def model_2d_test():
    """tests/2D_test.jl:17-24: f is 30 x 40 [row, col], background 1/6 with two blocks (1-based ranges)."""
    f = np.ones((30, 40)) / 6.0
    f[15:20, 19:24] = 1.0 / 5.0
    f[7:14, 9:18] = 1.0 / 7.0
    return f


def model_test3d():
    """tests/test3d.jl:13-26: 51^3, f = 1 with f[5:8,5:8,5:8] = 2 (1-based), source node (10,10,10) 1-based,
    h = 5 (line 24 overrides line 16), u0 = 1000 elsewhere."""
    m = n = l = 51
    f = np.ones((m, n, l))
    f[4:8, 4:8, 4:8] = 2.0
    u0 = 1000.0 * np.ones((m, n, l))
    u0[9, 9, 9] = 0.0
    return u0, f, 5.0
