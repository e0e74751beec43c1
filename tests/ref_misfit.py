"""numpy restatement of the drivers' receiver sampling + misfit (scripts/inversion.jl:64-105) and of
its derivative with respect to the travel-time field -- test infrastructure (checker only)."""
import math

import numpy as np


def sample(u, p):
    """Trilinear rule with the degenerate-axis shortcuts of inversion.jl:72-93 (0-based coordinates)."""
    jx, jy, jz = p
    x1, x2 = math.floor(jx), math.ceil(jx)
    y1, y2 = math.floor(jy), math.ceil(jy)
    z1, z2 = math.floor(jz), math.ceil(jz)
    if x1 == x2:
        tx11, tx12, tx21, tx22 = u[x1, y1, z1], u[x1, y1, z2], u[x1, y2, z1], u[x1, y2, z2]
    else:
        tx11 = (x2 - jx) * u[x1, y1, z1] + (jx - x1) * u[x2, y1, z1]
        tx12 = (x2 - jx) * u[x1, y1, z2] + (jx - x1) * u[x2, y1, z2]
        tx21 = (x2 - jx) * u[x1, y2, z1] + (jx - x1) * u[x2, y2, z1]
        tx22 = (x2 - jx) * u[x1, y2, z2] + (jx - x1) * u[x2, y2, z2]
    if y1 == y2:
        txy1, txy2 = tx11, tx12
    else:
        txy1 = (y2 - jy) * tx11 + (jy - y1) * tx21
        txy2 = (y2 - jy) * tx12 + (jy - y1) * tx22
    if z1 == z2:
        return txy1
    return (z2 - jz) * txy1 + (jz - z1) * txy2


def weights(p):
    """[(node (x,y,z), weight)] such that sample(u,p) == sum w*u[node]."""
    out = []
    axes = []
    for c in p:
        a, b = math.floor(c), math.ceil(c)
        axes.append([(a, 1.0)] if a == b else [(a, b - c), (b, c - a)])
    for (x, wx) in axes[0]:
        for (y, wy) in axes[1]:
            for (z, wz) in axes[2]:
                out.append(((x, y, z), wx * wy * wz))
    return out


def misfit_and_grad_u(u, rcv, uobs, qua):
    """misfit = sum_e qua*(uobs - t_e)^2 over uobs != -1; returns (misfit, d misfit / d u)."""
    g = np.zeros_like(u)
    mis = 0.0
    for e, p in enumerate(rcv):
        if uobs[e] == -1:
            continue
        t = sample(u, p)
        r = uobs[e] - t
        mis += qua[e] * r * r
        for node, w in weights(p):
            g[node] += -2.0 * qua[e] * r * w
    return mis, g
