#!/usr/bin/env python
"""Single-source forward (+ adjoint) timing of large grids (BASELINE config C5) under the kernel selected by
the environment (ADTOMO_TEAM / ADTOMO_TEAM_R).  Prints one JSON line per size incl. a sha1 of the field, so runs
with different kernels can be compared bit for bit.   python benchmarks/team_probe.py 256 384 512 [--adj]"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import adtomo_jl_b200 as A
    from adtomo_jl_b200 import synthetic as syn
    sizes = [int(a) for a in sys.argv[1:] if a.isdigit()]
    adj = "--adj" in sys.argv
    model = "checker" if "--checker" in sys.argv else "gil7"
    ctx = A.Context(0)
    dev = torch.device("cuda", 0)
    for sz in sizes:
        m = n = l = sz
        hh = 25.0 / l
        vel = syn.gil7_velocity(m, n, l, hh)
        if model == "checker":
            vel = syn.checkerboard(vel, max(4, sz // 12), 0.8)
        N = m * n * l
        d_u0 = torch.full((1, N), 1000.0, dtype=torch.float64, device=dev)
        d_u0[0, ((m // 2) * n + n // 2) * l + 0] = 0.0
        d_f = torch.from_numpy(np.ascontiguousarray(1.0 / vel).ravel()).to(dev)
        d_u = torch.empty_like(d_u0)
        rounds = np.zeros(1, dtype=np.int32)
        ts = []
        for it in range(4):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            ctx.forward3d_batch(d_u, d_u0, d_f, hh, (m, n, l), 1e-6, 1, rounds=rounds, loc=A.DEVICE)
            ctx.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        kern_ms = ctx.phase_ms(0)
        K = int(rounds[0])
        out = {"size": sz, "model": model, "fwd_ms": float(np.median(ts[1:])), "fwd_kernel_ms": kern_ms, "rounds": K,
               "fwd_alg_gbs": 8.0 * N * (2 + 24 * abs(K)) / 1e6 / float(np.median(ts[1:])),
               "sha1": hashlib.sha1(d_u.cpu().numpy().tobytes()).hexdigest()[:16],
               "env": {k: v for k, v in os.environ.items() if k.startswith("ADTOMO_")}}
        if adj:
            d_g = torch.ones_like(d_u0)
            d_gs = torch.empty(N, dtype=torch.float64, device=dev)
            ta = []
            for it in range(3):
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                ctx.backward3d_batch(None, None, d_gs, d_g, d_u, d_u0, d_f, hh, (m, n, l), 1, loc=A.DEVICE)
                ctx.synchronize()
                ta.append(1e3 * (time.perf_counter() - t0))
            out["adj_ms"] = float(np.median(ta[1:]))
            out["adj_phases_ms"] = {"setup": ctx.phase_ms(2), "wavefront": ctx.phase_ms(3), "finish": ctx.phase_ms(4)}
            out["adj_sha1"] = hashlib.sha1(d_gs.cpu().numpy().tobytes()).hexdigest()[:16]
        print(json.dumps(out), flush=True)
        del d_u0, d_f, d_u
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
