#!/usr/bin/env python
"""Forward + adjoint time of mid-size batches on the C3 grid (128x128x64, checkerboard model, tol 1e-3) under the kernel
selected by the environment (ADTOMO_TEAM, ADTOMO_FORCE_V2, ADTOMO_ADJ_TEAM): data for the dispatch thresholds.
  python benchmarks/batch_probe.py 16 32 64 128"""
import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import adtomo_jl_b200 as A
    from adtomo_jl_b200 import synthetic as syn
    m, n, l, h = 128, 128, 64, 1.0
    vel0 = syn.gil7_velocity(m, n, l, h)
    f = 1.0 / syn.checkerboard(vel0, 10, 0.8)
    ctx = A.Context(0)
    dev = torch.device("cuda", 0)
    N = m * n * l
    d_f = torch.from_numpy(np.ascontiguousarray(f).ravel()).to(dev)
    for S in [int(a) for a in sys.argv[1:]]:
        sta, _ = syn.stations_events(m, n, l, S, 1)
        ptr, idx, val = A.corner_sources(sta, h, vel0)
        u0 = np.full((S, N), 1000.0)
        for s in range(S):
            u0[s, idx[ptr[s]:ptr[s + 1]]] = val[ptr[s]:ptr[s + 1]]
        d_u0 = torch.from_numpy(u0).to(dev)
        d_u = torch.empty_like(d_u0)
        d_g = torch.ones_like(d_u0)
        d_gs = torch.empty(N, dtype=torch.float64, device=dev)
        rounds = np.zeros(S, dtype=np.int32)
        fw, ad = [], []
        for it in range(3):
            ctx.forward3d_batch(d_u, d_u0, d_f, h, (m, n, l), 1e-3, S, rounds=rounds, loc=A.DEVICE)
            fw.append(ctx.phase_ms(0))
            ctx.backward3d_batch(None, None, d_gs, d_g, d_u, d_u0, d_f, h, (m, n, l), S, loc=A.DEVICE)
            ad.append(ctx.phase_ms(2) + ctx.phase_ms(3) + ctx.phase_ms(4))
        print(json.dumps({"S": S, "fwd_ms": round(min(fw[1:]), 2), "adj_ms": round(min(ad[1:]), 2),
                          "rounds_mean": float(rounds.mean()),
                          "sha1": hashlib.sha1(d_u.cpu().numpy().tobytes() + d_gs.cpu().numpy().tobytes()).hexdigest()[:12],
                          "env": {k: v for k, v in os.environ.items() if k.startswith("ADTOMO_")}}), flush=True)
        del d_u0, d_u, d_g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
