#!/usr/bin/env python
"""Secondary measurements on the other BASELINE.json configurations (bench.py is the contract line).

  python benchmarks/configs.py [--out profiles/r01_configs.json]

C1: 2D 30x40 model, 40 sources, forward + adjoint           (tests/2D_test.jl shape)
C2: 3D 64^3 layered, 1 source and tests/test3d.jl 51^3       (latency-bound single source)
C4: 3D 200x200x80, S sources on this GPU (one rank's shard of the 2048), fused misfit + gradient
C5: 3D 256^3 / 384^3 / 512^3 single source forward + adjoint (team kernels: the whole GPU on one source)
Each entry: milliseconds (CUDA events of the library's stream), source-solves/s, rounds, and the
algorithmic GB/s of SURVEY 8(d).  Inputs are device resident (loc=DEVICE) unless noted."""
import argparse
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
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--c4-sources", type=int, default=256)
    ap.add_argument("--skip", default="")
    args = ap.parse_args()
    ctx = A.Context(0)
    dev = torch.device("cuda", 0)
    res = {}

    def timeit(fn, reps=3, warm=1):
        """Median wall time of `reps` synchronised calls (single-source cases are milliseconds long and the first
        calls after an allocation are 2x slower, so a mean of 3 was noise)."""
        for _ in range(warm):
            fn()
        ctx.synchronize()
        ts = []
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            ctx.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        return float(np.median(ts))

    # ---- C1
    if "c1" not in args.skip:
        f = syn.model_2d_test()
        rng = np.random.default_rng(233)
        S = 40
        ix = rng.integers(0, 40, S).astype(np.int32)
        jx = rng.integers(0, 30, S).astype(np.int32)
        U = np.empty((S, 30, 40))
        G = rng.standard_normal(U.shape)
        gs = np.empty((30, 40))
        rounds = np.zeros(S, dtype=np.int32)

        def c1():
            ctx.forward2d_batch(U, f, 39, 29, 1.0, ix, jx, rounds=rounds)
            ctx.backward2d_batch(None, gs, G, U, f, 39, 29, 1.0, ix, jx)
        ms = timeit(c1, reps=20, warm=3)
        res["C1_2d_30x40_S40_host_buffers"] = {"ms": ms, "solves_per_s": S / ms * 1e3, "rounds_mean": float(rounds.mean())}

    def single_source(name, u0, f, h, tol, reps=9):
        m, n, l = u0.shape
        N = u0.size
        d_u0 = torch.from_numpy(u0.reshape(1, -1).copy()).to(dev)
        d_f = torch.from_numpy(np.ascontiguousarray(f).ravel()).to(dev)
        d_u = torch.empty_like(d_u0)
        d_g = torch.ones_like(d_u0)
        d_gs = torch.empty(N, dtype=torch.float64, device=dev)
        rounds = np.zeros(1, dtype=np.int32)
        fw = lambda: ctx.forward3d_batch(d_u, d_u0, d_f, h, (m, n, l), tol, 1, rounds=rounds, loc=A.DEVICE)
        bw = lambda: ctx.backward3d_batch(None, None, d_gs, d_g, d_u, d_u0, d_f, h, (m, n, l), 1, loc=A.DEVICE)
        # device time of the call = the library's CUDA events around its kernels (conversion + sweeps; setup + wavefront +
        # finish); the wall-clock median is kept beside it (for multi-GB single sources it is noisy: allocator, host sync)
        wall_f = timeit(fw, reps=reps, warm=2)
        ms_f = min(ctx.phase_ms(0) + ctx.phase_ms(5) for _ in [fw() for _ in range(3)])
        wall_b = timeit(bw, reps=reps, warm=2)
        ms_b = min(ctx.phase_ms(2) + ctx.phase_ms(3) + ctx.phase_ms(4) for _ in [bw() for _ in range(3)])
        K = int(rounds[0])
        alg = 8.0 * N * (8 + 48 * K)
        res[name] = {"forward_ms": ms_f, "adjoint_ms": ms_b, "forward_wall_ms": wall_f, "adjoint_wall_ms": wall_b,
                     "solves_per_s": 1e3 / (ms_f + ms_b), "rounds": K, "alg_gbs": alg / 1e6 / (ms_f + ms_b)}

    # ---- C2
    if "c2" not in args.skip:
        u0, f, h = syn.model_test3d()
        single_source("C2_test3d_jl_51cubed", u0, f, h, 1e-6)
        m = n = l = 64
        vel = syn.gil7_velocity(m, n, l, 1.0)
        u0 = np.full((m, n, l), 1000.0)
        u0[32, 32, 2] = 0.0
        single_source("C2_64cubed_GIL7", u0, 1.0 / vel, 1.0, 1e-6)

    # ---- C5
    if "c5" not in args.skip:
        for sz in (256, 384, 512):
            m = n = l = sz
            hh = 25.0 / l
            vel = syn.gil7_velocity(m, n, l, hh)
            u0 = np.full((m, n, l), 1000.0)
            u0[m // 2, n // 2, 0] = 0.0
            single_source(f"C5_{sz}cubed_GIL7", u0, 1.0 / vel, hh, 1e-6, reps=3)
            del vel, u0
            torch.cuda.empty_cache()

    # ---- C4: one rank's shard of the joint inversion
    if "c4" not in args.skip:
        m, n, l, h = 200, 200, 80, 1.0
        S, E = args.c4_sources, 1024
        vel0 = syn.gil7_velocity(m, n, l, h)
        f = 1.0 / syn.checkerboard(vel0, 10, 0.8)
        sta, eve = syn.stations_events(m, n, l, S, E, h)
        rng = np.random.default_rng(5)
        d = np.linalg.norm(sta[:, None, :] - eve[None, :, :], axis=2)
        uobs = d / 5.5 + 0.05 * rng.standard_normal(d.shape)
        qua = np.ones_like(uobs)
        ptr, idx, val = A.corner_sources(sta, h, vel0)
        t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
        N = m * n * l
        d_f, d_ptr, d_idx, d_val = t(f, torch.float64), t(ptr, torch.int32), t(idx, torch.int32), t(val, torch.float64)
        d_rcv, d_obs, d_qua = t(eve, torch.float64), t(uobs, torch.float64), t(qua, torch.float64)
        d_packed = torch.zeros(N + 1, dtype=torch.float64, device=dev)
        rounds = np.zeros(S, dtype=np.int32)
        step = lambda: ctx.misfit_grad(d_packed, d_f, h, (m, n, l), 1e-3, S, d_ptr, d_idx, d_val, 1000.0, E, d_rcv,
                                       d_obs, d_qua, rounds=rounds, loc=A.DEVICE)
        ms = timeit(step, reps=1, warm=1)
        alg = float((8.0 * N * (8 + 48 * rounds.astype(np.float64))).sum())
        res[f"C4_200x200x80_S{S}_fused_step"] = {"ms": ms, "solves_per_s": S / ms * 1e3, "rounds_mean": float(rounds.mean()),
                                                 "alg_gbs": alg / 1e6 / ms,
                                                 "phases_ms": [ctx.phase_ms(p) for p in range(6)]}
    print(json.dumps(res, indent=1))
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
