#!/usr/bin/env python
"""bench.py -- source solves/s (forward + adjoint, 3D grid) of the Eikonal hot path.

Workload (config.workload): BASELINE.json configs[2] "C3": one inversion step on a 128x128x64 grid
with 256 sources per GPU and 512 receivers -- forward fast-sweeping solve of every source,
receiver sampling + weighted misfit, adjoint solve, slowness gradient summed over sources.
Model: GIL7 layers + checkerboard (len 10, +-0.8 km/s), tol = 1e-3 (scripts/inversion.jl:61).
A "step" is one such evaluation over the batch.  N GPUs: weak scaling, 256 sources per GPU
(source sharding) + ONE all-reduce of the packed [gradient | misfit] buffer per step.

  python bench.py [--gpus N] [--steps K] [--warmup W]          our CUDA path
  python bench.py --config c4 --no-cpu ...                      BASELINE configs[3] (200x200x80, 2048 sources over the GPUs)
  python bench.py --impl reference ...                          CPU arm: the oracle port of the
        reference algorithm on the host cores (pinned bit for bit to the reference's own C++, oracle/_ref;
        the reference's adjoint is a sparse LU that is infeasible at this size, the port back-substitutes)

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "source solves/sec (fwd+adjoint, 3D grid)"
UNIT = "source-solves/s"
GRID = (128, 128, 64)
S_PER_GPU = 256
E_RCV = 512
C4_GRID, C4_SOURCES, C4_RCV = (200, 200, 80), 2048, 1024
TOL = 1e-3
H = 1.0


def workload(world_size, rank, s_per_gpu=S_PER_GPU, grid=GRID, e_rcv=E_RCV):
    """Synthetic inputs of SURVEY 8(d) C3; rank r owns sources r::world (scripts/inversion.jl:36-38)."""
    import adtomo_jl_b200 as A
    from adtomo_jl_b200 import synthetic as syn
    m, n, l = grid
    vel0 = syn.gil7_velocity(m, n, l, H)
    vel = syn.checkerboard(vel0, 10, 0.8)
    sta, eve = syn.stations_events(m, n, l, s_per_gpu * world_size, e_rcv, H, seed=233)
    mine = A.shard_sources(len(sta), rank, world_size)
    rng = np.random.default_rng(1000 + rank)
    sta = sta[mine]
    # observations: straight-ray times through the layered start model plus noise (bench only needs
    # non-trivial residuals; parity of the misfit itself is covered by tests/)
    d = np.linalg.norm(sta[:, None, :] - eve[None, :, :], axis=2) * H
    uobs = d / 5.5 + 0.05 * rng.standard_normal(d.shape)
    uobs[rng.random(d.shape) < 0.05] = -1.0            # missing picks (inversion.jl:100-102)
    qua = 0.5 + rng.random(d.shape)
    return dict(dims=grid, h=H, vel0=vel0, f=1.0 / vel, sta=sta, eve=eve, uobs=uobs, qua=qua)


def b_alg(N, rounds):
    """Algorithmic bytes of SURVEY 8(d): forward 8N(2+24K), adjoint 8N(6+24K), per source."""
    K = np.asarray(rounds, dtype=np.float64)
    return float((8.0 * N * (2 + 24 * K)).sum()), float((8.0 * N * (6 + 24 * K)).sum())


# ------------------------------------------------------------------------------ CPU arm
def _cpu_one(args):
    import oracle
    import ref_misfit as rm
    u0, f, h, tol, eve, uobs, qua = args
    u, rounds, _ = oracle.eikonal3d_forward(u0, f, h, tol)
    mis, gu = rm.misfit_and_grad_u(u, eve, uobs, qua)
    _, gf, _ = oracle.eikonal3d_backward(gu, u, u0, f, h)
    return rounds, mis, gf


def cpu_arm(w, n_sources, procs):
    """Times the oracle (CPU port of the reference algorithm; adjoint by back-substitution, i.e.
    FASTER than the reference's SparseLU) on `n_sources` sources with `procs` worker processes,
    rank r of P taking sources r::P like `mpirun -n P` (scripts/inversion.jl:36-38)."""
    import multiprocessing as mp
    import adtomo_jl_b200 as A
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    oracle.build()
    m, n, l = w["dims"]
    ptr, idx, val = A.corner_sources(w["sta"][:n_sources], w["h"], w["vel0"])
    jobs = []
    for s in range(n_sources):
        u0 = np.full((m, n, l), 1000.0)
        u0.ravel()[idx[ptr[s]:ptr[s + 1]]] = val[ptr[s]:ptr[s + 1]]
        jobs.append((u0, w["f"], w["h"], TOL, w["eve"], w["uobs"][s], w["qua"][s]))
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(procs) as pool:
        res = pool.map(_cpu_one, jobs, chunksize=1)
    grad = np.zeros((m, n, l))
    mis = 0.0
    for r in res:
        mis += r[1]
        grad += r[2]
    dt = time.perf_counter() - t0
    return dict(seconds=dt, rounds=[r[0] for r in res], misfit=mis, grad=grad)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(1, 0)
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    n_src = 2 * procs                  # two sources per worker process per step
    times = []
    for it in range(args.warmup + args.steps):
        r = cpu_arm(w, n_src, procs)
        if it >= args.warmup:
            times.append(r["seconds"])
    tot = float(np.sum(times))
    val = n_src * args.steps / tot
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": 0, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "C3 inversion step: 128x128x64 grid, GIL7+checkerboard(len 10, +-0.8 km/s), tol 1e-3, "
                               "512 receivers; CPU arm evaluates a bounded sample of the 256-source batch",
                   "sources_per_step": n_src},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": procs, "kind": "port",
                         "sample": f"{n_src} of the 256 sources per step x {args.steps} steps, two sources per worker "
                                   f"process; oracle port of the reference sweeps, adjoint by back-substitution "
                                   f"(faster than the reference's Eigen SparseLU)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------ GPU arm
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=10)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            c = [x.strip() for x in ln.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": float(max(mx)) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def run_ours(args):
    import torch
    import torch.distributed as dist
    import adtomo_jl_b200 as A

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the CUDA path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    if not os.path.exists(A.LIB_PATH):
        if rank == 0:
            A.build_library()
        if world > 1:
            dist.barrier()
    ctx = A.Context(local)
    if world > 1:
        # the data-path collective goes through the library's own NCCL entry point (what a Julia driver
        # would call); torch.distributed only carries the 128-byte id, the barrier and the max-over-ranks time
        uid = torch.zeros(128, dtype=torch.uint8, device=torch.device("cuda", local))
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(A.Context.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        ctx.nccl_init(bytes(uid.cpu().numpy().tobytes()), rank, world)
    if args.config == "c4":
        # BASELINE configs[3]: 200x200x80, 2048 sources in total sharded over the GPUs (strong scaling), 1024 receivers
        if C4_SOURCES % world:
            raise SystemExit(f"--config c4 needs a GPU count that divides {C4_SOURCES}")
        w = workload(world, rank, s_per_gpu=C4_SOURCES // world, grid=C4_GRID, e_rcv=C4_RCV)
    else:
        w = workload(world, rank, s_per_gpu=args.sources)
    m, n, l = w["dims"]
    N = m * n * l
    S, E = len(w["sta"]), len(w["eve"])
    ptr, idx, val = A.corner_sources(w["sta"], w["h"], w["vel0"])
    dev = torch.device("cuda", local)

    # ---- device-resident inputs (kernel-throughput measurement) ----
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    d_f = t(w["f"], torch.float64)
    d_ptr, d_idx, d_val = t(ptr, torch.int32), t(idx, torch.int32), t(val, torch.float64)
    d_rcv, d_obs, d_qua = t(w["eve"], torch.float64), t(w["uobs"], torch.float64), t(w["qua"], torch.float64)
    d_packed = torch.zeros(N + 1, dtype=torch.float64, device=dev)
    rounds = np.zeros(S, dtype=np.int32)

    def step_device():
        mis, rc = ctx.misfit_grad(d_packed, d_f, w["h"], w["dims"], TOL, S, d_ptr, d_idx, d_val, 1000.0, E, d_rcv,
                                  d_obs, d_qua, rounds=rounds, loc=A.DEVICE)
        if world > 1:
            ctx.nccl_allreduce_sum(d_packed, N + 1, loc=A.DEVICE)
        return mis, rc

    # ---- host-buffer inputs through the public API (e2e) ----
    pin = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).pin_memory()
    h_f = pin(w["f"], torch.float64)
    h_ptr, h_idx, h_val = pin(ptr, torch.int32), pin(idx, torch.int32), pin(val, torch.float64)
    h_rcv, h_obs, h_qua = pin(w["eve"], torch.float64), pin(w["uobs"], torch.float64), pin(w["qua"], torch.float64)
    h_packed = torch.zeros(N + 1, dtype=torch.float64).pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in (h_f, h_ptr, h_idx, h_val, h_rcv, h_obs, h_qua))
    d2h = h_packed.numel() * 8 + 8 + 2 * 4 * S

    def step_host():
        mis, rc = ctx.misfit_grad(h_packed, h_f, w["h"], w["dims"], TOL, S, h_ptr, h_idx, h_val, 1000.0, E, h_rcv,
                                  h_obs, h_qua, rounds=rounds, loc=A.HOST)
        if world > 1:
            ctx.nccl_allreduce_sum(h_packed, N + 1, loc=A.HOST)   # host buffer summed over ranks (staged + NCCL)
        return mis, rc

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(max(warmup, 3) + 1):     # >= 3 warm-up steps (+1: first-touch of workspaces and clocks settle)
            fn()
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lc0 = ctx.launch_count
        e0.record()
        ph = np.zeros(6)
        for _ in range(steps):
            fn()
            ph += [ctx.phase_ms(p) for p in range(6)]
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        timed.launches = ctx.launch_count - lc0
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, ph / steps

    sampler = ClockSampler(local)
    if rank == 0 and not os.environ.get("ADTOMO_BENCH_NO_SMI"):
        sampler.start()
    ms_dev, phases = timed(step_device, args.steps, args.warmup)
    launches = int(timed.launches)
    clocks = sampler.stop() if rank == 0 else None
    rounds_dev = rounds.copy()
    mis_dev = float(d_packed[N].item())
    ms_e2e, _ = timed(step_host, args.steps, args.warmup)
    mis_e2e = float(h_packed[N].item())

    total_sources = S * world
    value = total_sources * args.steps / (ms_dev * 1e-3)
    e2e_val = total_sources * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches; HBM-bound stencil) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bf, ba = b_alg(N, rounds_dev)
    fwd_ms, adj_ms = phases[0], phases[3]
    kern = {"forward_sweeps": {"ms": fwd_ms, "alg_gb": bf / 1e9, "gbs": bf / 1e6 / max(fwd_ms, 1e-9)},
            "adjoint_sweeps": {"ms": adj_ms, "alg_gb": ba / 1e9, "gbs": ba / 1e6 / max(adj_ms, 1e-9)},
            "misfit_ms": phases[1], "adjoint_setup_ms": phases[2], "finish_ms": phases[4], "layout_convert_ms": phases[5]}
    dom = "forward_sweeps" if fwd_ms >= adj_ms else "adjoint_sweeps"
    achieved = kern[dom]["gbs"]
    # DRAM bytes of that kernel from the committed ncu --set full capture of this same launch (profiles/)
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_traffic_v2.json")   # k_fwd3d_v2<512,2>, the kernel this batch runs on
    if dom == "forward_sweeps" and S == S_PER_GPU and (m, n, l) == GRID and os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_gb_per_launch")
    step_alg_gbs = (bf + ba) / 1e6 / (ms_dev / args.steps)
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write)",
                "algorithmic_gb_per_launch": (bf if dom == "forward_sweeps" else ba) / 1e9, "peak_source": peak_src,
                "whole_step_alg_gbs": step_alg_gbs, "whole_step_frac": step_alg_gbs / peak,
                "note": "achieved = algorithmic bytes 8N(2+24K) fwd / 8N(6+24K) adj summed over the batch "
                        "(K = rounds each source ran) / CUDA-event time of that kernel on its launch stream"}

    if world > 1:
        ctx.nccl_finalize()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload ----
    cpu = None
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        procs = min(cores, 32)
        n_src = min(S, 4 * procs)          # ~10-20 s of CPU work
        r = cpu_arm(w, n_src, procs)
        cpu = {"value": n_src / r["seconds"], "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"first {n_src} of the {S} sources of this workload (forward+misfit+adjoint each) dealt to "
                         f"{procs} worker processes, {r['seconds']:.1f} s; oracle port, adjoint by back-substitution (faster "
                         f"than the reference's SparseLU)",
               "rounds_match_gpu": bool(list(r["rounds"]) == list(rounds_dev[:n_src]))}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.config == "c4" else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config.upper()} inversion step: {m}x{n}x{l} grid, {S} sources/GPU x {world} GPU, {E} receivers, "
                               "GIL7+checkerboard(len 10, +-0.8 km/s) model, tol 1e-3, misfit + slowness gradient",
                   "sources_per_gpu": S, "rounds_mean": float(np.mean(rounds_dev)),
                   "rounds_hist": {str(int(k)): int(v) for k, v in zip(*np.unique(np.abs(rounds_dev), return_counts=True))},
                   "l2": "inputs larger than L2 (travel-time fields of the batch: %.1f GB)" % (S * N * 8 / 1e9),
                   "parallelism": f"source-shard x{world}" + (" + 1 NCCL all-reduce of N+1 fp64 per step (adtomo_nccl_allreduce_sum)" if world > 1 else "")},
        "roofline": roofline, "kernels": kern, "cpu_baseline": cpu,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches, "clocks": clocks,
        "misfit": mis_dev, "misfit_e2e": mis_e2e,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sources", type=int, default=S_PER_GPU, help="sources per GPU (default: the C3 batch of 256)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--config", default="c3", choices=["c3", "c4"],
                    help="c3 (default, the contract line): 128x128x64, 256 sources per GPU, weak scaling; "
                         "c4: 200x200x80, 2048 sources in total sharded over the GPUs, strong scaling")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
