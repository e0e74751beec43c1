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
_CPU = {}      # inputs of the CPU arm: set in the parent BEFORE the pool forks, inherited copy-on-write by the workers


def _cpu_one(s):
    """One source on one worker: forward, receiver sampling + misfit, adjoint; the slowness gradient is added into this
    worker's own slice of a shared-memory array (nothing but three scalars is pickled)."""
    import multiprocessing as mp
    import oracle
    import ref_misfit as rm
    c = _CPU
    w, N = c["w"], c["N"]
    m, n, l = w["dims"]
    u0 = np.full((m, n, l), 1000.0)
    u0.ravel()[c["idx"][c["ptr"][s]:c["ptr"][s + 1]]] = c["val"][c["ptr"][s]:c["ptr"][s + 1]]
    u, rounds, _ = oracle.eikonal3d_forward(u0, w["f"], w["h"], TOL)
    mis, gu = rm.misfit_and_grad_u(u, w["eve"], w["uobs"][s], w["qua"][s])
    _, gf, _ = oracle.eikonal3d_backward(gu, u, u0, w["f"], w["h"])
    k = (mp.current_process()._identity[0] - 1) % c["procs"]
    np.frombuffer(c["G"], dtype=np.float64, count=N, offset=8 * N * k)[:] += gf.ravel()
    return rounds, mis


class CpuArm:
    """The oracle (CPU port of the reference algorithm; adjoint by back-substitution, i.e. FASTER than the reference's
    SparseLU) on the first `n_sources` sources of workload `w` with `procs` worker processes, one source at a time per
    worker like `mpirun -n P` over `rank+1:nproc:numsta` (scripts/inversion.jl:36-38).  The pool is created once
    (outside any timed region); the slowness field reaches the workers by fork inheritance, gradients come back through
    shared memory."""

    def __init__(self, w, n_sources, procs):
        import multiprocessing as mp
        import adtomo_jl_b200 as A
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle
        oracle.build()
        m, n, l = w["dims"]
        self.N, self.n_sources, self.procs, self.dims = m * n * l, n_sources, procs, (m, n, l)
        ptr, idx, val = A.corner_sources(w["sta"][:n_sources], w["h"], w["vel0"])
        ctx = mp.get_context("fork")
        self.G = ctx.RawArray("d", procs * self.N)
        _CPU.update(w=w, N=self.N, ptr=ptr, idx=idx, val=val, G=self.G, procs=procs)
        self.pool = ctx.Pool(procs)
        self.pool.map(_noop, range(4 * procs))          # workers up and imported

    def step(self):
        """One evaluation of the sample: returns dict(seconds, rounds, misfit, grad)."""
        slices = np.frombuffer(self.G, dtype=np.float64).reshape(self.procs, self.N)
        t0 = time.perf_counter()
        slices[:] = 0.0
        res = self.pool.map(_cpu_one, range(self.n_sources), chunksize=1)
        grad = slices.sum(axis=0).reshape(self.dims)
        mis = float(sum(r[1] for r in res))
        dt = time.perf_counter() - t0
        return dict(seconds=dt, rounds=[r[0] for r in res], misfit=mis, grad=grad)

    def close(self):
        self.pool.close()
        self.pool.join()


def _noop(_):
    import oracle          # noqa: F401  (first import in the worker)
    import ref_misfit      # noqa: F401
    return 0


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = workload(1, 0)
    cores = os.cpu_count() or 1
    procs = min(cores, 64)
    n_src = 2 * procs                  # two sources per worker process per step
    arm = CpuArm(w, n_src, procs)
    times = []
    for it in range(args.warmup + args.steps):
        r = arm.step()
        if it >= args.warmup:
            times.append(r["seconds"])
    arm.close()
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
                                   f"process (persistent pool, inputs by fork inheritance, gradients through shared "
                                   f"memory); oracle port of the reference sweeps, adjoint by back-substitution "
                                   f"(faster than the reference's Eigen SparseLU)"},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


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


def perturbed_models(f, k=3):
    """k slowness models around f: the timed steps rotate through them (what an optimiser does between evaluations:
    the model drifts a little, so the rounds every source needs -- and with them the batch kernel's placement memo of
    the PREVIOUS evaluation -- are close but not identical).  Model 0 is f itself."""
    m, n, l = f.shape
    i, j, kk = np.meshgrid(np.arange(m), np.arange(n), np.arange(l), indexing="ij")
    out = [f]
    for q in range(1, k):
        bump = np.sin(2 * np.pi * (q * i / m + 0.37 * q)) * np.cos(2 * np.pi * (j / n) * (q + 1)) * np.cos(np.pi * kk / l)
        out.append(f * (1.0 + 0.004 * bump))
    return out


def measure(args, ctx, A, torch, dist, world, rank, dev, w, steps, warmup, tag, sampler=None, n_models=3, min_warm=3, host_leg=True):
    """Device-resident and host-buffer (e2e) timing of the fused step on workload w.  Returns a dict of raw results."""
    m, n, l = w["dims"]
    N = m * n * l
    S, E = len(w["sta"]), len(w["eve"])
    ptr, idx, val = A.corner_sources(w["sta"], w["h"], w["vel0"])
    models = perturbed_models(w["f"], n_models)

    # ---- device-resident inputs (kernel-throughput measurement) ----
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).to(dev)
    d_f = [t(f, torch.float64) for f in models]
    d_ptr, d_idx, d_val = t(ptr, torch.int32), t(idx, torch.int32), t(val, torch.float64)
    d_rcv, d_obs, d_qua = t(w["eve"], torch.float64), t(w["uobs"], torch.float64), t(w["qua"], torch.float64)
    d_packed = torch.zeros(N + 1, dtype=torch.float64, device=dev)
    rounds = np.zeros(S, dtype=np.int32)
    torch.cuda.synchronize()          # the library runs on its own stream: inputs must be complete before the first call

    def step_device(k):
        mis, rc = ctx.misfit_grad(d_packed, d_f[k % n_models], w["h"], w["dims"], TOL, S, d_ptr, d_idx, d_val, 1000.0, E,
                                  d_rcv, d_obs, d_qua, rounds=rounds, loc=A.DEVICE)
        if world > 1:
            ctx.nccl_allreduce_sum(d_packed, N + 1, loc=A.DEVICE)
        return mis

    # ---- host-buffer inputs through the public API (e2e) ----
    pin = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dt).pin_memory()
    h_f = [pin(f, torch.float64) for f in models]
    h_ptr, h_idx, h_val = pin(ptr, torch.int32), pin(idx, torch.int32), pin(val, torch.float64)
    h_rcv, h_obs, h_qua = pin(w["eve"], torch.float64), pin(w["uobs"], torch.float64), pin(w["qua"], torch.float64)
    h_packed = torch.zeros(N + 1, dtype=torch.float64).pin_memory()
    h2d = sum(x.numel() * x.element_size() for x in (h_f[0], h_ptr, h_idx, h_val, h_rcv, h_obs, h_qua))
    d2h = h_packed.numel() * 8 + 8 + 2 * 4 * S

    def step_host(k):
        mis, rc = ctx.misfit_grad(h_packed, h_f[k % n_models], w["h"], w["dims"], TOL, S, h_ptr, h_idx, h_val, 1000.0, E,
                                  h_rcv, h_obs, h_qua, rounds=rounds, loc=A.HOST)
        if world > 1:
            ctx.nccl_allreduce_sum(h_packed, N + 1, loc=A.HOST)   # host buffer summed over ranks (staged + NCCL)
        return mis

    def sync_all():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn):
        for k in range(max(warmup, min_warm) + 1):     # >= 3 warm-up steps (+1: first-touch of workspaces and clocks settle)
            fn(k)
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        lc0 = ctx.launch_count
        ctx.phase_accumulate(True)              # per-kernel event pairs are kept and read AFTER the timed region
        rr, mm = [], []
        e0.record()
        for k in range(steps):
            mm.append(fn(k))
            rr.append(rounds.copy())            # host array the call has just filled: no device synchronisation
        e1.record()
        sync_all()
        ms = e0.elapsed_time(e1)
        ph = np.array([ctx.phase_ms(p) for p in range(6)]) / steps
        ctx.phase_accumulate(False)
        launches = ctx.launch_count - lc0
        if world > 1:
            tt = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        return ms, ph, rr, mm, launches

    if sampler is not None:
        sampler.start()
    ms_dev, phases, rounds_steps, mis_dev, launches = timed(step_device)
    clocks = sampler.stop() if sampler is not None else None
    red_dev = float(d_packed[N].item())
    ms_e2e, mis_e2e = float("nan"), []
    if host_leg:
        ms_e2e, _, _, mis_e2e, _ = timed(step_host)
    out = dict(N=N, S=S, E=E, dims=(m, n, l), ms_dev=ms_dev, ms_e2e=ms_e2e, phases=phases, rounds_steps=rounds_steps,
               mis_dev=[float(x) for x in mis_dev], mis_e2e=[float(x) for x in mis_e2e], launches=int(launches), clocks=clocks,
               h2d=int(h2d), d2h=int(d2h), reduced_misfit_last=red_dev)

    # ---- N > 1: the NCCL-reduced packed buffer equals the sum of the ranks' own results (scripts/inversion.jl:44,123) ----
    if world > 1 and tag == "c3":
        own = torch.zeros(N + 1, dtype=torch.float64, device=dev)
        mis, _ = ctx.misfit_grad(own, d_f[0], w["h"], w["dims"], TOL, S, d_ptr, d_idx, d_val, 1000.0, E, d_rcv, d_obs,
                                 d_qua, rounds=rounds, loc=A.DEVICE)
        red = own.clone()
        ctx.synchronize()
        torch.cuda.synchronize()
        ctx.nccl_allreduce_sum(red, N + 1, loc=A.DEVICE)
        ctx.synchronize()
        gathered = [torch.zeros_like(own) for _ in range(world)]
        dist.all_gather(gathered, own)                     # torch.distributed as the independent second path
        ref = torch.stack(gathered).sum(dim=0)
        scale = float(ref[:N].abs().max().item())
        out["allreduce_check"] = {
            "ok": bool(abs(float(red[N] - ref[N])) <= 1e-12 * abs(float(ref[N])) and
                       float((red[:N] - ref[:N]).abs().max()) <= 1e-12 * scale),
            "misfit_rel": abs(float(red[N] - ref[N])) / abs(float(ref[N])),
            "grad_rel": float((red[:N] - ref[:N]).abs().max()) / scale,
            "what": "adtomo_nccl_allreduce_sum of the packed [grad | misfit] buffers vs the sum of the per-rank buffers "
                    "all-gathered with torch.distributed"}
    return out


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
    dev = torch.device("cuda", local)
    c4_main = args.config == "c4"
    if c4_main:
        # BASELINE configs[3]: 200x200x80, 2048 sources in total sharded over the GPUs (strong scaling), 1024 receivers
        if C4_SOURCES % world:
            raise SystemExit(f"--config c4 needs a GPU count that divides {C4_SOURCES}")
        w = workload(world, rank, s_per_gpu=C4_SOURCES // world, grid=C4_GRID, e_rcv=C4_RCV)
    else:
        w = workload(world, rank, s_per_gpu=args.sources)
    sampler = ClockSampler(local) if rank == 0 and not os.environ.get("ADTOMO_BENCH_NO_SMI") else None
    ctx.set_batch_id(1)
    R = measure(args, ctx, A, torch, dist, world, rank, dev, w, args.steps, args.warmup, "c4" if c4_main else "c3", sampler)
    m, n, l = R["dims"]
    N, S, E = R["N"], R["S"], R["E"]
    ms_dev, ms_e2e, phases = R["ms_dev"], R["ms_e2e"], R["phases"]

    total_sources = S * world
    value = total_sources * args.steps / (ms_dev * 1e-3)
    e2e_val = total_sources * args.steps / (ms_e2e * 1e-3)

    # ---- roofline of the dominant kernel (this rank's launches; HBM-bound stencil) ----
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bfs, bas = zip(*[b_alg(N, np.abs(r)) for r in R["rounds_steps"]])
    bf, ba = float(np.mean(bfs)), float(np.mean(bas))           # per step (launch)
    fwd_ms, adj_ms = phases[0], phases[3]
    # the adjoint is ONE O(N) wavefront pass, not the reference's K-round structure: its own compulsory bytes are the
    # fields it must touch once per node and source (u, grad_u / x record, diagonal, code) = 8N * 4 + 2N per source;
    # the measured DRAM traffic of the kernel is in profiles/r02_ncu_summary_adj*.json
    adj_comp = float(S * N * (8 * 4 + 2))
    kern = {"forward_sweeps": {"ms": fwd_ms, "alg_gb": bf / 1e9, "gbs": bf / 1e6 / max(fwd_ms, 1e-9)},
            "adjoint_sweeps": {"ms": adj_ms, "compulsory_gb": adj_comp / 1e9, "compulsory_gbs": adj_comp / 1e6 / max(adj_ms, 1e-9),
                               "survey_formula_gb": ba / 1e9,
                               "note": "single O(N) wavefront pass over the ancestors of the receiver nodes; compulsory = (8*4+2) B "
                                       "per node and source (dense upper bound); survey_formula_gb = 8N(6+24K), the reference's "
                                       "K-round streaming model, listed for comparison only (not used for any fraction)"},
            "misfit_ms": phases[1], "adjoint_setup_ms": phases[2], "finish_ms": phases[4], "layout_convert_ms": phases[5]}
    dom = "forward_sweeps"
    achieved = kern[dom]["gbs"]
    # DRAM bytes of that kernel from the committed ncu --set full capture of this same launch (profiles/)
    traffic = None
    for name in ("r02_traffic_v3.json", "r01_traffic_v2.json"):
        tpath = os.path.join(ROOT, "profiles", name)
        if S == S_PER_GPU and (m, n, l) == GRID and os.path.exists(tpath):
            traffic = json.load(open(tpath)).get("dram_gb_per_launch")
            break
    step_ms = ms_dev / args.steps
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_unit": "GB per launch (ncu dram__bytes_read+write)",
                "algorithmic_gb_per_launch": bf / 1e9, "peak_source": peak_src,
                "kernel_share_of_step": fwd_ms / step_ms,
                "whole_step": {"forward_formula_plus_adjoint_compulsory_gbs": (bf + adj_comp) / 1e6 / step_ms,
                               "frac": (bf + adj_comp) / 1e6 / step_ms / peak,
                               "survey_formula_gbs": (bf + ba) / 1e6 / step_ms,
                               "survey_formula_frac": (bf + ba) / 1e6 / step_ms / peak,
                               "note": "frac credits the adjoint its compulsory bytes only; survey_formula_* credits it the "
                                       "reference's K-round model 8N(6+24K), which the O(N) adjoint does not move"},
                "note": "achieved = algorithmic bytes 8N(2+24K) of the forward solves summed over the batch (K = rounds each "
                        "source ran in that step, mean over the timed steps) / mean CUDA-event time of the forward kernel on "
                        "its launch stream"}

    # ---- BASELINE configs[3] (C4) sub-measurement: 200x200x80, 2048 sources in total over the GPUs, strong scaling ----
    c4 = None
    if not c4_main and not args.no_c4 and C4_SOURCES % world == 0:
        w4 = workload(world, rank, s_per_gpu=C4_SOURCES // world, grid=C4_GRID, e_rcv=C4_RCV)
        ctx.set_batch_id(2)
        R4 = measure(args, ctx, A, torch, dist, world, rank, dev, w4, args.c4_steps, 1, "c4", None, n_models=2, min_warm=1,
                     host_leg=False)
        r4 = np.abs(np.concatenate(R4["rounds_steps"]))
        c4 = {"workload": "C4: 200x200x80 grid, 2048 sources in total over %d GPU(s), 1024 receivers, tol 1e-3 (BASELINE configs[3])" % world,
              "scaling": "strong", "value": C4_SOURCES * args.c4_steps / (R4["ms_dev"] * 1e-3), "unit": UNIT,
              "ms_per_step": R4["ms_dev"] / args.c4_steps, "steps": args.c4_steps, "warmup": 2, "sources_per_gpu": R4["S"],
              "forward_ms": float(R4["phases"][0]), "adjoint_ms": float(R4["phases"][3]), "rounds_mean": float(r4.mean()),
              "rounds_hist": {str(int(k)): int(v) for k, v in zip(*np.unique(r4, return_counts=True))},
              "misfit": R4["mis_dev"][0]}

    # ---- BASELINE configs[1] (C2, one 64^3 source) and configs[4] (C5, one 256^3 / 512^3 source): single-source solves on
    # the whole GPU (team kernels), forward + adjoint, device time of the library's kernels; N = 1 only (replicas otherwise)
    single = None
    if world == 1 and not c4_main and not args.no_single:
        from adtomo_jl_b200 import synthetic as syn
        single = {}
        for name, sz in (("c2_64cubed", 64), ("c5_256cubed", 256), ("c5_512cubed", 512)):
            hh = 1.0 if sz == 64 else 25.0 / sz
            vel = syn.gil7_velocity(sz, sz, sz, hh)
            u0 = np.full((sz, sz, sz), 1000.0)
            u0[sz // 2, sz // 2, 2 if sz == 64 else 0] = 0.0
            Ns = sz ** 3
            d_u0 = torch.from_numpy(u0.reshape(1, -1)).to(dev)
            d_fs = torch.from_numpy(np.ascontiguousarray(1.0 / vel).ravel()).to(dev)
            del u0, vel
            d_u, d_g = torch.empty_like(d_u0), torch.ones_like(d_u0)
            d_gs = torch.empty(Ns, dtype=torch.float64, device=dev)
            rs = np.zeros(1, dtype=np.int32)
            torch.cuda.synchronize()
            best = None
            for it in range(4):                  # first call: workspaces; best of the next three
                ctx.forward3d_batch(d_u, d_u0, d_fs, hh, (sz, sz, sz), 1e-6, 1, rounds=rs, loc=A.DEVICE)
                ms_f = ctx.phase_ms(0) + ctx.phase_ms(5)
                ctx.backward3d_batch(None, None, d_gs, d_g, d_u, d_u0, d_fs, hh, (sz, sz, sz), 1, loc=A.DEVICE)
                ms_b = ctx.phase_ms(2) + ctx.phase_ms(3) + ctx.phase_ms(4)
                if it > 0 and (best is None or ms_f + ms_b < best[0] + best[1]):
                    best = (ms_f, ms_b)
            K = int(abs(rs[0]))
            fwd_alg = 8.0 * Ns * (2 + 24 * K)
            single[name] = {"grid": f"{sz}^3", "forward_ms": best[0], "adjoint_ms": best[1], "solves_per_s": 1e3 / (best[0] + best[1]),
                            "rounds": K, "forward_alg_gbs": fwd_alg / 1e6 / best[0], "forward_frac_of_hbm_peak": fwd_alg / 1e6 / best[0] / peak,
                            "kernel": ctx.last_kernel() if hasattr(ctx, "last_kernel") else None}
            del d_u0, d_fs, d_u, d_g, d_gs
            torch.cuda.empty_cache()
        single["parity"] = ("tests/test_gpu_parity.py: 64^3 and 256^3 bit-exact (oracle / team == cluster kernels); 512^3 by properties only "
                            "(idempotence, residual, sign): the oracle needs minutes there")

    if world > 1:
        ctx.nccl_finalize()
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- CPU baseline (rank 0, N = 1 only): bounded sample of the same workload, and parity of the GPU on that sample ----
    cpu, parity = None, None
    rounds0 = np.abs(R["rounds_steps"][0])              # model 0
    if world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        procs = min(cores, 32)
        n_src = min(S, 4 * procs)          # ~10-20 s of CPU work
        arm = CpuArm(w, n_src, procs)
        r = arm.step()
        arm.close()
        cpu = {"value": n_src / r["seconds"], "unit": UNIT, "cores": procs, "kind": "port",
               "sample": f"first {n_src} of the {S} sources of this workload (forward+misfit+adjoint each) dealt to "
                         f"{procs} worker processes, {r['seconds']:.1f} s; oracle port, adjoint by back-substitution (faster "
                         f"than the reference's SparseLU)",
               "rounds_match_gpu": bool(list(r["rounds"]) == list(rounds0[:n_src]))}
        # the GPU on exactly that sample, through the host-buffer C-ABI call
        ws = dict(w)
        ws["sta"], ws["uobs"], ws["qua"] = w["sta"][:n_src], w["uobs"][:n_src], w["qua"][:n_src]
        ptr, idx, val = A.corner_sources(ws["sta"], w["h"], w["vel0"])
        packed = np.zeros(N + 1)
        rs = np.zeros(n_src, dtype=np.int32)
        mis, _ = ctx.misfit_grad(packed, w["f"], w["h"], w["dims"], TOL, n_src, ptr, idx, val, 1000.0, E, w["eve"],
                                 ws["uobs"], ws["qua"], rounds=rs, loc=A.HOST)
        gmax = float(np.abs(r["grad"]).max())
        parity = {"sources": n_src, "misfit_rel": abs(mis - r["misfit"]) / abs(r["misfit"]),
                  "grad_rel": float(np.abs(packed[:N].reshape(m, n, l) - r["grad"]).max()) / gmax,
                  "rounds_equal": bool(list(rs) == list(r["rounds"])),
                  "bars": "misfit 1e-12, gradient 1e-10 of max|grad| (tests/test_gpu_fullsize.py holds the same bars on the "
                          "batch kernel at C3 and C4 size; 512^3 is checked by properties only)"}

    rounds_all = np.abs(np.concatenate(R["rounds_steps"]))
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_dev / args.steps, "higher_is_better": True,
        "scaling": "strong" if c4_main else "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"{args.config.upper()} inversion step: {m}x{n}x{l} grid, {S} sources/GPU x {world} GPU, {E} receivers, "
                               "GIL7+checkerboard(len 10, +-0.8 km/s) model, tol 1e-3, misfit + slowness gradient",
                   "sources_per_gpu": S, "rounds_mean": float(np.mean(rounds_all)),
                   "rounds_hist": {str(int(k)): int(v) for k, v in zip(*np.unique(rounds_all, return_counts=True))},
                   "models": "the timed steps rotate through 3 slowness models (the checkerboard model and two +-0.4 % smooth "
                             "perturbations of it): the batch kernel's memo of rounds / SM placement is one evaluation old, as in "
                             "an optimiser loop",
                   "l2": "inputs larger than L2 (travel-time fields of the batch: %.1f GB)" % (S * N * 8 / 1e9),
                   "parallelism": f"source-shard x{world}" + (" + 1 NCCL all-reduce of N+1 fp64 per step (adtomo_nccl_allreduce_sum)" if world > 1 else "")},
        "roofline": roofline, "kernels": kern, "cpu_baseline": cpu, "parity_sample": parity,
        "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": R["h2d"], "d2h_bytes_per_step": R["d2h"],
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": R["launches"], "clocks": R["clocks"],
        "misfit": R["mis_dev"][0], "misfit_e2e": R["mis_e2e"][0],
        "misfit_models": R["mis_dev"][: min(3, len(R["mis_dev"]))],
        "c4": c4, "single_source": single,
    }
    if "allreduce_check" in R:
        line["allreduce_check"] = R["allreduce_check"]
    emit(line)
    if world > 1:
        dist.destroy_process_group()


_OUT_FD = None


def emit(line):
    """The ONE JSON line of the contract, on the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _OUT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_OUT_FD, data)


def main():
    global _OUT_FD
    # libraries (NCCL prints its version banner from C) must not add lines to stdout: everything but the JSON line goes
    # to stderr
    sys.stdout.flush()
    _OUT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sources", type=int, default=S_PER_GPU, help="sources per GPU (default: the C3 batch of 256)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-c4", action="store_true", help="skip the C4 (200x200x80, 2048 sources) strong-scaling sub-measurement")
    ap.add_argument("--c4-steps", type=int, default=2, help="timed steps of the C4 sub-measurement")
    ap.add_argument("--no-single", action="store_true", help="skip the single-source (C2 64^3, C5 256^3 / 512^3) sub-measurements")
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
