"""GPU (-m gpu), needs >= 2 devices: source sharding + the library's own NCCL all-reduce
(adtomo_nccl_*), i.e. what replaces `mpirun -n P` + mpi_bcast/mpi_sum of the reference drivers.
Skipped on a single-GPU box; `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs it."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_WORKER = r'''
import os, sys, time, numpy as np
root, rank, world, idfile, outfile = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
sys.path.insert(0, root)
import adtomo_jl_b200 as A
from adtomo_jl_b200 import synthetic as syn
ctx = A.Context(rank)
if rank == 0:
    uid = A.Context.nccl_unique_id()
    open(idfile + ".tmp", "wb").write(uid); os.replace(idfile + ".tmp", idfile)
else:
    while not os.path.exists(idfile): time.sleep(0.05)
    uid = open(idfile, "rb").read()
ctx.nccl_init(uid, rank, world)
m, n, l, S, E, h = 24, 20, 14, 7, 9, 1.0
vel0 = syn.gil7_velocity(m, n, l, h)
f = 1.0 / syn.checkerboard(vel0, 5, 0.8)
sta, eve = syn.stations_events(m, n, l, S, E, h)
rng = np.random.default_rng(3)
uobs = 1.0 + rng.random((S, E)); qua = 0.5 + rng.random((S, E))
mine = A.shard_sources(S, rank, world)                      # rank+1:nproc:numsta
prob = A.InversionProblem(ctx, (m, n, l), h, sta[mine], eve, uobs[mine], qua[mine], vel0, tol=1e-3)
prob.loss_and_grad(f)
packed = prob.packed.copy()
ctx.nccl_allreduce_sum(packed)                               # ONE collective: [grad | misfit]
if rank == 0:
    full = A.InversionProblem(ctx, (m, n, l), h, sta, eve, uobs, qua, vel0, tol=1e-3)
    full.loss_and_grad(f)
    np.savez(outfile, reduced=packed, full=full.packed)
ctx.nccl_finalize(); ctx.close()
print("worker", rank, "ok")
'''


def test_sharded_misfit_grad_nccl_allreduce(tmp_path):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    script = tmp_path / "w.py"
    script.write_text(_WORKER)
    idfile, outfile = str(tmp_path / "nccl.id"), str(tmp_path / "out.npz")
    ps = [subprocess.Popen([sys.executable, str(script), ROOT, str(r), "2", idfile, outfile],
                           stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=300)[0] for p in ps]
    for p, o in zip(ps, outs):
        assert p.returncode == 0, o
    d = np.load(outfile)
    red, full = d["reduced"], d["full"]
    # fp64 sums in a different order: compare at 1e-12, misfit included
    assert np.abs(red[:-1] - full[:-1]).max() <= 1e-12 * np.abs(full[:-1]).max()
    assert abs(red[-1] - full[-1]) <= 1e-12 * abs(full[-1])
