"""Host-side mirror of the per-source work pattern of the reference's inversion drivers
(scripts/inversion.jl:36-105): round-robin source sharding, 8-corner source initialisation, and
the fused device evaluation of misfit + slowness gradient for a shard of sources.

Everything numerical runs in libadtomo_b200.so (adtomo_eikonal3d_misfit_grad); this module only
prepares the small host tables (source corners, receiver coordinates) and owns the sharding /
all-reduce plumbing that replaces mpi_bcast / mpi_sum (scripts/inversion.jl:44,123).
"""
import math

import numpy as np

from . import capi


def shard_sources(num_sources, rank, world_size):
    """Indices of the sources owned by `rank`: the reference's `rank+1:nproc:numsta`
    (scripts/inversion.jl:36-38), 0-based."""
    return np.arange(rank, num_sources, world_size)


def corner_sources(xyz, h, vel0):
    """8-corner source initialisation of scripts/inversion.jl:48-60.

    xyz: (S, 3) fractional 0-based node coordinates of the stations; vel0: (m, n, l) velocity used
    for the corner times.  Returns CSR-style (src_ptr int32[S+1], src_idx int32[nnz], src_val f64[nnz])
    in the Julia assignment order (ceil before floor on every axis)."""
    vel0 = np.asarray(vel0, dtype=np.float64)
    m, n, l = vel0.shape
    ptr, idx, val = [0], [], []
    for x, y, z in np.asarray(xyz, dtype=np.float64):
        xs = (math.ceil(x), math.floor(x))
        ys = (math.ceil(y), math.floor(y))
        zs = (math.ceil(z), math.floor(z))
        for cx in xs:
            for cy in ys:
                for cz in zs:
                    t = math.sqrt((x - cx) ** 2 + (y - cy) ** 2 + (z - cz) ** 2) * h / vel0[cx, cy, cz]
                    idx.append((cx * n + cy) * l + cz)
                    val.append(t)
        ptr.append(len(idx))
    return np.asarray(ptr, dtype=np.int32), np.asarray(idx, dtype=np.int32), np.asarray(val, dtype=np.float64)


class InversionProblem:
    """One rank's shard of a travel-time inversion: stations (sources), events (receivers),
    observations and weights.  `loss_and_grad(f)` evaluates the data misfit and its gradient with
    respect to the slowness field f for this shard on this rank's GPU; `allreduce` sums the packed
    [grad | misfit] buffer over ranks with ONE collective (replaces mpi_sum + the backward of
    mpi_bcast)."""

    _next_batch_id = 1

    def __init__(self, ctx, dims, h, sta_xyz, eve_xyz, uobs, qua, vel0, tol=1e-3, u0_fill=1000.0, max_rounds=0):
        self.ctx = ctx
        self.batch_id = InversionProblem._next_batch_id          # names this source set for the batch kernel's placement memo
        InversionProblem._next_batch_id += 1
        self.dims = tuple(int(d) for d in dims)
        self.N = self.dims[0] * self.dims[1] * self.dims[2]
        self.h = float(h)
        self.tol = float(tol)
        self.u0_fill = float(u0_fill)
        self.max_rounds = int(max_rounds)
        self.S = len(sta_xyz)
        self.E = len(eve_xyz)
        self.src_ptr, self.src_idx, self.src_val = corner_sources(sta_xyz, h, vel0)
        self.rcv = capi.f64(eve_xyz).reshape(self.E, 3)
        self.uobs = capi.f64(uobs).reshape(self.S, self.E)
        self.qua = capi.f64(qua).reshape(self.S, self.E)
        self.rounds = np.zeros(self.S, dtype=np.int32)

    def loss_and_grad(self, f, want_grad=True):
        """f: (m,n,l) host slowness.  Returns (misfit, grad_f (m,n,l) or None, rc)."""
        f = capi.f64(f).reshape(self.dims)
        packed = np.empty(self.N + 1, dtype=np.float64) if want_grad else None
        self.ctx.set_batch_id(self.batch_id)
        mis, rc = self.ctx.misfit_grad(packed, f, self.h, self.dims, self.tol, self.S, self.src_ptr, self.src_idx,
                                       self.src_val, self.u0_fill, self.E, self.rcv, self.uobs, self.qua,
                                       max_rounds=self.max_rounds, rounds=self.rounds, loc=capi.HOST)
        self.packed = packed
        return mis, (packed[: self.N].reshape(self.dims) if want_grad else None), rc

    @staticmethod
    def allreduce(packed):
        """Sum the packed [grad | misfit] host buffer over ranks (gloo on CPU tensors, NCCL on CUDA
        tensors).  No-op without an initialised process group."""
        import torch
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
            return packed
        t = packed if isinstance(packed, torch.Tensor) else torch.from_numpy(packed)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return packed
