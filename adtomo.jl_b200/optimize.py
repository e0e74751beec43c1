"""Host-side mirror of src/mpi_optimize.jl (closure form, :3-72) and of the model parametrisation /
regulariser of the inversion drivers (scripts/inversion.jl:42-43,107-121).

Not part of the hot path: the optimiser stays on the host (as Optim.jl does in the reference) and
calls the device evaluation `InversionProblem.loss_and_grad`.  What replaces the MPI machinery:
  - no worker spin-loop / flag broadcast (mpi_optimize.jl:11-33,52-69): every rank runs the same
    deterministic optimiser on the all-reduced loss/gradient, so all ranks stay in lock-step;
  - `mpi_bcast` of the model + `mpi_sum` of loss and gradient -> one all-reduce of [grad | misfit].
Logging mirrors the reference ("iter k, current loss=", "===== STEP k =====", :15-25); checkpoints
are written every `steps` GRADIENT evaluations (:22-29) by rank 0 as `iter_<k>.h5`, HDF5 dataset "data" holding
the raw optimiser vector like the reference's (hdf5_min.py: the package's own minimal HDF5 writer/reader, h5py is
not available here).  `VelocityModel` evaluates the parametrisation on the host (numpy); `DeviceVelocityModel` does
the same on the device through adtomo_model_* and also covers the joint P+S driver (inversion_joint.jl).
"""
import os

import numpy as np


def _backtracking(f, x, p, fx, gtp, alpha0=1.0, c1=1e-4, rho_hi=0.5, rho_lo=0.1, max_iter=1000):
    """LineSearches.BackTracking (order 3) with InitialStatic(alpha = 1): the reference's choice
    (mpi_optimize.jl:35-39).  Returns (alpha, f(x + alpha p))."""
    a1, a2 = alpha0, alpha0
    phi0, dphi0 = fx, gtp
    phi1 = phi2 = f(x + a2 * p)
    it = 0
    while not np.isfinite(phi2) and it < max_iter:     # shrink until finite
        it += 1
        a1 = a2
        a2 = a1 / 2
        phi1 = phi2 = f(x + a2 * p)
    k = 0
    while phi2 > phi0 + c1 * a2 * dphi0:
        k += 1
        if k > max_iter:
            raise RuntimeError("line search failed")
        if k == 1 or a1 == a2:
            atmp = -(dphi0 * a2 ** 2) / (2 * (phi2 - phi0 - dphi0 * a2))        # quadratic fit
        else:
            div = 1.0 / (a1 ** 2 * a2 ** 2 * (a2 - a1))
            a = (a1 ** 2 * (phi2 - phi0 - dphi0 * a2) - a2 ** 2 * (phi1 - phi0 - dphi0 * a1)) * div
            b = (-a1 ** 3 * (phi2 - phi0 - dphi0 * a2) + a2 ** 3 * (phi1 - phi0 - dphi0 * a1)) * div
            if abs(a) < 1e-300:
                atmp = dphi0 / (2 * b)
            else:
                disc = max(b * b - 3 * a * dphi0, 0.0)
                atmp = (-b + np.sqrt(disc)) / (3 * a)                          # cubic fit
        a1 = a2
        atmp = min(atmp, a2 * rho_hi)
        a2 = max(atmp, a2 * rho_lo)
        phi1, phi2 = phi2, f(x + a2 * p)
    return a2, phi2


def _rank():
    """Rank of this process in the torch.distributed group (0 without one): mpi_rank() of the reference."""
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank()
    except Exception:
        pass
    return int(os.environ.get("RANK", "0"))


def save_checkpoint(loc, k, x):
    """iter_<k> checkpoint of the raw optimiser vector (mpi_optimize.jl:26-28 writes it as HDF5 dataset "data").
    Written to a temporary name and renamed, so a reader never sees a partial file.  The format is HDF5
    (`iter_<k>.h5`, dataset "data", what scripts/post_rect.jl:31-43 reads) through the package's own minimal
    writer (hdf5_min.py; h5py is not available here)."""
    from . import hdf5_min
    os.makedirs(loc, exist_ok=True)
    path = os.path.join(loc, f"iter_{k}.h5")
    tmp = path + f".tmp{os.getpid()}"
    hdf5_min.write_dataset(tmp, "data", np.asarray(x, dtype=np.float64))
    os.replace(tmp, path)
    return path


def gpu_optimize(_f, _g, x0, method="LBFGS", iterations=1000, loc=None, steps=10, verbose=True, m=10, g_tol=1e-8,
                 rank=None):
    """Closure form of mpi_optimize (src/mpi_optimize.jl:3-72): _f(x) -> loss, _g(x) -> gradient (same shape as x).
    Every rank runs the same deterministic optimiser on the all-reduced loss/gradient; like the reference only
    rank 0 logs and writes checkpoints (:15-25).  method: "LBFGS" (memory m) or "BFGS" (dense inverse Hessian as
    in Optim.BFGS -- small problems only), both with InitialStatic + BackTracking (:35-45).
    Returns (x_min, history of accepted losses)."""
    if method not in ("LBFGS", "BFGS"):
        raise ValueError(f"Method {method} not implemented.")       # mpi_optimize.jl:46-48
    rank = _rank() if rank is None else int(rank)
    cnt = {"f": 0, "g": 0}

    def f(x):
        L = float(_f(x))
        cnt["f"] += 1
        if verbose and rank == 0:
            print(f"iter {cnt['f']}, current loss=", L)
        return L

    def g(x):
        cnt["g"] += 1
        if rank == 0:
            if verbose:
                print(f"================== STEP {cnt['g']} ==================")
            if loc is not None and cnt["g"] % steps == 0:
                save_checkpoint(loc, cnt["g"], x)
        return np.asarray(_g(x), dtype=np.float64)

    x = np.array(x0, dtype=np.float64).ravel().copy()
    shape = np.shape(x0)
    wrap = lambda v: v.reshape(shape)
    if rank == 0 and verbose:
        print("[ranks = %d] Optimization starts..." % int(os.environ.get("WORLD_SIZE", "1")))     # mpi_optimize.jl:53
    fx = f(wrap(x))
    gx = g(wrap(x)).ravel()
    hist = [fx]
    if method == "BFGS":
        if x.size > 20000:
            raise ValueError("BFGS keeps a dense %d x %d inverse Hessian; use LBFGS for models of this size" % (x.size, x.size))
        Hinv = np.eye(x.size)
        for _ in range(iterations):
            if np.abs(gx).max() <= g_tol:
                break
            p = -Hinv.dot(gx)
            gtp = gx.dot(p)
            if gtp >= 0:                 # lost positive definiteness: restart from the identity
                Hinv = np.eye(x.size)
                p = -gx
                gtp = gx.dot(p)
            alpha, fnew = _backtracking(lambda z: f(wrap(z)), x, p, fx, gtp)
            xn = x + alpha * p
            gn = g(wrap(xn)).ravel()
            sk, yk = xn - x, gn - gx
            ys = yk.dot(sk)
            if ys > 1e-300:              # standard inverse BFGS update
                Hy = Hinv.dot(yk)
                Hinv += ((ys + yk.dot(Hy)) / ys ** 2) * np.outer(sk, sk) - (np.outer(Hy, sk) + np.outer(sk, Hy)) / ys
            x, gx, fx = xn, gn, fnew
            hist.append(fx)
        return wrap(x), hist
    S, Y = [], []
    for _ in range(iterations):
        if np.abs(gx).max() <= g_tol:
            break
        q = gx.copy()
        al = []
        for s, y in zip(reversed(S), reversed(Y)):
            a = s.dot(q) / y.dot(s)
            al.append(a)
            q -= a * y
        if S:
            q *= S[-1].dot(Y[-1]) / Y[-1].dot(Y[-1])
        for (s, y), a in zip(zip(S, Y), reversed(al)):
            b = y.dot(q) / y.dot(s)
            q += (a - b) * s
        p = -q
        gtp = gx.dot(p)
        if gtp >= 0:                     # not a descent direction: restart with steepest descent
            S, Y = [], []
            p = -gx
            gtp = gx.dot(p)
        alpha, fnew = _backtracking(lambda z: f(wrap(z)), x, p, fx, gtp)
        xn = x + alpha * p
        gn = g(wrap(xn)).ravel()
        s, y = xn - x, gn - gx
        if y.dot(s) > 1e-300:
            S.append(s)
            Y.append(y)
            if len(S) > m:
                S.pop(0)
                Y.pop(0)
        x, gx, fx = xn, gn, fnew
        hist.append(fx)
    return wrap(x), hist


def box_filter_periodic(a, sh, sv):
    """conv3d(VALID) of the periodically padded field with ones(sh,sh,sv)/(sh*sh*sv): scripts/inversion.jl:107-118."""
    out = np.zeros_like(a)
    h2, v2 = (sh - 1) // 2, (sv - 1) // 2
    for di in range(-h2, h2 + 1):
        ai = np.roll(a, di, axis=0)
        for dj in range(-h2, h2 + 1):
            aj = np.roll(ai, dj, axis=1)
            for dk in range(-v2, v2 + 1):
                out += np.roll(aj, dk, axis=2)
    return out / (sh * sh * sv)


def gpu_optimize_model(model, method="LBFGS", iterations=1000, loc=None, steps=10, verbose=True, rank=None, x0=None):
    """Session form of mpi_optimize (src/mpi_optimize.jl:76-143): there the trainable variables of a TensorFlow graph
    are flattened into one vector, optimised through the closure form, assigned back, and their values returned by
    rank 0 (`nothing` on the other ranks).  Here the graph is a model object (`VelocityModel` / `DeviceVelocityModel`:
    `.loss(x)`, `.grad(x)`, and for the device model `.x0()` / `.n_vars`): returns on rank 0 the list of variable
    arrays -- [var_change (m, n, l)] and, for the joint driver with free scales, [var_change, scales] -- else None."""
    rank = _rank() if rank is None else int(rank)
    if x0 is None:
        x0 = model.x0() if hasattr(model, "x0") else np.zeros(model.vel0.size)
    x, hist = gpu_optimize(model.loss, model.grad, np.asarray(x0, dtype=np.float64).ravel(), method=method, iterations=iterations,
                           loc=loc, steps=steps, verbose=verbose, rank=rank)
    if rank != 0:
        return None
    N = model.vel0.size
    out = [np.array(x[:N]).reshape(model.vel0.shape)]
    if x.size > N:
        out.append(np.array(x[N:]))
    return out


class VelocityModel:
    """x -> fvar = 2*sigmoid(x) - 1 + vel0 -> slowness 1/fvar (inversion.jl:42-43,61) and the L1
    box-filter regulariser lambda * sum|fvar - smooth(fvar)| (:107-121).  `problems`: this rank's
    InversionProblem(s).  The regulariser is added ONCE (the reference adds it on every rank before
    mpi_sum, i.e. nproc times -- SURVEY 5)."""

    def __init__(self, vel0, problems, lam=0.0, smooth_hor=5, smooth_ver=3):
        self.vel0 = np.asarray(vel0, dtype=np.float64)
        self.problems = list(problems)
        self.lam, self.sh, self.sv = float(lam), int(smooth_hor), int(smooth_ver)
        self._cache = None

    def _eval(self, x, want_grad):
        from .inversion import InversionProblem
        x = np.asarray(x, dtype=np.float64).reshape(self.vel0.shape)
        key = (x.tobytes(), want_grad)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        sig = 1.0 / (1.0 + np.exp(-x))
        fvar = 2.0 * sig - 1.0 + self.vel0
        f = 1.0 / fvar
        N = f.size
        packed = np.zeros(N + 1)
        for p in self.problems:
            mis, gf, rc = p.loss_and_grad(f, want_grad=want_grad)
            self._note_status(rc)
            if want_grad:
                packed += p.packed
            else:
                packed[N] += mis
        InversionProblem.allreduce(packed)
        loss = packed[N]
        grad_x = None
        nvel = box_filter_periodic(fvar, self.sh, self.sv) if self.lam else None
        if self.lam:
            loss += self.lam * np.abs(fvar - nvel).sum()
        if want_grad:
            g_fvar = -packed[:N].reshape(f.shape) * f * f              # d(1/fvar) = -1/fvar^2
            if self.lam:
                s = np.sign(fvar - nvel)
                g_fvar += self.lam * (s - box_filter_periodic(s, self.sh, self.sv))   # symmetric periodic kernel
            grad_x = g_fvar * 2.0 * sig * (1.0 - sig)
        self._cache = (key, (loss, grad_x))
        return loss, grad_x

    status = 0        # worst positive status seen (1: a forward solve hit the round cap, 2: adjoint flagged)

    def _note_status(self, rc):
        """Positive status flags of the device evaluation are kept (self.status) and reported once each."""
        if rc > 0 and rc > self.status:
            self.status = rc
            if _rank() == 0:
                print("[adtomo] warning: device evaluation returned status %d (%s)" %
                      (rc, "forward solve hit the round cap" if rc == 1 else "adjoint flagged"))

    def loss(self, x):
        return self._eval(x, False)[0]

    def grad(self, x):
        return self._eval(x, True)[1]

    def velocity(self, x):
        x = np.asarray(x, dtype=np.float64).reshape(self.vel0.shape)
        return 2.0 / (1.0 + np.exp(-x)) - 1.0 + self.vel0


class DeviceVelocityModel:
    """VelocityModel with the parametrisation, the chain rule and the regulariser ON THE DEVICE
    (adtomo_model_begin / add_phase / finish): an evaluation ships the N optimiser variables in and N+1 doubles out;
    the slowness fields and their gradients never cross PCIe.

    phases: list of (InversionProblem, scale) -- the single-phase driver (scripts/inversion.jl) is [(P, 1.0)], the
    joint driver (inversion_joint.jl:49-51,80) [(P, 1.0), (S, pvs)] with `pvs` an optimiser variable: pass
    optimise_scales=True and the optimiser vector becomes [x (N) | scale of every phase with a non-unit start].
    add_reg: this rank adds the regulariser (exactly one rank of a multi-GPU run must)."""

    def __init__(self, ctx, vel0, phases, lam=0.0, smooth_hor=5, smooth_ver=3, add_reg=True, optimise_scales=False):
        self.ctx = ctx
        self.vel0 = np.ascontiguousarray(vel0, dtype=np.float64)
        self.dims = self.vel0.shape
        self.N = self.vel0.size
        self.phases = [(p, float(s)) for p, s in phases]
        self.lam, self.sh, self.sv = float(lam), int(smooth_hor), int(smooth_ver)
        self.add_reg = bool(add_reg)
        self.free = [i for i, (_, s) in enumerate(self.phases) if s != 1.0] if optimise_scales else []
        self.status = 0
        self._cache = None

    @property
    def n_vars(self):
        return self.N + len(self.free)

    def x0(self):
        """Start vector: var_change = 0 (inversion.jl:42) and the start scales (inversion_joint.jl:50)."""
        return np.concatenate([np.zeros(self.N), [self.phases[i][1] for i in self.free]])

    def _eval(self, z, want_grad):
        from . import capi
        from .inversion import InversionProblem
        z = np.ascontiguousarray(z, dtype=np.float64).ravel()
        key = (z.tobytes(), want_grad)
        if self._cache is not None and self._cache[0] == key:
            return self._cache[1]
        x = z[: self.N]
        scales = [s for _, s in self.phases]
        for j, i in enumerate(self.free):
            scales[i] = float(z[self.N + j])
        ctx = self.ctx
        ctx.model_begin(x, self.vel0, self.dims)
        gs = np.zeros(len(self.phases))
        for i, (p, _) in enumerate(self.phases):
            ctx.set_batch_id(p.batch_id)
            _, gs[i], rc = ctx.model_add_phase(scales[i], p.h, p.tol, p.S, p.src_ptr, p.src_idx, p.src_val, p.u0_fill, p.E,
                                               p.rcv, p.uobs, p.qua, max_rounds=p.max_rounds, rounds=p.rounds,
                                               want_grad=want_grad)
            VelocityModel._note_status(self, rc)
        # [d loss / d x | loss | d loss / d scale of the free phases]: ONE all-reduce sums all of it
        packed = np.zeros(self.N + 1 + len(self.free))
        if want_grad:
            loss, _ = ctx.model_finish(self.lam, self.sh, self.sv, self.add_reg, packed[: self.N + 1])
            packed[self.N + 1:] = [gs[i] for i in self.free]
        else:
            loss, _ = ctx.model_finish(self.lam, self.sh, self.sv, self.add_reg, None)
            packed[self.N] = loss
        InversionProblem.allreduce(packed)
        loss = packed[self.N]
        grad = np.concatenate([packed[: self.N], packed[self.N + 1:]]) if want_grad else None
        self._cache = (key, (loss, grad))
        return loss, grad

    def loss(self, z):
        return self._eval(z, False)[0]

    def grad(self, z):
        return self._eval(z, True)[1]

    def velocity(self, z):
        x = np.asarray(z, dtype=np.float64).ravel()[: self.N].reshape(self.dims)
        return 2.0 / (1.0 + np.exp(-x)) - 1.0 + self.vel0
