"""Host-side mirror of src/mpi_optimize.jl (closure form, :3-72) and of the model parametrisation /
regulariser of the inversion drivers (scripts/inversion.jl:42-43,107-121).

Not part of the hot path: the optimiser stays on the host (as Optim.jl does in the reference) and
calls the device evaluation `InversionProblem.loss_and_grad`.  What replaces the MPI machinery:
  - no worker spin-loop / flag broadcast (mpi_optimize.jl:11-33,52-69): every rank runs the same
    deterministic optimiser on the all-reduced loss/gradient, so all ranks stay in lock-step;
  - `mpi_bcast` of the model + `mpi_sum` of loss and gradient -> one all-reduce of [grad | misfit].
Logging mirrors the reference ("iter k, current loss=", "===== STEP k =====", :15-25); checkpoints
are written every `steps` GRADIENT evaluations (:22-29) as `iter_<k>.npy` holding the raw optimiser
vector (the reference writes the same vector as HDF5 dataset "data"; h5py is not available here).
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


def gpu_optimize(_f, _g, x0, method="LBFGS", iterations=1000, loc=None, steps=10, verbose=True, m=10, g_tol=1e-8):
    """Closure form of mpi_optimize: _f(x) -> loss, _g(x) -> gradient (same shape as x).
    Returns (x_min, history of accepted losses)."""
    if method != "LBFGS":
        raise ValueError(f"Method {method} not implemented.")       # the reference also offers BFGS
    cnt = {"f": 0, "g": 0}

    def f(x):
        L = float(_f(x))
        cnt["f"] += 1
        if verbose:
            print(f"iter {cnt['f']}, current loss=", L)
        return L

    def g(x):
        cnt["g"] += 1
        if verbose:
            print(f"================== STEP {cnt['g']} ==================")
        if loc is not None and cnt["g"] % steps == 0:
            os.makedirs(loc, exist_ok=True)
            np.save(os.path.join(loc, f"iter_{cnt['g']}.npy"), np.asarray(x))
        return np.asarray(_g(x), dtype=np.float64)

    x = np.array(x0, dtype=np.float64).ravel().copy()
    shape = np.shape(x0)
    wrap = lambda v: v.reshape(shape)
    fx = f(wrap(x))
    gx = g(wrap(x)).ravel()
    hist = [fx]
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
            mis, gf, _ = p.loss_and_grad(f, want_grad=want_grad)
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

    def loss(self, x):
        return self._eval(x, False)[0]

    def grad(self, x):
        return self._eval(x, True)[1]

    def velocity(self, x):
        x = np.asarray(x, dtype=np.float64).reshape(self.vel0.shape)
        return 2.0 / (1.0 + np.exp(-x)) - 1.0 + self.vel0
