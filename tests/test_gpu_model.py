"""-m gpu: the on-device model parametrisation / chain rule / regulariser (adtomo_model_*) against the
host mirror `VelocityModel` (numpy restatement of scripts/inversion.jl:42-43,61,107-121) and against finite differences;
the joint P+S form (inversion_joint.jl:49-51,80,140-166) against the same host pieces composed by hand."""
import numpy as np
import pytest

from test_gpu_parity import _inversion_case
import ref_misfit

pytestmark = pytest.mark.gpu


def _problem(lib, oracle, ctx, m, n, l, S, E, seed, scale=1.0):
    h, vel0, ftrue, f0, sta, eve = _inversion_case(lib, m, n, l, S, E, seed=seed)
    ptr, idx, val = lib.corner_sources(sta, h, vel0 / scale)
    uobs = np.zeros((S, E))
    for s in range(S):
        u0 = np.full((m, n, l), 1000.0)
        u0.ravel()[idx[ptr[s]:ptr[s + 1]]] = val[ptr[s]:ptr[s + 1]]
        ut, _, _ = oracle.eikonal3d_forward(u0, ftrue * scale, h, 1e-9)
        uobs[s] = [ref_misfit.sample(ut, p) for p in eve]
    uobs[0, 3] = -1.0                                   # a missing pick (inversion.jl:89)
    prob = lib.InversionProblem(ctx, (m, n, l), h, sta, eve, uobs, np.ones((S, E)), vel0 / scale, tol=1e-9)
    return prob, vel0


def test_device_model_matches_host_model(lib, oracle, ctx):
    m, n, l, S, E = 20, 18, 12, 5, 40
    prob, vel0 = _problem(lib, oracle, ctx, m, n, l, S, E, seed=3)
    host = lib.VelocityModel(vel0, [prob], lam=1e-3, smooth_hor=3, smooth_ver=3)
    dev = lib.DeviceVelocityModel(ctx, vel0, [(prob, 1.0)], lam=1e-3, smooth_hor=3, smooth_ver=3)
    rng = np.random.default_rng(1)
    for x in (np.zeros((m, n, l)), 0.3 * rng.standard_normal((m, n, l))):
        Lh, gh = host.loss(x), host.grad(x)
        Ld, gd = dev.loss(x.ravel()), dev.grad(x.ravel())
        assert abs(Ld - Lh) <= 1e-12 * abs(Lh)
        assert np.abs(gd - gh.ravel()).max() <= 1e-12 * np.abs(gh).max()
        # loss-only evaluation (want_grad = 0) gives the same loss
        dev._cache = None
        assert abs(dev.loss(x.ravel()) - Lh) <= 1e-12 * abs(Lh)
    # one-call form
    packed = np.zeros(m * n * l + 1)
    x = 0.1 * rng.standard_normal((m, n, l))
    loss, rc = ctx.model_loss_grad(packed, x, vel0, 1e-3, 3, 3, True, prob.h, (m, n, l), prob.tol, prob.S, prob.src_ptr,
                                   prob.src_idx, prob.src_val, prob.u0_fill, prob.E, prob.rcv, prob.uobs, prob.qua)
    assert rc == 0 and packed[-1] == loss
    assert abs(loss - host.loss(x)) <= 1e-12 * abs(loss)
    assert np.abs(packed[:-1] - host.grad(x).ravel()).max() <= 1e-12 * np.abs(packed[:-1]).max()
    # no regulariser on this rank: the plain data term
    dev0 = lib.DeviceVelocityModel(ctx, vel0, [(prob, 1.0)], lam=1e-3, smooth_hor=3, smooth_ver=3, add_reg=False)
    host0 = lib.VelocityModel(vel0, [prob], lam=0.0)
    assert abs(dev0.loss(x.ravel()) - host0.loss(x)) <= 1e-12 * abs(host0.loss(x))


def test_device_model_joint_p_and_s(lib, oracle, ctx):
    """Two phases sharing one model: f_P = 1 / fvar, f_S = pvs / fvar with pvs an optimiser variable."""
    m, n, l, S, E = 16, 14, 10, 4, 30
    pvs = 1.73
    probP, vel0 = _problem(lib, oracle, ctx, m, n, l, S, E, seed=5)
    probS, _ = _problem(lib, oracle, ctx, m, n, l, S, E, seed=5, scale=pvs)
    dev = lib.DeviceVelocityModel(ctx, vel0, [(probP, 1.0), (probS, pvs)], lam=2e-3, smooth_hor=3, smooth_ver=3,
                                  optimise_scales=True)
    N = m * n * l
    assert dev.n_vars == N + 1
    rng = np.random.default_rng(2)
    z = dev.x0()
    z[:N] = 0.2 * rng.standard_normal(N)
    z[N] = 1.6
    L, g = dev.loss(z), dev.grad(z)
    # host composition of the same evaluation
    x = z[:N].reshape(m, n, l)
    sig = 1.0 / (1.0 + np.exp(-x))
    fvar = 2.0 * sig - 1.0 + vel0
    mp, gp, _ = probP.loss_and_grad(1.0 / fvar)
    ms, gs, _ = probS.loss_and_grad(z[N] / fvar)
    gp, gs = np.array(gp).reshape(fvar.shape), np.array(gs).reshape(fvar.shape)
    box = lib.box_filter_periodic
    d = fvar - box(fvar, 3, 3)
    Lh = mp + ms + 2e-3 * np.abs(d).sum()
    sgn = np.sign(d)
    g_fvar = -gp / fvar ** 2 - gs * z[N] / fvar ** 2 + 2e-3 * (sgn - box(sgn, 3, 3))
    gh = np.concatenate([(g_fvar * 2.0 * sig * (1.0 - sig)).ravel(), [(gs / fvar).sum()]])
    assert abs(L - Lh) <= 1e-12 * abs(Lh)
    assert np.abs(g - gh).max() <= 1e-11 * np.abs(gh).max()
    # finite differences along a random direction of all N + 1 variables (tol 1e-9 solves: the loss is smooth enough)
    v = rng.standard_normal(N + 1)
    eps = 1e-6
    fd = (dev.loss(z + eps * v) - dev.loss(z - eps * v)) / (2 * eps)
    assert abs(fd - g.dot(v)) <= 5e-4 * abs(fd)


def test_device_model_argument_errors(lib, ctx):
    with pytest.raises(lib.AdtomoError):
        ctx.model_finish(0.0, 3, 3, False, None)                   # nothing open
    vel0 = np.ones((4, 4, 4))
    ctx.model_begin(np.zeros((4, 4, 4)), vel0, (4, 4, 4))
    with pytest.raises(lib.AdtomoError):
        ctx.model_finish(1e-3, 4, 3, True, None)                   # even window
    loss, rc = ctx.model_finish(0.0, 3, 3, False, None)
    assert loss == 0.0 and rc == 0


def test_device_tables_are_validated(lib, ctx):
    """Index / coordinate tables that live on the device are checked there (the host cannot see them): an index outside the
    grid or a receiver outside it is an argument error, not an out-of-bounds store."""
    import torch
    m, n, l, S, E = 8, 8, 8, 2, 3
    dev = torch.device("cuda", 0)
    N = m * n * l
    f = torch.ones(N, dtype=torch.float64, device=dev)
    ptr = torch.tensor([0, 1, 2], dtype=torch.int32, device=dev)
    val = torch.zeros(2, dtype=torch.float64, device=dev)
    rcv = torch.tensor([[1.0, 2.0, 3.0], [4.5, 4.5, 4.5], [7.0, 7.0, 7.0]], dtype=torch.float64, device=dev)
    obs = torch.ones(S * E, dtype=torch.float64, device=dev)
    qua = torch.ones(S * E, dtype=torch.float64, device=dev)
    packed = torch.zeros(N + 1, dtype=torch.float64, device=dev)
    good = torch.tensor([10, 100], dtype=torch.int32, device=dev)
    mis, rc = ctx.misfit_grad(packed, f, 1.0, (m, n, l), 1e-6, S, ptr, good, val, 1000.0, E, rcv, obs, qua, loc=lib.DEVICE)
    assert rc == 0 and np.isfinite(mis)
    bad_idx = torch.tensor([10, N + 5], dtype=torch.int32, device=dev)
    with pytest.raises(lib.AdtomoError, match="src_idx"):
        ctx.misfit_grad(packed, f, 1.0, (m, n, l), 1e-6, S, ptr, bad_idx, val, 1000.0, E, rcv, obs, qua, loc=lib.DEVICE)
    bad_rcv = rcv.clone()
    bad_rcv[1, 2] = 7.5
    with pytest.raises(lib.AdtomoError, match="receiver"):
        ctx.misfit_grad(packed, f, 1.0, (m, n, l), 1e-6, S, ptr, good, val, 1000.0, E, bad_rcv, obs, qua, loc=lib.DEVICE)
    # the context is still usable
    mis2, rc = ctx.misfit_grad(packed, f, 1.0, (m, n, l), 1e-6, S, ptr, good, val, 1000.0, E, rcv, obs, qua, loc=lib.DEVICE)
    assert rc == 0 and abs(mis2 - mis) <= 1e-12 * abs(mis)
