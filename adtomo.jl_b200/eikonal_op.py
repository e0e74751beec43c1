"""Host-side mirror of the reference's operator interface, src/eikonal_op.jl:3-32.

`eikonal(f, srcx, srcy, h)` and `eikonal3d(u0, f, h, m, n, l, tol, verbose)` keep the reference's
argument meaning and return shapes; given torch tensors they are differentiable (the custom
gradient calls the library's adjoint, as ADCME's load_op_and_grad wires eikonal_grad /
eikonal_three_d_grad).  numpy in -> numpy out (no gradient).  All arithmetic happens in the CUDA
library through the C ABI; nothing here computes a solve on the CPU.
"""
import numpy as np

from . import capi


def _lib():
    return capi.load_library()


# ---- plain array interface (1:1 with the C ABI, host arrays) --------------------------------
def eikonal_forward(f, srcx, srcy, h):
    """f: (n_, m_) slowness [row, col]; srcx = column, srcy = row, both 1-BASED like the reference
    (src/eikonal_op.jl:3-21, shift at Eikonal.cpp:125).  Returns u of shape (n_, m_)."""
    f = capi.f64(f)
    n_, m_ = f.shape
    u = np.empty_like(f)
    rc = capi.check(_lib().adtomo_eikonal2d_forward(capi.ptr(u), capi.ptr(f), m_ - 1, n_ - 1, float(h),
                                                    int(srcx) - 1, int(srcy) - 1), "adtomo_eikonal2d_forward")
    return u, rc


def eikonal_backward(grad_u, u, f, srcx, srcy, h):
    f, u, grad_u = capi.f64(f), capi.f64(u), capi.f64(grad_u)
    n_, m_ = f.shape
    gf = np.empty_like(f)
    rc = capi.check(_lib().adtomo_eikonal2d_backward(capi.ptr(gf), capi.ptr(grad_u), capi.ptr(u), capi.ptr(f),
                                                     m_ - 1, n_ - 1, float(h), int(srcx) - 1, int(srcy) - 1),
                    "adtomo_eikonal2d_backward")
    return gf, rc


def eikonal3d_forward(u0, f, h, m, n, l, tol, verbose=False):
    u0 = capi.f64(u0).reshape(m, n, l)
    f = capi.f64(f).reshape(m, n, l)
    u = np.empty_like(u0)
    rc = capi.check(_lib().adtomo_eikonal3d_forward(capi.ptr(u), capi.ptr(u0), capi.ptr(f), float(h), int(m), int(n),
                                                    int(l), float(tol), int(bool(verbose))), "adtomo_eikonal3d_forward")
    return u, rc


def eikonal3d_backward(grad_u, u, u0, f, h, m, n, l):
    grad_u, u, u0, f = (capi.f64(a).reshape(m, n, l) for a in (grad_u, u, u0, f))
    gu0 = np.empty_like(u)
    gf = np.empty_like(u)
    rc = capi.check(_lib().adtomo_eikonal3d_backward(capi.ptr(gu0), capi.ptr(gf), capi.ptr(grad_u), capi.ptr(u),
                                                     capi.ptr(u0), capi.ptr(f), float(h), int(m), int(n), int(l)),
                    "adtomo_eikonal3d_backward")
    return gu0, gf, rc


# ---- differentiable interface -------------------------------------------------------------------
def _torch():
    import torch
    return torch


def eikonal(f, srcx, srcy, h):
    """Reference signature (src/eikonal_op.jl:3).  torch tensor in -> differentiable tensor out."""
    if isinstance(f, np.ndarray):
        return eikonal_forward(f, srcx, srcy, h)[0]
    torch = _torch()

    class _Eikonal(torch.autograd.Function):
        @staticmethod
        def forward(ctx, f_):
            fn = f_.detach().cpu().numpy()
            u, _ = eikonal_forward(fn, srcx, srcy, h)
            ctx.save_for_backward(f_)
            ctx.u = u
            return torch.from_numpy(u).to(f_.device)

        @staticmethod
        def backward(ctx, gu):
            (f_,) = ctx.saved_tensors
            gf, _ = eikonal_backward(gu.detach().cpu().numpy(), ctx.u, f_.detach().cpu().numpy(), srcx, srcy, h)
            return torch.from_numpy(gf).to(f_.device)

    return _Eikonal.apply(f)


def eikonal3d(u0, f, h, m, n, l, tol, verbose):
    """Reference signature (src/eikonal_op.jl:24).  Gradients flow to u0 and f
    (EikonalThreeDGrad writes grad_u0 and grad_f only, EikonalThreeD.cpp:245-248)."""
    if isinstance(f, np.ndarray) and isinstance(u0, np.ndarray):
        return eikonal3d_forward(u0, f, h, m, n, l, tol, verbose)[0]
    torch = _torch()
    u0 = torch.as_tensor(u0, dtype=torch.float64)
    f = torch.as_tensor(f, dtype=torch.float64)

    class _Eikonal3D(torch.autograd.Function):
        @staticmethod
        def forward(ctx, u0_, f_):
            u, _ = eikonal3d_forward(u0_.detach().cpu().numpy(), f_.detach().cpu().numpy(), h, m, n, l, tol, verbose)
            ctx.save_for_backward(u0_, f_)
            ctx.u = u
            return torch.from_numpy(u).to(f_.device)

        @staticmethod
        def backward(ctx, gu):
            u0_, f_ = ctx.saved_tensors
            gu0, gf, _ = eikonal3d_backward(gu.detach().cpu().numpy(), ctx.u, u0_.detach().cpu().numpy(),
                                            f_.detach().cpu().numpy(), h, m, n, l)
            return (torch.from_numpy(gu0).to(u0_.device).reshape(u0_.shape),
                    torch.from_numpy(gf).to(f_.device).reshape(f_.shape))

    return _Eikonal3D.apply(u0, f)
