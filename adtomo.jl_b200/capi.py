"""ctypes binding of include/adtomo_b200.h (the same symbols a Julia `ccall` binds)."""
import ctypes
import os
import re
import subprocess

import numpy as np

HOST, DEVICE = 0, 1
_HERE = os.path.dirname(os.path.abspath(__file__))
# ADTOMO_LIB: A/B runs of differently built libraries (benchmarks only); the default is the in-tree build
LIB_PATH = os.environ.get("ADTOMO_LIB") or os.path.join(_HERE, "libadtomo_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "adtomo_b200.h")

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)
_vp = ctypes.c_void_p
_lib = None


class AdtomoError(RuntimeError):
    pass


def build_library(force=False, extra=""):
    """nvcc -> libadtomo_b200.so (sm_100a).  Cross-compiles without a GPU."""
    srcs = [os.path.join(_HERE, "csrc", n) for n in os.listdir(os.path.join(_HERE, "csrc"))
            if n.endswith((".cu", ".cuh", ".h"))] + [HEADER_PATH]
    if not force and os.path.exists(LIB_PATH) and all(os.path.getmtime(LIB_PATH) >= os.path.getmtime(s) for s in srcs):
        return LIB_PATH
    env = dict(os.environ)
    if extra:
        env["NVCC_EXTRA"] = extra
    subprocess.check_call(["bash", os.path.join(_HERE, "csrc", "build.sh")], env=env)
    return LIB_PATH


def exported_symbols():
    """Function names declared in include/adtomo_b200.h."""
    txt = open(HEADER_PATH).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(adtomo_[a-z0-9_]+)\s*\(", txt)))


def load_library():
    """Loads the CUDA library; raises (never falls back) when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise AdtomoError(f"{LIB_PATH} not built: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(there is no CPU fallback)")
    L = ctypes.CDLL(LIB_PATH)
    c_int, c_double = ctypes.c_int, ctypes.c_double
    L.adtomo_create.restype = c_int
    L.adtomo_create.argtypes = [ctypes.POINTER(_vp), c_int]
    L.adtomo_destroy.restype = c_int
    L.adtomo_destroy.argtypes = [_vp]
    L.adtomo_last_error.restype = ctypes.c_char_p
    L.adtomo_last_error.argtypes = []
    L.adtomo_version.restype = c_int
    L.adtomo_synchronize.restype = c_int
    L.adtomo_synchronize.argtypes = [_vp]
    L.adtomo_stream.restype = ctypes.c_ulonglong
    L.adtomo_stream.argtypes = [_vp]
    L.adtomo_last_kernel_ms.restype = c_double
    L.adtomo_last_kernel_ms.argtypes = [_vp]
    L.adtomo_last_phase_ms.restype = c_double
    L.adtomo_last_phase_ms.argtypes = [_vp, c_int]
    L.adtomo_launch_count.restype = ctypes.c_longlong
    L.adtomo_launch_count.argtypes = [_vp]
    L.adtomo_phase_accumulate.restype = c_int
    L.adtomo_phase_accumulate.argtypes = [_vp, c_int]
    L.adtomo_set_batch_id.restype = c_int
    L.adtomo_set_batch_id.argtypes = [_vp, ctypes.c_longlong]
    L.adtomo_last_forward_kernel.restype = ctypes.c_char_p
    L.adtomo_last_forward_kernel.argtypes = [_vp]
    L.adtomo_selftest_sqrt.restype = c_int
    L.adtomo_selftest_sqrt.argtypes = [_vp, ctypes.c_longlong, ctypes.c_ulonglong, ctypes.POINTER(ctypes.c_longlong)]
    L.adtomo_eikonal2d_forward.restype = c_int
    L.adtomo_eikonal2d_forward.argtypes = [_vp, _vp, c_int, c_int, c_double, c_int, c_int]
    L.adtomo_eikonal2d_backward.restype = c_int
    L.adtomo_eikonal2d_backward.argtypes = [_vp, _vp, _vp, _vp, c_int, c_int, c_double, c_int, c_int]
    L.adtomo_eikonal3d_forward.restype = c_int
    L.adtomo_eikonal3d_forward.argtypes = [_vp, _vp, _vp, c_double, c_int, c_int, c_int, c_double, c_int]
    L.adtomo_eikonal3d_backward.restype = c_int
    L.adtomo_eikonal3d_backward.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, c_double, c_int, c_int, c_int]
    L.adtomo_eikonal3d_forward_batch.restype = c_int
    L.adtomo_eikonal3d_forward_batch.argtypes = [_vp, _vp, _vp, _vp, c_double, c_int, c_int, c_int, c_double, c_int,
                                                 c_int, _vp, c_int]
    L.adtomo_eikonal3d_backward_batch.restype = c_int
    L.adtomo_eikonal3d_backward_batch.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _vp, c_double, c_int, c_int,
                                                  c_int, c_int, c_int]
    L.adtomo_eikonal2d_forward_batch.restype = c_int
    L.adtomo_eikonal2d_forward_batch.argtypes = [_vp, _vp, _vp, c_int, c_int, c_double, _vp, _vp, c_int, _vp, c_int]
    L.adtomo_eikonal2d_backward_batch.restype = c_int
    L.adtomo_eikonal2d_backward_batch.argtypes = [_vp, _vp, _vp, _vp, _vp, _vp, c_int, c_int, c_double, _vp, _vp,
                                                  c_int, c_int]
    L.adtomo_eikonal3d_misfit_grad.restype = c_int
    L.adtomo_eikonal3d_misfit_grad.argtypes = [_vp, _vp, _vp, _vp, c_double, c_int, c_int, c_int, c_double, c_int,
                                               c_int, _vp, _vp, _vp, c_double, c_int, _vp, _vp, _vp, _vp, c_int]
    L.adtomo_model_begin.restype = c_int
    L.adtomo_model_begin.argtypes = [_vp, _vp, _vp, c_int, c_int, c_int, c_int]
    L.adtomo_model_add_phase.restype = c_int
    L.adtomo_model_add_phase.argtypes = [_vp, c_double, c_double, c_double, c_int, c_int, _vp, _vp, _vp, c_double, c_int,
                                         _vp, _vp, _vp, _vp, _vp, _vp, c_int, c_int]
    L.adtomo_model_finish.restype = c_int
    L.adtomo_model_finish.argtypes = [_vp, c_double, c_int, c_int, c_int, c_int, _vp, _vp, c_int]
    L.adtomo_model_loss_grad.restype = c_int
    L.adtomo_model_loss_grad.argtypes = [_vp, _vp, _vp, _vp, _vp, c_double, c_int, c_int, c_int, c_double, c_int, c_int,
                                         c_int, c_double, c_int, c_int, _vp, _vp, _vp, c_double, c_int, _vp, _vp, _vp,
                                         _vp, c_int]
    L.adtomo_nccl_unique_id.restype = c_int
    L.adtomo_nccl_unique_id.argtypes = [ctypes.c_char_p]
    L.adtomo_nccl_init.restype = c_int
    L.adtomo_nccl_init.argtypes = [_vp, ctypes.c_char_p, c_int, c_int]
    L.adtomo_nccl_allreduce_sum.restype = c_int
    L.adtomo_nccl_allreduce_sum.argtypes = [_vp, _vp, ctypes.c_longlong, c_int]
    L.adtomo_nccl_finalize.restype = c_int
    L.adtomo_nccl_finalize.argtypes = [_vp]
    _lib = L
    return L


def check(rc, what=""):
    """Negative codes raise; positive flags are returned to the caller."""
    if rc < 0:
        raise AdtomoError(f"{what}: rc={rc}: {load_library().adtomo_last_error().decode()}")
    return rc


def ptr(a):
    """void* of a numpy array (host) or a torch tensor (host or CUDA); None -> NULL.
    The library works on its own stream: a CUDA tensor that torch kernels may still be writing (or reading) is made
    safe by synchronising torch's current stream of that device first -- a host-side wait of microseconds when nothing
    is pending, and the call that follows synchronises anyway."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    if getattr(a, "is_cuda", False):
        import torch
        torch.cuda.current_stream(a.device).synchronize()
    return a.data_ptr()   # torch.Tensor


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Context:
    """adtomo_ctx: one device, one stream, one workspace."""

    def __init__(self, device=-1):
        self._lib = load_library()
        h = _vp()
        check(self._lib.adtomo_create(ctypes.byref(h), int(device)), "adtomo_create")
        self.handle = h

    def close(self):
        if getattr(self, "handle", None):
            self._lib.adtomo_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(self._lib.adtomo_synchronize(self.handle), "adtomo_synchronize")

    @property
    def stream(self):
        return int(self._lib.adtomo_stream(self.handle))

    def phase_accumulate(self, on):
        """Keep the per-kernel event pairs of every following call (phase_ms then sums over all of them)."""
        check(self._lib.adtomo_phase_accumulate(self.handle, 1 if on else 0), "adtomo_phase_accumulate")

    def set_batch_id(self, batch_id):
        """Names the source batch of the following batched calls (placement memo of the batch kernel)."""
        check(self._lib.adtomo_set_batch_id(self.handle, int(batch_id)), "adtomo_set_batch_id")

    def last_kernel(self):
        """Name of the 3D forward sweep kernel of the last call."""
        return self._lib.adtomo_last_forward_kernel(self.handle).decode()

    @property
    def last_kernel_ms(self):
        return float(self._lib.adtomo_last_kernel_ms(self.handle))

    def phase_ms(self, phase):
        """0 forward sweeps, 1 misfit, 2 adjoint setup, 3 adjoint sweeps, 4 gradient finish."""
        return float(self._lib.adtomo_last_phase_ms(self.handle, int(phase)))

    @property
    def launch_count(self):
        return int(self._lib.adtomo_launch_count(self.handle))

    def selftest_sqrt(self, n, seed=1):
        """Mismatches of the library's call-free fp64 sqrt against CUDA's sqrt on n arguments (must be 0)."""
        bad = ctypes.c_longlong(-1)
        rc = self._lib.adtomo_selftest_sqrt(self.handle, int(n), int(seed), ctypes.byref(bad))
        if rc < 0:
            raise AdtomoError(f"adtomo_selftest_sqrt: rc={rc}: {self._lib.adtomo_last_error().decode()}")
        return int(bad.value)

    # ---- NCCL (one all-reduce of the packed [grad | misfit] buffer per evaluation) -------------
    @staticmethod
    def nccl_unique_id():
        buf = ctypes.create_string_buffer(128)
        check(load_library().adtomo_nccl_unique_id(buf), "adtomo_nccl_unique_id")
        return buf.raw

    def nccl_init(self, unique_id, rank, nranks):
        check(self._lib.adtomo_nccl_init(self.handle, unique_id, int(rank), int(nranks)), "adtomo_nccl_init")

    def nccl_allreduce_sum(self, buf, count=None, loc=HOST):
        n = int(count if count is not None else (buf.size if isinstance(buf, np.ndarray) else buf.numel()))
        check(self._lib.adtomo_nccl_allreduce_sum(self.handle, ptr(buf), n, loc), "adtomo_nccl_allreduce_sum")

    def nccl_finalize(self):
        check(self._lib.adtomo_nccl_finalize(self.handle), "adtomo_nccl_finalize")

    # ---- batched 3D --------------------------------------------------------------------------
    def forward3d_batch(self, u, u0, f, h, dims, tol, S, max_rounds=0, rounds=None, loc=HOST):
        m, n, l = dims
        return check(self._lib.adtomo_eikonal3d_forward_batch(self.handle, ptr(u), ptr(u0), ptr(f), float(h), m, n, l,
                                                              float(tol), int(max_rounds), int(S), ptr(rounds), loc),
                     "adtomo_eikonal3d_forward_batch")

    def backward3d_batch(self, grad_u0, grad_f, grad_f_sum, grad_u, u, u0, f, h, dims, S, loc=HOST):
        m, n, l = dims
        return check(self._lib.adtomo_eikonal3d_backward_batch(self.handle, ptr(grad_u0), ptr(grad_f), ptr(grad_f_sum),
                                                               ptr(grad_u), ptr(u), ptr(u0), ptr(f), float(h), m, n, l,
                                                               int(S), loc), "adtomo_eikonal3d_backward_batch")

    # ---- batched 2D --------------------------------------------------------------------------
    def forward2d_batch(self, u, f, m, n, h, ix, jx, rounds=None, loc=HOST):
        ix, jx = i32(ix), i32(jx)
        return check(self._lib.adtomo_eikonal2d_forward_batch(self.handle, ptr(u), ptr(f), int(m), int(n), float(h),
                                                              ptr(ix), ptr(jx), len(ix), ptr(rounds), loc),
                     "adtomo_eikonal2d_forward_batch")

    def backward2d_batch(self, grad_f, grad_f_sum, grad_u, u, f, m, n, h, ix, jx, loc=HOST):
        ix, jx = i32(ix), i32(jx)
        return check(self._lib.adtomo_eikonal2d_backward_batch(self.handle, ptr(grad_f), ptr(grad_f_sum), ptr(grad_u),
                                                               ptr(u), ptr(f), int(m), int(n), float(h), ptr(ix),
                                                               ptr(jx), len(ix), loc),
                     "adtomo_eikonal2d_backward_batch")

    # ---- fused inversion step -------------------------------------------------------------------
    def misfit_grad(self, grad_f, f, h, dims, tol, S, src_ptr, src_idx, src_val, u0_fill, E, rcv_xyz, uobs, qua,
                    max_rounds=0, rounds=None, loc=HOST):
        m, n, l = dims
        mis = ctypes.c_double(0.0)
        rc = check(self._lib.adtomo_eikonal3d_misfit_grad(self.handle, ctypes.addressof(mis), ptr(grad_f), ptr(f),
                                                          float(h), m, n, l, float(tol), int(max_rounds), int(S),
                                                          ptr(src_ptr), ptr(src_idx), ptr(src_val), float(u0_fill),
                                                          int(E), ptr(rcv_xyz), ptr(uobs), ptr(qua), ptr(rounds), loc),
                   "adtomo_eikonal3d_misfit_grad")
        return mis.value, rc

    # ---- model parametrisation + chain rule + regulariser on the device --------------------------
    def model_begin(self, x, vel0, dims, loc=HOST):
        m, n, l = dims
        check(self._lib.adtomo_model_begin(self.handle, ptr(x), ptr(vel0), m, n, l, loc), "adtomo_model_begin")

    def model_add_phase(self, scale, h, tol, S, src_ptr, src_idx, src_val, u0_fill, E, rcv_xyz, uobs, qua, max_rounds=0,
                        rounds=None, want_grad=True, loc=HOST):
        """Returns (misfit, d misfit / d scale, rc)."""
        mis, gs = ctypes.c_double(0.0), ctypes.c_double(0.0)
        rc = check(self._lib.adtomo_model_add_phase(self.handle, float(scale), float(h), float(tol), int(max_rounds), int(S),
                                                    ptr(src_ptr), ptr(src_idx), ptr(src_val), float(u0_fill), int(E),
                                                    ptr(rcv_xyz), ptr(uobs), ptr(qua), ptr(rounds), ctypes.addressof(mis),
                                                    ctypes.addressof(gs), 1 if want_grad else 0, loc),
                   "adtomo_model_add_phase")
        return mis.value, gs.value, rc

    def model_finish(self, lam, smooth_hor, smooth_ver, add_reg, packed, loc=HOST):
        """packed: N+1 doubles ([d loss / d x | loss]) or None for the loss only.  Returns (loss, rc)."""
        loss = ctypes.c_double(0.0)
        rc = check(self._lib.adtomo_model_finish(self.handle, float(lam), int(smooth_hor), int(smooth_ver),
                                                 1 if add_reg else 0, 0 if packed is None else 1, ctypes.addressof(loss),
                                                 ptr(packed), loc), "adtomo_model_finish")
        return loss.value, rc

    def model_loss_grad(self, packed, x, vel0, lam, smooth_hor, smooth_ver, add_reg, h, dims, tol, S, src_ptr, src_idx,
                        src_val, u0_fill, E, rcv_xyz, uobs, qua, max_rounds=0, rounds=None, loc=HOST):
        """adtomo_model_loss_grad: single-phase evaluation in one call.  Returns (loss, rc)."""
        m, n, l = dims
        loss = ctypes.c_double(0.0)
        rc = check(self._lib.adtomo_model_loss_grad(self.handle, ctypes.addressof(loss), ptr(packed), ptr(x), ptr(vel0),
                                                    float(lam), int(smooth_hor), int(smooth_ver), 1 if add_reg else 0,
                                                    float(h), m, n, l, float(tol), int(max_rounds), int(S), ptr(src_ptr),
                                                    ptr(src_idx), ptr(src_val), float(u0_fill), int(E), ptr(rcv_xyz),
                                                    ptr(uobs), ptr(qua), ptr(rounds), loc), "adtomo_model_loss_grad")
        return loss.value, rc
