"""Golden vectors computed by the REFERENCE'S OWN C++ solvers.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_ref.py
oracle/_ref/libref_eikonal.so is the reference's deps/CustomOps/Eikonal/Eikonal.h and
deps/CustomOps/Eikonal3D/Eikonal3D.cpp compiled UNMODIFIED from where they lie (oracle/Makefile) against a stub
of the absent third-party Eigen (oracle/eigen_stub: forward solvers do not use Eigen; the adjoints' SparseLU is
a dense LU there, hence adjoint cases stay below ~1500 unknowns).  Inputs come from tests/golden/ref_cases.py
(seeds + the reference's own test configurations); this script stores the OUTPUTS in tests/golden/ref_cpp.npz:
full fields for small cases, SHA-256 of the bytes (+ a strided sample) for large ones -- the forward parity bar
is bit-exactness, for which a hash is as strict as the field."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from oracle import ref          # noqa: E402
import ref_cases                # noqa: E402

FULL_LIMIT = 3000               # store the whole field up to this many nodes


def main():
    ref.build(force=True)
    out = {}
    for name, c in ref_cases.cases3d().items():
        u = ref.eikonal3d_forward(c["u0"], c["f"], c["h"], c["tol"])
        out[f"3d/{name}/u_sha256"] = np.array(ref_cases.sha(u))
        out[f"3d/{name}/u_sample"] = u.ravel()[::97].copy()
        if u.size <= FULL_LIMIT:
            out[f"3d/{name}/u"] = u
        if c["grad_u"] is not None:
            gu0, gf = ref.eikonal3d_backward(c["grad_u"], u, c["u0"], c["f"], c["h"])
            out[f"3d/{name}/grad_u0"] = gu0
            out[f"3d/{name}/grad_f"] = gf
    for name, c in ref_cases.cases2d().items():
        u = ref.eikonal2d_forward(c["f"], c["h"], c["ix"], c["jx"])
        out[f"2d/{name}/u"] = u
        out[f"2d/{name}/grad_f"] = ref.eikonal2d_backward(c["grad_u"], u, c["f"], c["h"], c["ix"], c["jx"])
    path = os.path.join(HERE, "ref_cpp.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes,", len(out), "arrays")


if __name__ == "__main__":
    main()
