"""Generate golden vectors from the REFERENCE'S OWN Python statements of the sweeps.

Run in the build container only (needs /root/reference):  python tests/golden/make_golden.py
It extracts the function definitions (not the module-level demo code, which needs matplotlib)
from
    /root/reference/tests/Eikonal3D/prototype.py      (3D sweeps, same order as Eikonal3D.cpp:59-68)
    /root/reference/tests/Eikonal3D/prototype2d.py    (2D sweeps, same order as Eikonal.h:73-77)
executes them unmodified on small seeded inputs and stores inputs + outputs in
tests/golden/*.npz.  The prototypes square with `**2` and form f**2*h**2, so they agree with the
C++ (and with oracle/) to rounding, not bit for bit; tests compare at rtol 1e-11.
"""
import ast
import io
import contextlib
import os
import sys

import numpy as np

REF = "/root/reference/tests/Eikonal3D"
OUT = os.path.dirname(os.path.abspath(__file__))


def load_functions(path):
    src = open(path).read()
    tree = ast.parse(src)
    tree.body = [n for n in tree.body if isinstance(n, ast.FunctionDef)]
    ns = {"np": np}
    exec(compile(tree, path, "exec"), ns)
    return ns


def main():
    rng = np.random.default_rng(233)
    p3 = load_functions(os.path.join(REF, "prototype.py"))
    p2 = load_functions(os.path.join(REF, "prototype2d.py"))
    sink = io.StringIO()

    # 3D case A: the prototype's own configuration (21^3, f = 1, h = 0.01, centre source),
    # which is also deps/CustomOps/Eikonal3D/gradtest.jl:14-23 and the old main() in tests/Eikonal3D/Eikonal3D.cpp.
    m = n = l = 21
    f = np.ones((m, n, l))
    h = 0.01
    u0 = 1000 * np.ones((m, n, l))
    u0[m // 2, n // 2, l // 2] = 0.0
    with contextlib.redirect_stdout(sink):
        u = p3["eikonal_solve"](u0.copy(), f, h)
    np.savez_compressed(os.path.join(OUT, "proto3d_21.npz"), u0=u0, f=f, h=h, u=u)

    # 3D case B: ragged dims, random slowness, two source nodes with different start times.
    m, n, l = 9, 7, 6
    f = 0.5 + rng.random((m, n, l))
    h = 0.25
    u0 = 1000 * np.ones((m, n, l))
    u0[2, 4, 1] = 0.0
    u0[3, 4, 1] = 0.11
    with contextlib.redirect_stdout(sink):
        u = p3["eikonal_solve"](u0.copy(), f, h)
    np.savez_compressed(os.path.join(OUT, "proto3d_ragged.npz"), u0=u0, f=f, h=h, u=u)

    # 3D case C: ONE single (+,+,+) sweep and one (-,+,-) sweep from a non-trivial state
    # (pins loop order / mirror boundaries, not just the fixed point).
    m, n, l = 6, 5, 7
    f = 0.5 + rng.random((m, n, l))
    h = 0.5
    u_in = 3.0 * rng.random((m, n, l))
    I = list(range(m)); J = list(range(n)); K = list(range(l))
    s1 = p3["sweeping_over_I_J_K"](u_in.copy(), I, J, K, f, h)
    s7 = p3["sweeping_over_I_J_K"](u_in.copy(), I[::-1], J, K[::-1], f, h)
    np.savez_compressed(os.path.join(OUT, "proto3d_sweeps.npz"), u_in=u_in, f=f, h=h, s1=s1, s7=s7)

    # 2D: prototype2d on a 17 x 12 node grid, u[i, j] with i the OUTER loop (= x of Eikonal.h).
    mi, nj = 17, 12
    f2 = 0.5 + rng.random((mi, nj))
    h2 = 0.1
    u2 = 1000 * np.ones((mi, nj))
    u2[5, 3] = 0.0
    with contextlib.redirect_stdout(sink):
        u2s = p2["eikonal_solve"](u2.copy(), f2, h2)
    np.savez_compressed(os.path.join(OUT, "proto2d.npz"), f=f2, h=h2, src=np.array([5, 3]), u=u2s)
    print("wrote goldens to", OUT)


if __name__ == "__main__":
    main()
