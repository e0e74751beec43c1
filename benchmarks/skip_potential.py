#!/usr/bin/env python
"""How much of the batch forward solve could be skipped bit-exactly?  (CPU measurement, see skip_potential.c)

Result on the bench workload (C3, tol 1e-3), written to profiles/r02_skip_potential.json: only 5-9 % of the warp-slot
evaluations are skippable, because ~65 % of ALL node evaluations still lower the node's value (by amounts far below
the tolerance) -- fast sweeping on the checkerboard model keeps rippling tiny corrections through every box until
the L-inf test stops the rounds."""
import ctypes
import json
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)


def main():
    import bench
    import adtomo_jl_b200 as A
    so = "/tmp/libskip_potential.so"
    subprocess.check_call(["gcc", "-O3", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(HERE, "skip_potential.c"), "-lm"])
    L = ctypes.CDLL(so)
    dp = ctypes.POINTER(ctypes.c_double)
    w = bench.workload(1, 0)
    m, n, l = w["dims"]
    L.set_dims(m, n, l, ctypes.c_double(w["h"]))
    ptr, idx, val = A.corner_sources(w["sta"], w["h"], w["vel0"])
    f = np.ascontiguousarray(w["f"])

    def run(s, mode, box):
        u0 = np.full((m, n, l), 1000.0)
        u0.ravel()[idx[ptr[s]:ptr[s + 1]]] = val[ptr[s]:ptr[s + 1]]
        u = np.empty_like(u0)
        r = ctypes.c_int(0)
        L.reset_counts()
        L.skip_forward(u.ctypes.data_as(dp), u0.ctypes.data_as(dp), f.ctypes.data_as(dp), ctypes.c_double(bench.TOL), 20,
                       *box, mode, ctypes.byref(r))
        c = (ctypes.c_longlong * 3)()
        L.get_counts(c)
        return u, r.value, list(c)

    out = {"workload": "bench C3: 128x128x64, GIL7+checkerboard, tol 1e-3; warp slot = 4 A x 8 C pencils at one level",
           "rule": "evaluate a slot iff a box it touches, or a face neighbour of one, changed in the previous or current sweep",
           "sources": []}
    for s in [0, 5, 17, 100, 200, 255]:
        u_ref, r_ref, c_ref = run(s, 0, (4, 8, 8))
        row = {"source": s, "rounds": r_ref, "live_slots": c_ref[1],
               "fraction_of_node_evaluations_that_change_the_value": c_ref[2] / (m * n * l * 8.0 * r_ref), "boxes": {}}
        for box in [(4, 4, 8), (4, 8, 8), (4, 16, 8), (8, 8, 8), (4, 32, 8)]:
            u, r, c = run(s, 2, box)
            row["boxes"]["x".join(map(str, box))] = {"bit_exact": bool(np.array_equal(u, u_ref) and r == r_ref),
                                                     "evaluated_fraction": c[0] / c[1]}
        out["sources"].append(row)
        print(json.dumps(row), flush=True)
    json.dump(out, open(os.path.join(ROOT, "profiles", "r02_skip_potential.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
