"""GPU (-m gpu): parity AT THE SIZES THE NUMBERS ARE QUOTED ON ("parity gaps").

The bench line runs the batch forward kernel (kernels_fwd_v3.cuh) on 128x128x64 (C3) and bench --config c4 on
200x200x80 (C4).  With few sources the library would pick the team kernel, so the batch kernel -- and the other
kernel families at their natural sizes -- are forced through their knobs, in a fresh process each:
forward travel times and round counts BIT-EXACT against the oracle at the production tolerance (1e-3), fused misfit
<= 1e-12 relative, slowness gradient <= 1e-10 relative to max|grad|, on the first sources of the bench workload.
"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_SCRIPT = r'''
import sys, numpy as np
root, grid, nsrc = sys.argv[1], sys.argv[2], int(sys.argv[3])
sys.path.insert(0, root); sys.path.insert(0, root + "/tests")
import bench, oracle, ref_misfit, adtomo_jl_b200 as A
dims = bench.GRID if grid == "c3" else bench.C4_GRID
w = bench.workload(1, 0, s_per_gpu=8, grid=dims, e_rcv=64)
m, n, l = dims
sta = w["sta"][:nsrc]
ptr, idx, val = A.corner_sources(sta, w["h"], w["vel0"])
u0 = np.full((nsrc, m, n, l), 1000.0)
for s in range(nsrc):
    u0[s].ravel()[idx[ptr[s]:ptr[s + 1]]] = val[ptr[s]:ptr[s + 1]]
ctx = A.Context(0)
# forward through the batched entry point
u = np.empty_like(u0)
rounds = np.zeros(nsrc, dtype=np.int32)
rc = ctx.forward3d_batch(u, u0, w["f"], w["h"], dims, bench.TOL, nsrc, rounds=rounds)
assert rc == 0, rc
# fused step (what the bench times): misfit + summed gradient
packed = np.zeros(m * n * l + 1)
r2 = np.zeros(nsrc, dtype=np.int32)
mis, rc = ctx.misfit_grad(packed, w["f"], w["h"], dims, bench.TOL, nsrc, ptr, idx, val, 1000.0, len(w["eve"]),
                          w["eve"], w["uobs"][:nsrc], w["qua"][:nsrc], rounds=r2)
assert rc == 0, rc
assert ctx.last_kernel().startswith(sys.argv[4]), (ctx.last_kernel(), sys.argv[4])
mis_ref, g_ref = 0.0, np.zeros((m, n, l))
for s in range(nsrc):
    ur, rr, _ = oracle.eikonal3d_forward(u0[s], w["f"], w["h"], bench.TOL)
    assert rr == rounds[s] == r2[s], (s, rr, rounds[s], r2[s])
    assert np.array_equal(ur, u[s]), "forward differs from the oracle, source %d" % s
    ms, gu = ref_misfit.misfit_and_grad_u(ur, w["eve"], w["uobs"][s], w["qua"][s])
    mis_ref += ms
    g_ref += oracle.eikonal3d_backward(gu, ur, u0[s], w["f"], w["h"])[1]
assert abs(mis - mis_ref) <= 1e-12 * abs(mis_ref), (mis, mis_ref)
assert packed[-1] == mis
err = np.abs(packed[:-1].reshape(m, n, l) - g_ref).max() / np.abs(g_ref).max()
assert err <= 1e-10, err
print("fullsize ok", grid, nsrc, list(rounds), "grad rel err %.1e" % err)
'''


@pytest.mark.parametrize("grid,nsrc,env,kernel", [
    ("c3", 4, {"ADTOMO_FORCE_V2": "1"}, "k_fwd3d_v3"),                               # the kernel of the bench line
    ("c3", 4, {"ADTOMO_FORCE_V2": "1", "ADTOMO_ADJ_SPARSE": "1"}, "k_fwd3d_v3"),      # ... with the active-set adjoint the bench batch runs on
    ("c3", 4, {"ADTOMO_FORCE_V2": "1", "ADTOMO_V3_STAGED": "1"}, "k_fwd3d_v3"),       # its one-CTA-per-SM variant
    ("c3", 4, {"ADTOMO_FORCE_V2": "1", "ADTOMO_V4": "1"}, "k_fwd3d_v4"),              # slot-block sweep (register hand-over; opt-in)
    ("c3", 2, {"ADTOMO_FORCE_V2": "1", "ADTOMO_V3": "2"}, "k_fwd3d_v3"),              # run-time row pitch
    ("c3", 2, {"ADTOMO_FORCE_V2": "1", "ADTOMO_V3": "0"}, "k_fwd3d_v2"),              # the round-1 sweep loop
    ("c3", 2, {"ADTOMO_TEAM": "0"}, "k_fwd3d_v1"),                                    # level-major kernel, one SM per source
    ("c3", 2, {}, "k_fwd3d_team"),                                                    # what the library picks for 2 sources
    ("c4", 4, {"ADTOMO_FORCE_V2": "1", "ADTOMO_ADJ_SPARSE": "1"}, "k_fwd3d_v3"),      # the kernels of bench --config c4
    ("c4", 2, {"ADTOMO_FORCE_V2": "1", "ADTOMO_V4": "1"}, "k_fwd3d_v4"),
    ("c4", 1, {"ADTOMO_TEAM": "0"}, "k_fwd3d_v1"),                                    # cluster kernel at its natural size
])
def test_bench_size_parity(tmp_path, grid, nsrc, env, kernel):
    script = tmp_path / "fs.py"
    script.write_text(_SCRIPT)
    p = subprocess.run([sys.executable, str(script), ROOT, grid, str(nsrc), kernel], env=dict(os.environ, **env),
                       capture_output=True, text=True, timeout=900)
    assert p.returncode == 0 and "fullsize ok" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]
