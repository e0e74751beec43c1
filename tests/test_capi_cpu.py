"""CPU: the C-ABI library builds, loads and exports every symbol include/adtomo_b200.h declares;
host-side logic (sharding, corner sources, synthetic models, gloo all-reduce plumbing)."""
import ctypes
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_header_symbols(lib):
    L = lib.load_library()
    names = lib.exported_symbols()
    assert len(names) >= 15
    for name in names:
        assert hasattr(L, name), f"{name} declared in include/adtomo_b200.h but not exported"
    assert L.adtomo_version() >= 100


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lib.AdtomoError):
        lib.Context(0)
    u = np.zeros((3, 3, 3))
    with pytest.raises(lib.AdtomoError):
        lib.eikonal3d_forward(u, u, 1.0, 3, 3, 3, 1e-6)


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "adtomo.jl_b200")
    for dp, _, fs in os.walk(pkg):
        for fn in fs:
            if fn.endswith((".py", ".cu", ".cuh", ".h", ".sh")):
                txt = open(os.path.join(dp, fn)).read()
                assert "oracle" not in txt.lower().replace("no oracle", ""), f"{fn} mentions the oracle"


def test_shard_sources(lib):
    S = 11
    seen = np.concatenate([lib.shard_sources(S, r, 4) for r in range(4)])
    assert sorted(seen) == list(range(S))
    assert list(lib.shard_sources(S, 1, 4)) == [1, 5, 9]     # rank+1:nproc:numsta, 0-based


def test_corner_sources(lib):
    vel = np.full((6, 5, 4), 2.0)
    ptr, idx, val = lib.corner_sources([[1.25, 2.0, 0.5], [3.0, 3.0, 2.0]], 0.5, vel)
    assert list(ptr) == [0, 8, 16]
    # station 0: x in {2,1}, y in {2,2}, z in {1,0}
    nodes = {(2, 2, 1), (2, 2, 0), (1, 2, 1), (1, 2, 0)}
    got = {(i // 20, (i // 4) % 5, i % 4) for i in idx[:8]}
    assert got == nodes
    k = list(idx[:8]).index((1 * 5 + 2) * 4 + 0)
    assert val[k] == pytest.approx(np.sqrt(0.25 ** 2 + 0.5 ** 2) * 0.5 / 2.0)
    # integer station: all 8 entries are the node itself with time 0
    assert set(idx[8:]) == {(3 * 5 + 3) * 4 + 2} and np.all(val[8:] == 0.0)


def test_synthetic_models(lib):
    from adtomo_jl_b200 import synthetic as syn
    v = syn.gil7_velocity(4, 4, 64, 1.0)
    assert v[0, 0, 0] == 3.20 and v[0, 0, 1] == 3.20 and v[0, 0, 2] == 4.50   # (k1-2)*h >= 1 first at k1 = 3
    c = syn.checkerboard(v, 10, 0.8)
    assert c[0, 0, 0] == pytest.approx(3.20 + 0.8) and c[0, 0, 10] == pytest.approx(v[0, 0, 10] - 0.8)
    f = syn.model_2d_test()
    assert f.shape == (30, 40) and f[15, 19] == 0.2 and f[7, 9] == pytest.approx(1 / 7)
    sta, eve = syn.stations_events(32, 32, 16, 5, 7)
    assert sta.shape == (5, 3) and eve.shape == (7, 3)
    assert (eve >= 0).all() and (eve[:, 2] <= 15).all()


_GLOO = r'''
import os, sys
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
import adtomo_jl_b200 as A
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
r = dist.get_rank()
S, N = 7, 10
mine = A.shard_sources(S, r, 2)
# stand-in for the per-rank packed [grad | misfit] buffer: every source s contributes (s+1) to every entry
packed = np.zeros(N + 1)
for s in mine:
    packed += (s + 1)
A.InversionProblem.allreduce(packed)
assert np.all(packed == sum(range(1, S + 1))), packed
dist.destroy_process_group()
print("ok", r)
'''


def test_gloo_allreduce_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(_GLOO)
    port = str(29500 + os.getpid() % 2000)
    ps = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                           stderr=subprocess.STDOUT) for r in range(2)]
    outs = [p.communicate(timeout=180)[0].decode() for p in ps]
    for p, o in zip(ps, outs):
        assert p.returncode == 0, o


def test_lbfgs_driver_on_quadratic(lib, tmp_path, capsys):
    """gpu_optimize (mirror of mpi_optimize's closure form): converges on a convex quadratic, logs like the
    reference and checkpoints every `steps` gradient evaluations."""
    rng = np.random.default_rng(0)
    Q = rng.standard_normal((12, 12))
    Q = Q @ Q.T + 12 * np.eye(12)
    b = rng.standard_normal(12)
    f = lambda x: 0.5 * x @ Q @ x - b @ x
    g = lambda x: Q @ x - b
    x, hist = lib.gpu_optimize(f, g, np.zeros(12), iterations=60, loc=str(tmp_path), steps=5)
    assert np.abs(x - np.linalg.solve(Q, b)).max() < 1e-6
    assert all(h2 <= h1 + 1e-12 for h1, h2 in zip(hist, hist[1:]))
    out = capsys.readouterr().out
    assert "iter 1, current loss=" in out and "================== STEP 1 ==================" in out
    # checkpoints: HDF5 dataset "data" holding the raw optimiser vector (mpi_optimize.jl:26-28), no temporary left behind
    from adtomo_jl_b200 import hdf5_min
    ck = hdf5_min.read_dataset(str(tmp_path / "iter_5.h5"), "data")
    assert ck.shape == (12,) and np.isfinite(ck).all()
    assert not [n for n in os.listdir(tmp_path) if ".tmp" in n]
    with pytest.raises(ValueError):
        lib.gpu_optimize(f, g, np.zeros(12), method="NelderMead")
    # the reference's other method (mpi_optimize.jl:40-45); rank != 0 neither logs nor writes
    xb, hb = lib.gpu_optimize(f, g, np.zeros(12), method="BFGS", iterations=60, loc=str(tmp_path / "r1"), steps=1, rank=1)
    assert np.abs(xb - np.linalg.solve(Q, b)).max() < 1e-6
    assert capsys.readouterr().out == "" and not os.path.exists(tmp_path / "r1")


def test_session_form_returns_variable_list(lib, tmp_path):
    """gpu_optimize_model = the session form of mpi_optimize (:76-143): variables flattened, optimised, returned as a
    list by rank 0 and as None by the others."""
    class Model:            # a quadratic "graph": var_change (2,3,2) and one free scale
        vel0 = np.ones((2, 3, 2))
        c = np.arange(13, dtype=np.float64) / 7.0
        def x0(self): return np.zeros(13)
        def loss(self, z): return float(((np.asarray(z).ravel() - self.c) ** 2).sum())
        def grad(self, z): return 2.0 * (np.asarray(z).ravel() - self.c)
    out = lib.gpu_optimize_model(Model(), iterations=50, verbose=False, rank=0)
    assert [o.shape for o in out] == [(2, 3, 2), (1,)]
    assert np.abs(np.concatenate([o.ravel() for o in out]) - Model.c).max() < 1e-8
    assert lib.gpu_optimize_model(Model(), iterations=5, verbose=False, rank=1) is None


def test_hdf5_min_roundtrip_and_real_file(lib, tmp_path):
    """The checkpoint writer produces what its reader -- checked here against a file written by libhdf5 itself --
    reads back bit for bit; the datatype message equals the library's byte for byte."""
    from adtomo_jl_b200 import hdf5_min as H
    rng = np.random.default_rng(3)
    x = rng.standard_normal(1537)
    H.write_dataset(str(tmp_path / "a.h5"), "data", x)
    assert H.list_names(str(tmp_path / "a.h5")) == ["data"]
    assert np.array_equal(H.read_dataset(str(tmp_path / "a.h5"), "data"), x)
    many = {"data": x.reshape(29, 53), "matrix": np.arange(12, dtype=np.int32).reshape(3, 4), "idx": np.arange(5, dtype=np.int64),
            "b": np.float32([1.5, -2.0])}
    H.write_datasets(str(tmp_path / "b.h5"), many)
    assert H.list_names(str(tmp_path / "b.h5")) == sorted(many)
    for k, v in many.items():
        r = H.read_dataset(str(tmp_path / "b.h5"), k)
        assert r.dtype == v.dtype and np.array_equal(r, v)
    raw = open(tmp_path / "a.h5", "rb").read()
    assert raw[:8] == H.SIGNATURE and len(raw) == int.from_bytes(raw[40:48], "little")      # end-of-file address
    # a file written by the HDF5 library (MATLAB 7.3 container, 512-byte user block) shipped with scipy
    import scipy.io
    real = os.path.join(os.path.dirname(scipy.io.__file__), "matlab", "tests", "data", "testhdf5_7.4_GLNX86.mat")
    if os.path.exists(real):
        assert H.list_names(real) == ["testdouble"]
        d = H.read_dataset(real, "testdouble")
        assert np.allclose(d.ravel(), np.arange(9) * np.pi / 4, rtol=0, atol=1e-15)
        f = H._File(real)
        dt = [b for t, _, b in f.messages(f.links(f.root["oh"])["testdouble"]) if t == 0x0003][0]
        assert bytes(dt[:20]) == H._DTYPES["<f8"]


def test_box_filter_is_periodic_mean(lib):
    a = np.random.default_rng(1).random((7, 6, 5))
    s = lib.box_filter_periodic(a, 5, 3)
    # brute force at one node with wrap-around indices
    i, j, k = 0, 5, 4
    ref = np.mean([a[(i + di) % 7, (j + dj) % 6, (k + dk) % 5] for di in range(-2, 3) for dj in range(-2, 3)
                   for dk in range(-1, 2)])
    assert s[i, j, k] == pytest.approx(ref, rel=1e-13)
    assert s.sum() == pytest.approx(a.sum(), rel=1e-12)
