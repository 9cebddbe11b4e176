"""PLOT3D output interop (host-side): byte layout of the gfortran records SaveStdVarsP3D / SaveGrid2DP3D
write (src/file_manip.f:1024-1173, :319-357), variable order, name file and round trip."""
import struct

import numpy as np
import pytest

from wolfd2_b200 import deck as dk, plot3d


def _fields(d, n, seed=0):
    rng = np.random.default_rng(seed)
    return [rng.uniform(-1, 1, (d.mny + 1, d.mnx + 1)) for _ in range(n)]


def test_unformatted_layout_cold_flow(tmp_path):
    d = dk.cavity(12, re=100.0, dt=0.01, ny=9, mnx=20, mny=15)
    u, v, p = _fields(d, 3)
    pre = str(tmp_path / "run")
    gx, gy = d.node_arrays()
    assert plot3d.save_std_vars_p3d(pre, d.nx, d.ny, u, v, p, grid=(gx, gy)) == 4
    raw = open(pre + ".qqq", "rb").read()
    n = d.nx * d.ny
    assert struct.unpack("<iiiii", raw[:20]) == (12, d.nx, d.ny, 4, 12)          # record 1: nx, ny, nVars
    assert struct.unpack("<i", raw[20:24])[0] == 4 * n * 8 and len(raw) == 20 + 8 + 4 * n * 8
    body = np.frombuffer(raw[24:24 + 4 * n * 8], dtype="<f8").reshape(4, d.ny, d.nx)
    W = (slice(1, d.ny + 1), slice(1, d.nx + 1))                                 # nodes i = 1..nx, j = 1..ny
    assert np.array_equal(body[0], p[W]) and np.array_equal(body[1], u[W]) and np.array_equal(body[2], v[W])
    assert not body[3].any()                                                      # W = 0 in 2-D files
    # grid file: REAL*4 coordinates, x plane then y plane
    g = open(pre + ".xyz", "rb").read()
    assert struct.unpack("<iiii", g[:16]) == (8, d.nx, d.ny, 8)
    assert struct.unpack("<i", g[16:20])[0] == 2 * n * 4 and len(g) == 16 + 8 + 2 * n * 4
    xy = np.frombuffer(g[20:20 + 2 * n * 4], dtype="<f4").reshape(2, d.ny, d.nx)
    assert np.array_equal(xy[0], gx[W].astype(np.float32)) and np.array_equal(xy[1], gy[W].astype(np.float32))
    assert xy[0][0, 0] == 0.0 and xy[0][0, -1] == 1.0 and xy[1][-1, 0] == 1.0
    assert [l.strip() for l in open(pre + ".nam")] == ["Complete Pressure", "Complete U ; Complete Velocity",
                                                        "Complete V", "Complete W"]
    nx, ny, planes = plot3d.read_std_vars_p3d(pre + ".qqq")
    assert (nx, ny, len(planes)) == (d.nx, d.ny, 4) and np.array_equal(planes[1], u[W])


def test_thermal_and_small_scale_variable_order(tmp_path):
    d = dk.cavity(10, re=100.0, dt=0.01, ny=8)
    u, v, p, t, us, vs, ps, ts = _fields(d, 8, seed=3)
    pre = str(tmp_path / "ss")
    W = (slice(1, d.ny + 1), slice(1, d.nx + 1))
    assert plot3d.save_std_vars_p3d(pre, d.nx, d.ny, u, v, p, t=t, us=us, vs=vs, ps=ps, ts=ts) == 10
    _, _, pl = plot3d.read_std_vars_p3d(pre + ".qqq")
    for k, a in enumerate([p, u, v, None, t, ps, us, vs, None, ts]):
        assert (not pl[k].any()) if a is None else np.array_equal(pl[k], a[W]), k
    assert [l.strip() for l in open(pre + ".nam")][4::5] == ["Complete Temperature", "Small-Scale Temperature"]
    # small scales without thermal energy: 8 variables (nLamV = 4)
    assert plot3d.save_std_vars_p3d(pre, d.nx, d.ny, u, v, p, us=us, vs=vs, ps=ps, names=False) == 8
    _, _, pl = plot3d.read_std_vars_p3d(pre + ".qqq")
    assert np.array_equal(pl[4], ps[W]) and np.array_equal(pl[5], us[W]) and not pl[7].any()


def test_formatted_files_round_trip(tmp_path):
    d = dk.cavity(9, re=100.0, dt=0.01, ny=7)
    u, v, p = _fields(d, 3, seed=5)
    pre = str(tmp_path / "fmt")
    gx, gy = d.node_arrays()
    plot3d.save_std_vars_p3d(pre, d.nx, d.ny, u, v, p, form=plot3d.FT_FORMATTED, grid=(gx, gy))
    nx, ny, pl = plot3d.read_std_vars_p3d(pre + ".qqq", form=plot3d.FT_FORMATTED)
    W = (slice(1, d.ny + 1), slice(1, d.nx + 1))
    assert (nx, ny) == (d.nx, d.ny) and np.array_equal(pl[0], p[W]) and np.array_equal(pl[2], v[W])   # 17 digits: exact
    lines = open(pre + ".xyz").read().splitlines()
    assert lines[0].split() == [str(d.nx), str(d.ny)]
    assert all(len(l) == 70 for l in lines[1:-1])                       # 5(e14.6)
    vals = np.array([float(l[k:k + 14]) for l in lines[1:] for k in range(0, len(l), 14)])
    assert np.allclose(vals[:nx * ny], gx[W].ravel(), rtol=1e-6, atol=1e-12)
    assert plot3d._e14_6(0.5) == "  0.500000E+00" and plot3d._e14_6(-1.25e-3) == " -0.125000E-02"
    with pytest.raises(ValueError, match="Wrong nForm"):
        plot3d.save_std_vars_p3d(pre, d.nx, d.ny, u, v, p, form=7)
