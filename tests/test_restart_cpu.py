"""Restart-file interop (host-side): byte layout of the gfortran records and round trip."""
import struct

import numpy as np
import pytest

from wolfd2_b200 import deck as dk, restart


def test_record_layout_and_roundtrip(tmp_path):
    d = dk.cavity(12, re=100.0, dt=0.01, ny=9, mnx=20, mny=15)
    rng = np.random.default_rng(0)
    u, v, p = (rng.uniform(-1, 1, (d.mny + 1, d.mnx + 1)) for _ in range(3))
    path = tmp_path / "restart.out"
    restart.save_restart(str(path), d, 37, 0.37, u, v, p)
    raw = path.read_bytes()
    n = (d.nx + 2) * (d.ny + 2)
    # record 1: k (i4), dtime (r8) -> 12 bytes framed by markers; record 2: nx, ny; record 3: 8 fields
    assert struct.unpack("<i", raw[:4])[0] == 12 and struct.unpack("<i", raw[16:20])[0] == 12
    assert struct.unpack("<id", raw[4:16]) == (37, 0.37)
    assert struct.unpack("<iiii", raw[20:36]) == (8, d.nx, d.ny, 8)
    assert struct.unpack("<i", raw[36:40])[0] == 8 * n * 8
    assert len(raw) == 20 + 16 + 8 + 8 * n * 8
    first_p = np.frombuffer(raw[40:40 + 8 * (d.nx + 2)], dtype="<f8")
    assert np.array_equal(first_p, p[0, :d.nx + 2])            # p first, i fastest, row j = 0
    k, t, f = restart.read_restart(str(path), d)
    assert (k, t) == (37, 0.37)
    W = (slice(0, d.ny + 2), slice(0, d.nx + 2))
    assert np.array_equal(f["u"][W], u[W]) and np.array_equal(f["v"][W], v[W]) and np.array_equal(f["p"][W], p[W])
    assert not f["t"].any() and not f["tss"].any()
    assert f["u"][d.ny + 2:, :].max(initial=0.0) == 0.0           # outside the window stays zero
    with pytest.raises(ValueError, match="Index mismatch"):
        restart.read_restart(str(path), dk.cavity(13, re=100.0, dt=0.01, ny=9, mnx=20, mny=15))


def test_subrecord_framing(monkeypatch, tmp_path):
    """Records above the marker limit are split into sub-records with signed markers."""
    monkeypatch.setattr(restart, "_MAXREC", 1000)
    d = dk.cavity(12, re=100.0, dt=0.01, ny=9)
    u = np.arange((d.mny + 1) * (d.mnx + 1), dtype=float).reshape(d.mny + 1, d.mnx + 1)
    path = tmp_path / "r.out"
    restart.save_restart(str(path), d, 1, 0.5, u, 2 * u, 3 * u)
    raw = path.read_bytes()
    assert struct.unpack("<i", raw[36:40])[0] == -1000            # first sub-record: more follow
    assert struct.unpack("<i", raw[1040:1044])[0] == 1000          # its trailing marker: none precedes
    assert struct.unpack("<i", raw[1044:1048])[0] == -1000         # second sub-record
    assert struct.unpack("<i", raw[2048:2052])[0] == -1000         # trailing: one precedes
    k, t, f = restart.read_restart(str(path), d)
    assert np.array_equal(f["v"][:d.ny + 2, :d.nx + 2], 2 * u[:d.ny + 2, :d.nx + 2])
