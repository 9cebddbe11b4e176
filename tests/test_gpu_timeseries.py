"""Time-series monitor points sampled on the device (SaveTimeSrs case 1, src/file_manip.f:806-834; main.f:984-995):
every record must hold exactly the values of the resident fields at the sampled step, and the `.ts` files written from
them must be the ones a host-side sampling of downloaded fields gives."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


def test_probe_records_equal_the_fields(api, tmp_path):
    from wolfd2_b200 import deck as dk
    from wolfd2_b200.timeseries import TimeSeriesWriter
    d = dk.cavity(48, re=100.0, dt=0.005, ny=40)
    d.msorit = 80
    iTS, jTS = [5, 24, 47, 99], [5, 20, 39, 3]          # the last one lies outside the grid: the reference falls back to (1,1)
    z = d.new_field()
    with api.Context(d) as ctx:
        for w in (api.F_U, api.F_V, api.F_P):
            ctx.upload(w, z)
        ctx.coldstart()
        ctx.set_probes(iTS, jTS, freq=2)
        want = []
        for k in range(1, 8):
            ctx.step(1)
            if k % 2 == 0:
                f = [ctx.download(w) for w in (api.F_U, api.F_V, api.F_P)]
                want.append((k, f))
        steps, rec = ctx.probe_records()
        assert list(steps) == [2, 4, 6] and rec.shape == (3, 4, 8)
        pts = [(5, 5), (24, 20), (47, 39), (1, 1)]
        for (k, f), r in zip(want, rec):
            for q, (i, j) in enumerate(pts):
                assert r[q, 0] == f[0][j, i] and r[q, 1] == f[1][j, i] and r[q, 2] == f[2][j, i]
                assert (r[q, 3:] == 0.0).all()             # cold flow without the small-scale model
        assert np.abs(rec[:, :3, :3]).max() > 0
        s2, r2 = ctx.probe_records()
        assert len(s2) == 0                                # handed out once
        # more samples than the device buffer holds between two reads
        ctx.set_probes(iTS[:2], jTS[:2], freq=1)
        ctx.set_params(msorit=2, mqiter=1)
        ctx.step(1100)
        s3, r3 = ctx.probe_records()
        assert list(s3) == list(range(1, 1101)) and r3.shape == (1100, 2, 8)
        u = ctx.download(api.F_U)
        assert r3[-1, 1, 0] == u[20, 24]
    with TimeSeriesWriter(str(tmp_path / "cav"), iTS, jTS) as w:
        for k, r in zip(steps, rec):
            w.write(k * d.dk * d.dlref / d.uref, r)
    lines = open(tmp_path / "cav002.ts").read().split("\n")
    assert lines[2] == "# Location:   24  20" and len(lines) == 5 + 3 + 1
    assert abs(float(lines[5][14:28]) - rec[0, 1, 0]) <= 1e-6 * max(abs(rec[0, 1, 0]), 1e-30) + 1e-300
