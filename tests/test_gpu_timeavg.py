"""The reference's time-averaging mode (-D_TIMEAVG_, src/main.f:510-541, :1107-1208, :1239-1297) on the device against
the oracle's restatement: two passes over the same run -- means first, then the fluctuation statistics about them."""
import numpy as np
import pytest

from oracle import get_oracle
from util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


def _decks():
    from wolfd2_b200 import deck as dk
    d1 = dk.cavity(44, re=200.0, dt=0.004, ny=38)
    d2 = dk.heated_cavity(40, re=150.0, dt=0.004, ny=36)
    d3 = dk.backward_step(48, re=80.0, dt=0.003, ny=40)
    return [d1, d2, d3]


@pytest.mark.parametrize("k", range(3))
def test_two_pass_time_average(api, k, tmp_path):
    from wolfd2_b200 import plot3d
    d = _decks()[k]
    d.msorit = 120
    nts = 7
    o = get_oracle()
    o.config(d.mnx, d.mny)
    rng = np.random.default_rng(3 + k)
    u0, v0, p0, t0 = (d.new_field() for _ in range(4))
    for f in (u0, v0):
        f[:d.ny + 2, :d.nx + 2] = 0.02 * rng.standard_normal((d.ny + 2, d.nx + 2))
    if d.thermal:
        t0[:d.ny + 2, :d.nx + 2] = 0.5 + 0.1 * rng.standard_normal((d.ny + 2, d.nx + 2))
    names = api.Context.TIMEAVG_NAMES
    # ---- oracle: two runs from the same state
    acc = [d.new_field() for _ in range(19)]
    for npass in (1, 2):
        uo, vo, po, to, do = u0.copy(), v0.copy(), p0.copy(), t0.copy(), d.new_field()
        o.coldstart(d, uo, vo, po)
        us, vs, ts, pn = (d.new_field() for _ in range(4))
        for _ in range(nts):
            rc, _lg = o.step(d, uo, vo, po, 1, t=to, d=do)
            assert rc == 0
            o.timeavg_accumulate(d, npass, uo, vo, po, to, us, vs, ts, pn, acc)
        o.timeavg_finish(d, npass, nts, acc)
    # ---- device
    with api.Context(d) as ctx:
        ctx.timeavg("begin")
        for npass in (1, 2):
            for w, f in ((api.F_U, u0), (api.F_V, v0), (api.F_P, p0), (api.F_T, t0), (api.F_D, d.new_field())):
                ctx.upload(w, f)
            ctx.coldstart()
            ctx.timeavg(npass)
            ctx.step(nts)
            ctx.timeavg_finish(npass, nts)
        got = {n: ctx.timeavg_get(n) for n in names}
        ctx.timeavg("release")
    scale = max(np.abs(acc[0]).max(), np.abs(acc[1]).max())
    assert scale > 1e-4 and np.abs(acc[7]).max() > 0          # a flow developed and it fluctuated in time
    for n, ref in zip(names, acc):
        g = got[n]
        assert np.isfinite(g).all(), n
        if n in ("upb", "vpb", "tpb"):      # mean fluctuation about the mean: zero up to rounding, on both sides
            assert np.abs(g - ref).max() <= 1e-10 * max(scale, 1.0), (n, np.abs(g - ref).max())
            continue
        tol = 1e-9 if n in ("ubar", "vbar", "tbar", "pbar") else 1e-6     # fluctuations are differences of nearly equal numbers
        if np.abs(ref).max() > 0:
            assert rel_l2(g, ref) <= tol, (n, rel_l2(g, ref))
    # SaveTmAvgP3D layout: 16 single-precision planes
    pre = str(tmp_path / "tavg")
    assert plot3d.save_tmavg_p3d(pre, d.nx, d.ny, got) == 16
    import struct
    raw = open(pre + ".qqq", "rb").read()
    assert struct.unpack("<iiiii", raw[:20]) == (12, d.nx, d.ny, 16, 12)
    assert struct.unpack("<i", raw[20:24])[0] == 16 * d.nx * d.ny * 4
    first = np.frombuffer(raw[24:24 + 4 * d.nx * d.ny], dtype="<f4").reshape(d.ny, d.nx)
    assert np.array_equal(first, got["ubar"][1:d.ny + 1, 1:d.nx + 1].astype(np.float32))
    assert open(pre + ".nam").read().split("\n")[2] == " Avgd. W-vel (null)"
