"""The compiled C++ host driver (wolfd2_b200/host/wolfd2_host.cpp: set-up + time loop over the C ABI, no
Python in the loop) must reproduce the oracle's PrintDiff lines for the lid-driven cavity."""
import subprocess

import numpy as np
import pytest

from oracle import get_oracle

pytestmark = pytest.mark.gpu


def test_compiled_host_driver_matches_oracle():
    from wolfd2_b200 import build, deck as dk
    exe = build.build_host()
    n, re, dt, nsteps, solver, sorrel, msorit = 40, 100.0, 0.01, 5, 5, 1.5, 400
    out = subprocess.run([exe, str(n), str(re), str(dt), str(nsteps), str(solver), str(sorrel), str(msorit)],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    lines = [l.split() for l in out.stdout.splitlines() if l.strip() and not l.startswith("checksum")]
    d = dk.cavity(n, re=re, dt=dt)
    d.sorrel, d.msorit = sorrel, msorit
    o = get_oracle()
    u, v, p = d.new_field(), d.new_field(), d.new_field()
    o.coldstart(d, u, v, p)
    rc, lo = o.step(d, u, v, p, nsteps)
    assert len(lines) == nsteps
    for l, r in zip(lines, lo):
        assert int(l[2].rstrip("*")) == r["nQLiter"] and int(l[-4]) == r["nSorConv"]
        np.testing.assert_allclose([float(x) for x in l[-3:]], r["dif"][:3], rtol=1e-5)
    chk = [float(x) for x in out.stdout.splitlines()[-1].split()[1:]]
    np.testing.assert_allclose(chk, [u.sum(), v.sum(), p.sum()], rtol=1e-9)
