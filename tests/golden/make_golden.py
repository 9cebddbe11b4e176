"""Regenerates tests/golden/*.npz from the CPU oracle.

The reference ships no golden vectors and cannot be built here (no Fortran compiler), so these fixtures
are REGRESSION PINS of the oracle (oracle/wolfd2_oracle.c at -O2 -ffp-contract=off), not outputs of the
reference.  They let the GPU tests check the CUDA path against committed numbers and catch any later
drift of the oracle itself.  tests/test_oracle_numpy_crosscheck.py reproduces every fixture bit for bit with the second
(numpy / Python) restatement, without calling the oracle.   Run:  python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from oracle import get_oracle  # noqa: E402
from wolfd2_b200 import deck as dk  # noqa: E402


def cases():
    c = {}
    d = dk.cavity(24, re=100.0, dt=0.02, ny=20); d.msorit = 150; d.sorrel = 1.5
    c["cavity24x20"] = d
    d = dk.channel(22, re=100.0, dt=0.01, ny=18, fully_dev=True); d.msorit = 400; d.sorrel = 1.6
    c["channel22x18_fd"] = d
    d = dk.channel(22, re=100.0, dt=0.01, ny=18, fully_dev=False); d.msorit = 400; d.sorrel = 1.6
    c["channel22x18_mc"] = d
    d = dk.backward_step(26, re=100.0, dt=0.01, ny=20); d.msorit = 400; d.sorrel = 1.6
    c["bstep26x20"] = d
    d = dk.heated_cavity(26, re=100.0, dt=0.01, ny=22); d.msorit = 300; d.sorrel = 1.5   # thermal energy + EqState, 3 M-E iterations
    c["heated_cavity26x22"] = d
    return c


def run(d, nsteps=4):
    o = get_oracle()
    u, v, p, t, den = (d.new_field() for _ in range(5))
    thermal = getattr(d, "thermal", False)
    if thermal:
        t[:d.ny + 2, :d.nx + 2] = 0.5
    ncold = o.coldstart(d, u, v, p)
    out = {"ncold": np.array(ncold)}
    nql, nsor, dif = [], [], []
    for k in range(nsteps):
        rc, lg = o.step(d, u, v, p, 1, t=t, d=den)
        assert rc == 0
        nql.append(lg[0]["nQLiter"]); nsor.append(lg[0]["nSorConv"]); dif.append(lg[0]["dif"][:4 if thermal else 3])
        out[f"u{k}"], out[f"v{k}"], out[f"p{k}"] = u.copy(), v.copy(), p.copy()
        if thermal:
            out[f"t{k}"], out[f"d{k}"] = t.copy(), den.copy()
    out["nql"], out["nsor"], out["dif"] = np.array(nql), np.array(nsor), np.array(dif)
    return out


if __name__ == "__main__":
    for name, d in cases().items():
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **run(d))
        print("wrote", name)
