"""Parity with the CPU oracle AT THE BASELINE SIZES (BASELINE.json configs[1] 1024^2, the metric grid 4096^2,
configs[2] channel / backward step 4096^2).  The chain solver's third level, 64-bit offsets and the one-wave band
logic of the fused SOR only occur at these sizes, so the small-grid parity tests cannot stand in for them.

The oracle (oracle/wolfd2_oracle.c, -O2 -ffp-contract=off, one core) takes about 1 s per fixed-work step at 1024^2
and 6-16 s at 4096^2, which is affordable for ONE step per deck.  Bars (north star): identical QL and SOR counts,
rel-L2 <= 1e-10 on u, v, p per step, the PrintDiff tuple to 1e-9 relative.
"""
import numpy as np
import pytest

from util import rel_l2

pytestmark = pytest.mark.gpu

TOL = 1e-10


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


def _run_both(api, d, nsteps, state=None):
    import bench
    from oracle import get_oracle
    o = get_oracle()
    f0 = state if state is not None else bench.developed_state(d)
    uo, vo, po = (a.copy() for a in f0)
    nco = o.coldstart(d, uo, vo, po)
    rc, lo = o.step(d, uo, vo, po, nsteps)
    assert rc == 0
    with api.Context(d) as ctx:
        for w, f in zip((api.F_U, api.F_V, api.F_P), f0):
            ctx.upload(w, f)
        ncg = ctx.coldstart()
        lg = ctx.step(nsteps)
        ug, vg, pg = ctx.download(api.F_U), ctx.download(api.F_V), ctx.download(api.F_P)
    assert ncg == nco, f"cold-start SOR count {ncg} vs oracle {nco}"
    for k, (g, r) in enumerate(zip(lg, lo)):
        assert g["nQLiter"] == r["nQLiter"] and g["nSorConv"] == r["nSorConv"], (k, g, r)
        for a, b in zip(g["dif"][:3], r["dif"][:3]):
            assert abs(a - b) <= 1e-9 * max(abs(b), 1e-300), (k, g["dif"], r["dif"])
    errs = {}
    for nm, a, b in (("u", ug, uo), ("v", vg, vo), ("p", pg, po)):
        assert np.isfinite(a).all(), nm
        errs[nm] = rel_l2(a, b)
        assert errs[nm] <= TOL, f"{d.name}: field {nm} rel-L2 {errs[nm]:.3e} vs the oracle after {nsteps} step(s)"
    return lg, errs


def test_cavity_1024_fixed_work_two_steps(api):
    """configs[1]: lid-driven cavity Re=1000 on 1024^2, the bench's fixed work (Q=2, S=100)."""
    import bench
    d = bench.make_deck("cavity", 1024, True, 2, 100)
    _run_both(api, d, 2)


def test_cavity_1024_converging_sor(api):
    """Same grid with a tolerance the SOR reaches (count decided by the on-device test, mid-pass repeats included)
    and a QL loop that stops on its own."""
    import bench
    d = bench.make_deck("cavity", 1024, False, 0, 0)
    d.sortol, d.msorit, d.sorrel, d.qtol = 1e-3, 400, 1.9, 1e-4
    lg, _ = _run_both(api, d, 2)
    assert any(0 < l["nSorConv"] < d.msorit for l in lg), [l["nSorConv"] for l in lg]   # it did converge
    assert all(0 < l["nQLiter"] < d.mqiter for l in lg), [l["nQLiter"] for l in lg]


def test_cavity_4096_fixed_work_one_step(api):
    """The metric grid of BASELINE.json: exactly the bench workload (Q=2, S=100), one step."""
    import bench
    d = bench.make_deck("cavity", 4096, True, 2, 100)
    _run_both(api, d, 1)


@pytest.mark.parametrize("workload,outlet", [("channel", "fully_dev"), ("channel", "mass_cons"), ("bstep", "fully_dev")])
def test_inflow_outflow_4096_one_step(api, workload, outlet):
    """configs[2]: channel (both outlet types) and backward-facing step with a blockage on 4096^2, plug-flow start,
    fixed work Q=2, S=40."""
    import bench
    d = bench.make_deck(workload, 4096, True, 2, 40, outlet=outlet)
    _run_both(api, d, 1)


@pytest.mark.parametrize("outlet", ["fully_dev", "mass_cons"])
def test_channel_512_converged(api, outlet):
    """configs[2] in converged mode at a size where the point SOR does converge: the iteration counts are the
    reference's, not a cap."""
    import bench
    d = bench.make_deck("channel", 512, False, 0, 0, outlet=outlet)
    d.sortol, d.msorit, d.sorrel = 1e-6, 6000, 1.97
    lg, _ = _run_both(api, d, 2)
    assert all(l["nSorConv"] < d.msorit for l in lg), [l["nSorConv"] for l in lg]
