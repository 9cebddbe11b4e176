"""GPU parity of the thermal energy row: TempBoundCond, ThermEnergy, EqState, Filter(_T_)
and the momentum-energy iteration loop of the step body, through the C ABI against the CPU oracle.
Bit-exact for the ghost fills, EqState and the filter; the ThermEnergy solve uses the parallel elimination
order of the momentum solves and is held to the same tolerances."""
import numpy as np
import pytest

from oracle import get_oracle
from util import rand_field, rel_l2

pytestmark = pytest.mark.gpu

TOL_SOLVE = 1e-12
TOL_STEP = 1e-10


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


@pytest.fixture(scope="module")
def orc():
    return get_oracle()


def _thermal_decks():
    from wolfd2_b200 import deck as dk
    out = [dk.heated_cavity(37, re=100.0, dt=0.005, ny=29), dk.heated_cavity(64, re=400.0, dt=0.004)]
    # 2x2 regions: a heat source, a fixed-temperature block, flux and temperature faces, inflow/outflow
    reg = dk.RegionTables(44, 36, 2, 2, (20,), (16,))
    reg.heat_generation(1, 1, 2.5).fixed_temperature_region(2, 2, 0.8)
    reg.wall_temperature(1, 1, "w", 1.0).wall_heat_flux(1, 2, "w", 0.05).wall_temperature(2, 1, "s", 0.2)
    reg.wall_heat_flux(1, 2, "n", -0.02).wall(1, 2, "n", tangent_vel=1.0)
    out.append(dk._mk("thermal_2x2", 44, 36, reg, 100.0, 0.004, thermal=True, eqstate=True, nmeiter=2))
    reg = dk.RegionTables(40, 30, 2, 1, (18,), ())
    reg.inlet(1, 1, "w", normal_vel=1.0).outlet(2, 1, "e", fully_dev=True)
    reg.wall_temperature(1, 1, "w", 0.0).wall_temperature(1, 1, "s", 1.0).wall_temperature(2, 1, "s", 1.0)
    reg.heat_generation(2, 1, 1.0)
    out.append(dk._mk("thermal_channel", 40, 30, reg, 80.0, 0.004, thermal=True, eqstate=False, nmeiter=3,
                      nfiltt=1, fpt=200.0))
    return out


DECKS = _thermal_decks()
IDS = [d.name for d in DECKS]


def _cfg(api, orc, d):
    api.config(d.mnx, d.mny)
    orc.config(d.mnx, d.mny)


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_tempboundcond_eqstate_filter_bitwise(api, orc, d):
    _cfg(api, orc, d)
    rng = np.random.default_rng(2024)
    r = d.regions
    t = rand_field(d, rng, 0.0, 1.0)
    tg, to = t.copy(), t.copy()
    args = (d.nx, d.ny, r.nReg, r.nRegBrd, r.nTRgType, r.nTemBdTp, r.dTRgVal, r.dBCVal)
    api.TempBoundCond(*args, tg)
    orc.tempboundcond(*args, to)
    assert np.array_equal(tg, to)
    api.TempBoundCond(*args, tg)          # a second application (corner cells read ghosts written before)
    orc.tempboundcond(*args, to)
    assert np.array_equal(tg, to)
    p = rand_field(d, rng, -0.3, 0.3)
    dg, do = rand_field(d, rng), None
    do = dg.copy()
    es = (d.nx, d.ny, d.uref, d.densref, d.tmax, d.tref, d.rconst, p, t)
    api.EqState(*es, dg)
    orc.eqstate(*es, do)
    assert np.array_equal(dg, do)
    # a state at reference conditions gives exactly zero (the 1e-10 clip, thermal.f:320)
    z = d.new_field()
    dz = rand_field(d, rng)
    api.EqState(d.nx, d.ny, d.uref, d.densref, d.tmax, d.tref, d.rconst, z, z, dz)
    assert np.all(dz[2:d.ny + 1, 2:d.nx + 1] == 0.0)
    fg, fo = t.copy(), t.copy()
    fa = (d.nx, d.ny, 4, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.nTRgType, 300.0)
    api.Filter(*fa, fg)
    orc.filter(*fa, fo)
    assert np.array_equal(fg, fo)


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_thermenergy(api, orc, d):
    """One ThermEnergy call (TempBoundCond + both split steps + update) on random fields."""
    _cfg(api, orc, d)
    rng = np.random.default_rng(99)
    r, m = d.regions, d.metrics
    un, vn, u, v = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    tn, t = rand_field(d, rng, 0.0, 1.0), rand_field(d, rng, 0.0, 1.0)
    mm = [m[n] for n in "rau rbu rbv rgv djc xeu yeu xzv yzv xec yec xzc yzc".split()]
    tg, to = t.copy(), t.copy()
    args = (d.nx, d.ny, r.nReg, r.nRegBrd, r.nTRgType, r.nTemBdTp, d.dk, d.pe, r.dTRgVal, r.dHGSTval, r.dBCVal,
            *mm, un, vn, u, v, tn)
    api.ThermEnergy(*args, tg)
    orc.thermenergy(*args, to)
    assert np.isfinite(to).all()
    assert rel_l2(tg, to) <= TOL_SOLVE, rel_l2(tg, to)
    # cells ThermEnergy does not own (ghost ring written by TempBoundCond only) are bit-identical
    ring = np.ones_like(t, dtype=bool)
    ring[2:d.ny + 1, 2:d.nx + 1] = False
    assert np.array_equal(tg[ring], to[ring])


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_thermal_time_steps(api, orc, d):
    """Cold start + 6 steps with the momentum-energy iterations: u, v, p, t, d per step, identical QL / SOR
    counts and PrintDiff tuple (4 columns)."""
    d.msorit = 300
    orc.config(d.mnx, d.mny)
    uo, vo, po, to, do = (d.new_field() for _ in range(5))
    to[:d.ny + 2, :d.nx + 2] = 0.5
    t0 = to.copy()
    nso = orc.coldstart(d, uo, vo, po)
    worst = 0.0
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P, api.F_D):
            ctx.upload(w, z)
        ctx.upload(api.F_T, t0)
        assert ctx.coldstart() == nso
        for step in range(6):
            lg = ctx.step(1)[0]
            rc, lo = orc.step(d, uo, vo, po, 1, t=to, d=do)
            assert rc == 0
            assert lg["nQLiter"] == lo[0]["nQLiter"] and lg["nSorConv"] == lo[0]["nSorConv"]
            got = [ctx.download(w) for w in (api.F_U, api.F_V, api.F_P, api.F_T, api.F_D)]
            errs = [rel_l2(g, o) for g, o in zip(got, (uo, vo, po, to, do))]
            worst = max(worst, *errs)
            assert max(errs) <= TOL_STEP, f"{d.name} step {step}: rel-L2 (u,v,p,t,d) = {errs}"
            np.testing.assert_allclose(lg["dif"], lo[0]["dif"], rtol=1e-9, atol=1e-14)
        assert np.abs(to[2:d.ny + 1, 2:d.nx + 1] - 0.5).max() > 1e-3     # the temperature field did evolve
    print(f"{d.name}: worst rel-L2 over 6 steps = {worst:.2e}")


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_taveraged_bitwise_and_resident_dump(api, orc, d, tmp_path):
    """TAveraged (utility.f:668-740) through the shim, both scales, sentinel values outside the loops; then the
    resident variant after a few thermal steps and the 5-variable `.qqq` dump."""
    from wolfd2_b200 import plot3d
    _cfg(api, orc, d)
    rng = np.random.default_rng(11)
    r = d.regions
    t = rand_field(d, rng)
    for nscale in (0, 1):
        s = rand_field(d, rng)
        g, o = s.copy(), s.copy()
        api.TAveraged(d.nx, d.ny, nscale, r.nReg, r.nRegBrd, r.nTRgType, r.dTRgVal, t, g)
        orc.taveraged(d.nx, d.ny, nscale, r.nReg, r.nRegBrd, r.nTRgType, r.dTRgVal, t, o)
        assert np.array_equal(g, o) and not np.array_equal(g, s)
    d.msorit = 200
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P, api.F_D):
            ctx.upload(w, z)
        t0 = d.new_field()
        t0[:d.ny + 2, :d.nx + 2] = 0.5
        ctx.upload(api.F_T, t0)
        ctx.coldstart()
        ctx.step(2)
        tg = ctx.download(api.F_T)
        util, vbar, pav, tav = ctx.node_averages(temperature=True)
        ref = d.new_field()
        orc.taveraged(d.nx, d.ny, 0, r.nReg, r.nRegBrd, r.nTRgType, r.dTRgVal, tg, ref)
        assert np.array_equal(tav, ref)
        pre = str(tmp_path / "hot")
        assert plot3d.save_std_vars_p3d(pre, d.nx, d.ny, util, vbar, pav, t=tav) == 5
        _, _, pl = plot3d.read_std_vars_p3d(pre + ".qqq")
        assert np.array_equal(pl[4], ref[1:d.ny + 1, 1:d.nx + 1])


def test_thermal_off_is_the_cold_path(api, orc):
    """A context whose thermal switch was set and cleared again steps exactly like a cold one."""
    from wolfd2_b200 import deck as dk
    import dataclasses
    hot = dk.heated_cavity(40, re=100.0, dt=0.005, ny=32)
    hot.msorit = 200
    cold = dataclasses.replace(hot, thermal=False, eqstate=False)
    res = []
    for deck_, toggle in ((cold, False), (hot, True)):
        with api.Context(deck_) as ctx:
            if toggle:
                ctx.set_thermal(nthermen=0, neqstate=0)
            z = deck_.new_field()
            for w in (api.F_U, api.F_V, api.F_P):
                ctx.upload(w, z)
            ctx.coldstart()
            ctx.step(3)
            res.append([ctx.download(w) for w in (api.F_U, api.F_V, api.F_P)])
    for a, b in zip(*res):
        assert np.array_equal(a, b)


def test_buoyancy_without_the_energy_equation(api, orc):
    """`buoyancy` without `thermal_energy` (parse.f sets neqstate independently of nthermen): the reference still
    calls EqState every step (main.f:853), so d follows p at a frozen temperature and feeds the YMomentum buoyancy
    term through d and dn.  Steps against the oracle, d included."""
    from wolfd2_b200 import deck as dk
    import dataclasses
    d = dataclasses.replace(dk.heated_cavity(40, re=100.0, dt=0.005, ny=34), thermal=False, eqstate=True, nmeiter=1)
    d.msorit = 300
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(5)
    uo, vo, po, do = (d.new_field() for _ in range(4))
    to = d.new_field()
    yy, xx = np.mgrid[0:d.ny + 2, 0:d.nx + 2]
    to[:d.ny + 2, :d.nx + 2] = 0.5 + 0.3 * np.sin(0.2 * xx) * np.cos(0.15 * yy) + 0.01 * rng.standard_normal(xx.shape)
    t0 = to.copy()
    nso = orc.coldstart(d, uo, vo, po)
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P, api.F_D):
            ctx.upload(w, z)
        ctx.upload(api.F_T, t0)
        assert ctx.coldstart() == nso
        for step in range(4):
            lg = ctx.step(1)[0]
            rc, lo = orc.step(d, uo, vo, po, 1, t=to, d=do)
            assert rc == 0
            assert lg["nQLiter"] == lo[0]["nQLiter"] and lg["nSorConv"] == lo[0]["nSorConv"]
            got = [ctx.download(w) for w in (api.F_U, api.F_V, api.F_P, api.F_D)]
            errs = [rel_l2(g, o) for g, o in zip(got, (uo, vo, po, do))]
            assert max(errs) <= TOL_STEP, f"step {step}: rel-L2 (u,v,p,d) = {errs}"
        assert np.array_equal(ctx.download(api.F_T), t0)       # the temperature itself is frozen
    assert np.abs(do).max() > 0 and np.abs(vo[2:d.ny, 2:d.nx]).max() > 1e-6    # buoyancy did drive a flow
