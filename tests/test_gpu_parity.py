"""GPU parity tests: every routine of SURVEY.md §8a through the C ABI (libwolfd2_b200.so) against the
CPU oracle on the same seeded inputs.  Bit-exact for the streaming stencils, ghost fills, SOR and
norms (the library is built with -fmad=false and keeps the reference's operation order); the
momentum solves use a parallel elimination order and are held to rel-L2 <= 1e-12 per call and
<= 1e-10 per step (north-star tolerance)."""
import numpy as np
import pytest

from oracle import get_oracle
from util import rand_field, region_args, rel_l2, make_test_decks

pytestmark = pytest.mark.gpu

TOL_SOLVE = 1e-12   # one momentum component solve, relative L2
TOL_STEP = 1e-10    # u, v, p after each time step, relative L2 (BASELINE.json north_star)


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


@pytest.fixture(scope="module")
def orc():
    return get_oracle()


DECKS = make_test_decks() + make_test_decks(64, 64)[:2]
IDS = [f"{d.name}-{d.nx}x{d.ny}" for d in DECKS]


def _cfg(api, orc, d):
    api.config(d.mnx, d.mny)
    orc.config(d.mnx, d.mny)


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_ghost_fills_bitwise(api, orc, d):
    _cfg(api, orc, d)
    rng = np.random.default_rng(12345)
    r = d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for gf, of in ((api.VelBoundCond, orc.velboundcond), (api.VelOutflowBCs, orc.veloutflowbcs)):
        ug, vg, uo, vo = u.copy(), v.copy(), u.copy(), v.copy()
        gf(d.nx, d.ny, r.nReg, r.nRegBrd, r.nMomBdTp, r.dBCVal, ug, vg)
        of(d.nx, d.ny, r.nReg, r.nRegBrd, r.nMomBdTp, r.dBCVal, uo, vo)
        assert np.array_equal(ug, uo) and np.array_equal(vg, vo)
    pg, po = p.copy(), p.copy()
    api.PresBoundCond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.dBCVal, pg)
    orc.presboundcond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.dBCVal, po)
    assert np.array_equal(pg, po)


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_divergence_project_filter_norms_bitwise(api, orc, d):
    _cfg(api, orc, d)
    rng = np.random.default_rng(777)
    r, m = d.regions, d.metrics
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for nloc, names in ((1, "xeu yeu xzv yzv"), (2, "xev yev xzu yzu")):
        ms = [m[n] for n in names.split()]
        dg, do = d.new_field(), d.new_field()
        api.Divergence(d.nx, d.ny, nloc, *ms, u, v, dg)
        orc.divergence(d.nx, d.ny, nloc, *ms, u, v, do)
        assert np.array_equal(dg, do)
    pm = [m[n] for n in "dju djv yeu xzv yzu xev".split()]
    ug, vg, uo, vo = u.copy(), v.copy(), u.copy(), v.copy()
    api.Project(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, *pm, p, ug, vg)
    orc.project(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, *pm, p, uo, vo)
    assert np.array_equal(ug, uo) and np.array_equal(vg, vo)
    ntr = np.zeros(200, np.int32)
    for comp in (1, 2):
        qg, qo = u.copy(), u.copy()
        api.Filter(d.nx, d.ny, comp, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, ntr, 5.0, qg)
        orc.filter(d.nx, d.ny, comp, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, ntr, 5.0, qo)
        assert np.array_equal(qg, qo)
    assert api.DiffMaxNorm(d.nx, d.ny, u, v) == orc.diffmaxnorm(d.nx, d.ny, u, v)
    assert api.DMaxNorm(d.nx, d.ny, u) == orc.dmaxnorm(d.nx, d.ny, u)
    # DMaxNorm's odd seed (utility.f:493): a spike at (5,5) must be seen, one at (1,1) must not
    w = d.new_field(); w[5, 5] = -7.0; w[1, 1] = 9.0
    assert api.DMaxNorm(d.nx, d.ny, w) == 7.0 == orc.dmaxnorm(d.nx, d.ny, w)


@pytest.mark.parametrize("d", DECKS, ids=IDS)
def test_momentum_components(api, orc, d):
    _cfg(api, orc, d)
    rng = np.random.default_rng(4242)
    r, m = d.regions, d.metrics
    us, vs, un, vn = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    xm = [m[n] for n in "rbn rgn rac rbc dju xec yec xzn yzn xeu yeu xzu yzu".split()]
    dg, do = d.new_field(), d.new_field()
    api.XMomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2, *xm, us, vs, un, vn, dg)
    orc.xmomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2, *xm, us, vs, un, vn, do)
    assert rel_l2(dg, do) <= TOL_SOLVE
    assert np.array_equal(dg == 0.0, do == 0.0)        # same sparsity: identity rows, untouched ghosts
    ym = [m[n] for n in "ran rbn rbc rgc djv xen yen xzc yzc xev yev xzv yzv".split()]
    dd, dn = rand_field(d, rng, 0.0, 0.01), rand_field(d, rng, 0.0, 0.01)   # exercise the buoyancy term
    dg, do = d.new_field(), d.new_field()
    api.YMomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, dd, dn, us, vs, un, vn, dg)
    orc.ymomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, dd, dn, us, vs, un, vn, do)
    assert rel_l2(dg, do) <= TOL_SOLVE
    assert np.array_equal(dg == 0.0, do == 0.0)


@pytest.mark.parametrize("d", DECKS[:4], ids=IDS[:4])
def test_nauxmomentum(api, orc, d):
    _cfg(api, orc, d)
    rng = np.random.default_rng(99)
    r, m = d.regions, d.metrics
    un, vn = rand_field(d, rng, -0.3, 0.3), rand_field(d, rng, -0.3, 0.3)
    names = ("ran rbn rgn rac rbc rgc dju djv xec yec xzn yzn xen yen xzc yzc "
             "xeu yeu xzu yzu xev yev xzv yzv").split()
    mm = [m[n] for n in names]
    z = d.new_field()
    out = []
    for f in (api.nAuxMomentum, orc.nauxmomentum):
        us, vs = rand_field(d, np.random.default_rng(5)), rand_field(d, np.random.default_rng(6))
        n = f(d.nx, d.ny, 6, *region_args(d), d.dk, d.re, d.fr, 1e-4, r.dPRporos, r.dPRporc1, r.dPRporc2,
              r.dBCVal, *mm, z, z, un, vn, us, vs)
        out.append((n, us, vs))
    (ng, ug, vg), (no, uo, vo) = out
    assert ng == no
    assert rel_l2(ug, uo) <= 1e-11 and rel_l2(vg, vo) <= 1e-11


@pytest.mark.parametrize("n", [7, 64, 4095, 4096, 4097, 70001, 1000003])
def test_alttridlu(api, orc, n):
    """AltTridLU incl. the first-row quirk (momentum.f:1319), one chain of n unknowns; sizes straddle
    the 4096-unknown segment and exercise 1, 2 and 3 levels of the partitioned solver."""
    api.config(1200, 1200); orc.config(1200, 1200)
    rng = np.random.default_rng(n)
    a = np.zeros((n, 3))
    a[:, 0] = rng.uniform(-1, 1, n); a[:, 2] = rng.uniform(-1, 1, n)
    a[:, 1] = 2.2 + rng.uniform(0, 1, n)
    b = rng.uniform(-1, 1, n)
    ag, bg, ao, bo = a.copy().reshape(-1), b.copy(), a.copy().reshape(-1), b.copy()
    api.AltTridLU(n, ag, bg)
    orc.alttridlu(n, ao, bo)
    assert rel_l2(bg, bo) <= 1e-13
    # slowly decaying coupling (|off-diagonals| close to the diagonal): spikes reach far
    a[:, 0] = -1.0; a[:, 2] = -1.0; a[:, 1] = 2.0 + 1e-4
    ag, bg, ao, bo = a.copy().reshape(-1), b.copy(), a.copy().reshape(-1), b.copy()
    api.AltTridLU(n, ag, bg)
    orc.alttridlu(n, ao, bo)
    assert rel_l2(bg, bo) <= 1e-9   # condition number ~ 4e4: both solves carry ~1e-12 relative error


@pytest.mark.parametrize("d", DECKS, ids=IDS)
@pytest.mark.parametrize("cart", [1, 0])
@pytest.mark.parametrize("solver", [5, 6])
def test_ppe_bitwise(api, orc, d, cart, solver):
    """Ppe with the red/black point SOR: identical iterate path => identical iteration count and p."""
    _cfg(api, orc, d)
    rng = np.random.default_rng(31337)
    r, m = d.regions, d.metrics
    u, v, p = rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01)
    pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
    for msorit, tol in ((3000, 1e-8), (25, 1e-8), (60, 0.0)):
        pg, po = p.copy(), p.copy()
        ng = api.Ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, solver, msorit, d.dk, tol, 1.4, *pm8, u, v, pg)
        no = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, solver, msorit, d.dk, tol, 1.4, *pm8, u, v, po)
        assert ng == no
        assert np.array_equal(pg, po)
    assert ng == 60  # tolerance 0 never converges: nSorConv = msorit (pressure.f:242-246)


STEP_DECKS = None


def _step_decks():
    from wolfd2_b200 import deck as dk
    ds = [dk.cavity(64, re=100.0, dt=0.01), dk.cavity(37, re=400.0, dt=0.02, ny=29),
          dk.channel(48, re=100.0, dt=0.005, ny=40), dk.channel(48, re=100.0, dt=0.005, ny=40, fully_dev=False),
          dk.backward_step(64, re=100.0, dt=0.005, ny=48), dk.cavity(130, re=1000.0, dt=0.004, ny=66)]
    for d in ds:
        d.msorit = 300
    return ds


@pytest.mark.parametrize("k", range(6))
def test_time_steps(api, orc, k):
    """Cold start + 8 time steps: per-step rel-L2 of u, v, p, identical QL / SOR iteration counts and
    identical PrintDiff tuple to 1e-9; drift over the run reported."""
    d = _step_decks()[k]
    orc.config(d.mnx, d.mny)
    uo, vo, po = d.new_field(), d.new_field(), d.new_field()
    nso = orc.coldstart(d, uo, vo, po)
    worst = 0.0
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P):
            ctx.upload(w, z)
        assert ctx.coldstart() == nso
        for step in range(8):
            lg = ctx.step(1)[0]
            rc, lo = orc.step(d, uo, vo, po, 1)
            assert rc == 0
            assert lg["nQLiter"] == lo[0]["nQLiter"] and lg["nSorConv"] == lo[0]["nSorConv"]
            ug, vg, pg = ctx.download(api.F_U), ctx.download(api.F_V), ctx.download(api.F_P)
            errs = [rel_l2(ug, uo), rel_l2(vg, vo), rel_l2(pg, po)]
            worst = max(worst, *errs)
            assert max(errs) <= TOL_STEP, f"{d.name} step {step}: rel-L2 (u,v,p) = {errs}"
            np.testing.assert_allclose(lg["dif"][:3], lo[0]["dif"][:3], rtol=1e-9, atol=1e-14)
    print(f"{d.name}: worst rel-L2 over 8 steps = {worst:.2e}")


RAGGED = [(6, 6), (7, 131), (131, 7), (1025, 9), (9, 1030), (129, 33), (2049, 6)]


@pytest.mark.parametrize("size", RAGGED, ids=[f"{a}x{b}" for a, b in RAGGED])
@pytest.mark.parametrize("kind", ["cavity", "channel_mc"])
def test_ragged_and_minimal_grids(api, orc, size, kind):
    """Smallest legal grid (6x6, CheckGridSize), one-cell-wide extremes and row lengths around the solver's
    thread-block stride (128) and segment size (1024): chain segments then cover many grid rows or straddle
    them at every offset, the chain walker takes its multi-row stride, the last segment is mostly padding."""
    from wolfd2_b200 import deck as dk
    nx, ny = size
    h = 1.0 / (max(nx, ny) - 1)
    if kind == "cavity":
        d = dk.cavity(nx, re=100.0, dt=min(0.01, 20.0 * h * h), ny=ny)
    else:
        d = dk.channel(nx, re=100.0, dt=min(0.01, 20.0 * h * h), ny=ny, fully_dev=False)
    d.msorit = 60
    orc.config(d.mnx, d.mny)
    uo, vo, po = d.new_field(), d.new_field(), d.new_field()
    nso = orc.coldstart(d, uo, vo, po)
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P):
            ctx.upload(w, z)
        assert ctx.coldstart() == nso
        for step in range(3):
            lg = ctx.step(1)[0]
            rc, lo = orc.step(d, uo, vo, po, 1)
            assert rc == 0
            assert lg["nQLiter"] == lo[0]["nQLiter"] and lg["nSorConv"] == lo[0]["nSorConv"]
            for w, ref in ((api.F_U, uo), (api.F_V, vo), (api.F_P, po)):
                assert rel_l2(ctx.download(w), ref) <= TOL_STEP, (size, kind, step, w)


def test_step_host_equals_resident(api):
    """The host-buffer entry point (e2e path) gives the same fields as the resident one."""
    from wolfd2_b200 import deck as dk
    d = dk.channel(48, re=100.0, dt=0.005, ny=40)
    d.msorit = 100
    with api.Context(d) as c1, api.Context(d) as c2:
        z = d.new_field()
        for c in (c1, c2):
            for w in (api.F_U, api.F_V, api.F_P):
                c.upload(w, z)
            c.coldstart()
        c1.step(3)
        u, v, p = c2.download(api.F_U), c2.download(api.F_V), c2.download(api.F_P)
        for _ in range(3):
            c2.step_host(u, v, p, 1)
        assert np.array_equal(u, c1.download(api.F_U)) and np.array_equal(p, c1.download(api.F_P))


@pytest.mark.parametrize("d", DECKS[:5], ids=IDS[:5])
@pytest.mark.parametrize("cart", [1, 0])
@pytest.mark.parametrize("solver", [1, 2, 3, 4])
def test_ppe_other_solvers(api, orc, d, cart, solver):
    """ids 1-4.  Lexicographic point SOR (1) is bit-exact (same data dependences, same arithmetic); the
    line solvers (2-4) differ from the oracle only by the elimination order of each line solve."""
    _cfg(api, orc, d)
    rng = np.random.default_rng(555)
    r, m = d.regions, d.metrics
    u, v, p = rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01)
    pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
    for msorit, tol in ((400, 1e-8), (9, 0.0)):
        pg, po = p.copy(), p.copy()
        ng = api.Ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, solver, msorit, d.dk, tol, 1.3, *pm8, u, v, pg)
        no = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, solver, msorit, d.dk, tol, 1.3, *pm8, u, v, po)
        if solver == 1:
            assert ng == no and np.array_equal(pg, po)
        else:
            assert abs(ng - no) <= 1          # a near-threshold exit may move by one iteration
            if ng != no:                      # then hold the iterate path itself to the tolerance: the oracle's
                po = p.copy()                 # field after exactly as many iterations as the device ran
                n2 = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, solver, ng, d.dk, 0.0, 1.3, *pm8, u, v, po)
                assert n2 == ng
            assert rel_l2(pg, po) <= 1e-11, (solver, ng, no)


def test_default_solver_runs_whole_steps(api, orc):
    """The reference's default deck (ppe_solver sor, id 1) through the step driver."""
    from wolfd2_b200 import deck as dk
    d = dk.cavity(40, re=100.0, dt=0.01, ny=36)
    d.ppe_solver, d.msorit, d.sorrel = "sor", 500, 1.5
    orc.config(d.mnx, d.mny)
    uo, vo, po = d.new_field(), d.new_field(), d.new_field()
    orc.coldstart(d, uo, vo, po)
    rc, lo = orc.step(d, uo, vo, po, 3)
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P):
            ctx.upload(w, z)
        ctx.coldstart()
        lg = ctx.step(3)
        for g, o_ in zip(lg, lo):
            assert g["nQLiter"] == o_["nQLiter"] and g["nSorConv"] == o_["nSorConv"]
        assert rel_l2(ctx.download(api.F_P), po) <= TOL_STEP and rel_l2(ctx.download(api.F_U), uo) <= TOL_STEP


def _porous_decks():
    from wolfd2_b200 import deck as dk
    out = []
    reg = dk.RegionTables(40, 32, 2, 1, (18,), ()).porous(2, 1, 0.7, 5.0, 2.0)
    reg.wall(1, 1, "n", tangent_vel=1.0).wall(2, 1, "n", tangent_vel=1.0)
    out.append(dk._mk("porous_2x1", 40, 32, reg, 100.0, 0.005))
    reg = dk.RegionTables(36, 40, 1, 2, (), (20,)).porous(1, 2, 0.6, 8.0, 1.5)
    reg.inlet(1, 1, "w", normal_vel=1.0).inlet(1, 2, "w", normal_vel=1.0).outlet(1, 1, "e").outlet(1, 2, "e")
    out.append(dk._mk("porous_1x2_channel", 36, 40, reg, 100.0, 0.005))
    # two porous regions sharing borders: points on the shared border are divided by both porosities
    reg = dk.RegionTables(44, 40, 2, 2, (22,), (20,))
    reg.porous(1, 1, 0.8, 3.0, 1.0).porous(2, 1, 0.5, 6.0, 2.5).porous(2, 2, 0.9, 1.0, 0.5)
    reg.wall(1, 2, "n", tangent_vel=1.0).wall(2, 2, "n", tangent_vel=-0.5)
    out.append(dk._mk("porous_2x2", 44, 40, reg, 100.0, 0.005))
    return out


@pytest.mark.parametrize("k", range(3))
def test_porous_regions(api, orc, k):
    """PorosCoef + porosity scaling of the convective coefficients (momentum.f:296-343, 1115-1226)."""
    d = _porous_decks()[k]
    _cfg(api, orc, d)
    rng = np.random.default_rng(77 + k)
    r, m = d.regions, d.metrics
    us, vs, un, vn = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    xm = [m[n] for n in "rbn rgn rac rbc dju xec yec xzn yzn xeu yeu xzu yzu".split()]
    dg, do = d.new_field(), d.new_field()
    api.XMomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2, *xm, us, vs, un, vn, dg)
    orc.xmomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2, *xm, us, vs, un, vn, do)
    assert rel_l2(dg, do) <= TOL_SOLVE
    ym = [m[n] for n in "ran rbn rbc rgc djv xen yen xzc yzc xev yev xzv yzv".split()]
    z = d.new_field()
    dg, do = d.new_field(), d.new_field()
    api.YMomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, z, z, us, vs, un, vn, dg)
    orc.ymomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, z, z, us, vs, un, vn, do)
    assert rel_l2(dg, do) <= TOL_SOLVE
    # and a few whole steps
    d.msorit, d.sorrel = 300, 1.6
    uo, vo, po = d.new_field(), d.new_field(), d.new_field()
    nso = orc.coldstart(d, uo, vo, po)
    with api.Context(d) as ctx:
        zz = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P):
            ctx.upload(w, zz)
        assert ctx.coldstart() == nso
        for step in range(4):
            lg = ctx.step(1)[0]
            rc, lo = orc.step(d, uo, vo, po, 1)
            assert lg["nQLiter"] == lo[0]["nQLiter"] and lg["nSorConv"] == lo[0]["nSorConv"]
            for w, ref in ((api.F_U, uo), (api.F_V, vo), (api.F_P, po)):
                assert rel_l2(ctx.download(w), ref) <= TOL_STEP


def _fused_decks():
    from wolfd2_b200 import deck as dk
    return [dk.cavity(300, re=100.0, dt=0.001, ny=40), dk.channel(600, re=100.0, dt=0.001, ny=90),
            dk.backward_step(520, re=100.0, dt=0.001, ny=64), dk.cavity(254, re=100.0, dt=0.001, ny=8),
            dk.cavity(1030, re=100.0, dt=0.001, ny=37)]


@pytest.mark.parametrize("k", range(5))
@pytest.mark.parametrize("T", [0, 1, 2])
def test_ppe_fused_pipeline_bitwise(api, orc, k, T):
    """The fused red/black pipeline (several strips, several bands, blockage sentinel, T iterations per
    pass incl. convergence in the middle of a pass) must reproduce SorRB bit for bit."""
    d = _fused_decks()[k]
    _cfg(api, orc, d)
    api.set_option("sor_fused_T", T)
    try:
        rng = np.random.default_rng(2024 + k)
        r, m = d.regions, d.metrics
        u, v, p = rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01)
        pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
        seen = set()
        for msorit, tol in ((40, 1e-3), (41, 2e-3), (400, 3e-4), (401, 1e-4), (7, 0.0), (8, 0.0), (1, 1.0), (2, 1.0), (3, 1.0)):
            pg, po = p.copy(), p.copy()
            ng = api.Ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 5, msorit, d.dk, tol, 1.7, *pm8, u, v, pg)
            no = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 5, msorit, d.dk, tol, 1.7, *pm8, u, v, po)
            assert ng == no, (msorit, tol)
            assert np.array_equal(pg, po), (msorit, tol)
            seen.add(no % 2)
        assert seen == {0, 1}   # both odd and even convergence points were exercised
    finally:
        api.set_option("sor_fused_T", -1)


@pytest.mark.parametrize("name", ["cavity24x20", "channel22x18_fd", "channel22x18_mc", "bstep26x20", "heated_cavity26x22"])
def test_time_steps_against_committed_golden(api, name):
    """CUDA path vs the committed fixtures (tests/golden/*.npz, made by make_golden.py) -- no oracle
    involved at run time."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden
    d = make_golden.cases()[name]
    ref = np.load(os.path.join(here, "golden", name + ".npz"))
    thermal = "t0" in ref
    fields = [("u", api.F_U), ("v", api.F_V), ("p", api.F_P)] + ([("t", api.F_T), ("d", api.F_D)] if thermal else [])
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P):
            ctx.upload(w, z)
        if thermal:
            t0 = d.new_field()
            t0[:d.ny + 2, :d.nx + 2] = 0.5
            ctx.upload(api.F_T, t0)
        assert ctx.coldstart() == int(ref["ncold"])
        for k in range(4):
            lg = ctx.step(1)[0]
            assert lg["nQLiter"] == int(ref["nql"][k]) and lg["nSorConv"] == int(ref["nsor"][k])
            for f, w in fields:
                assert rel_l2(ctx.download(w), ref[f"{f}{k}"]) <= TOL_STEP, (name, f, k)
            n = ref["dif"].shape[1]
            np.testing.assert_allclose(lg["dif"][:n], ref["dif"][k], rtol=1e-9, atol=1e-14)
