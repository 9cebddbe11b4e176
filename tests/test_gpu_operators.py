"""GPU parity of the operators below XMomentum / YMomentum / Ppe in the reference's call tree (SURVEY.md 8b,
"internal but worth exporting"): ConvCoef (all six cases), DConvU/V, DDiffU/V, PorosCoef, RhsPpe.  On the
production path these are evaluated per unknown inside the fused momentum kernels; the shims run the same device
functions one operator at a time, so every comparison here is BIT-EXACT against the oracle, including the cells
outside each routine's loop range (they must keep the caller's values)."""
import numpy as np
import pytest

from oracle import get_oracle
from util import rand_field, make_test_decks

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


@pytest.fixture(scope="module")
def orc():
    return get_oracle()


def _deck(nx=37, ny=29):
    return make_test_decks(nx, ny)[0]


def _cfg(api, orc, d):
    api.config(d.mnx, d.mny)
    orc.config(d.mnx, d.mny)


# metric arrays at the reference's call sites (momentum.f:282-290, :607-615, thermal.f:111-116)
CALL_SITES = {1: "xzn xec yzn yec", 2: "xzc xen yzc yen", 3: "xzv xeu yzv yeu",
              4: "xzu xeu yzu yeu", 5: "xzv xev yzv yev", 6: "xzc xec yzc yec"}


@pytest.mark.parametrize("size", [(37, 29), (64, 64), (257, 130)])
@pytest.mark.parametrize("ncomp", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("njacob", [0, 1])
def test_convcoef_bitwise(api, orc, size, ncomp, njacob):
    d = _deck(*size)
    _cfg(api, orc, d)
    rng = np.random.default_rng(12345 + 10 * ncomp + njacob)
    ms = [d.metrics[n] for n in CALL_SITES[ncomp].split()]
    u, v = rand_field(d, rng), rand_field(d, rng)
    c1, c2 = rand_field(d, rng), rand_field(d, rng)   # sentinel values outside the written ranges
    g1, g2, o1, o2 = c1.copy(), c2.copy(), c1.copy(), c2.copy()
    api.ConvCoef(d.nx, d.ny, ncomp, njacob, *ms, u, v, g1, g2)
    orc.convcoef(d.nx, d.ny, ncomp, njacob, *ms, u, v, o1, o2)
    assert np.array_equal(g1, o1) and np.array_equal(g2, o2)
    assert not np.array_equal(g1, c1)


@pytest.mark.parametrize("size", [(37, 29), (257, 130)])
def test_convcoef_with_foreign_metrics(api, orc, size):
    """Any four arrays may be passed (the routine does not know which metric set it gets)."""
    d = _deck(*size)
    _cfg(api, orc, d)
    rng = np.random.default_rng(99)
    ms = [rand_field(d, rng) for _ in range(4)]
    u, v = rand_field(d, rng), rand_field(d, rng)
    for ncomp in range(1, 7):
        g1, g2, o1, o2 = (d.new_field() for _ in range(4))
        api.ConvCoef(d.nx, d.ny, ncomp, 1 if ncomp in (4, 5) else 0, *ms, u, v, g1, g2)
        orc.convcoef(d.nx, d.ny, ncomp, 1 if ncomp in (4, 5) else 0, *ms, u, v, o1, o2)
        assert np.array_equal(g1, o1) and np.array_equal(g2, o2), ncomp


@pytest.mark.parametrize("size", [(37, 29), (64, 64), (257, 130)])
def test_dconv_ddiff_bitwise(api, orc, size):
    d = _deck(*size)
    _cfg(api, orc, d)
    rng = np.random.default_rng(4242)
    m = d.metrics
    q, c1, c2, s = (rand_field(d, rng) for _ in range(4))
    for gf, of in ((api.DConvU, orc.dconvu), (api.DConvV, orc.dconvv)):
        g, o = s.copy(), s.copy()
        gf(d.nx, d.ny, c1, c2, q, g)
        of(d.nx, d.ny, c1, c2, q, o)
        assert np.array_equal(g, o) and not np.array_equal(g, s)
    for gf, of, names in ((api.DDiffU, orc.ddiffu, "rac rbc rbn rgn"), (api.DDiffV, orc.ddiffv, "ran rbc rbn rgc")):
        ms = [m[n] for n in names.split()]
        g, o = s.copy(), s.copy()
        gf(d.nx, d.ny, *ms, q, g)
        of(d.nx, d.ny, *ms, q, o)
        assert np.array_equal(g, o) and not np.array_equal(g, s)


def test_convection_operator_composes_like_xmomentum(api, orc):
    """ConvCoef case 1 followed by DConvU is the convective term of XMomentum's right-hand side (momentum.f:285-330)."""
    d = _deck(64, 64)
    _cfg(api, orc, d)
    rng = np.random.default_rng(5)
    ms = [d.metrics[n] for n in CALL_SITES[1].split()]
    u, v = rand_field(d, rng), rand_field(d, rng)
    res = []
    for cc, dc in ((api.ConvCoef, api.DConvU), (orc.convcoef, orc.dconvu)):
        c1, c2, c = d.new_field(), d.new_field(), d.new_field()
        cc(d.nx, d.ny, 1, 0, *ms, u, v, c1, c2)
        dc(d.nx, d.ny, c1, c2, u, c)
        res.append(c)
    assert np.array_equal(res[0], res[1])


def _porous_decks():
    from wolfd2_b200 import deck as dk
    out = []
    reg = dk.RegionTables(40, 32, 2, 1, (18,), ()).porous(2, 1, 0.7, 5.0, 2.0)
    out.append(dk._mk("porous_2x1", 40, 32, reg, 100.0, 0.005))
    reg = dk.RegionTables(44, 40, 2, 2, (22,), (20,))
    reg.porous(1, 1, 0.8, 3.0, 1.0).porous(2, 1, 0.5, 6.0, 2.5).porous(2, 2, 0.9, 1.0, 0.5)
    out.append(dk._mk("porous_2x2", 44, 40, reg, 100.0, 0.005))
    reg = dk.RegionTables(37, 29, 2, 2, (18,), (14,))
    reg.blockage(1, 1)
    out.append(dk._mk("no_porous_2x2", 37, 29, reg, 100.0, 0.005))
    return out


@pytest.mark.parametrize("k", range(3))
@pytest.mark.parametrize("ncomp", [1, 2])
@pytest.mark.parametrize("njacob", [0, 1])
def test_poroscoef_bitwise(api, orc, k, ncomp, njacob):
    d = _porous_decks()[k]
    _cfg(api, orc, d)
    rng = np.random.default_rng(31 + k)
    r = d.regions
    u, v = rand_field(d, rng, -0.5, 0.5), rand_field(d, rng, -0.5, 0.5)
    u[3:6, 3:9] = 0.0
    v[2:7, 2:10] = 0.0    # |velocity| below the 1e-8 threshold of :1196 somewhere
    s = rand_field(d, rng)
    g, o = s.copy(), s.copy()
    api.PorosCoef(d.nx, d.ny, ncomp, njacob, r.nReg, r.nRegBrd, r.nRegType, r.dPRporos, r.dPRporc1, r.dPRporc2, u, v, g)
    orc.poroscoef(d.nx, d.ny, ncomp, njacob, r.nReg, r.nRegBrd, r.nRegType, r.dPRporos, r.dPRporc1, r.dPRporc2, u, v, o)
    assert np.array_equal(g, o) and not np.array_equal(g, s)


@pytest.mark.parametrize("size", [(37, 29), (64, 64), (257, 130)])
@pytest.mark.parametrize("cartes", [1, 0])
def test_rhsppe_bitwise(api, orc, size, cartes):
    d = _deck(*size)
    _cfg(api, orc, d)
    rng = np.random.default_rng(2024)
    m = d.metrics
    div, p = rand_field(d, rng), rand_field(d, rng)
    rbu, rbv = (rand_field(d, rng), rand_field(d, rng)) if not cartes else (m["rbu"], m["rbv"])
    n = (d.nx - 1) * (d.ny - 1)
    s = rng.uniform(-1, 1, size=d.mnx * d.mny)
    g, o = s.copy(), s.copy()
    api.RhsPpe(d.nx, d.ny, cartes, d.dk, rbu, rbv, div, p, g)
    orc.rhsppe(d.nx, d.ny, cartes, d.dk, rbu, rbv, div, p, o)
    assert np.array_equal(g, o)
    assert np.array_equal(g[n:], s[n:]) and not np.array_equal(g[:n], s[:n])
