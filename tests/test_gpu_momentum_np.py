"""The QL loop keeps the time-level-n halves of the explicit momentum terms (cnvn, difn: momentum.f:327-330, :652-655)
across its iterations and, on the first iteration of a deck without outlets, takes them from the starred fields, which
are then bitwise copies of un, vn.  Both shortcuts must give THE SAME BITS as evaluating them from un, vn every time
(option mom_np_cache 0), on every deck family, including those where the shortcut must not be taken."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api():
    from wolfd2_b200 import api as a
    a.lib()
    return a


def _decks():
    from wolfd2_b200 import deck as dk
    from util import make_test_decks
    out = list(make_test_decks(70, 45))
    out.append(dk.cavity(300, re=400.0, dt=2e-3, ny=130))
    out.append(dk.channel(1100, re=100.0, dt=2e-4, ny=40, fully_dev=False))
    return out


@pytest.mark.parametrize("k", range(9))
def test_np_cache_is_bit_identical(api, k):
    decks = _decks()
    if k >= len(decks):
        pytest.skip("no such deck")
    d = decks[k]
    d.msorit, d.mqiter, d.qtol = 60, 5, 1e-9          # several QL iterations per step
    rng = np.random.default_rng(11 + k)
    f0 = [d.new_field() for _ in range(3)]
    for f in f0[:2]:
        f[:d.ny + 2, :d.nx + 2] = 0.002 * rng.standard_normal((d.ny + 2, d.nx + 2))
    res = []
    for cache in (0, 1):
        api.set_option("mom_np_cache", cache)
        try:
            with api.Context(d) as ctx:
                for w, f in zip((api.F_U, api.F_V, api.F_P), f0):
                    ctx.upload(w, f)
                ctx.coldstart()
                lg = ctx.step(3)
                res.append((lg, [ctx.download(w) for w in (api.F_U, api.F_V, api.F_P)]))
        finally:
            api.set_option("mom_np_cache", 1)
    (l0, f_off), (l1, f_on) = res
    assert [(a["nQLiter"], a["nSorConv"], a["dif"]) for a in l0] == [(a["nQLiter"], a["nSorConv"], a["dif"]) for a in l1]
    assert any(a["nQLiter"] == -1 or a["nQLiter"] > 1 for a in l1)        # several QL iterations: the cached path did run
    for a, b in zip(f_off, f_on):
        assert np.array_equal(a, b, equal_nan=True), d.name     # (the mixed-face deck may blow up from this start: then both do)
    assert any(np.isfinite(a).all() for a in f_on) or d.name == "mixed"


def _cart_decks():
    from wolfd2_b200 import deck as dk
    from util import make_test_decks
    out = list(make_test_decks(70, 45))[:5]
    out.append(dk.cavity(300, re=400.0, dt=2e-3, ny=130))
    out.append(dk.channel(1100, re=100.0, dt=2e-4, ny=40, fully_dev=False))
    # a skewed, stretched grid: the verification of the metric classes must reject it (the run is then the general variant
    # twice, trivially identical -- the point is that nothing breaks and nothing is mis-classified)
    x, y = dk.stretched_grid(60, 44)
    reg = dk.RegionTables(60, 44).wall(1, 1, "n", tangent_vel=1.0)
    out.append(dk._mk("cavity_stretched", 60, 44, reg, 100.0, 0.004, x=x, y=y, cartesian=False))
    return out


@pytest.mark.parametrize("k", range(8))
def test_cartesian_metric_variant_is_bit_identical(api, k):
    """On a Cartesian grid 14 of the 24 metric arrays of a QL iteration are bitwise one-dimensional or constant; the
    momentum kernels then read them as such (option mom_cart, checked on the uploaded arrays).  Same bits as reading
    the 2-D arrays, on every deck family."""
    d = _cart_decks()[k]
    d.msorit, d.mqiter, d.qtol = 60, 4, 1e-9
    rng = np.random.default_rng(23 + k)
    f0 = [d.new_field() for _ in range(3)]
    for f in f0[:2]:
        f[:d.ny + 2, :d.nx + 2] = 0.002 * rng.standard_normal((d.ny + 2, d.nx + 2))
    res = []
    for cart in (0, 1):
        api.set_option("mom_cart", cart)
        try:
            with api.Context(d) as ctx:
                for w, f in zip((api.F_U, api.F_V, api.F_P), f0):
                    ctx.upload(w, f)
                ctx.coldstart()
                lg = ctx.step(3)
                res.append((lg, [ctx.download(w) for w in (api.F_U, api.F_V, api.F_P)]))
        finally:
            api.set_option("mom_cart", 1)
    (l0, f_off), (l1, f_on) = res
    assert [(a["nQLiter"], a["nSorConv"], a["dif"]) for a in l0] == [(a["nQLiter"], a["nSorConv"], a["dif"]) for a in l1]
    for a, b in zip(f_off, f_on):
        assert np.array_equal(a, b, equal_nan=True), d.name


@pytest.mark.parametrize("k", [0, 1, 3, 5, 6])
def test_two_stream_momentum_is_bit_identical(api, k):
    """One GPU: XMomentum and YMomentum of a QL iteration are independent and run on two streams with their own work
    arrays (option mom_two_streams).  Same bits as one after the other on one stream -- from the very first step of a
    fresh context, when the second set of arrays is allocated and zero-filled on demand (a zero fill on the default
    stream once landed after the first kernels: w2_context.cu dalloc)."""
    d = _cart_decks()[k]
    d.msorit, d.mqiter, d.qtol = 60, 4, 1e-9
    rng = np.random.default_rng(41 + k)
    f0 = [d.new_field() for _ in range(3)]
    for f in f0[:2]:
        f[:d.ny + 2, :d.nx + 2] = 0.002 * rng.standard_normal((d.ny + 2, d.nx + 2))
    res = []
    for two in (0, 1, 1):                 # the two-stream run twice, each in a fresh context
        api.set_option("mom_two_streams", two)
        try:
            with api.Context(d) as ctx:
                for w, f in zip((api.F_U, api.F_V, api.F_P), f0):
                    ctx.upload(w, f)
                ctx.coldstart()
                lg = ctx.step(3)
                res.append((lg, [ctx.download(w) for w in (api.F_U, api.F_V, api.F_P)]))
        finally:
            api.set_option("mom_two_streams", 1)
    key = lambda lg: [(a["nQLiter"], a["nSorConv"], a["dif"]) for a in lg]
    for lg, fields in res[1:]:
        assert key(lg) == key(res[0][0])
        for a, b in zip(res[0][1], fields):
            assert np.array_equal(a, b, equal_nan=True), d.name
