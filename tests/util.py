"""Shared helpers for the parity tests (test infrastructure)."""
import numpy as np

from wolfd2_b200 import deck as dk
from wolfd2_b200._abi import METRIC_NAMES


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64); b = np.asarray(b, dtype=np.float64)
    den = np.sqrt(np.sum(b * b))
    num = np.sqrt(np.sum((a - b) ** 2))
    return num / den if den > 0 else num


def rand_field(deck, rng, lo=-1.0, hi=1.0):
    """uniform(lo,hi) on 0..nx+1 x 0..ny+1, zero elsewhere (as the reference's static arrays)."""
    f = deck.new_field()
    f[0:deck.ny + 2, 0:deck.nx + 2] = rng.uniform(lo, hi, size=(deck.ny + 2, deck.nx + 2))
    return f


def region_args(deck):
    r = deck.regions
    return r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp


def make_test_decks(nx=37, ny=29, **kw):
    """Small decks covering every boundary type and a blockage."""
    out = []
    out.append(dk.cavity(nx, re=100.0, dt=0.01, ny=ny, **kw))
    out.append(dk.channel(nx, re=100.0, dt=0.01, ny=ny, fully_dev=True, **kw))
    out.append(dk.channel(nx, re=100.0, dt=0.01, ny=ny, fully_dev=False, **kw))
    out.append(dk.backward_step(nx, re=100.0, dt=0.01, ny=ny, **kw))
    # every face type somewhere: no-stress walls, inlet from south, mass-cons outlet north, fully-dev west
    reg = dk.RegionTables(nx, ny, 2, 2, (nx // 2,), (ny // 2,))
    reg.wall(1, 1, "s", no_stress=True).wall(2, 1, "s", tangent_vel=0.5, press=0.1)
    reg.inlet(2, 1, "e", normal_vel=-0.7, tangent_vel=0.1).outlet(1, 1, "w", fully_dev=True, press=0.2)
    reg.outlet(1, 2, "w", fully_dev=False).outlet(1, 2, "n", fully_dev=True).outlet(2, 2, "n", fully_dev=False)
    reg.outlet(2, 2, "e", fully_dev=False)
    out.append(dk._mk("mixed", nx, ny, reg, 100.0, 0.01, **kw))
    # outlets on south / east with both types
    reg = dk.RegionTables(nx, ny, 2, 1, (nx // 2,), ())
    reg.outlet(1, 1, "s", fully_dev=True).outlet(2, 1, "s", fully_dev=False).inlet(1, 1, "n", normal_vel=-1.0)
    reg.outlet(2, 1, "e", fully_dev=True)
    out.append(dk._mk("south_out", nx, ny, reg, 100.0, 0.01, **kw))
    return out


def metric_list(deck, names):
    return [deck.metrics[n] for n in names]
