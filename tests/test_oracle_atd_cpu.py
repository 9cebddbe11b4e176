"""CPU checks of the oracle's restatement of the optional rows: SmallScale / SmlSclBC,
Traject, VelAvg / PTDAvg.  The reference holds no fixtures for them (parity unpinned), so these are invariants,
hand-computed cases and the quirks listed in oracle/wolfd2_oracle_atd.inc."""
import numpy as np
import pytest

from oracle import Oracle, get_oracle
from util import rand_field
from wolfd2_b200 import deck as dk
from wolfd2_b200.deck import PPE_SOLVERS

SS_KW = dict(smallscale=True, ss_cu0=1.0, ss_bncrit=2.0, ss_rmpmax=0.95, ss_ppe_solver="rb_sor", ss_msorit=400,
             ss_sortol=1e-10, ss_sorrel=1.6)


def _ss_args(d, initflg):
    r, m = d.regions, d.metrics
    fp = np.array(d.ss_filt, dtype=np.float64)
    return (d.nx, d.ny, initflg, int(d.thermal), int(d.cartesian), r.nReg, r.nRegBrd, r.nRegType, r.nTRgType, r.nMomBdTp,
            r.nTemBdTp, PPE_SOLVERS[d.ss_ppe_solver], d.ss_msorit, d.dlref, d.uref, d.tref, d.tmax, d.dk, d.re, d.pe,
            d.ss_sortol, d.ss_sorrel, fp, d.ss_cu0, d.ss_tscoef, d.ss_hscoef, d.ss_temcoef, d.ss_bncrit, d.ss_rmpmax,
            d.ss_rmpexp, r.dTRgVal, r.dBCVal, m["rau"], m["rbu"], m["rbv"], m["rgv"], m["dju"], m["djv"], m["djc"],
            m["xeu"], m["yeu"], m["xzv"], m["yzv"], m["xzu"], m["yzu"], m["xev"], m["yev"], m["xec"], m["yec"],
            m["xzc"], m["yzc"])


def test_map_seeding_is_one_recurrence_through_the_grid():
    """small_scale.f:203-220: i outer, j inner, three iterates per cell, seeds 0.92 / 0.31 / 0.50 at r = rc."""
    o = get_oracle()
    d = dk.cavity(12, re=1000.0, dt=0.01, ny=9, **SS_KW)
    o.config(d.mnx, d.mny)
    z = d.new_field()
    out = [rand_field(d, np.random.default_rng(0)) for _ in range(4)]
    pss0 = out[2].copy()
    o.smallscale(*_ss_args(d, -1), z, z, z, *out)
    dAr, dAm, rc = 4.82842712474, 1.47839783948, 0.20710678119
    for fam, seed in enumerate((0.92, 0.31, 0.50)):
        m = seed
        planes = [o.ss_map(d, fam, l) for l in (1, 2, 3)]
        for i in range(0, d.nx + 2):
            for j in range(0, d.ny + 2):
                for l in range(3):
                    m = rc * dAr * m * (1.0 - dAm * abs(m))
                    assert planes[l][j, i] == m
    assert not out[0].any() and not out[1].any() and not out[3].any()     # uss, vss, tss zeroed on 0..nx+1, 0..ny+1
    assert np.array_equal(out[2], pss0)                                    # initflg < 0 returns before touching pss
    assert o.lib.orc_ss_tarea() == pytest.approx(d.nx * d.ny / ((d.nx - 1) * (d.ny - 1)), rel=1e-12)


def test_smallscale_zero_cu0_changes_nothing_but_solves():
    """cu0 = 0: nmap = 0 (the maps stay), the amplitude factors vanish, so uss = vss = tss = 0 inside."""
    o = get_oracle()
    d = dk.cavity(24, re=1000.0, dt=0.01, ny=20, **dict(SS_KW, ss_cu0=0.0))
    o.config(d.mnx, d.mny)
    rng = np.random.default_rng(3)
    u, v, t = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng, 0.0, 1.0)
    out = [d.new_field() for _ in range(4)]
    o.smallscale(*_ss_args(d, 0), u, v, t, *out)
    m0 = [o.ss_map(d, f, l) for f in range(3) for l in (1, 2, 3)]
    o.smallscale(*_ss_args(d, 1), u, v, t, *out)
    m1 = [o.ss_map(d, f, l) for f in range(3) for l in (1, 2, 3)]
    for a, b in zip(m0, m1):
        assert np.array_equal(a, b)
    for a in out:
        assert not a.any()


def test_smallscale_output_is_discretely_divergence_free():
    """The model's own Ppe + Project (small_scale.f:552-577) leave uss, vss solenoidal to the SOR tolerance."""
    o = get_oracle()
    d = dk.cavity(30, re=1000.0, dt=0.01, ny=26, **SS_KW)
    o.config(d.mnx, d.mny)
    rng = np.random.default_rng(4)
    u, v, t = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng, 0.0, 1.0)
    out = [d.new_field() for _ in range(4)]
    o.smallscale(*_ss_args(d, 0), u, v, t, *out)
    uss, vss, pss, tss = out
    assert np.abs(uss).max() > 1e-3 and np.abs(vss).max() > 1e-3 and np.abs(tss).max() > 1e-6
    m = d.metrics
    div = d.new_field()
    o.divergence(d.nx, d.ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], uss, vss, div)
    # Away from the walls only: the pressure ghosts are frozen during the SOR sweeps (pressure.f:431) at the value
    # PresBoundCond gave them from pss = 0, so the wall-adjacent rows carry a flux through the wall face that
    # Project never applies -- the Neumann condition lags one solve, as in the large-scale step.
    inner = div[3:d.ny, 3:d.nx]
    scale = np.abs(uss).max() * (d.nx - 1)
    assert np.abs(inner).max() < 1e-9 * scale
    assert np.abs(div[2:d.ny + 1, 2:d.nx + 1]).max() > 1e-6 * scale      # ... and it is visible next to them
    # no-slip walls: normal component zero on the wall faces (bound_cond.f:1275-1277; Project leaves them alone)
    assert not uss[1:d.ny + 1, 1].any() and not uss[1:d.ny + 1, d.nx].any()


def test_smallscale_map_iterates_only_where_active_and_clamps_nmap():
    """Cells with peh <= 3 keep their maps; elsewhere nmap = int(1 + dk/ts) capped at 50 (small_scale.f:362-387)."""
    o = get_oracle()
    d = dk.cavity(20, re=1000.0, dt=0.5, ny=18, **SS_KW)     # a huge step: dk/ts >> 50 wherever the model is on
    o.config(d.mnx, d.mny)
    u, v, t = d.new_field(), d.new_field(), d.new_field()
    u[1:12, 1:d.nx + 1] = np.linspace(0.0, 3.0, 11)[:, None]    # shear in the lower half only
    out = [d.new_field() for _ in range(4)]
    o.smallscale(*_ss_args(d, -1), u, v, t, *out)
    before = o.ss_map(d, 0, 1)
    o.smallscale(*_ss_args(d, 1), u, v, t, *out)
    after = o.ss_map(d, 0, 1)
    changed = after != before
    assert changed[2:10, 2:d.nx].all()           # sheared cells iterate
    assert not changed[13:, :].any()             # quiescent cells (delu2n = 0, peh = 0) do not
    # 50 iterates of the map from the seeded value at the cell's own rmap stay inside the map's invariant interval
    assert np.all(np.abs(after) < 0.8166 + 1e-12)


def test_smlsclbc_faces_by_hand():
    o = get_oracle()
    reg = dk.RegionTables(16, 12, 1, 1)
    reg.wall(1, 1, "s", no_stress=True).outlet(1, 1, "e", fully_dev=True).inlet(1, 1, "w", normal_vel=1.0)
    d = dk._mk("ssbc", 16, 12, reg, 100.0, 0.01)
    o.config(d.mnx, d.mny)
    rng = np.random.default_rng(5)
    u, v, p, t = (rand_field(d, rng) for _ in range(4))
    u0, v0, t0 = u.copy(), v.copy(), t.copy()
    r = d.regions
    o.smlsclbc(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.nTRgType, r.nTemBdTp, r.dBCVal, u, v, p, t)
    nx, ny = d.nx, d.ny
    assert not u[1:ny + 1, 1].any()                                   # inlet: homogeneous for the fluctuation
    assert np.array_equal(v[2:ny + 1, 1], -v[2:ny + 1, 2])
    assert np.array_equal(u[1, 2:nx + 1], u[2, 2:nx + 1])             # no-stress south: even mirror
    assert np.array_equal(u[2:ny, nx + 1], u0[2:ny, nx + 1])          # east OUTLT1: u(iE+1,j) = u(iE+1,j), a no-op as written
    assert np.array_equal(v[2:ny, nx + 1], -v[2:ny, nx])             # (row ny: the north wall zeroes v(iE,jN) afterwards)
    # default thermal faces are adiabatic: even mirror, no boundary value
    assert np.array_equal(t[2:ny + 1, 1], t[2:ny + 1, 2]) and np.array_equal(t[ny + 1, 2:nx + 1], t[ny, 2:nx + 1])
    assert np.array_equal(t[2:ny, 2:nx], t0[2:ny, 2:nx])


def test_node_averages_blockage_order():
    """VelAvg zeroes a blockage INCLUDING its border nodes, later regions overwrite the shared ones; PTDAvg
    zeroes the blockage interior only (utility.f:548-552, 614-620)."""
    o = get_oracle()
    d = dk.backward_step(20, re=100.0, dt=0.01, ny=16)
    o.config(d.mnx, d.mny)
    r = d.regions
    u, v, p = d.new_field(), d.new_field(), d.new_field()
    u[:] = 2.0; v[:] = 3.0; p[:] = 5.0
    ut, vb, pa = d.new_field(), d.new_field(), d.new_field()
    o.velavg(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, u, v, ut, vb)
    o.ptdavg(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, p, pa)
    ib, jb = 5, 8                                     # borders of the blockage region (1,1): i = 1..5, j = 1..8
    assert not ut[1:jb, 1:ib].any()                   # inside and on the outer (W, S) borders: zero
    assert np.all(ut[1:jb + 1, ib] == 2.0) and np.all(ut[jb, 1:ib + 1] == 2.0)   # shared borders: the neighbours win
    assert np.all(vb[jb + 1:d.ny + 1, 1:d.nx + 1] == 3.0)
    assert not pa[2:jb, 2:ib].any() and np.all(pa[1, 1:ib] == 0.0)                # border nodes: never written by PTDAvg
    assert np.all(pa[jb:d.ny + 1, 1:d.nx + 1] == 5.0)


def _uniform_traj_case(n=7):
    d = dk.cavity(21, re=100.0, dt=0.05, ny=17)
    gx, gy = d.node_arrays()
    uf = d.new_field(); uf[:] = 0.5
    vf = d.new_field(); vf[:] = -0.25
    return d, gx, gy, uf, vf


def test_traject_fwdeuler_by_hand_and_out_of_bounds_flags():
    o = get_oracle()
    d, gx, gy, uf, vf = _uniform_traj_case()
    o.config(d.mnx, d.mny)
    z = d.new_field()
    xp = np.array([0.5, -0.1, 1.2, 0.5, 0.5, 0.0, 0.5])
    yp = np.array([0.5, 0.5, 0.5, -0.1, 1.2, 0.5, 1.0])
    up, vp = np.full(7, 0.1), np.full(7, 0.05)
    cx, cy, repc = np.full(7, 2.0), np.full(7, 1.5), np.full(7, 10.0)
    out = np.zeros(7, dtype=np.int32)
    fr, densref = 4.0, 1.2
    x0, y0, u0, v0 = xp[0], yp[0], up[0], vp[0]
    o.traject(d.nx, d.ny, 7, 1, 2, 1, 1, out, d.dk, densref, fr, 1e-10, 1.0, cx, cy, repc, gx, gy, uf, vf, uf, vf, z, z,
              xp, yp, up, vp)
    # iFindPos: left of the first node -> 1; right of the last -> 2; below -> 3; above -> 2; x == x(1) exactly is
    # inside (x(2) > xp first matches at ip = 2) ; y == y(ny) exactly has no node above -> 2
    assert list(out) == [0, 1, 2, 3, 2, 0, 2]
    rep = 10.0 * np.sqrt((0.5 - u0) ** 2 + (-0.25 - v0) ** 2)
    cd = 24.0 / rep
    cpx, cpy = cd * 2.0 * densref, cd * 1.5 * densref
    h = d.dk
    u1 = u0 + h * (cpx * abs(0.5 - u0) * (0.5 - u0))
    v1 = v0 + h * (cpy * abs(-0.25 - v0) * (-0.25 - v0) - 1.0 / fr)
    assert up[0] == u1 and vp[0] == v1
    assert xp[0] == x0 + h * u1 and yp[0] == y0 + h * v1      # semi-implicit: the NEW velocity moves the particle
    assert xp[1] == -0.1 and up[1] == 0.1                      # flagged particles are never updated


def test_traject_heuntrap_converges_to_trapezoidal_rule():
    """With a constant fluid state the Newton iterations of HeunTrap (h = sub-step) solve
    w1 = w0 + h/2 (f(w0) + f(w1)); check the residual of that equation."""
    o = get_oracle()
    d, gx, gy, uf, vf = _uniform_traj_case()
    o.config(d.mnx, d.mny)
    z = d.new_field()
    xp, yp = np.array([0.4]), np.array([0.6])
    up, vp = np.array([0.1]), np.array([0.05])
    cx, cy, repc = np.array([2.0]), np.array([1.5]), np.array([10.0])
    out = np.zeros(1, dtype=np.int32)
    fr, densref = 4.0, 1.0
    w0 = np.array([xp[0], up[0], yp[0], vp[0]])
    o.traject(d.nx, d.ny, 1, 1, 1, 1, 20, out, d.dk, densref, fr, 1e-13, 1.0, cx, cy, repc, gx, gy, uf, vf, uf, vf, z, z,
              xp, yp, up, vp)
    w1 = np.array([xp[0], up[0], yp[0], vp[0]])
    rep = 10.0 * np.sqrt((0.5 - w0[1]) ** 2 + (-0.25 - w0[3]) ** 2)
    cpx, cpy = 24.0 / rep * 2.0, 24.0 / rep * 1.5

    def f(w):
        return np.array([w[1], cpx * abs(0.5 - w[1]) * (0.5 - w[1]), w[3], cpy * abs(-0.25 - w[3]) * (-0.25 - w[3]) - 1.0 / fr])
    res = w1 - w0 - 0.5 * d.dk * (f(w0) + f(w1))
    assert np.abs(res).max() < 1e-12
    assert out[0] == 0 and w1[1] > w0[1]
    # maxit = 1 is Heun's predictor only (traject.f:384-390): explicit Euler with the old-level fluid state
    xp2, yp2, up2, vp2 = np.array([0.4]), np.array([0.6]), np.array([0.1]), np.array([0.05])
    o.traject(d.nx, d.ny, 1, 1, 1, 1, 1, out, d.dk, densref, fr, 1e-13, 1.0, cx, cy, repc, gx, gy, uf, vf, uf, vf, z, z,
              xp2, yp2, up2, vp2)
    assert np.allclose([xp2[0], up2[0], yp2[0], vp2[0]], w0 + d.dk * f(w0), rtol=0, atol=1e-15)


def test_bilinear_interpolation_reproduces_linear_fields():
    o = get_oracle()
    x, y = dk.uniform_grid(15, 13, 2.0, 1.0)
    d = dk._mk("lin", 15, 13, dk.RegionTables(15, 13), 100.0, 1e-3, x=x, y=y)
    o.config(d.mnx, d.mny)
    gx, gy = d.node_arrays()
    uf = 0.3 + 0.2 * gx - 0.4 * gy
    z = d.new_field()
    rng = np.random.default_rng(6)
    n = 50
    xp, yp = rng.uniform(0.1, 1.9, n), rng.uniform(0.1, 0.9, n)
    up, vp = np.zeros(n), np.zeros(n)
    out = np.zeros(n, dtype=np.int32)
    # huge drag, one tiny step: up moves towards the interpolated fluid velocity by a known fraction
    cx = np.full(n, 1.0); cy = np.full(n, 1.0); repc = np.full(n, 1.0)
    x0, y0 = xp.copy(), yp.copy()
    o.traject(d.nx, d.ny, n, 1, 2, 1, 1, out, 1e-6, 1.0, 1e30, 1e-10, 1.0, cx, cy, repc, gx, gy, uf, z, uf, z, z, z,
              xp, yp, up, vp)
    ufl = 0.3 + 0.2 * x0 - 0.4 * y0          # exact for a linear field
    rep = np.abs(ufl)
    expect = 1e-6 * (24.0 / rep * np.abs(ufl) * ufl)
    assert np.allclose(up, expect, rtol=1e-11)
    assert not out.any()


def test_atd_and_trajectory_paths_stay_in_bounds():
    """The subscript-checked build aborts on any out-of-range array reference."""
    o = Oracle("liboracle_chk.so")
    d = dk.backward_step(24, re=800.0, dt=0.004, ny=20, **SS_KW)
    o.config(d.mnx, d.mny)
    rng = np.random.default_rng(8)
    u, v, t = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng, 0.0, 1.0)
    out = [d.new_field() for _ in range(4)]
    o.smallscale(*_ss_args(d, 0), u, v, t, *out)
    o.smallscale(*_ss_args(d, 1), u, v, t, *out)
    gx, gy = d.node_arrays()
    n = 40
    xp, yp = rng.uniform(-0.1, 1.1, n), rng.uniform(-0.1, 1.1, n)
    up, vp = rng.uniform(-1, 1, n), rng.uniform(-1, 1, n)
    outb = np.zeros(n, dtype=np.int32)
    one = np.ones(n)
    for method in (1, 2):
        o.traject(d.nx, d.ny, n, 2, method, 3, 4, outb, d.dk, 1.2, d.fr, 1e-10, 1.0, one, one, 10 * one, gx, gy, u, v, u, v,
                  t, t, xp, yp, up, vp)
    p, dd = d.new_field(), d.new_field()
    import ctypes as C
    from wolfd2_b200 import _abi
    tr = _abi.Traject()
    tr.ntr, tr.ntsubstp, tr.nTrMethod, tr.nTrCdEq, tr.mTrHTmit = n, 1, 2, 1, 1
    tr.densref, tr.dTrHTtol, tr.dTrHTdel = 1.0, 1e-10, 1.0
    part = dict(tr=tr, gx=gx, gy=gy, cpartx=one, cparty=one, repc=one, xp=xp, yp=yp, up=up, vp=vp, out=outb)
    u2, v2 = d.new_field(), d.new_field()
    o.coldstart(d, u2, v2, p)
    tz = d.new_field()
    o.atd_init(d, u2, v2, tz, *out)
    rc, logs = o.step_full(d, u2, v2, p, tz, dd, ss_fields=out, particles=part, nsteps=2)
    assert rc == 0 and len(logs) == 2
