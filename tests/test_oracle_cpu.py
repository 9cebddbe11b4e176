"""CPU tests of the oracle (test infrastructure) and of the host-side set-up code.

The reference has no tests or golden vectors, so the oracle is pinned by: (1) quirk tests
that fail if a reference idiosyncrasy is "fixed", (2) invariants, (3) an external physical check against
Ghia, Ghia & Shin (1982), (4) regression fixtures under tests/golden/ produced by make_golden.py."""
import os
import sys

import numpy as np
import pytest

from oracle import Oracle, get_oracle
from util import rand_field, region_args, rel_l2, make_test_decks
from wolfd2_b200 import deck as dk

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))


@pytest.fixture(scope="module")
def orc():
    return get_oracle()


# ------------------------------------------------------------------ host set-up code vs oracle
@pytest.mark.parametrize("grid", ["uniform", "stretched"])
def test_metrics_numpy_equals_oracle_bitwise(orc, grid):
    nx, ny = 41, 33
    x, y = dk.uniform_grid(nx, ny, 2.0, 1.0) if grid == "uniform" else dk.stretched_grid(nx, ny)
    m_np = dk.metrics_from_grid(x, y, nx + 3, ny + 2, dlref=1.3)
    m_c = orc.grid_metrics(x, y, nx + 3, ny + 2, dlref=1.3)
    for n, a in m_c.items():
        assert np.array_equal(a, m_np[n]), n
    # zero padding outside 1..nx,1..ny is load-bearing
    assert m_np["rac"][:, nx + 1].max() == 0.0 and m_np["rgn"][ny + 1, :].max() == 0.0
    if grid == "uniform":
        assert np.all(m_np["rbu"][1:ny + 1, 1:nx + 1] == 0.0)     # Cartesian: cross metrics vanish exactly


def test_region_tables_match_oracle_setup(orc):
    """Python RegionTables == InitBCFlags + `blockage` statement + SetUpBCs completion in the oracle."""
    import ctypes as C
    d = dk.backward_step(40, re=100.0, dt=0.01, ny=30)
    r = d.regions
    orc.config(d.mnx, d.mny)
    L = orc.lib
    nReg = r.nReg.copy()
    brd = np.zeros_like(r.nRegBrd); typ = np.zeros_like(r.nRegType); mom = np.zeros_like(r.nMomBdTp)
    val = np.ones_like(r.dBCVal)
    brd[0, 0, 1] = r.i_borders[0]; brd[2, 1, 0] = r.j_borders[0]
    ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
    L.orc_initbcflags(ip(nReg), ip(typ), ip(mom), val.ctypes.data_as(C.POINTER(C.c_double)))
    L.orc_bc_blockage(ip(nReg), ip(typ), ip(mom), 1, 1)
    mom[0, 1, 0] = dk.BM_INLET; mom[1, 0, 1] = dk.BM_OUTLT1; mom[1, 1, 1] = dk.BM_OUTLT1
    L.orc_setupbcs_complete(d.nx, d.ny, ip(nReg), ip(brd), ip(mom))
    assert np.array_equal(brd, r.nRegBrd) and np.array_equal(typ, r.nRegType) and np.array_equal(mom, r.nMomBdTp)


def test_deck_writer_emits_reference_syntax(tmp_path):
    d = dk.backward_step(40, re=100.0, dt=0.01, ny=30)
    d.write_reference_files(str(tmp_path), n_time_steps=7)
    txt = (tmp_path / "input.dat").read_text()
    for kw in ("section input_parameters", "time_step_size", "ppe_solver rb_sor", "section boundary_conditions",
               "number_of_regions 2 2", "blockage 1 1", "inlet 1 2 w normal_vel", "outlet 2 1 e fully_dev"):
        assert kw in txt
    assert all(len(l) <= 80 for l in txt.splitlines())          # mlinelgt = 80, include/config.f
    first = (tmp_path / "grid.dat").read_text().splitlines()[0].split()
    assert first == ["40", "30"]


# ------------------------------------------------------------------ quirks
def test_alttridlu_first_row_quirk(orc):
    """Q1: row 1 divides by a(2,2); equals plain Thomas on a system whose c1 is c1*d1/d2."""
    orc.config(302, 302)
    rng = np.random.default_rng(1)
    n = 50
    a = np.zeros((n, 3)); a[:, 0] = rng.uniform(-1, 1, n); a[:, 2] = rng.uniform(-1, 1, n); a[:, 1] = 3 + rng.uniform(0, 1, n)
    b = rng.uniform(-1, 1, n)
    A = np.diag(a[:, 1]) + np.diag(a[1:, 0], -1) + np.diag(a[:-1, 2], 1)
    Aq = A.copy(); Aq[0, 1] = a[0, 2] * a[0, 1] / a[1, 1]
    x_plain, x_quirk = np.linalg.solve(A, b), np.linalg.solve(Aq, b)
    bo = b.copy(); orc.alttridlu(n, a.copy().reshape(-1), bo)
    assert rel_l2(bo, x_quirk) < 1e-13
    assert rel_l2(bo, x_plain) > 1e-6          # a "fixed" solver would land here


def test_momentum_is_one_coupled_chain_not_independent_lines(orc):
    """Q2 / F4: zeroing the inter-line couplings changes the step-1 result far above tolerance; we detect
    it through XMomentum's sensitivity to a perturbation placed at the END of the previous grid line."""
    d = dk.cavity(33, re=100.0, dt=0.05)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(3)
    r, m = d.regions, d.metrics
    us, vs, un, vn = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    ym = [m[n] for n in "ran rbn rbc rgc djv xen yen xzc yzc xev yev xzv yzv".split()]
    z = d.new_field()

    def solve(vn_):
        out = d.new_field()
        orc.ymomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, z, z,
                      us, vs, un, vn_, out)
        return out
    base = solve(vn)
    vn2 = vn.copy(); vn2[10, d.nx] += 1.0      # last unknown (i=nx) of line j=10
    pert = solve(vn2)
    # the first unknowns (i=2,3) of the NEXT line j=11 feel it through a(1,.) of the first unknown
    assert abs(pert[11, 2] - base[11, 2]) > 1e-6
    # and nothing leaks two lines up at i=2 beyond what the stencil (j+-1) explains
    assert abs(pert[14, 2] - base[14, 2]) < 1e-9


def test_norm_index_sets(orc):
    """Q6: DMaxNorm seeds with |u(5,5)| and scans 2..nx-1, 2..ny-1 only."""
    d = dk.cavity(12, re=100.0, dt=0.01)
    orc.config(d.mnx, d.mny)
    w = d.new_field(); w[5, 5] = -3.0; w[d.ny, 3] = 50.0; w[1, 1] = 9.0; w[3, d.nx] = 70.0
    assert orc.dmaxnorm(d.nx, d.ny, w) == 3.0
    w2 = d.new_field(); w2[2, 2] = 4.0
    assert orc.diffmaxnorm(d.nx, d.ny, w, w2) == 4.0


def test_sorrb_equals_sorrbp_and_min_two_iterations(orc):
    """Q10: ids 5 and 6 are numerically identical; at least 2 iterations are always run."""
    d = dk.backward_step(30, re=100.0, dt=0.01, ny=26)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(5)
    r, m = d.regions, d.metrics
    u, v, p = rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01)
    pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
    res = []
    for s in (5, 6):
        pp = p.copy()
        n = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, s, 2000, d.dk, 1e-9, 1.6, *pm8, u, v, pp)
        res.append((n, pp))
    assert res[0][0] == res[1][0] and np.array_equal(res[0][1], res[1][1])
    pp = p.copy()
    assert orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 5, 2000, d.dk, 1e30, 1.0, *pm8, u, v, pp) == 2


@pytest.mark.parametrize("solver", [1, 2, 3, 4, 5, 6])
def test_all_six_ppe_solvers_reach_the_same_solution(orc, solver):
    d = dk.cavity(21, re=100.0, dt=0.01, ny=19)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(8)
    r, m = d.regions, d.metrics
    u, v = rand_field(d, rng, -0.01, 0.01), rand_field(d, rng, -0.01, 0.01)
    pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
    sols = {}
    for s in (5, solver):
        pp = d.new_field()
        n = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, s, 20000, d.dk, 1e-13, 1.5, *pm8, u, v, pp)
        assert n < 20000
        q = pp[2:d.ny + 1, 2:d.nx + 1]
        sols[s] = q - q.mean()       # pure-Neumann problem: compare up to a constant
    assert np.abs(sols[solver] - sols[5]).max() < 1e-8 * max(1.0, np.abs(sols[5]).max())


def test_projection_removes_divergence(orc):
    """Invariant: after a converged Ppe + Project the discrete divergence vanishes at every cell whose
    equation does not touch a ghost pressure (ghost p is frozen during the solve, pressure.f:431-446, so
    wall-adjacent cells keep a residual of the size of the pressure change -- as in the reference)."""
    d = dk.cavity(26, re=100.0, dt=0.01, ny=22)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(11)
    r, m = d.regions, d.metrics
    u, v, p = rand_field(d, rng, -0.1, 0.1), rand_field(d, rng, -0.1, 0.1), d.new_field()
    orc.velboundcond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nMomBdTp, r.dBCVal, u, v)
    dm = [m[n] for n in "xeu yeu xzv yzv".split()]
    div0 = d.new_field(); orc.divergence(d.nx, d.ny, 1, *dm, u, v, div0)
    pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
    orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 5, 50000, d.dk, 1e-13, 1.7, *pm8, u, v, p)
    orc.presboundcond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.dBCVal, p)
    pj = [m[n] for n in "dju djv yeu xzv yzu xev".split()]
    orc.project(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, *pj, p, u, v)
    div1 = d.new_field(); orc.divergence(d.nx, d.ny, 1, *dm, u, v, div1)
    I = (slice(3, d.ny), slice(3, d.nx))
    assert np.abs(div1[I]).max() < 1e-8 * np.abs(div0[I]).max()


def test_subscript_checked_build_runs_clean():
    """Every A(f,i,j) of the oracle stays inside (0:mnx,0:mny) with the smallest legal mnx=nx+1."""
    o = Oracle("liboracle_chk.so")
    for d in make_test_decks(19, 15)[:4]:
        d.msorit = 30
        u, v, p = d.new_field(), d.new_field(), d.new_field()
        o.coldstart(d, u, v, p)
        rc, _ = o.step(d, u, v, p, 2)
        assert rc == 0 and o.lib.orc_get_errflag() == 0
    d = dk.heated_cavity(19, re=100.0, dt=0.005, ny=15, nfiltt=1)       # ThermEnergy, TempBoundCond, EqState, Filter(_T_)
    d.msorit = 30
    u, v, p, t, den = (d.new_field() for _ in range(5))
    rc, _ = o.step(d, u, v, p, 2, t=t, d=den)
    assert rc == 0 and o.lib.orc_get_errflag() == 0


# ------------------------------------------------------------------ external physical check
@pytest.mark.slow
def test_cavity_matches_ghia_at_doubled_reynolds(orc):
    """Ghia, Ghia & Shin (1982), Re=100 u on the vertical centreline.  As written, DConvU
    (momentum.f:1002-1004) multiplies the face flux coefficient by the SUM of the two neighbouring
    velocities (not their mean): convection is counted twice, so a deck with Re=50 reproduces the
    published Re=100 profile (error 4.8e-3 at 48^2, 2.2e-3 at 64^2: second order), while Re=100 does not
    (error 0.08 at every resolution).  This pins momentum + PPE + projection + BCs physically."""
    n = 48
    h = 1.0 / (n - 1)
    gy = [0.9766, 0.9688, 0.9609, 0.9531, 0.8516, 0.7344, 0.6172, 0.5, 0.4531, 0.2813, 0.1719, 0.1016, 0.0703, 0.0625, 0.0547]
    gu = [0.84123, 0.78871, 0.73722, 0.68717, 0.23151, 0.00332, -0.13641, -0.20581, -0.21090, -0.15662, -0.10150,
          -0.06434, -0.04775, -0.04192, -0.03717]
    err = {}
    for re in (50.0, 100.0):
        d = dk.cavity(n, re=re, dt=0.2 * re * h * h)
        d.sorrel, d.sortol = 2.0 / (1.0 + np.sin(np.pi * h)), 1e-7
        u, v, p = d.new_field(), d.new_field(), d.new_field()
        orc.coldstart(d, u, v, p)
        for _ in range(60):
            rc, lg = orc.step(d, u, v, p, 100)
            assert rc == 0
            if max(lg[-1]["dif"][1:3]) < 2e-7:
                break
        xi = (np.arange(d.nx + 2) - 1.0) * h
        yc = (np.arange(d.ny + 2) - 1.5) * h          # u(i,j) sits at (x_i, y_{j-1/2}), src/grid.f:337-343
        col = np.array([np.interp(0.5, xi[1:d.nx + 1], u[j, 1:d.nx + 1]) for j in range(d.ny + 2)])
        err[re] = max(abs(np.interp(y, yc[1:d.ny + 2], col[1:d.ny + 2]) - g) for y, g in zip(gy, gu))
    assert err[50.0] < 8e-3
    assert err[100.0] > 5e-2


@pytest.mark.slow
def test_cavity_matches_ghia_re400_at_nominal_200(orc):
    """The same check one octave up: Ghia, Ghia & Shin's Re=400 centreline profile is reproduced by a deck with
    Re=200 (convection counted twice, see above), not by one with Re=400.  At this Reynolds number the O(dt)
    splitting error of the steady state is visible (Poiseuille pin): the profile is extrapolated to
    dt = 0 from dt = 0.2 h and 0.1 h.  40^2 grid: 0.023 against 0.118 (64^2 with dt -> 0: 0.007)."""
    n = 40
    h = 1.0 / (n - 1)
    gy = [0.9766, 0.9688, 0.9609, 0.9531, 0.8516, 0.7344, 0.6172, 0.5, 0.4531, 0.2813, 0.1719, 0.1016, 0.0703, 0.0625, 0.0547]
    gu = np.array([0.75837, 0.68439, 0.61756, 0.55892, 0.29093, 0.16256, 0.02135, -0.11477, -0.17119, -0.32726, -0.24299,
                   -0.14612, -0.10338, -0.09266, -0.08186])
    err = {}
    for re in (200.0, 400.0):
        prof = {}
        for dtf in (0.2, 0.1):
            d = dk.cavity(n, re=re, dt=dtf * h)
            d.sorrel, d.sortol = 2.0 / (1.0 + np.sin(np.pi * h)), 1e-7
            u, v, p = d.new_field(), d.new_field(), d.new_field()
            orc.coldstart(d, u, v, p)
            for _ in range(300):
                rc, lg = orc.step(d, u, v, p, 100)
                assert rc == 0
                if max(lg[-1]["dif"][1:3]) < 2e-7:
                    break
            xi = (np.arange(d.nx + 2) - 1.0) * h
            yc = (np.arange(d.ny + 2) - 1.5) * h
            col = np.array([np.interp(0.5, xi[1:d.nx + 1], u[j, 1:d.nx + 1]) for j in range(d.ny + 2)])
            prof[dtf] = np.array([np.interp(y, yc[1:d.ny + 2], col[1:d.ny + 2]) for y in gy])
        err[re] = np.abs(2.0 * prof[0.1] - prof[0.2] - gu).max()
    assert err[200.0] < 0.03
    assert err[400.0] > 0.09


@pytest.mark.slow
def test_poiseuille_channel_converges_to_the_discrete_parabola(orc):
    """Plane channel 6 x 1, uniform inlet (`inlet 1 1 w normal_vel 1`), `outlet 1 1 e mass_cons`, no-slip walls, Re = 1.
    Far from both ends the flow is fully developed: convection vanishes (so the doubled convection of DConvU does
    not matter), v = 0, and the discrete steady x-momentum balance is  (1/Re) d2u/dy2 = dp/dx  with the mirrored
    wall ghosts u_g = -u_1.  Its exact solution on this staggered grid is
        u_j = C (y_j (1 - y_j) + h^2/4),  C = 1 / (1/6 + h^2/3)  (unit flux),   Re dp/dx = -2 C.
    The steady state the REFERENCE SCHEME reaches depends on the time step: the second split step carries no
    j-coupling, which leaves an O(dt) splitting error (3.1e-3, 1.5e-3, 7.7e-4 for dt = 0.2,
    0.1, 0.05 Re h^2: ratios 2.00).  Extrapolated to dt = 0 the profile matches the formula to 1.2e-8 and the
    pressure gradient to 7e-8.  Pins the inlet and mass-conserving outlet fills, the no-slip ghosts, the
    y-diffusion operator, the PPE / projection scaling and global mass conservation physically."""
    nx, ny, length, re = 49, 13, 6.0, 1.0
    x, y = dk.uniform_grid(nx, ny, length, 1.0)
    h, hx = 1.0 / (ny - 1), length / (nx - 1)
    yc = (np.arange(ny + 2) - 1.5) * h
    c = 1.0 / (1.0 / 6.0 + h * h / 3.0)
    exact = (c * (yc * (1.0 - yc) + h * h / 4.0))[2:ny + 1]
    i = nx // 2
    prof, grad = {}, {}
    for dtf in (0.2, 0.1, 0.05):
        reg = dk.RegionTables(nx, ny).inlet(1, 1, "w", normal_vel=1.0).outlet(1, 1, "e", fully_dev=False)
        d = dk._mk("poiseuille", nx, ny, reg, re, dtf * re * h * h, x=x, y=y)
        d.sorrel, d.sortol, d.msorit, d.qtol = 1.7, 1e-13, 20000, 1e-11
        u, v, p = d.new_field(), d.new_field(), d.new_field()
        orc.coldstart(d, u, v, p)
        for _ in range(200):
            rc, lg = orc.step(d, u, v, p, 100)
            assert rc == 0
            if max(lg[-1]["dif"][1:3]) < 2e-14:
                break
        assert abs(u[2:ny + 1, i].sum() * h - 1.0) < 1e-11          # every section carries the inlet flux
        assert abs(u[2:ny + 1, nx].sum() * h - 1.0) < 1e-11         # ... the outlet face too (mass_cons)
        assert np.abs(v[1:ny + 1, i]).max() < 1e-7 and np.abs(u[1, i] + u[2, i]) == 0.0
        prof[dtf] = u[2:ny + 1, i].copy()
        grad[dtf] = re * (p[ny // 2, i + 1] - p[ny // 2, i]) / hx
    err = {k: np.abs(prof[k] - exact).max() for k in prof}
    assert 1.9 < err[0.2] / err[0.1] < 2.1 and 1.9 < err[0.1] / err[0.05] < 2.1     # first order in dt
    u0 = (prof[0.2] - 6.0 * prof[0.1] + 8.0 * prof[0.05]) / 3.0                     # quadratic extrapolation to dt = 0
    g0 = (grad[0.2] - 6.0 * grad[0.1] + 8.0 * grad[0.05]) / 3.0
    assert np.abs(u0 - exact).max() < 1e-7
    assert abs(g0 + 2.0 * c) < 1e-6


def test_natural_convection_matches_de_vahl_davis_ra1000(orc):
    """Differentially heated square cavity, Pr = 0.71, Ra = 10^3: de Vahl Davis' benchmark solution (1983) has a
    mean Nusselt number of 1.118 and mid-plane velocity maxima of 3.649 (u on x = 1/2) and 3.697 (v on y = 1/2) in
    units of alpha/L.  Here the buoyancy comes from the ideal-gas EqState with (Tmax - Tref)/Tref = 1/30 (close to
    Boussinesq), Ra = (dT/T) Re^2 Pr / Fr, and the steady state is reached through the momentum-energy iterations.
    24^2 grid: Nu = 1.115 on BOTH walls (energy conservation), u_max = 3.54, v_max = 3.59.  At this Rayleigh number
    inertia is negligible, so the doubled momentum convection does not matter (at Ra = 10^4 it costs 2 % in Nu: 2.20
    against 2.243).  Pins ThermEnergy, TempBoundCond, EqState, the buoyancy term of YMomentum and their coupling."""
    n, ra, re, pr = 24, 1.0e3, 10.0, 0.71
    h = 1.0 / (n - 1)
    eps = (310.0 - 300.0) / 300.0
    fr = eps * re * re * pr / ra
    uref = np.sqrt(fr * 9.81)
    dt_nd = 0.05 * re * pr * h * h
    d = dk.heated_cavity(n, re=re, dt=dt_nd / uref, uref=uref, nmeiter=3)
    d.sorrel, d.sortol, d.msorit = 2.0 / (1.0 + np.sin(np.pi * h)), 1e-8, 2000
    assert abs(d.fr - fr) < 1e-12 * fr and abs(d.dk - dt_nd) < 1e-15 and abs(d.pe - re * pr) < 1e-12
    u, v, p, t, den = (d.new_field() for _ in range(5))
    t[:d.ny + 2, :d.nx + 2] = 0.5
    orc.coldstart(d, u, v, p)
    for _ in range(100):
        rc, lg = orc.step(d, u, v, p, 100, t=t, d=den)
        assert rc == 0
        if max(lg[-1]["dif"][1:4]) < 2e-8:
            break
    nu_w = ((1.0 - t[2:n + 1, 2]) / (0.5 * h)).mean()          # wall value 1, first cell centre h/2 away
    nu_e = ((t[2:n + 1, n] - 0.0) / (0.5 * h)).mean()
    xi = (np.arange(n + 2) - 1.0) * h
    umax = np.abs([np.interp(0.5, xi[1:n + 1], u[j, 1:n + 1]) for j in range(n + 2)]).max() * re * pr
    vmax = np.abs([np.interp(0.5, xi[1:n + 1], v[1:n + 1, i]) for i in range(n + 2)]).max() * re * pr
    assert abs(nu_w - 1.118) < 0.006 and abs(nu_w - nu_e) < 2e-5
    assert abs(umax - 3.649) < 0.15 and abs(vmax - 3.697) < 0.15


# ------------------------------------------------------------------ regression fixtures
@pytest.mark.parametrize("name", ["cavity24x20", "channel22x18_fd", "channel22x18_mc", "bstep26x20", "heated_cavity26x22"])
def test_oracle_reproduces_golden_fixtures(orc, name):
    import make_golden
    d = make_golden.cases()[name]
    got = make_golden.run(d)
    ref = np.load(os.path.join(HERE, "golden", name + ".npz"))
    assert int(got["ncold"]) == int(ref["ncold"])
    assert np.array_equal(got["nql"], ref["nql"]) and np.array_equal(got["nsor"], ref["nsor"])
    for k in range(4):
        for f in ("uvptd" if f"t0" in ref else "uvp"):
            assert np.array_equal(got[f"{f}{k}"], ref[f"{f}{k}"]), (name, f, k)


# ---------------------------------------------------------------------------------------------------------
# Thermal energy row: analytic pins of ThermEnergy + TempBoundCond

def _conduction_deck(regions, n=24, ny=20, re=10.0):
    pe = re * 0.71
    h = 1.0 / (n - 1)
    dt = 0.2 * pe * h * h          # the split steps give no implicit j-coupling: keep dt/(Pe h^2) small
    d = dk._mk("conduction", n, ny, regions, re, dt, thermal=True, eqstate=False, nmeiter=1, sortol=1e-8, msorit=50)
    return d, pe, h, int(6 * pe / dt)


def test_pure_conduction_reaches_the_linear_profile(orc):
    """Fluid at rest between a hot (T=1) and a cold (T=0) wall, adiabatic top and bottom: the steady state of
    the thermal energy equation is the linear profile, whatever the split steps do on the way."""
    reg = dk.RegionTables(24, 20).wall_temperature(1, 1, "w", 1.0).wall_temperature(1, 1, "e", 0.0)
    d, pe, h, steps = _conduction_deck(reg)
    u, v, p, t, den = (d.new_field() for _ in range(5))
    t[:d.ny + 2, :d.nx + 2] = 0.5
    rc, logs = orc.step(d, u, v, p, steps, t=t, d=den)
    assert rc == 0 and np.abs(u).max() == 0.0 and np.abs(v).max() == 0.0
    x = (np.arange(2, d.nx + 1) - 1.5) / (d.nx - 1)        # cell centres between the walls at i = 1 and i = nx
    assert np.abs(t[2:d.ny + 1, 2:d.nx + 1] - (1.0 - x)[None, :]).max() < 1e-12
    assert logs[-1]["dif"][3] < 1e-14


def test_uniform_heat_source_gives_the_discrete_parabola(orc):
    """Both walls at T=0, uniform source s: steady T = (s Pe / 4) (x(1-x) + h^2/4) exactly on this grid -- the
    h^2/4 is the linear ghost extrapolation at the walls, the 1/4 (not 1/2) is the reference's dk*s/2 source
    term entering one split step only (thermal.f:198)."""
    s = 3.0
    reg = dk.RegionTables(24, 20).wall_temperature(1, 1, "w", 0.0).wall_temperature(1, 1, "e", 0.0)
    reg.heat_generation(1, 1, s)
    d, pe, h, steps = _conduction_deck(reg)
    u, v, p, t, den = (d.new_field() for _ in range(5))
    rc, _ = orc.step(d, u, v, p, steps, t=t, d=den)
    assert rc == 0
    x = (np.arange(2, d.nx + 1) - 1.5) / (d.nx - 1)
    exact = s * pe / 4.0 * (x * (1.0 - x) + h * h / 4.0)
    assert np.abs(t[2:d.ny + 1, 2:d.nx + 1] - exact[None, :]).max() < 1e-11
