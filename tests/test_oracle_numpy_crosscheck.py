"""A SECOND, independent restatement of the streaming operators, vectorised in numpy straight from the Fortran text,
cross-checked against the C oracle.  The reference ships no vectors and cannot be compiled here ("parity unpinned",
DESIGN.md 1c); two restatements written separately (C loops, statement by statement; numpy slices, loop nest by loop
nest) that agree bit for bit make a transcription slip in either one visible.  Arrays are f[j, i] (row-major
(0:mny, 0:mnx)); `W(f, di, dj)` is the operand f(i+di, j+dj) over the loop range."""
import numpy as np
import pytest

from oracle import get_oracle
from util import make_test_decks, rand_field


@pytest.fixture(scope="module")
def orc():
    return get_oracle()


def _deck(k=0, nx=37, ny=29):
    d = make_test_decks(nx, ny)[k]
    return d


class Rng:
    """f(i+di, j+dj) for i in i0..i1, j in j0..j1 (inclusive, Fortran indices)."""
    def __init__(self, i0, i1, j0, j1):
        self.i0, self.i1, self.j0, self.j1 = i0, i1, j0, j1

    def __call__(self, f, di=0, dj=0):
        return f[self.j0 + dj:self.j1 + dj + 1, self.i0 + di:self.i1 + di + 1]

    def put(self, f, val):
        f[self.j0:self.j1 + 1, self.i0:self.i1 + 1] = val


def np_convcoef(nx, ny, ncomp, njacob, xzi, xet, yzi, yet, u, v, cc1, cc2):
    """src/momentum.f:895-972."""
    djac = 2.0 if njacob == 1 else 1.0
    big, small = Rng(1, nx + 1, 1, ny + 1), Rng(1, nx, 1, ny)
    if ncomp == 1:
        W = big
        W.put(cc1, (djac * W(yet) * (W(u) + W(u, -1, 0)) - W(xet) * (W(v) + W(v, 0, -1))) * 0.5)
        W = small
        W.put(cc2, (W(xzi) * (W(v, 1, 0) + W(v)) - djac * W(yzi) * (W(u, 0, 1) + W(u))) * 0.5)
    elif ncomp == 2:
        W = small
        W.put(cc1, (W(yet) * (W(u, 0, 1) + W(u)) - djac * W(xet) * (W(v, 1, 0) + W(v))) * 0.5)
        W = big
        W.put(cc2, (djac * W(xzi) * (W(v) + W(v, 0, -1)) - W(yzi) * (W(u) + W(u, -1, 0))) * 0.5)
    elif ncomp == 3:
        W = small
        W.put(cc1, (W(yet) * W(u) - W(xet) * (W(v, 1, 0) + W(v) + W(v, 1, -1) + W(v, 0, -1)) / 4.0) * 0.5)
        W.put(cc2, (W(xzi) * W(v) - W(yzi) * (W(u, 0, 1) + W(u, -1, 1) + W(u) + W(u, -1, 0)) / 4.0) * 0.5)
    elif ncomp == 4:
        W = Rng(0, nx, 1, ny)
        s4 = W(v, 1, 0) + W(v) + W(v, 1, -1) + W(v, 0, -1)
        W.put(cc1, djac * W(yet) * W(u) - W(xet) * s4 / 4.0)
        W.put(cc2, W(xzi) * s4 / 4.0 - djac * W(yzi) * W(u))
    elif ncomp == 5:
        W = Rng(1, nx, 0, ny)
        s4 = W(u, 0, 1) + W(u, -1, 1) + W(u) + W(u, -1, 0)
        W.put(cc1, W(yet) * s4 / 4.0 - djac * W(xet) * W(v))
        W.put(cc2, djac * W(xzi) * W(v) - W(yzi) * s4 / 4.0)
    elif ncomp == 6:
        W = big
        W.put(cc1, (W(yet) * (W(u) + W(u, -1, 0)) - W(xet) * (W(v) + W(v, 0, -1))) * 0.5)
        W.put(cc2, (W(xzi) * (W(v) + W(v, 0, -1)) - W(yzi) * (W(u) + W(u, -1, 0))) * 0.5)


def np_dconvu(nx, ny, c1, c2, u, c):
    """src/momentum.f:1000-1006."""
    W = Rng(1, nx, 2, ny)
    W.put(c, -W(c2, 0, -1) * W(u, 0, -1) - W(c1) * W(u, -1, 0)
          + (W(c1, 1, 0) - W(c1) + W(c2) - W(c2, 0, -1)) * W(u)
          + W(c1, 1, 0) * W(u, 1, 0) + W(c2) * W(u, 0, 1))


def np_dconvv(nx, ny, c1, c2, v, c):
    """src/momentum.f:1064-1070."""
    W = Rng(2, nx, 1, ny)
    W.put(c, -W(c2) * W(v, 0, -1) - W(c1, -1, 0) * W(v, -1, 0)
          + (W(c1) - W(c1, -1, 0) + W(c2, 0, 1) - W(c2)) * W(v)
          + W(c1) * W(v, 1, 0) + W(c2, 0, 1) * W(v, 0, 1))


def np_ddiffu(nx, ny, ac, bc, bn, gn, u, d):
    """src/momentum.f:1028-1042."""
    W = Rng(1, nx, 2, ny)
    s1 = (W(ac, 1, 0) * (W(u, 1, 0) - W(u)) - W(ac) * (W(u) - W(u, -1, 0))
          + W(bc, 1, 0) * (W(u, 1, 1) + W(u, 0, 1) - W(u, 1, -1) - W(u, 0, -1))
          - W(bc) * (W(u, 0, 1) + W(u, -1, 1) - W(u, 0, -1) - W(u, -1, -1)))
    s2 = (W(bn) * (W(u, 1, 1) + W(u, 1, 0) - W(u, -1, 1) - W(u, -1, 0))
          - W(bn, 0, -1) * (W(u, 1, 0) + W(u, 1, -1) - W(u, -1, 0) - W(u, -1, -1))
          + W(gn) * (W(u, 0, 1) - W(u)) - W(gn, 0, -1) * (W(u) - W(u, 0, -1)))
    W.put(d, s1 + s2)


def np_ddiffv(nx, ny, an, bc, bn, gc, v, d):
    """src/momentum.f:1092-1106."""
    W = Rng(2, nx, 1, ny)
    s1 = (W(an) * (W(v, 1, 0) - W(v)) - W(an, -1, 0) * (W(v) - W(v, -1, 0))
          + W(bn) * (W(v, 1, 1) + W(v, 0, 1) - W(v, 1, -1) - W(v, 0, -1))
          - W(bn, -1, 0) * (W(v, 0, 1) + W(v, -1, 1) - W(v, 0, -1) - W(v, -1, -1)))
    s2 = (W(bc, 0, 1) * (W(v, 1, 1) + W(v, 1, 0) - W(v, -1, 1) - W(v, -1, 0))
          - W(bc) * (W(v, 1, 0) + W(v, 1, -1) - W(v, -1, 0) - W(v, -1, -1))
          + W(gc, 0, 1) * (W(v, 0, 1) - W(v)) - W(gc) * (W(v) - W(v, 0, -1)))
    W.put(d, s1 + s2)


def np_divergence(nx, ny, nloc, xet, yet, xzi, yzi, u, v, div):
    """src/pressure.f:284-315."""
    W = Rng(1, nx, 1, ny)
    if nloc == 1:
        ucij = W(yet) * W(u) - W(xet) * (W(v, 1, 0) + W(v) + W(v, 1, -1) + W(v, 0, -1)) / 4.0
        uci1j = W(yet, -1, 0) * W(u, -1, 0) - W(xet, -1, 0) * (W(v) + W(v, -1, 0) + W(v, 0, -1) + W(v, -1, -1)) / 4.0
        vcij = W(xzi) * W(v) - W(yzi) * (W(u, 0, 1) + W(u, -1, 1) + W(u) + W(u, -1, 0)) / 4.0
        vcij1 = W(xzi, 0, -1) * W(v, 0, -1) - W(yzi, 0, -1) * (W(u) + W(u, -1, 0) + W(u, 0, -1) + W(u, -1, -1)) / 4.0
        W.put(div, ucij - uci1j + vcij - vcij1)
    else:
        uci1j = W(yet, 1, 0) * (W(u, 0, 1) + W(u, 1, 1) + W(u) + W(u, 1, 0)) / 4.0 - W(xet, 1, 0) * W(v, 1, 0)
        ucij = W(yet) * (W(u, -1, 1) + W(u, 0, 1) + W(u, -1, 0) + W(u)) / 4.0 - W(xet) * W(v)
        vcij1 = W(xzi, 0, 1) * (W(v, 0, 1) + W(v, 1, 1) + W(v) + W(v, 1, 0)) / 4.0 - W(yzi, 0, 1) * W(u, 0, 1)
        vcij = W(xzi) * (W(v) + W(v, 1, 0) + W(v, 0, -1) + W(v, 1, -1)) / 4.0 - W(yzi) * W(u)
        W.put(div, uci1j - ucij + vcij1 - vcij)


def np_rhsppe(nx, ny, cartes, dk, rbu, rbv, div, p, b):
    """src/pressure.f:347-375; b is the vector b(ind), ind = (j-2)*(nx-1)+i-1."""
    W = Rng(2, nx, 2, ny)
    bb = W(div) / dk
    if not cartes:
        bb = bb - (W(rbu) * (W(p, 1, 1) + W(p, 0, 1) - W(p, 1, -1) - W(p, 0, -1))
                   - W(rbu, -1, 0) * (W(p, 0, 1) + W(p, -1, 1) - W(p, 0, -1) - W(p, -1, -1))
                   + W(rbv) * (W(p, 1, 1) + W(p, 1, 0) - W(p, -1, 1) - W(p, -1, 0))
                   - W(rbv, 0, -1) * (W(p, 1, 0) + W(p, 1, -1) - W(p, -1, 0) - W(p, -1, -1)))
    b[:(nx - 1) * (ny - 1)] = bb.ravel()


def np_sorrb_iteration(nx, ny, sorrel, rau, rgv, b, p):
    """One iteration of SorRB on a one-region grid without blockage: matrix of src/pressure.f:97-106, colour
    sweeps :505-531 (black: i starts at 2+mod(j,2)); each colour is a vector update because a colour's points
    only read the other colour.  Returns dif = max |sum|."""
    W = Rng(2, nx, 2, ny)
    a1, a2, a4, a5 = W(rgv, 0, -1), W(rau, -1, 0), W(rau), W(rgv)
    a3 = -W(rau) - W(rau, -1, 0) - W(rgv) - W(rgv, 0, -1)
    bb = b[:(nx - 1) * (ny - 1)].reshape(ny - 1, nx - 1)
    jj, ii = np.meshgrid(np.arange(2, ny + 1), np.arange(2, nx + 1), indexing="ij")
    dif = 0.0
    for colour in (0, 1):          # black: (i - 2 - mod(j,2)) even  <=>  (i + j) even
        mask = ((ii + jj) % 2) == colour
        s = bb - a1 * W(p, 0, -1) - a2 * W(p, -1, 0) - a4 * W(p, 1, 0) - a5 * W(p, 0, 1)
        s = s / a3 - W(p)
        new = W(p) + sorrel * s
        W(p)[mask] = new[mask]
        dif = max(dif, np.abs(s[mask]).max())
    return dif


SIZES = [(37, 29), (64, 64)]


@pytest.mark.parametrize("size", SIZES)
def test_convcoef_all_cases(orc, size):
    d = _deck(0, *size)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(1)
    ms = [rand_field(d, rng) for _ in range(4)]
    u, v = rand_field(d, rng), rand_field(d, rng)
    for ncomp in range(1, 7):
        for njacob in (0, 1):
            s1, s2 = rand_field(d, rng), rand_field(d, rng)
            a, b, c, e = s1.copy(), s2.copy(), s1.copy(), s2.copy()
            np_convcoef(d.nx, d.ny, ncomp, njacob, *ms, u, v, a, b)
            orc.convcoef(d.nx, d.ny, ncomp, njacob, *ms, u, v, c, e)
            assert np.array_equal(a, c) and np.array_equal(b, e), (ncomp, njacob)


@pytest.mark.parametrize("size", SIZES)
def test_conv_diff_divergence_rhs(orc, size):
    d = _deck(0, *size)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(2)
    f = [rand_field(d, rng) for _ in range(8)]
    for mine, theirs, nargs in ((np_dconvu, orc.dconvu, 3), (np_dconvv, orc.dconvv, 3), (np_ddiffu, orc.ddiffu, 5),
                                (np_ddiffv, orc.ddiffv, 5)):
        s = rand_field(d, rng)
        a, b = s.copy(), s.copy()
        mine(d.nx, d.ny, *f[:nargs], a)
        theirs(d.nx, d.ny, *f[:nargs], b)
        assert np.array_equal(a, b), mine.__name__
    for nloc in (1, 2):
        s = rand_field(d, rng)
        a, b = s.copy(), s.copy()
        np_divergence(d.nx, d.ny, nloc, *f[:6], a)
        orc.divergence(d.nx, d.ny, nloc, *f[:6], b)
        assert np.array_equal(a, b), nloc
    for cartes in (1, 0):
        a = rng.uniform(-1, 1, d.mnx * d.mny)
        b = a.copy()
        np_rhsppe(d.nx, d.ny, cartes, d.dk, f[0], f[1], f[2], f[3], a)
        orc.rhsppe(d.nx, d.ny, cartes, d.dk, f[0], f[1], f[2], f[3], b)
        assert np.array_equal(a, b), cartes


@pytest.mark.parametrize("size", SIZES)
def test_one_red_black_sor_iteration_through_ppe(orc, size):
    """Ppe with max_sor_iter = 1 on the cavity deck == Divergence + RhsPpe + one SorRB iteration, restated in numpy."""
    d = _deck(0, *size)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(3)
    m, r = d.metrics, d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    po = p.copy()
    nconv = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 5, 1, d.dk, 0.0, 1.7, m["rau"], m["rbu"], m["rbv"],
                    m["rgv"], m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, po)
    div = d.new_field()
    np_divergence(d.nx, d.ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, div)
    b = np.zeros(d.mnx * d.mny)
    np_rhsppe(d.nx, d.ny, 1, d.dk, m["rbu"], m["rbv"], div, p, b)
    pn = p.copy()
    np_sorrb_iteration(d.nx, d.ny, 1.7, m["rau"], m["rgv"], b, pn)
    assert nconv == 1 and np.array_equal(pn, po)
    assert not np.array_equal(pn, p)


# ------------------------------------------------------------------ XMomentum / YMomentum, non-porous decks
def py_alttridlu(a, b):
    """src/momentum.f:1307-1339 (a is (n, 3): a(1,i), a(2,i), a(3,i); both modified in place)."""
    n = len(b)
    a[0][2] = a[0][2] / a[1][1]
    b[0] = b[0] / a[0][1]
    for i in range(1, n - 1):
        a[i][1] = a[i][1] - (a[i][0] * a[i - 1][2])
        a[i][2] = a[i][2] / a[i][1]
        b[i] = (b[i] - a[i][0] * b[i - 1]) / a[i][1]
    a[n - 1][1] = a[n - 1][1] - (a[n - 1][0] * a[n - 2][2])
    b[n - 1] = (b[n - 1] - a[n - 1][0] * b[n - 2]) / a[n - 1][1]
    for i in range(n - 2, -1, -1):
        b[i] = b[i] - (a[i][2] * b[i + 1])


def _regions(d):
    from wolfd2_b200 import deck as dk
    r = d.regions
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            brd = {k: int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH)}
            bd = {k: int(r.nMomBdTp[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH)}
            yield brd[dk.WEST], brd[dk.EAST], brd[dk.SOUTH], brd[dk.NORTH], int(r.nRegType[jr, ir]), bd


def _upwind_rows(W, rkj, cj, cjm, cjp, re1, dm, dp, diag_first, por=0.0):
    """The two branches of the implicit operators (momentum.f:365-375, :411-421, :689-699, :737-747): cj is the
    local coefficient that switches, cjm / cjp the neighbours' coefficients on the side the flow comes from,
    dm / dp the diffusion coefficients towards the previous / next unknown."""
    up = cj >= 0.0
    a1 = np.where(up, rkj * (-cjm - re1 * dm), rkj * (-re1 * dm))
    if diag_first:   # XMomentum: dOne + rkj*(...) + dk2*cpj
        a2 = np.where(up, 1.0 + rkj * (cj + re1 * (dp + dm)), 1.0 + rkj * (-cj + re1 * (dp + dm)))
    else:            # YMomentum: rkj*(...) + dk2*cpj + dOne
        a2 = np.where(up, rkj * (cj + re1 * (dp + dm)) + por + 1.0, rkj * (-cj + re1 * (dp + dm)) + por + 1.0)
    a3 = np.where(up, rkj * (-re1 * dp), rkj * (cjp - re1 * dp))
    return a1, a2, a3


def _porous_divide(d, ncomp, arrays):
    """"Divide convective terms by porosity" (src/momentum.f:296-324, :621-649): region after region, so a point on
    a border shared by two porous regions is divided twice."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            if int(r.nRegType[jr, ir]) != dk.RM_POROUS:
                continue
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            sl = (slice(jS + 1, jN + 1), slice(iW, iE + 1)) if ncomp == 1 else (slice(jS, jN + 1), slice(iW + 1, iE + 1))
            for f in arrays:
                f[sl] = f[sl] / r.dPRporos[jr, ir]


def py_poroscoef(d, ncomp, njacob, u, v):
    """PorosCoef, src/momentum.f:1152-1222 (the x case keeps the reference's `/dFour` on the last v only)."""
    from wolfd2_b200 import deck as dk
    import math
    r = d.regions
    cp = d.new_field()
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            if int(r.nRegType[jr, ir]) != dk.RM_POROUS:
                cp[jS:jN + 1, iW:iE + 1] = 0.0
                continue
            porc1, porc2 = r.dPRporc1[jr, ir], r.dPRporc2[jr, ir]
            if ncomp == 1:
                for j in range(jS + 1, jN + 1):
                    for i in range(iW, iE + 1):
                        unorm = math.sqrt(u[j, i] ** 2 + (v[j, i] + v[j, i + 1] + v[j - 1, i] + v[j - 1, i + 1] / 4.0) ** 2)
                        unrm1 = (u[j, i] ** 2) / unorm if unorm > 1.0e-8 else 0.0
                        if njacob == 1:
                            unorm = unrm1 + unorm
                        cp[j, i] = porc1 + porc2 * unorm
            else:
                for j in range(jS, jN + 1):
                    for i in range(iW + 1, iE + 1):
                        unorm = math.sqrt(((u[j + 1, i - 1] + u[j + 1, i] + u[j, i - 1] + u[j, i]) / 4.0) ** 2 + v[j, i] ** 2)
                        unrm1 = (v[j, i] ** 2) / unorm if unorm > 1.0e-8 else 0.0
                        if njacob == 1:
                            unorm = unrm1 + unorm
                        cp[j, i] = porc1 + porc2 * unorm
    return cp


def np_xmomentum(d, us, vs, un, vn):
    """src/momentum.f:278-510."""
    from wolfd2_b200 import deck as dk
    nx, ny, m = d.nx, d.ny, d.metrics
    re1, dk2 = 1.0 / d.re, d.dk * 0.5
    z = d.new_field
    cj1, cj2, c1s, c2s, c1n, c2n, cnvs, cnvn, difs, difn = (z() for _ in range(10))
    np_convcoef(nx, ny, 4, 1, m["xzu"], m["xeu"], m["yzu"], m["yeu"], us, vs, cj1, cj2)
    np_convcoef(nx, ny, 1, 0, m["xzn"], m["xec"], m["yzn"], m["yec"], us, vs, c1s, c2s)
    np_convcoef(nx, ny, 1, 0, m["xzn"], m["xec"], m["yzn"], m["yec"], un, vn, c1n, c2n)
    _porous_divide(d, 1, (cj1, cj2, c1s, c2s, c1n, c2n))
    cpj, cps, cpn = py_poroscoef(d, 1, 1, us, vs), py_poroscoef(d, 1, 0, us, vs), py_poroscoef(d, 1, 0, un, vn)
    np_dconvu(nx, ny, c1s, c2s, us, cnvs)
    np_dconvu(nx, ny, c1n, c2n, un, cnvn)
    np_ddiffu(nx, ny, m["rac"], m["rbc"], m["rbn"], m["rgn"], us, difs)
    np_ddiffu(nx, ny, m["rac"], m["rbc"], m["rbn"], m["rgn"], un, difn)
    W = Rng(1, nx, 2, ny)
    rkj = dk2 * W(m["dju"])
    a1, a2, a3 = _upwind_rows(W, rkj, W(cj1), W(cj1, -1, 0), W(cj1, 1, 0), re1, W(m["rac"]), W(m["rac"], 1, 0), True)
    a2 = a2 + dk2 * W(cpj)
    b = (W(un) - W(us) + rkj * (-W(cnvs) - W(cnvn)) + rkj * re1 * (W(difs) + W(difn))
         - dk2 * (W(cps) * W(us) + W(cpn) * W(un)))
    a = np.stack([a1.ravel(), a2.ravel(), a3.ravel()], axis=1).tolist()
    b = b.ravel().tolist()
    py_alttridlu(a, b)
    a1, a2, a3 = _upwind_rows(W, rkj, W(cj2), W(cj2, 0, -1), W(cj2, 0, 1), re1, W(m["rgn"], 0, -1), W(m["rgn"]), True)
    a2 = a2 + dk2 * W(cpj)
    a = np.stack([a1.ravel(), a2.ravel(), a3.ravel()], axis=1)
    b = np.array(b)
    ind = lambda i, j: (j - 2) * nx + i - 1          # 0-based ind of (i, j)
    ident = []
    for iW, iE, jS, jN, typ, bd in _regions(d):
        if typ == dk.RM_BLOCKG:
            ident += [ind(i, j) for j in range(jS + 1, jN + 1) for i in range(iW, iE + 1)]
        for face, col in ((dk.WEST, iW), (dk.EAST, iE)):
            if bd[face] in (dk.BM_WALL1, dk.BM_WALL2, dk.BM_INLET):
                ident += [ind(col, j) for j in range(jS + 1, jN + 1)]
    if ident:
        a[ident] = (0.0, 1.0, 0.0)
        b[ident] = 0.0
    a, b = a.tolist(), b.tolist()
    py_alttridlu(a, b)
    dus = d.new_field()
    W.put(dus, np.array(b).reshape(ny - 1, nx))
    return dus


def np_ymomentum(d, us, vs, un, vn, dens, densn):
    """src/momentum.f:603-834."""
    from wolfd2_b200 import deck as dk
    nx, ny, m = d.nx, d.ny, d.metrics
    re1, dk2 = 1.0 / d.re, d.dk * 0.5
    z = d.new_field
    cj1, cj2, c1s, c2s, c1n, c2n, cnvs, cnvn, difs, difn = (z() for _ in range(10))
    np_convcoef(nx, ny, 5, 1, m["xzv"], m["xev"], m["yzv"], m["yev"], us, vs, cj1, cj2)
    np_convcoef(nx, ny, 2, 0, m["xzc"], m["xen"], m["yzc"], m["yen"], us, vs, c1s, c2s)
    np_convcoef(nx, ny, 2, 0, m["xzc"], m["xen"], m["yzc"], m["yen"], un, vn, c1n, c2n)
    _porous_divide(d, 2, (cj1, cj2, c1s, c2s, c1n, c2n))
    cpj, cps, cpn = py_poroscoef(d, 2, 1, us, vs), py_poroscoef(d, 2, 0, us, vs), py_poroscoef(d, 2, 0, un, vn)
    np_dconvv(nx, ny, c1s, c2s, vs, cnvs)
    np_dconvv(nx, ny, c1n, c2n, vn, cnvn)
    np_ddiffv(nx, ny, m["ran"], m["rbc"], m["rbn"], m["rgc"], vs, difs)
    np_ddiffv(nx, ny, m["ran"], m["rbc"], m["rbn"], m["rgc"], vn, difn)
    W = Rng(2, nx, 1, ny)
    rkj = dk2 * W(m["djv"])
    a1, a2, a3 = _upwind_rows(W, rkj, W(cj1), W(cj1, -1, 0), W(cj1, 1, 0), re1, W(m["ran"], -1, 0), W(m["ran"]), False,
                              dk2 * W(cpj))
    buoy = d.dk * (W(dens, 0, 1) + W(dens) + W(densn, 0, 1) + W(densn)) / (4.0 * d.fr)
    b = (W(vn) - W(vs) + rkj * (-W(cnvs) - W(cnvn)) + rkj * re1 * (W(difs) + W(difn))
         - dk2 * (W(cps) * W(vs) + W(cpn) * W(vn)) - buoy)
    a = np.stack([a1.ravel(), a2.ravel(), a3.ravel()], axis=1).tolist()
    b = b.ravel().tolist()
    py_alttridlu(a, b)
    a1, a2, a3 = _upwind_rows(W, rkj, W(cj2), W(cj2, 0, -1), W(cj2, 0, 1), re1, W(m["rgc"]), W(m["rgc"], 0, 1), False,
                              dk2 * W(cpj))
    a = np.stack([a1.ravel(), a2.ravel(), a3.ravel()], axis=1)
    b = np.array(b)
    ind = lambda i, j: (j - 1) * (nx - 1) + i - 2    # 0-based ind of (i, j)
    ident = []
    for iW, iE, jS, jN, typ, bd in _regions(d):
        if typ == dk.RM_BLOCKG:
            ident += [ind(i, j) for j in range(jS, jN + 1) for i in range(iW + 1, iE + 1)]
        for face, row in ((dk.SOUTH, jS), (dk.NORTH, jN)):
            if bd[face] in (dk.BM_WALL1, dk.BM_WALL2, dk.BM_INLET):
                ident += [ind(i, row) for i in range(iW + 1, iE + 1)]
    if ident:
        a[ident] = (0.0, 1.0, 0.0)
        b[ident] = 0.0
    a, b = a.tolist(), b.tolist()
    py_alttridlu(a, b)
    dvs = d.new_field()
    W.put(dvs, np.array(b).reshape(ny, nx - 1))
    return dvs


@pytest.mark.parametrize("k", range(6))
def test_xmomentum_ymomentum_second_restatement(orc, k):
    """The whole of XMomentum and YMomentum (coefficients, both split steps, identity rows per region and face type,
    AltTridLU with its first-row quirk) on the six small decks: cavity, channels with both outlet types, backward
    step with a blockage, the mixed-face 2x2 deck, outlets on south/east.  Bit for bit."""
    d = make_test_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(40 + k)
    r, m = d.regions, d.metrics
    us, vs, un, vn = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    dens, densn = rand_field(d, rng, 0.9, 1.1), rand_field(d, rng, 0.9, 1.1)
    xm = [m[n] for n in "rbn rgn rac rbc dju xec yec xzn yzn xeu yeu xzu yzu".split()]
    do = d.new_field()
    orc.xmomentum(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2,
                  *xm, us, vs, un, vn, do)
    assert np.array_equal(np_xmomentum(d, us, vs, un, vn), do)
    ym = [m[n] for n in "ran rbn rbc rgc djv xen yen xzc yzc xev yev xzv yzv".split()]
    do = d.new_field()
    orc.ymomentum(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1,
                  r.dPRporc2, *ym, dens, densn, us, vs, un, vn, do)
    assert np.array_equal(np_ymomentum(d, us, vs, un, vn, dens, densn), do)
    assert np.abs(do).max() > 1e-6


# ------------------------------------------------------------------ a whole time step of the lid-driven cavity
def np_dmaxnorm(nx, ny, u):
    """src/utility.f:479-507 (seed |u(5,5)|, scan 2..nx-1, 2..ny-1)."""
    return max(abs(u[5, 5]), np.abs(u[2:ny, 2:nx]).max())


def np_diffmaxnorm(nx, ny, un, u):
    """src/utility.f:446-473 (seed (2,2), same scan)."""
    return max(abs(un[2, 2] - u[2, 2]), np.abs(un[2:ny, 2:nx] - u[2:ny, 2:nx]).max())


def np_velboundcond_walls(d, u, v):
    """src/bound_cond.f:555-847, one region whose four faces are no-slip walls (BM_WALL1): W, E, S, N in that order."""
    from wolfd2_b200 import deck as dk
    nx, ny, bc = d.nx, d.ny, d.regions.dBCVal
    val = lambda face, var: bc[var - 1, face - 1, 0, 0]        # dBCVal(1,1,face,var); _U_ = 1, _V_ = 2
    iW, iE, jS, jN = 1, nx, 1, ny
    u[jS:jN + 1, iW] = 0.0
    v[jS + 1:jN + 1, iW] = 2.0 * val(dk.WEST, 2) - v[jS + 1:jN + 1, iW + 1]
    u[jS:jN + 1, iE] = 0.0
    v[jS + 1:jN + 1, iE + 1] = 2.0 * val(dk.EAST, 2) - v[jS + 1:jN + 1, iE]
    u[jS, iW + 1:iE + 1] = 2.0 * val(dk.SOUTH, 1) - u[jS + 1, iW + 1:iE + 1]
    v[jS, iW:iE + 1] = 0.0
    u[jN + 1, iW + 1:iE + 1] = 2.0 * val(dk.NORTH, 1) - u[jN, iW + 1:iE + 1]
    v[jN, iW:iE + 1] = 0.0


def np_presboundcond_walls(d, p):
    """src/bound_cond.f:940-1015, walls on all four faces: Neumann ghosts val + inner."""
    from wolfd2_b200 import deck as dk
    nx, ny, bc = d.nx, d.ny, d.regions.dBCVal
    val = lambda face: bc[2, face - 1, 0, 0]                   # _P_ = 3
    iW, iE, jS, jN = 1, nx, 1, ny
    p[jS + 1:jN + 1, iW] = val(dk.WEST) + p[jS + 1:jN + 1, iW + 1]
    p[jS + 1:jN + 1, iE + 1] = val(dk.EAST) + p[jS + 1:jN + 1, iE]
    p[jS, iW + 1:iE + 1] = val(dk.SOUTH) + p[jS + 1, iW + 1:iE + 1]
    p[jN + 1, iW + 1:iE + 1] = val(dk.NORTH) + p[jN, iW + 1:iE + 1]


def np_project_walls(d, p, u, v):
    """src/utility.f:305-392, one flow region with wall faces: interior points only."""
    nx, ny, m = d.nx, d.ny, d.metrics
    W = Rng(2, nx - 1, 2, ny)                                  # u: j = jS+1..jN, i = iW+1..iE-1
    pzi = W(p, 1, 0) - W(p)
    pet = (W(p, 1, 1) + W(p, 0, 1) - W(p, 1, -1) - W(p, 0, -1)) / 4.0
    W.put(u, W(u) - W(m["dju"]) * d.dk * (W(m["yeu"]) * pzi - W(m["yzu"]) * pet))
    W = Rng(2, nx, 2, ny - 1)                                  # v: j = jS+1..jN-1, i = iW+1..iE
    pzi = (W(p, 1, 1) + W(p, 1, 0) - W(p, -1, 1) - W(p, -1, 0)) / 4.0
    pet = W(p, 0, 1) - W(p)
    W.put(v, W(v) - W(m["djv"]) * d.dk * (-W(m["xev"]) * pzi + W(m["xzv"]) * pet))


def np_ppe_rb(d, u, v, p):
    """Ppe with nPpeSolver = 5 on a Cartesian one-region grid (src/pressure.f:90-106, :197-246, :478-541)."""
    nx, ny, m = d.nx, d.ny, d.metrics
    div = d.new_field()
    np_divergence(nx, ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, div)
    b = np.zeros(d.mnx * d.mny)
    np_rhsppe(nx, ny, 1, d.dk, m["rbu"], m["rbv"], div, p, b)
    for it in range(1, d.msorit + 1):
        dif = np_sorrb_iteration(nx, ny, d.sorrel, m["rau"], m["rgv"], b, p)
        if it > 1 and dif < d.sortol:
            return it
    return d.msorit          # "did not converge": nSorConv = msorit (:242-246)


def np_nauxmomentum(d, un, vn, us, vs, dens, densn):
    """src/momentum.f:111-190 (no outlets on this deck: VelOutflowBCs does nothing)."""
    nx, ny = d.nx, d.ny
    us[1:ny + 2, 1:nx + 2] = un[1:ny + 2, 1:nx + 2]
    vs[1:ny + 2, 1:nx + 2] = vn[1:ny + 2, 1:nx + 2]
    for it in range(1, d.mqiter + 1):
        dus = np_xmomentum(d, us, vs, un, vn)
        dvs = np_ymomentum(d, us, vs, un, vn, dens, densn)
        us[1:ny + 1, 1:nx + 1] += dus[1:ny + 1, 1:nx + 1]
        vs[1:ny + 1, 1:nx + 1] += dvs[1:ny + 1, 1:nx + 1]
        if max(np_dmaxnorm(nx, ny, dus), np_dmaxnorm(nx, ny, dvs)) <= d.qtol:
            return it
    return -1


def np_cavity_step(d, u, v, p):
    """src/main.f:690-972, cold flow (one momentum-energy iteration), no filter, no small scales."""
    nx, ny = d.nx, d.ny
    pn, un, vn = p.copy(), u.copy(), v.copy()
    us, vs = u.copy(), v.copy()
    zero = d.new_field()
    nql = np_nauxmomentum(d, un, vn, us, vs, zero, zero)
    np_velboundcond_walls(d, us, vs)
    np_presboundcond_walls(d, p)
    nsor = np_ppe_rb(d, us, vs, p)
    np_presboundcond_walls(d, p)
    np_project_walls(d, p, us, vs)
    np_velboundcond_walls(d, us, vs)
    np_presboundcond_walls(d, p)
    u[:ny + 2, :nx + 2] = us[:ny + 2, :nx + 2]
    v[:ny + 2, :nx + 2] = vs[:ny + 2, :nx + 2]
    np_velboundcond_walls(d, u, v)
    np_presboundcond_walls(d, p)
    dif = [np_diffmaxnorm(nx, ny, pn, p), np_diffmaxnorm(nx, ny, un, u), np_diffmaxnorm(nx, ny, vn, v)]
    return nql, nsor, dif


@pytest.mark.parametrize("size", [(20, 16), (33, 24)])
def test_cavity_time_steps_second_restatement(orc, size):
    """Four time steps of the lid-driven cavity (the bench workload's deck family) through the numpy restatement of the
    whole step body -- QL loop with both momentum solves, wall ghost fills in the reference's call order, Ppe with
    the red/black solver and its convergence rule, projection, the PrintDiff norms -- against orc.step: fields bit
    for bit, identical QL and SOR counts, identical norms."""
    from wolfd2_b200 import deck as dk
    nx, ny = size
    d = dk.cavity(nx, re=100.0, dt=0.01, ny=ny)
    d.msorit, d.sortol, d.sorrel, d.mqiter, d.qtol = 80, 1e-6, 1.6, 6, 1e-5
    d.ppe_solver = "rb_sor"
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(8)
    u, v, p = (0.05 * rand_field(d, rng) for _ in range(3))
    np_velboundcond_walls(d, u, v)
    np_presboundcond_walls(d, p)
    uo, vo, po = u.copy(), v.copy(), p.copy()
    seen = set()
    for step in range(4):
        nql, nsor, dif = np_cavity_step(d, u, v, p)
        rc, lg = orc.step(d, uo, vo, po, 1)
        assert rc == 0
        assert (nql, nsor) == (lg[0]["nQLiter"], lg[0]["nSorConv"]), step
        assert np.array_equal(u, uo) and np.array_equal(v, vo) and np.array_equal(p, po), step
        assert dif == list(lg[0]["dif"][:3]), step
        seen.add(nsor < d.msorit)
    assert nql >= 1


# ------------------------------------------------------------------ ghost fills, every face type, plain loops
def py_velbc(d, u, v, outflow_only):
    """VelBoundCond (src/bound_cond.f:549-847) or, with outflow_only, VelOutflowBCs (:1697-1874): region by region,
    faces W, E, S, N, statement by statement (the mass-conserving outlets are recurrences along the face)."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    U, V, P_ = 1, 2, 3
    WALL1, WALL2, INLET, OUT1, OUT2 = dk.BM_WALL1, dk.BM_WALL2, dk.BM_INLET, dk.BM_OUTLT1, dk.BM_OUTLT2
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            bd = lambda face: int(r.nMomBdTp[face - 1, jr, ir])
            val = lambda face, var: r.dBCVal[var - 1, face - 1, jr, ir]
            # ---- west
            t = bd(dk.WEST)
            if t in (WALL1, WALL2, INLET) and not outflow_only:
                for j in range(jS, jN + 1):
                    u[j, iW] = val(dk.WEST, U) if t == INLET else 0.0
                for j in range(jS + 1, jN + 1):
                    v[j, iW] = v[j, iW + 1] if t == WALL2 else 2.0 * val(dk.WEST, V) - v[j, iW + 1]
            elif t == OUT1:
                for j in range(jS, jN + 1):
                    u[j, iW - 1] = val(dk.WEST, U) + u[j, iW]
                for j in range(jS + 1, jN + 1):
                    v[j, iW] = -v[j, iW + 1]
            elif t == OUT2:
                for j in range(jS + 1, jN + 1):
                    if outflow_only:
                        u[j, iW] = u[j, iW + 1] - v[j, iW + 1] + v[j - 1, iW + 1]     # :1722
                    else:
                        u[j, iW] = u[j, iW + 1] + v[j, iW + 1] - v[j - 1, iW + 1]     # :606
                for j in range(jS + 1, jN + 1):
                    v[j, iW] = -v[j - 1, iW] + 5.0 * (v[j, iW + 1] - v[j - 1, iW + 1]) + 8.0 * (u[j, iW + 1] - u[j, iW])
            # ---- east
            t = bd(dk.EAST)
            if t in (WALL1, WALL2, INLET) and not outflow_only:
                for j in range(jS, jN + 1):
                    u[j, iE] = val(dk.EAST, U) if t == INLET else 0.0
                for j in range(jS + 1, jN + 1):
                    v[j, iE + 1] = v[j, iE] if t == WALL2 else 2.0 * val(dk.EAST, V) - v[j, iE]
            elif t == OUT1:
                for j in range(jS, jN + 1):
                    u[j, iE + 1] = val(dk.EAST, U) + u[j, iE]
                for j in range(jS + 1, jN + 1):
                    v[j, iE + 1] = -v[j, iE]
            elif t == OUT2:
                for j in range(jS + 1, jN + 1):
                    u[j, iE] = u[j, iE - 1] - (v[j, iE] - v[j - 1, iE])
                for j in range(jS + 1, jN):                                           # stops at jN-1
                    v[j, iE + 1] = v[j - 1, iE + 1] + 3.0 * (v[j - 1, iE] - v[j, iE]) - 4.0 * (u[j, iE] - u[j, iE - 1])
            # ---- south
            t = bd(dk.SOUTH)
            if t in (WALL1, WALL2, INLET) and not outflow_only:
                for i in range(iW + 1, iE + 1):
                    u[jS, i] = u[jS + 1, i] if t == WALL2 else 2.0 * val(dk.SOUTH, U) - u[jS + 1, i]
                for i in range(iW, iE + 1):
                    v[jS, i] = val(dk.SOUTH, V) if t == INLET else 0.0
            elif t == OUT1:
                for i in range(iW + 1, iE + 1):
                    u[jS, i] = -u[jS + 1, i]
                for i in range(iW, iE + 1):
                    v[jS, i] = val(dk.SOUTH, V) + v[jS, i]                            # self-reference, as written
            elif t == OUT2:
                for i in range(iW + 1, iE + 1):
                    v[jS, i] = v[jS + 1, i] + (u[jS + 1, i] - u[jS + 1, i - 1])
                for i in range(iW + 1, iE):                                           # stops at iE-1
                    u[jS, i] = u[jS, i - 1] + 3.0 * (u[jS + 1, i - 1] - u[jS + 1, i]) - 4.0 * (v[jS + 1, i] - v[jS, i])
            # ---- north
            t = bd(dk.NORTH)
            if t in (WALL1, WALL2, INLET) and not outflow_only:
                for i in range(iW + 1, iE + 1):
                    u[jN + 1, i] = u[jN, i] if t == WALL2 else 2.0 * val(dk.NORTH, U) - u[jN, i]
                for i in range(iW, iE + 1):
                    v[jN, i] = val(dk.NORTH, V) if t == INLET else 0.0
            elif t == OUT1:
                for i in range(iW + 1, iE + 1):
                    u[jN + 1, i] = -u[jN, i]
                for i in range(iW, iE + 1):
                    v[jN + 1, i] = val(dk.NORTH, V) + v[jN, i]
            elif t == OUT2:
                for i in range(iW, iE + 1):
                    v[jN, i] = v[jN - 1, i] - (u[jN, i] - u[jN, i - 1])
                for i in range(iW + 1, iE):
                    u[jN + 1, i] = u[jN + 1, i - 1] + 3.0 * (u[jN, i - 1] - u[jN, i]) - 4.0 * (v[jN, i] - v[jN - 1, i])


def py_presbc(d, p):
    """PresBoundCond (src/bound_cond.f:884-1024): blockages zero their interior and copy the outward neighbours
    in; walls and inlets are Neumann ghosts (val + inner), outlets Dirichlet mirrors (2 val - inner)."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            val = lambda face: r.dBCVal[2, face - 1, jr, ir]
            if int(r.nRegType[jr, ir]) == dk.RM_BLOCKG:
                for j in range(jS + 1, jN + 1):
                    for i in range(iW + 1, iE + 1):
                        p[j, i] = 0.0
                for j in range(jS + 1, jN + 1):
                    p[j, iW + 1] = val(dk.WEST) + p[j, iW]
                for j in range(jS + 1, jN + 1):
                    p[j, iE] = val(dk.EAST) + p[j, iE + 1]
                for i in range(iW + 1, iE + 1):
                    p[jS + 1, i] = val(dk.SOUTH) + p[jS, i]
                for i in range(iW + 1, iE + 1):
                    p[jN, i] = val(dk.NORTH) + p[jN + 1, i]
                continue
            neu = (dk.BM_WALL1, dk.BM_WALL2, dk.BM_INLET)
            out = (dk.BM_OUTLT1, dk.BM_OUTLT2)
            t = int(r.nMomBdTp[dk.WEST - 1, jr, ir])
            for j in range(jS + 1, jN + 1):
                if t in neu: p[j, iW] = val(dk.WEST) + p[j, iW + 1]
                elif t in out: p[j, iW] = 2.0 * val(dk.WEST) - p[j, iW + 1]
            t = int(r.nMomBdTp[dk.EAST - 1, jr, ir])
            for j in range(jS + 1, jN + 1):
                if t in neu: p[j, iE + 1] = val(dk.EAST) + p[j, iE]
                elif t in out: p[j, iE + 1] = 2.0 * val(dk.EAST) - p[j, iE]
            t = int(r.nMomBdTp[dk.SOUTH - 1, jr, ir])
            for i in range(iW + 1, iE + 1):
                if t in neu: p[jS, i] = val(dk.SOUTH) + p[jS + 1, i]
                elif t in out: p[jS, i] = 2.0 * val(dk.SOUTH) - p[jS + 1, i]
            t = int(r.nMomBdTp[dk.NORTH - 1, jr, ir])
            for i in range(iW + 1, iE + 1):
                if t in neu: p[jN + 1, i] = val(dk.NORTH) + p[jN, i]
                elif t in out: p[jN + 1, i] = 2.0 * val(dk.NORTH) - p[jN, i]


@pytest.mark.parametrize("k", range(6))
def test_ghost_fills_second_restatement(orc, k):
    """All six face types on all four faces (the `mixed` and `south_out` decks put every type somewhere), blockage
    included: VelBoundCond, VelOutflowBCs and PresBoundCond as plain loops from the Fortran == the oracle, bit for bit,
    on random fields with non-zero boundary values."""
    d = make_test_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(60 + k)
    r = d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for outflow_only, ofn in ((False, orc.velboundcond), (True, orc.veloutflowbcs)):
        a, b, c, e = u.copy(), v.copy(), u.copy(), v.copy()
        py_velbc(d, a, b, outflow_only)
        ofn(d.nx, d.ny, r.nReg, r.nRegBrd, r.nMomBdTp, r.dBCVal, c, e)
        assert np.array_equal(a, c) and np.array_equal(b, e), outflow_only
        if not outflow_only:
            assert not np.array_equal(a, u)
    a, c = p.copy(), p.copy()
    py_presbc(d, a)
    orc.presboundcond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.dBCVal, c)
    assert np.array_equal(a, c) and not np.array_equal(a, p)


# ------------------------------------------------------------------ a whole time step on any of the small decks
def np_ppe_general(d, u, v, p):
    """Ppe, nPpeSolver = 5, Cartesian grid (src/pressure.f:90-246, :478-541) with the blockage rows of :108-191."""
    from wolfd2_b200 import deck as dk
    nx, ny, m, r = d.nx, d.ny, d.metrics, d.regions
    div = d.new_field()
    np_divergence(nx, ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, div)
    W = Rng(2, nx, 2, ny)
    rau, rgv = m["rau"], m["rgv"]
    A = [W(rgv, 0, -1).copy(), W(rau, -1, 0).copy(), (-W(rau) - W(rau, -1, 0) - W(rgv) - W(rgv, 0, -1)), W(rau).copy(),
         W(rgv).copy()]

    def ident(i, j):                       # row of cell (i, j) becomes the identity, its divergence zero
        for k, val in enumerate((0.0, 0.0, 1.0, 0.0, 0.0)):
            A[k][j - 2, i - 2] = val
        div[j, i] = 0.0
    nI, nJ = int(r.nReg[0]), int(r.nReg[1])
    blk = lambda ir, jr: int(r.nRegType[jr, ir]) == dk.RM_BLOCKG
    for jr in range(nJ):
        for ir in range(nI):
            if not blk(ir, jr):
                continue
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            for j in range(jS + 2, jN):
                for i in range(iW + 2, iE):
                    ident(i, j)
            if ir > 0 and blk(ir - 1, jr):
                for j in range(jS + 1, jN + 1): ident(iW + 1, j)
            if ir < nI - 1 and blk(ir + 1, jr):
                for j in range(jS + 1, jN + 1): ident(iE, j)
            if jr > 0 and blk(ir, jr - 1):
                for i in range(iW + 1, iE + 1): ident(i, jS + 1)
            if jr < nJ - 1 and blk(ir, jr + 1):
                for i in range(iW + 1, iE + 1): ident(i, jN)
    b = np.zeros(d.mnx * d.mny)
    np_rhsppe(nx, ny, 1, d.dk, m["rbu"], m["rbv"], div, p, b)
    bb = b[:(nx - 1) * (ny - 1)].reshape(ny - 1, nx - 1)
    jj, ii = np.meshgrid(np.arange(2, ny + 1), np.arange(2, nx + 1), indexing="ij")
    for it in range(1, d.msorit + 1):
        dif = 0.0
        for colour in (0, 1):
            mask = ((ii + jj) % 2) == colour
            s = bb - A[0] * W(p, 0, -1) - A[1] * W(p, -1, 0) - A[3] * W(p, 1, 0) - A[4] * W(p, 0, 1)
            s = s / A[2] - W(p)
            new = W(p) + d.sorrel * s
            W(p)[mask] = new[mask]
            dif = max(dif, np.abs(s[mask]).max())
        if it > 1 and dif < d.sortol:
            return it
    return d.msorit


def py_project(d, p, u, v):
    """Project (src/utility.f:290-434): per non-blockage region the interior, the east / north face when it is an
    interface or a fully-developed outlet, the west / south face only when it is a fully-developed outlet."""
    from wolfd2_b200 import deck as dk
    m, r = d.metrics, d.regions
    dju, djv, yeu, yzu, xev, xzv = (m[n] for n in "dju djv yeu yzu xev xzv".split())

    def pu(i, j):
        pzi = p[j, i + 1] - p[j, i]
        pet = (p[j + 1, i + 1] + p[j + 1, i] - p[j - 1, i + 1] - p[j - 1, i]) / 4.0
        u[j, i] = u[j, i] - dju[j, i] * d.dk * (yeu[j, i] * pzi - yzu[j, i] * pet)

    def pv(i, j):
        pzi = (p[j + 1, i + 1] + p[j, i + 1] - p[j + 1, i - 1] - p[j, i - 1]) / 4.0
        pet = p[j + 1, i] - p[j, i]
        v[j, i] = v[j, i] - djv[j, i] * d.dk * (-xev[j, i] * pzi + xzv[j, i] * pet)
    regs = [(ir, jr) for jr in range(int(r.nReg[1])) for ir in range(int(r.nReg[0]))]
    for comp in (0, 1):                         # the u sweep over all regions comes first, then the v sweep
        for ir, jr in regs:
            if int(r.nRegType[jr, ir]) == dk.RM_BLOCKG:
                continue
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            bd = lambda face: int(r.nMomBdTp[face - 1, jr, ir])
            if comp == 0:
                for j in range(jS + 1, jN + 1):
                    for i in range(iW + 1, iE):
                        pu(i, j)
                if bd(dk.WEST) == dk.BM_OUTLT1:
                    for j in range(jS + 1, jN + 1): pu(iW, j)
                if bd(dk.EAST) in (dk.BM_INTERN, dk.BM_OUTLT1):
                    for j in range(jS + 1, jN + 1): pu(iE, j)
            else:
                for j in range(jS + 1, jN):
                    for i in range(iW + 1, iE + 1):
                        pv(i, j)
                if bd(dk.SOUTH) == dk.BM_OUTLT1:
                    for i in range(iW + 1, iE + 1): pv(i, jS)
                if bd(dk.NORTH) in (dk.BM_INTERN, dk.BM_OUTLT1):
                    for i in range(iW + 1, iE + 1): pv(i, jN)


def np_step_general(d, u, v, p):
    """src/main.f:690-972 (cold flow, no filter) with the general ghost fills, Ppe and Project above."""
    nx, ny = d.nx, d.ny
    pn, un, vn = p.copy(), u.copy(), v.copy()
    us, vs = u.copy(), v.copy()
    zero = d.new_field()
    us[1:ny + 2, 1:nx + 2] = un[1:ny + 2, 1:nx + 2]
    vs[1:ny + 2, 1:nx + 2] = vn[1:ny + 2, 1:nx + 2]
    nql = -1
    for it in range(1, d.mqiter + 1):              # nAuxMomentum, src/momentum.f:120-190
        py_velbc(d, us, vs, outflow_only=True)
        dus = np_xmomentum(d, us, vs, un, vn)
        dvs = np_ymomentum(d, us, vs, un, vn, zero, zero)
        us[1:ny + 1, 1:nx + 1] += dus[1:ny + 1, 1:nx + 1]
        vs[1:ny + 1, 1:nx + 1] += dvs[1:ny + 1, 1:nx + 1]
        if max(np_dmaxnorm(nx, ny, dus), np_dmaxnorm(nx, ny, dvs)) <= d.qtol:
            nql = it
            break
    py_velbc(d, us, vs, False)
    py_presbc(d, p)
    nsor = np_ppe_general(d, us, vs, p)
    py_presbc(d, p)
    py_project(d, p, us, vs)
    py_velbc(d, us, vs, False)
    py_presbc(d, p)
    u[:ny + 2, :nx + 2] = us[:ny + 2, :nx + 2]
    v[:ny + 2, :nx + 2] = vs[:ny + 2, :nx + 2]
    py_velbc(d, u, v, False)
    py_presbc(d, p)
    dif = [np_diffmaxnorm(nx, ny, pn, p), np_diffmaxnorm(nx, ny, un, u), np_diffmaxnorm(nx, ny, vn, v)]
    return nql, nsor, dif


@pytest.mark.parametrize("k", range(6))
def test_time_steps_second_restatement_all_decks(orc, k):
    """Cold start fields + 3 time steps on every small deck -- inflow / both outlet types on every side, no-stress
    walls, a blockage -- through the second restatement of the WHOLE cold-flow step (QL loop with VelOutflowBCs,
    both momentum solves, ghost fills in call order, Ppe with blockage rows, Project with its face rules, norms):
    fields bit for bit, QL / SOR counts and norms identical to orc.step."""
    d = make_test_decks()[k]
    d.msorit, d.sortol, d.sorrel, d.mqiter, d.qtol = 60, 1e-7, 1.5, 5, 1e-6
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(70 + k)
    u, v, p = (0.05 * rand_field(d, rng) for _ in range(3))
    py_velbc(d, u, v, False)
    py_presbc(d, p)
    uo, vo, po = u.copy(), v.copy(), p.copy()
    compared = 0
    for step in range(3):
        rc, lg = orc.step(d, uo, vo, po, 1)
        if not (np.isfinite(uo).all() and np.isfinite(po).all()):
            # the `mixed` deck (every face type at once, random start) blows up in its second step in the oracle;
            # the restatement must fail there too (a zero pivot raises in Python where C produces inf / NaN)
            with np.errstate(all="ignore"):
                try:
                    np_step_general(d, u, v, p)
                    assert not (np.isfinite(u).all() and np.isfinite(p).all())
                except ZeroDivisionError:
                    pass
            break
        nql, nsor, dif = np_step_general(d, u, v, p)
        assert rc == 0
        assert (nql, nsor) == (lg[0]["nQLiter"], lg[0]["nSorConv"]), (step, nql, nsor, lg[0])
        assert np.array_equal(u, uo) and np.array_equal(v, vo) and np.array_equal(p, po), step
        assert dif == list(lg[0]["dif"][:3]), step
        compared += 1
    assert compared >= 1


# ------------------------------------------------------------------ Filter, the lexicographic Sor, the cold start
def py_filter(d, ncomp, fp, qu):
    """Filter for u (ncomp 1) and v (ncomp 2), src/utility.f:72-203, :238-243: a full copy is filtered per
    non-blockage region -- interior, then the east / north face when interface or fully-developed outlet, the
    west / south face only when fully-developed outlet -- and copied back whole."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    nx, ny = d.nx, d.ny
    qh = qu.copy()

    def f(i, j):
        qh[j, i] = (qu[j - 1, i] + qu[j, i - 1] + qu[j + 1, i] + qu[j, i + 1] + fp * qu[j, i]) / (fp + 4.0)
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            if int(r.nRegType[jr, ir]) == dk.RM_BLOCKG:
                continue
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            bd = lambda face: int(r.nMomBdTp[face - 1, jr, ir])
            if ncomp == 1:
                for j in range(jS + 1, jN + 1):
                    for i in range(iW + 1, iE):
                        f(i, j)
                if bd(dk.WEST) == dk.BM_OUTLT1:
                    for j in range(jS + 1, jN + 1): f(iW, j)
                if bd(dk.EAST) in (dk.BM_INTERN, dk.BM_OUTLT1):
                    for j in range(jS + 1, jN + 1): f(iE, j)
            else:
                for j in range(jS + 1, jN):
                    for i in range(iW + 1, iE + 1):
                        f(i, j)
                if bd(dk.SOUTH) == dk.BM_OUTLT1:
                    for i in range(iW + 1, iE + 1): f(i, jS)
                if bd(dk.NORTH) in (dk.BM_INTERN, dk.BM_OUTLT1):
                    for i in range(iW + 1, iE + 1): f(i, jN)
    qu[:ny + 2, :nx + 2] = qh[:ny + 2, :nx + 2]


def py_sor_lex(d, rau, rgv, b, p, msorit):
    """Sor, the reference's default solver (src/pressure.f:411-446): lexicographic Gauss-Seidel sweeps with
    relaxation, one-region grid without blockage (matrix of :97-106)."""
    nx, ny = d.nx, d.ny
    for it in range(1, msorit + 1):
        dif = 0.0
        for j in range(2, ny + 1):
            for i in range(2, nx + 1):
                ind = (j - 2) * (nx - 1) + i - 2
                a1, a2, a4, a5 = rgv[j - 1, i], rau[j, i - 1], rau[j, i], rgv[j, i]
                a3 = -rau[j, i] - rau[j, i - 1] - rgv[j, i] - rgv[j - 1, i]
                s = b[ind] - a1 * p[j - 1, i] - a2 * p[j, i - 1] - a4 * p[j, i + 1] - a5 * p[j + 1, i]
                s = s / a3 - p[j, i]
                p[j, i] = p[j, i] + d.sorrel * s
                dif = max(dif, abs(s))
        if it > 1 and dif < d.sortol:
            return it
    return msorit


@pytest.mark.parametrize("k", range(6))
def test_filter_second_restatement(orc, k):
    d = make_test_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(90 + k)
    r = d.regions
    q = rand_field(d, rng)
    ntr = np.zeros(200, np.int32)
    for ncomp in (1, 2):
        a, b = q.copy(), q.copy()
        py_filter(d, ncomp, 5.0, a)
        orc.filter(d.nx, d.ny, ncomp, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, ntr, 5.0, b)
        assert np.array_equal(a, b) and not np.array_equal(a, q), ncomp


def test_default_lexicographic_sor_second_restatement(orc):
    """Ppe with ppe_solver sor (id 1) on the cavity deck: Divergence + RhsPpe + lexicographic sweeps in plain loops."""
    d = _deck(0, 24, 20)
    d.sorrel, d.sortol = 1.5, 1e-7
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(5)
    m, r = d.metrics, d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for msorit in (1, 7, 400):
        po = p.copy()
        nconv = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 1, msorit, d.dk, d.sortol, d.sorrel, m["rau"], m["rbu"],
                        m["rbv"], m["rgv"], m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, po)
        div = d.new_field()
        np_divergence(d.nx, d.ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, div)
        b = np.zeros(d.mnx * d.mny)
        np_rhsppe(d.nx, d.ny, 1, d.dk, m["rbu"], m["rbv"], div, p, b)
        pn = p.copy()
        n = py_sor_lex(d, m["rau"], m["rgv"], b, pn, msorit)
        assert n == nconv and np.array_equal(pn, po), msorit
    assert nconv < 400       # the last call converged before the cap


@pytest.mark.parametrize("k", [0, 1, 2, 3])
def test_cold_start_second_restatement(orc, k):
    """src/main.f:606-641: VelBoundCond, Ppe, PresBoundCond, Project, VelBoundCond on the quiescent initial field."""
    d = make_test_decks()[k]
    d.msorit, d.sortol, d.sorrel = 80, 1e-7, 1.5
    orc.config(d.mnx, d.mny)
    u, v, p = d.new_field(), d.new_field(), d.new_field()
    uo, vo, po = d.new_field(), d.new_field(), d.new_field()
    nso = orc.coldstart(d, uo, vo, po)
    py_velbc(d, u, v, False)
    ns = np_ppe_general(d, u, v, p)
    py_presbc(d, p)
    py_project(d, p, u, v)
    py_velbc(d, u, v, False)
    assert ns == nso
    assert np.array_equal(u, uo) and np.array_equal(v, vo) and np.array_equal(p, po)


# ------------------------------------------------------------------ thermal energy row
def _thermal_decks():
    from wolfd2_b200 import deck as dk
    out = [dk.heated_cavity(29, re=100.0, dt=0.005, ny=23)]
    reg = dk.RegionTables(34, 28, 2, 2, (16,), (12,))
    reg.heat_generation(1, 1, 2.5).fixed_temperature_region(2, 2, 0.8)
    reg.wall_temperature(1, 1, "w", 1.0).wall_heat_flux(1, 2, "w", 0.05).wall_temperature(2, 1, "s", 0.2)
    reg.wall_heat_flux(1, 2, "n", -0.02).wall(1, 2, "n", tangent_vel=1.0)
    out.append(dk._mk("thermal_2x2", 34, 28, reg, 100.0, 0.004, thermal=True, eqstate=True, nmeiter=2))
    return out


def py_tempbc(d, t):
    """TempBoundCond, src/bound_cond.f:1067-1204."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            val = lambda face: r.dBCVal[3, face - 1, jr, ir]              # _T_ = 4
            if int(r.nTRgType[jr, ir]) == dk.RT_TEMPER:
                for j in range(jS + 1, jN + 1):
                    for i in range(iW + 1, iE + 1):
                        t[j, i] = r.dTRgVal[jr, ir]
                for j in range(jS + 1, jN + 1):
                    t[j, iW + 1] = 2.0 * val(dk.WEST) - t[j, iW]
                for j in range(jS + 1, jN + 1):
                    t[j, iE] = 2.0 * val(dk.EAST) - t[j, iE + 1]
                for i in range(iW + 1, iE + 1):
                    t[jS + 1, i] = 2.0 * val(dk.SOUTH) - t[jS, i]
                for i in range(iW + 1, iE + 1):
                    t[jN, i] = 2.0 * val(dk.NORTH) - t[jN + 1, i]
                continue
            bt = lambda face: int(r.nTemBdTp[face - 1, jr, ir])
            for face, rng_, ghost, inner in (
                    (dk.WEST, range(jS + 1, jN + 1), lambda q: (q, iW), lambda q: (q, iW + 1)),
                    (dk.EAST, range(jS + 1, jN + 1), lambda q: (q, iE + 1), lambda q: (q, iE)),
                    (dk.SOUTH, range(iW + 1, iE + 1), lambda q: (jS, q), lambda q: (jS + 1, q)),
                    (dk.NORTH, range(iW + 1, iE + 1), lambda q: (jN + 1, q), lambda q: (jN, q))):
                if bt(face) == 1:       # BT_TEMPER: fixed temperature
                    for q in rng_:
                        t[ghost(q)] = 2.0 * val(face) - t[inner(q)]
                elif bt(face) == 2:     # BT_HTFLUX
                    for q in rng_:
                        t[ghost(q)] = val(face) + t[inner(q)]


def np_eqstate(d, p, t, den):
    """EqState, src/thermal.f:313-322."""
    W = Rng(2, d.nx, 2, d.ny)
    pref = d.densref * d.rconst * d.tref
    c1 = W(p) * d.densref * d.uref ** 2 + pref
    c2 = d.densref * d.rconst * (W(t) * (d.tmax - d.tref) + d.tref)
    r = c1 / c2 - 1.0
    r[np.abs(r) < 1.0e-10] = 0.0
    W.put(den, r)


def np_thermenergy(d, un, vn, u, v, tn, t):
    """ThermEnergy, src/thermal.f:97-272."""
    from wolfd2_b200 import deck as dk
    nx, ny, m, r = d.nx, d.ny, d.metrics, d.regions
    pe1 = 1.0 / d.pe
    py_tempbc(d, t)
    cu1, cv1, cun, cvn, s = (d.new_field() for _ in range(5))
    np_convcoef(nx, ny, 6, 0, m["xzc"], m["xec"], m["yzc"], m["yec"], u, v, cu1, cv1)
    np_convcoef(nx, ny, 3, 0, m["xzv"], m["xeu"], m["yzv"], m["yeu"], un, vn, cun, cvn)
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            s[jS + 1:jN + 1, iW + 1:iE + 1] = r.dHGSTval[jr, ir] if int(r.nTRgType[jr, ir]) == dk.RT_HEATGN else 0.0
    W = Rng(2, nx, 2, ny)
    rau, rbu, rbv, rgv = (m[n] for n in "rau rbu rbv rgv".split())
    rkj = d.dk * W(m["djc"]) * 0.5
    up = W(cu1) >= 0.0
    a1 = np.where(up, rkj * (-W(cu1, -1, 0) - pe1 * W(rau, -1, 0)), rkj * (-pe1 * W(rau, -1, 0)))
    a2 = np.where(up, 1.0 + rkj * (W(cu1) + pe1 * (W(rau) + W(rau, -1, 0))), 1.0 + rkj * (-W(cu1) + pe1 * (W(rau) + W(rau, -1, 0))))
    a3 = np.where(up, rkj * (-pe1 * W(rau)), rkj * (W(cu1, 1, 0) - pe1 * W(rau)))
    c = (-W(cvn, 0, -1) * W(tn, 0, -1) - W(cun, -1, 0) * W(tn, -1, 0)
         + (W(cun) - W(cun, -1, 0) + W(cvn) - W(cvn, 0, -1)) * W(tn)
         + W(cun) * W(tn, 1, 0) + W(cvn) * W(tn, 0, 1))
    dd = (W(rau) * (W(tn, 1, 0) - W(tn)) - W(rau, -1, 0) * (W(tn) - W(tn, -1, 0))
          + W(rbu) * (W(tn, 1, 1) + W(tn, 0, 1) - W(tn, 1, -1) - W(tn, 0, -1))
          - W(rbu, -1, 0) * (W(tn, 0, 1) + W(tn, -1, 1) - W(tn, 0, -1) - W(tn, -1, -1))
          + W(rbv) * (W(tn, 1, 1) + W(tn, 1, 0) - W(tn, -1, 1) - W(tn, -1, 0))
          - W(rbv, 0, -1) * (W(tn, 1, 0) + W(tn, 1, -1) - W(tn, -1, 0) - W(tn, -1, -1))
          + W(rgv) * (W(tn, 0, 1) - W(tn)) - W(rgv, 0, -1) * (W(tn) - W(tn, 0, -1)))
    b = d.dk * W(s) * 0.5 + 2.0 * rkj * (-c + pe1 * dd)
    a = np.stack([a1.ravel(), a2.ravel(), a3.ravel()], axis=1).tolist()
    b = b.ravel().tolist()
    py_alttridlu(a, b)
    up = W(cv1) >= 0.0
    a1 = np.where(up, rkj * (-W(cv1, 0, -1) - pe1 * W(rgv, 0, -1)), rkj * (-pe1 * W(rgv, 0, -1)))
    a2 = np.where(up, 1.0 + rkj * (W(cv1) + pe1 * (W(rgv) + W(rgv, 0, -1))), 1.0 + rkj * (-W(cv1) + pe1 * (W(rgv) + W(rgv, 0, -1))))
    a3 = np.where(up, rkj * (-pe1 * W(rgv)), rkj * (W(cv1, 0, 1) - pe1 * W(rgv)))
    a = np.stack([a1.ravel(), a2.ravel(), a3.ravel()], axis=1)
    b = np.array(b)
    ind = lambda i, j: (j - 2) * (nx - 1) + i - 2
    ident = []
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            if int(r.nTRgType[jr, ir]) == dk.RT_TEMPER:
                iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
                ident += [ind(i, j) for j in range(jS + 1, jN + 1) for i in range(iW + 1, iE + 1)]
    if ident:
        a[ident] = (0.0, 1.0, 0.0)
        b[ident] = 0.0
    a, b = a.tolist(), b.tolist()
    py_alttridlu(a, b)
    W.put(t, W(t) + np.array(b).reshape(ny - 1, nx - 1))


@pytest.mark.parametrize("k", range(2))
def test_thermal_row_second_restatement(orc, k):
    """TempBoundCond (fixed-temperature blocks, temperature and flux faces), EqState with its 1e-10 clip, and
    ThermEnergy (case 3 / 6 coefficients, heat sources, both split steps, fixed-T identity rows), bit for bit."""
    d = _thermal_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(300 + k)
    r, m = d.regions, d.metrics
    t = rand_field(d, rng, 0.0, 1.0)
    a, b = t.copy(), t.copy()
    py_tempbc(d, a)
    orc.tempboundcond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nTRgType, r.nTemBdTp, r.dTRgVal, r.dBCVal, b)
    assert np.array_equal(a, b) and not np.array_equal(a, t)
    p = rand_field(d, rng, -0.3, 0.3)
    s = rand_field(d, rng)
    a, b = s.copy(), s.copy()
    np_eqstate(d, p, t, a)
    orc.eqstate(d.nx, d.ny, d.uref, d.densref, d.tmax, d.tref, d.rconst, p, t, b)
    assert np.array_equal(a, b)
    un, vn, u, v = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    tn = rand_field(d, rng, 0.0, 1.0)
    mm = [m[n] for n in "rau rbu rbv rgv djc xeu yeu xzv yzv xec yec xzc yzc".split()]
    a, b = t.copy(), t.copy()
    np_thermenergy(d, un, vn, u, v, tn, a)
    orc.thermenergy(d.nx, d.ny, r.nReg, r.nRegBrd, r.nTRgType, r.nTemBdTp, d.dk, d.pe, r.dTRgVal, r.dHGSTval, r.dBCVal,
                    *mm, un, vn, u, v, tn, b)
    assert np.array_equal(a, b) and not np.array_equal(a, t)


def _porous_decks():
    from wolfd2_b200 import deck as dk
    out = []
    reg = dk.RegionTables(40, 32, 2, 1, (18,), ()).porous(2, 1, 0.7, 5.0, 2.0)
    reg.wall(1, 1, "n", tangent_vel=1.0).wall(2, 1, "n", tangent_vel=1.0)
    out.append(dk._mk("porous_2x1", 40, 32, reg, 100.0, 0.005))
    reg = dk.RegionTables(44, 40, 2, 2, (22,), (20,))
    reg.porous(1, 1, 0.8, 3.0, 1.0).porous(2, 1, 0.5, 6.0, 2.5).porous(2, 2, 0.9, 1.0, 0.5)
    reg.wall(1, 2, "n", tangent_vel=1.0).wall(2, 2, "n", tangent_vel=-0.5)
    out.append(dk._mk("porous_2x2", 44, 40, reg, 100.0, 0.005))
    return out


@pytest.mark.parametrize("k", range(2))
def test_porous_momentum_second_restatement(orc, k):
    """PorosCoef (both components, both njacob, the mis-parenthesised /dFour, the 1e-8 threshold), the porosity
    division with its double hit on shared region borders, and the Dupuit-Forchheimer terms of both momentum
    equations: XMomentum / YMomentum on decks with one and with three adjoining porous regions, bit for bit."""
    d = _porous_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(500 + k)
    r, m = d.regions, d.metrics
    us, vs, un, vn = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    us[3:6, 20:30] = 0.0
    vs[2:7, 19:31] = 0.0
    for ncomp in (1, 2):
        for njacob in (0, 1):
            o = d.new_field()
            orc.poroscoef(d.nx, d.ny, ncomp, njacob, r.nReg, r.nRegBrd, r.nRegType, r.dPRporos, r.dPRporc1, r.dPRporc2, us, vs, o)
            assert np.array_equal(py_poroscoef(d, ncomp, njacob, us, vs), o), (ncomp, njacob)
    zero = d.new_field()
    xm = [m[n] for n in "rbn rgn rac rbc dju xec yec xzn yzn xeu yeu xzu yzu".split()]
    do = d.new_field()
    orc.xmomentum(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2,
                  *xm, us, vs, un, vn, do)
    assert np.array_equal(np_xmomentum(d, us, vs, un, vn), do)
    ym = [m[n] for n in "ran rbn rbc rgc djv xen yen xzc yzc xev yev xzv yzv".split()]
    do = d.new_field()
    orc.ymomentum(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1,
                  r.dPRporc2, *ym, zero, zero, us, vs, un, vn, do)
    assert np.array_equal(np_ymomentum(d, us, vs, un, vn, zero, zero), do)


# ------------------------------------------------------------------ node averages, line SOR
def py_node_averages(d, u, v, p, util, vbar, pav):
    """VelAvg (src/utility.f:600-626) and PTDAvg (:536-560)."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            blk = int(r.nRegType[jr, ir]) == dk.RM_BLOCKG
            for j in range(jS, jN + 1):
                for i in range(iW, iE + 1):
                    util[j, i] = 0.0 if blk else (u[j, i] + u[j + 1, i]) / 2.0
                    vbar[j, i] = 0.0 if blk else (v[j, i] + v[j, i + 1]) / 2.0
            if blk:
                for j in range(jS + 1, jN):
                    for i in range(iW + 1, iE):
                        pav[j, i] = 0.0
                continue
            for j in range(jS, jN + 1):
                for i in range(iW, iE + 1):
                    s = (p[j, i] + p[j + 1, i] + p[j + 1, i + 1] + p[j, i + 1]) / 4.0
                    pav[j, i] = 0.0 if abs(s) < 1.0e-20 else s


def py_slor(d, rau, rgv, b, p, msorit):
    """Slor with ndir = 1 (lines along i, as Ppe calls it; src/pressure.f:704-799) on a one-region grid: each line
    is assembled from the 5-point matrix, solved with AltTridLU (first-row quirk included), relaxed and stored
    before the next line is assembled."""
    nx, ny = d.nx, d.ny
    for it in range(1, msorit + 1):
        dif = 0.0
        for j in range(2, ny + 1):
            al, bl, old = [], [], []
            for i in range(2, nx + 1):
                ind = (j - 2) * (nx - 1) + i - 2
                a1, a2, a4, a5 = rgv[j - 1, i], rau[j, i - 1], rau[j, i], rgv[j, i]
                a3 = -rau[j, i] - rau[j, i - 1] - rgv[j, i] - rgv[j - 1, i]
                old.append(p[j, i])
                al.append([a2, a3, a4])
                bl.append(b[ind] - (a1 * p[j - 1, i] + a5 * p[j + 1, i]))
            py_alttridlu(al, bl)
            for k in range(nx - 1):
                s = bl[k] - old[k]
                p[j, 2 + k] = old[k] + d.sorrel * s
                dif = max(dif, abs(s))
        if it > 1 and dif < d.sortol:
            return it
    return msorit


@pytest.mark.parametrize("k", range(6))
def test_node_averages_second_restatement(orc, k):
    d = make_test_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(700 + k)
    r = d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    p[3, 3:6] = 1e-21
    s = [rand_field(d, rng) for _ in range(3)]
    a, b = [x.copy() for x in s], [x.copy() for x in s]
    py_node_averages(d, u, v, p, *a)
    orc.velavg(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, u, v, b[0], b[1])
    orc.ptdavg(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, p, b[2])
    for x, y in zip(a, b):
        assert np.array_equal(x, y)


def test_line_sor_second_restatement(orc):
    """Ppe with ppe_solver lsor (id 2): whole lines solved by AltTridLU, Gauss-Seidel from line to line."""
    d = _deck(0, 24, 20)
    d.sorrel, d.sortol = 1.3, 1e-8
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(6)
    m, r = d.metrics, d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for msorit in (1, 5, 300):
        po = p.copy()
        nconv = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, 1, 2, msorit, d.dk, d.sortol, d.sorrel, m["rau"], m["rbu"],
                        m["rbv"], m["rgv"], m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, po)
        div = d.new_field()
        np_divergence(d.nx, d.ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, div)
        b = np.zeros(d.mnx * d.mny)
        np_rhsppe(d.nx, d.ny, 1, d.dk, m["rbu"], m["rbv"], div, p, b)
        pn = p.copy()
        n = py_slor(d, m["rau"], m["rgv"], b, pn, msorit)
        assert n == nconv and np.array_equal(pn, po), msorit
    assert nconv < 300


# ------------------------------------------------------------------ ATD small-scale model
def py_filter_t(d, fp, qu):
    """Filter, case(_T_) (src/utility.f:205-229): the region test compares nTRgType with BT_TEMPER (= 1), as written."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    qh = qu.copy()
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            if int(r.nTRgType[jr, ir]) == 1:
                continue
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            for j in range(jS + 1, jN + 1):
                for i in range(iW + 1, iE + 1):
                    qh[j, i] = (qu[j - 1, i] + qu[j, i - 1] + qu[j + 1, i] + qu[j, i + 1] + fp * qu[j, i]) / (fp + 4.0)
    qu[:d.ny + 2, :d.nx + 2] = qh[:d.ny + 2, :d.nx + 2]


class PySmallScale:
    """SmallScale (src/small_scale.f:160-581) with its `save`d state (map iterates, cell areas)."""
    AR, AM, RC = 4.82842712474, 1.47839783948, 0.20710678119

    def __init__(self, d):
        self.d = d
        self.umap = np.zeros((3,) + d.new_field().shape)
        self.vmap, self.tmap = self.umap.copy(), self.umap.copy()
        self.area = d.new_field()
        self.tarea = 0.0

    def call(self, initflg, u1, v1, t1, uss, vss, pss, tss):
        import math
        from wolfd2_b200 import deck as dk
        d = self.d
        nx, ny, m, r = d.nx, d.ny, d.metrics, d.regions
        AR, AM, RC = self.AR, self.AM, self.RC
        fp = d.ss_filt
        cu0, TsCoef, HsCoef, TemCoef = d.ss_cu0, d.ss_tscoef, d.ss_hscoef, d.ss_temcoef
        bnumc, rmax, rlc = d.ss_bncrit, d.ss_rmpmax, d.ss_rmpexp
        dlref, uref, tref, tmax, dka, re, pe = d.dlref, d.uref, d.tref, d.tmax, d.dk, d.re, d.pe
        piosr2 = math.acos(-1.0) / math.sqrt(2.0)
        pehmin = 3.0
        hs = dlref * HsCoef
        dk_ = dka * dlref / uref
        pr = pe / re
        rnu = uref * dlref / re
        dkappa = rnu / pr
        djc = m["djc"]
        if initflg <= 0:
            ump, vmp, tmp = 0.92, 0.31, 0.50
            for i in range(0, nx + 2):
                for j in range(0, ny + 2):
                    for l in range(3):
                        ump = RC * AR * ump * (1.0 - AM * abs(ump))
                        vmp = RC * AR * vmp * (1.0 - AM * abs(vmp))
                        tmp = RC * AR * tmp * (1.0 - AM * abs(tmp))
                        self.umap[l, j, i], self.vmap[l, j, i], self.tmap[l, j, i] = ump, vmp, tmp
                    uss[j, i] = vss[j, i] = tss[j, i] = 0.0
            self.tarea = 0.0
            for i in range(1, nx + 1):
                for j in range(1, ny + 1):
                    self.area[j, i] = 1.0 / djc[j, i]
                    self.tarea = self.tarea + self.area[j, i]
            if initflg < 0:
                return
        z = d.new_field
        xz, xe, yz, ye, rj, vl, ul, tl, uc, vc = (z() for _ in range(10))
        W = Rng(1, nx, 1, ny)
        W.put(xz, dlref * W(m["xzc"])); W.put(xe, dlref * W(m["xec"]))
        W.put(yz, dlref * W(m["yzc"])); W.put(ye, dlref * W(m["yec"]))
        W.put(rj, W(djc) / (dlref * dlref))
        W.put(vl, uref * (W(v1) + W(v1, 0, -1)) * 0.5)
        W.put(ul, uref * (W(u1) + W(u1, -1, 0)) * 0.5)
        W.put(tl, (tmax - tref) * W(t1) + tref)
        W.put(uc, W(ye) * W(ul) - W(xe) * W(vl))
        W.put(vc, W(xz) * W(vl) - W(yz) * W(ul))
        uf, vf, tf = z(), z(), z()
        W.put(uf, W(uc)); W.put(vf, W(vc)); W.put(tf, W(tl))
        # the work arrays are static in the reference: uf, vf, tf keep what earlier calls left outside 1..nx, 1..ny
        for name, f in (("uf", uf), ("vf", vf), ("tf", tf)):
            old = getattr(self, name, None)
            if old is not None:
                keep = np.ones_like(f, dtype=bool)
                keep[1:ny + 1, 1:nx + 1] = False
                f[keep] = old[keep]
        py_tempbc(d, tf)
        py_filter(d, 1, fp[0], uf)
        py_filter(d, 2, fp[1], vf)
        py_filter_t(d, fp[3], tf)
        W.put(uf, W(uc) - W(uf)); W.put(vf, W(vc) - W(vf)); W.put(tf, W(tl) - W(tf))
        self.uf, self.vf, self.tf = uf, vf, tf
        s15 = math.sqrt(15.0)
        for jr in range(int(r.nReg[1])):
            for ir in range(int(r.nReg[0])):
                iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
                if int(r.nRegType[jr, ir]) == dk.RM_BLOCKG or int(r.nTRgType[jr, ir]) == dk.RT_TEMPER:
                    continue
                for i in range(iW + 1, iE + 1):
                    for j in range(jS + 1, jN + 1):
                        XZ, YZ, XE, YE, RJ = xz[j, i], yz[j, i], xe[j, i], ye[j, i], rj[j, i]
                        hxy = math.sqrt(XZ * XZ + YZ * YZ + XE * XE + YE * YE)
                        uz = (ul[j, i + 1] - ul[j, i - 1]) * 0.5; ue = (ul[j + 1, i] - ul[j - 1, i]) * 0.5
                        vz = (vl[j, i + 1] - vl[j, i - 1]) * 0.5; ve = (vl[j + 1, i] - vl[j - 1, i]) * 0.5
                        sq = lambda x: x * x
                        uxsq = sq(RJ * (YE * uz - YZ * ue)); uysq = sq(RJ * (XZ * ue - XE * uz))
                        vxsq = sq(RJ * (YE * vz - YZ * ve)); vysq = sq(RJ * (XZ * ve - XE * vz))
                        delu2n = math.sqrt(uxsq + uysq + vxsq + vysq)
                        tz = (tl[j, i + 1] - tl[j, i - 1]) * 0.5; te = (tl[j + 1, i] - tl[j - 1, i]) * 0.5
                        txsq = sq(RJ * (YE * tz - YZ * te)); tysq = sq(RJ * (XZ * te - XE * tz))
                        delt2n = math.sqrt(txsq + tysq)
                        reh = delu2n * (hxy * hxy) / rnu
                        peh = pr * reh
                        if not peh > pehmin:
                            continue
                        cuU, cuV, cuT = cu0 * TemCoef, cu0, cu0
                        if cu0 > float(np.float32(1.e-10)):
                            ts = TsCoef * piosr2 * math.pow(reh, 1.0 / 3.0) / (cu0 * delu2n)
                            nmap = int(1.0 + dk_ / ts)
                        else:
                            nmap = 0
                        nmap = min(nmap, 50)
                        bnum = s15 * math.pow(hs * hs * delu2n / rnu, 1.0 / 6.0)
                        rmap = rmax * math.tanh(math.pow(bnum / bnumc, rlc) * math.atanh(RC / rmax))
                        uz = (uf[j, i + 1] - uf[j, i - 1]) * 0.5; ue = (uf[j + 1, i] - uf[j - 1, i]) * 0.5
                        vz = (vf[j, i + 1] - vf[j, i - 1]) * 0.5; ve = (vf[j + 1, i] - vf[j - 1, i]) * 0.5
                        tz = (tf[j, i + 1] - tf[j, i - 1]) * 0.5; te = (tf[j + 1, i] - tf[j - 1, i]) * 0.5
                        gu1 = RJ * (YE * uz - YZ * ue); gu2 = RJ * (XZ * ue - XE * uz)
                        gv1 = RJ * (YE * vz - YZ * ve); gv2 = RJ * (XZ * ve - XE * vz)
                        gt1 = RJ * (YE * tz - YZ * te); gt2 = RJ * (XZ * te - XE * tz)
                        grduf = math.sqrt(gu1 * gu1 + gu2 * gu2)
                        grdvf = math.sqrt(gv1 * gv1 + gv2 * gv2)
                        grdtf = math.sqrt(gt1 * gt1 + gt2 * gt2)
                        grd = math.sqrt(grduf * grduf + grdvf * grdvf + grdtf * grdtf)
                        with np.errstate(all="ignore"):
                            s1, s2 = np.float64(grduf) / grd, np.float64(grdvf) / grd
                            rnrmjs = np.sqrt(sq(XZ * s1 + XE * s2) + sq(YZ * s1 + YE * s2))
                            zeta1 = math.sqrt(2.0) * s1 / rnrmjs
                            zeta2 = math.sqrt(2.0) * s2 / rnrmjs
                        zeta3 = 1.0
                        a11, a12 = (gu1 / grduf, gu2 / grduf) if abs(grduf) > 0.0 else (0.0, 0.0)
                        a21, a22 = (gv1 / grdvf, gv2 / grdvf) if abs(grdvf) > 0.0 else (0.0, 0.0)
                        if abs(grdtf) > 0.0:
                            a31, a32, a33 = gu1 / grdtf, gu2 / grdtf, math.sqrt(gt1 * gt1 + gt2 * gt2) / grdtf
                        else:
                            a31 = a32 = a33 = 0.0
                        mp = [self.umap[l, j, i] for l in range(3)] + [self.vmap[l, j, i] for l in range(3)] + \
                             [self.tmap[l, j, i] for l in range(3)]
                        for _ in range(nmap):
                            mp = [rmap * AR * x * (1.0 - AM * abs(x)) for x in mp]
                        for l in range(3):
                            self.umap[l, j, i], self.vmap[l, j, i], self.tmap[l, j, i] = mp[l], mp[3 + l], mp[6 + l]
                        um = a11 * mp[0] + a12 * mp[1]
                        vm = a21 * mp[0] + a22 * mp[1]
                        tm = a31 * mp[6] + a32 * mp[7] + a33 * mp[8]
                        r16 = math.pow(reh, 1.0 / 6.0)
                        av = cuV * r16 * math.sqrt(rnu * delu2n)
                        at = math.pow(3.0 * math.pow(cuT, 4.0) * peh / pr, 1.0 / 6.0) * math.sqrt(dkappa)
                        with np.errstate(all="ignore"):
                            at = np.float64(at) * delt2n / math.sqrt(delu2n) * TemCoef
                        av = av * math.sqrt(self.area[j, i] / self.tarea) * math.pow(hxy, 1.0 / 3.0)
                        at = at * math.sqrt(self.area[j, i] / self.tarea) * math.pow(hxy, 1.0 / 3.0)
                        uscon = av * zeta1 * um
                        vscon = av * zeta2 * vm
                        uss[j, i] = RJ * (XZ * uscon + YE * vscon) / uref
                        vss[j, i] = RJ * (YE * vscon + YZ * uscon) / uref
                        tss[j, i] = (at * tm * zeta3) / (tmax - tref)
                        if abs(uss[j, i]) < 1.0e-14: uss[j, i] = 0.0
                        if abs(vss[j, i]) < 1.0e-14: vss[j, i] = 0.0
                        if abs(tss[j, i]) < 1.0e-14: tss[j, i] = 0.0
        py_smlsclbc(d, uss, vss, pss, tss)
        pss[0:ny + 1, 0:nx + 1] = 0.0
        py_smlsclbc(d, uss, vss, pss, tss)
        sv = (d.msorit, d.sortol, d.sorrel)
        d.msorit, d.sortol, d.sorrel = d.ss_msorit, d.ss_sortol, d.ss_sorrel
        try:
            np_ppe_general(d, uss, vss, pss)
        finally:
            d.msorit, d.sortol, d.sorrel = sv
        py_smlsclbc(d, uss, vss, pss, tss)
        py_project(d, pss, uss, vss)


def _ss_call_args(d, initflg):
    from wolfd2_b200.deck import PPE_SOLVERS
    r, m = d.regions, d.metrics
    fp = np.array(d.ss_filt, dtype=np.float64)
    return (d.nx, d.ny, initflg, int(d.thermal), int(d.cartesian), r.nReg, r.nRegBrd, r.nRegType, r.nTRgType, r.nMomBdTp,
            r.nTemBdTp, PPE_SOLVERS[d.ss_ppe_solver], d.ss_msorit, d.dlref, d.uref, d.tref, d.tmax, d.dk, d.re, d.pe,
            d.ss_sortol, d.ss_sorrel, fp, d.ss_cu0, d.ss_tscoef, d.ss_hscoef, d.ss_temcoef, d.ss_bncrit, d.ss_rmpmax,
            d.ss_rmpexp, r.dTRgVal, r.dBCVal, m["rau"], m["rbu"], m["rbv"], m["rgv"], m["dju"], m["djv"], m["djc"],
            m["xeu"], m["yeu"], m["xzv"], m["yzv"], m["xzu"], m["yzu"], m["xev"], m["yev"], m["xec"], m["yec"],
            m["xzc"], m["yzc"])


def _ss_decks():
    from wolfd2_b200 import deck as dk
    kw = dict(smallscale=True, ss_cu0=1.0, ss_bncrit=2.0, ss_rmpmax=0.95, ss_ppe_solver="rb_sor", ss_msorit=120,
              ss_sortol=1e-9, ss_sorrel=1.5, ss_filt=(4e2, 3e2, 0.0, 2e2))
    out = [dk.cavity(25, re=1000.0, dt=0.01, ny=21, **kw)]
    reg = dk.RegionTables(30, 26, 2, 2, (14,), (12,))
    reg.heat_generation(1, 1, 2.0).fixed_temperature_region(2, 2, 0.7)
    reg.wall_temperature(1, 1, "w", 1.0).wall_heat_flux(1, 2, "w", 0.05).wall(1, 2, "n", tangent_vel=1.0)
    out.append(dk._mk("atd_thermal_2x2", 30, 26, reg, 900.0, 0.004, thermal=True, nmeiter=2, **kw))
    reg = dk.RegionTables(30, 22, 2, 1, (14,), ())
    reg.inlet(1, 1, "w", normal_vel=1.0).outlet(2, 1, "e", fully_dev=False).outlet(2, 1, "n", fully_dev=True)
    out.append(dk._mk("atd_channel", 30, 22, reg, 700.0, 0.004, **kw))
    return out


@pytest.mark.parametrize("k", range(3))
def test_smallscale_second_restatement(orc, k):
    """SmallScale through three calls (seeding, then two advances on changing fields): the high-pass filter, the
    per-cell chaotic-map model with pow / tanh / atanh from the same libm, the clamp nmap <= 50, `av` used for both
    velocity components (:505-506), the 1e-14 flush, SmlSclBC, the small-scale Ppe and projection -- uss, vss, pss,
    tss and all nine saved map planes bit for bit.  Cold cavity; thermal 2x2 deck with a heat source, a
    fixed-temperature block, temperature and flux faces; channel with an inlet and both outlet types."""
    d = _ss_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(900 + k)
    mine = PySmallScale(d)
    g = [d.new_field() for _ in range(4)]
    o = [d.new_field() for _ in range(4)]
    for call, initflg in enumerate((0, 1, 1)):
        u, v, t = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng, 0.0, 1.0)
        mine.call(initflg, u, v, t, *g)
        orc.smallscale(*_ss_call_args(d, initflg), u, v, t, *o)
        assert orc.lib.orc_get_errflag() == 0
        for name, a, b in zip(("uss", "vss", "pss", "tss"), g, o):
            assert np.array_equal(a, b), (call, name, np.abs(a - b).max())
        n = d.new_field().size
        for fam, mp in enumerate((mine.umap, mine.vmap, mine.tmap)):
            for l in range(3):
                ref = np.ctypeslib.as_array(orc.lib.orc_ss_map(fam, l + 1), shape=(n,)).reshape(d.new_field().shape)
                assert np.array_equal(mp[l], ref), (call, fam, l)
    assert np.abs(g[0]).max() > 0 and np.abs(g[3]).max() >= 0


# ------------------------------------------------------------------ Lagrangian particles
def py_ifindpos(nx, ny, xp, yp, x, y):
    """iFindPos (src/traject.f:507-560): first node, j outer / i inner, with x > xp and y > yp."""
    ip = jp = -1
    xrel = yrel = 0.0
    for j in range(1, ny + 1):
        for i in range(1, nx + 1):
            if x[j, i] > xp and y[j, i] > yp:
                ip, jp = i, j
                xrel, yrel = x[j, i] - xp, y[j, i] - yp
                break
        if ip >= 0:
            break
    flag = 0
    if ip <= 1 and xrel >= 0.0: flag = 1
    if jp <= 1 and yrel >= 0.0: flag = 3
    if ip < 0 or jp < 0: flag = 2
    return flag, ip, jp


def py_bilin(i, j, xs, ys, x, y, u):
    """BiLinInterp (src/traject.f:586-625)."""
    x1, x2, x3, x4 = x[j - 1, i - 1], x[j - 1, i], x[j, i], x[j, i - 1]
    y1, y2, y3, y4 = y[j - 1, i - 1], y[j - 1, i], y[j, i], y[j, i - 1]
    f1, f2, f3, f4 = u[j - 1, i - 1], u[j - 1, i], u[j, i], u[j, i - 1]
    fa = f1 + (xs - x1) * (f2 - f1) / (x2 - x1)
    fb = f2 + (ys - y2) * (f3 - f2) / (y3 - y2)
    fc = f4 + (xs - x4) * (f3 - f4) / (x3 - x4)
    fd = f1 + (ys - y1) * (f4 - f1) / (y4 - y1)
    ya = y1 + (xs - x1) * (y2 - y1) / (x2 - x1)
    xb = x2 + (ys - y2) * (x3 - x2) / (y3 - y2)
    yc = y4 + (xs - x4) * (y3 - y4) / (x3 - x4)
    xd = x1 + (ys - y1) * (x4 - x1) / (y4 - y1)
    fsx = fd + (xs - xd) * (fb - fd) / (xb - xd)
    fsy = fa + (ys - ya) * (fc - fa) / (yc - ya)
    return (fsx + fsy) / 2.0


def _trajfunc(i, fr, uf, vf, cpx, cpy, w):
    if i == 0: return w[1]
    if i == 1: return cpx * abs(uf - w[1]) * (uf - w[1])
    if i == 2: return w[3]
    return cpy * abs(vf - w[3]) * (vf - w[3]) - 1.0 / fr


def py_gauss(a, b):
    """Gauss (src/traject.f:640-690): partial pivoting, in place; returns x."""
    n = len(b)
    for k in range(n - 1):
        amax, imax = abs(a[k][k]), k
        for i in range(k + 1, n):
            if abs(a[i][k]) > amax:
                amax, imax = abs(a[i][k]), i
        if imax != k:
            for j in range(k, n):
                a[k][j], a[imax][j] = a[imax][j], a[k][j]
            b[k], b[imax] = b[imax], b[k]
        for i in range(k + 1, n):
            dm = a[i][k] / a[k][k]
            b[i] = b[i] - dm * b[k]
            for j in range(k + 1, n):
                a[i][j] = a[i][j] - dm * a[k][j]
    x = [0.0] * n
    x[n - 1] = b[n - 1] / a[n - 1][n - 1]
    for i in range(n - 2, -1, -1):
        s = 0.0
        for j in range(i + 1, n):
            s = s + a[i][j] * x[j]
        x[i] = (b[i] - s) / a[i][i]
    return x


def py_heuntrap(maxit, h, toler, delta, fr, uf1, vf1, ufn, vfn, cpx1, cpxn, cpy1, cpyn, u):
    """HeunTrap (src/traject.f:336-415); h is the sub-step size (the definition of DESIGN.md for the reference's
    INTEGER-for-REAL argument at :281)."""
    h2 = h / 2.0
    g = [0.0] * 4
    us = [0.0] * 4
    for m in range(1, maxit + 1):
        if m == 1:
            for i in range(4):
                g[i] = _trajfunc(i, fr, ufn, vfn, cpxn, cpyn, u)
                us[i] = u[i] + h * g[i]
            if maxit == 1:
                u[:] = us
                return
            for i in range(4):          # us is updated in place while TrajFunc reads it, as written
                us[i] = u[i] + h2 * (g[i] + _trajfunc(i, fr, uf1, vf1, cpx1, cpy1, us))
        dj = [[0.0] * 4 for _ in range(4)]
        dj[0][1] = 1.0
        dj[1][1] = cpx1 * ((uf1 - us[1]) - abs(uf1 - us[1]))
        dj[2][3] = 1.0
        dj[3][3] = cpy1 * ((vf1 - us[3]) - abs(vf1 - us[3]))
        f = [0.0] * 4
        for i in range(4):
            f[i] = us[i] - h2 * _trajfunc(i, fr, uf1, vf1, cpx1, cpy1, us) - (u[i] + h2 * g[i])
            for j in range(4):
                dj[i][j] = (1.0 if i == j else 0.0) - h2 * dj[i][j]
        f = [-q for q in f]
        udel = py_gauss(dj, f)
        dumax = 0.0
        for i in range(4):
            dumax = max(dumax, abs(udel[i]))
            us[i] = us[i] + delta * udel[i]
        if dumax < toler:
            u[:] = us
            return


def py_traject(d, ntr, nsub, method, cdeq, maxit, out, dkflow, densref, fr, tol, delta, cx, cy, repc, x, y, u, v, un, vn,
               dens, densn, xp, yp, up, vp):
    """Traject (src/traject.f:195-300).  A particle whose first iFindPos is non-zero is only flagged (the reference
    integrates a copy and discards it at :296)."""
    import math
    dk_ = dkflow / float(nsub) if nsub != 1 else dkflow
    for k in range(1, nsub + 1):
        for l in range(ntr):
            if out[l] > 0:
                continue
            out[l], ip, jp = py_ifindpos(d.nx, d.ny, xp[l], yp[l], x, y)
            if out[l] > 0:
                continue
            uf1, vf1 = py_bilin(ip, jp, xp[l], yp[l], x, y, u), py_bilin(ip, jp, xp[l], yp[l], x, y, v)
            ufn, vfn = py_bilin(ip, jp, xp[l], yp[l], x, y, un), py_bilin(ip, jp, xp[l], yp[l], x, y, vn)
            df1, dfn = py_bilin(ip, jp, xp[l], yp[l], x, y, dens), py_bilin(ip, jp, xp[l], yp[l], x, y, densn)
            df1, dfn = densref * (df1 + 1.0), densref * (dfn + 1.0)
            rep = repc[l] * math.sqrt((uf1 - up[l]) * (uf1 - up[l]) + (vf1 - vp[l]) * (vf1 - vp[l]))
            dst = 24.0 / rep
            if cdeq == 1: cd = dst
            elif cdeq == 2: cd = dst * (1.0 + math.pow(rep, 2.0 / 3.0) / 6.0)
            elif cdeq == 3: cd = 0.40 + dst + 6.0 / (1.0 + math.sqrt(rep))
            else: cd = dst * (1.0 + 0.1970 * math.pow(rep, 0.63) + 0.26e-3 * math.pow(rep, 1.38))
            cpx1, cpxn, cpy1, cpyn = cd * cx[l] * df1, cd * cx[l] * dfn, cd * cy[l] * df1, cd * cy[l] * dfn
            w = [xp[l], up[l], yp[l], vp[l]]
            if method == 1:
                py_heuntrap(maxit, dk_, tol, delta, fr, uf1, vf1, ufn, vfn, cpx1, cpxn, cpy1, cpyn, w)
            else:           # FwdEuler, :305-320
                w[1] = w[1] + dk_ * (cpxn * abs(ufn - w[1]) * (ufn - w[1]))
                w[0] = w[0] + dk_ * w[1]
                w[3] = w[3] + dk_ * (cpyn * abs(vfn - w[3]) * (vfn - w[3]) - (1.0 / fr))
                w[2] = w[2] + dk_ * w[3]
            xp[l], up[l], yp[l], vp[l] = w


@pytest.mark.parametrize("method", [1, 2])
@pytest.mark.parametrize("cdeq", [1, 2, 3, 4])
def test_traject_second_restatement(orc, method, cdeq):
    """Traject with both integrators and all four drag laws on a stretched (non-uniform) grid: first-match cell
    search, bilinear interpolation, Heun / trapezoidal Newton iterations with Gauss elimination, forward Euler;
    particles starting outside the domain are flagged.  Positions, velocities and flags bit for bit."""
    from wolfd2_b200 import deck as dk
    x, y = dk.stretched_grid(30, 24)
    reg = dk.RegionTables(30, 24).wall(1, 1, "n", tangent_vel=1.0)
    d = dk._mk("traj", 30, 24, reg, 100.0, 0.02, x=x, y=y, cartesian=False)
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(1000 + 10 * method + cdeq)
    n = 40
    gx, gy = d.node_arrays()
    lx, ly = gx.max(), gy.max()
    xp, yp = rng.uniform(0.05 * lx, 0.95 * lx, n), rng.uniform(0.05 * ly, 0.95 * ly, n)
    xp[:4] = (-0.01, 1.5 * lx, 0.3 * lx, 0.5 * lx)
    yp[:4] = (0.4 * ly, 0.5 * ly, -0.2 * ly, 2.0 * ly)
    up, vp = rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n)
    cx, cy, repc = rng.uniform(0.5, 3.0, n), rng.uniform(0.5, 3.0, n), rng.uniform(5.0, 50.0, n)
    f = [rand_field(d, rng, -0.5, 0.5) for _ in range(4)] + [rand_field(d, rng, -0.05, 0.05) for _ in range(2)]
    mine = [a.copy() for a in (xp, yp, up, vp)]
    ref = [a.copy() for a in (xp, yp, up, vp)]
    out_m, out_r = np.zeros(n, np.int32), np.zeros(n, np.int32)
    for call in range(2):
        py_traject(d, n, 3, method, cdeq, 6, out_m, d.dk, 1.2, d.fr, 1e-10, 1.0, cx, cy, repc, gx, gy, *f, *mine)
        orc.traject(d.nx, d.ny, n, 3, method, cdeq, 6, out_r, d.dk, 1.2, d.fr, 1e-10, 1.0, cx, cy, repc, gx, gy, *f, *ref)
        assert np.array_equal(out_m, out_r), call
        for name, a, b in zip("xp yp up vp".split(), mine, ref):
            assert np.array_equal(a, b), (call, name, np.abs(a - b).max())
    assert set(out_m[:4]) >= {2, 3} and (out_m[4:] == 0).sum() > 20      # (on this skewed grid the first match of
    # the particle left of the domain is a node further along a lower row: flag 0, as the reference would)
    assert not np.array_equal(mine[0][4:], xp[4:])


# ------------------------------------------------------------------ time steps with the momentum-energy iterations
def np_step_thermal(d, u, v, p, t, den, ss=None):
    """src/main.f:690-972 with the thermal energy equation on: the momentum-energy iteration loop (:736-880) around
    nAuxMomentum (buoyancy from d, dn) ... Project, ThermEnergy, EqState, its convergence rule, Filter(_T_),
    TempBoundCond, four norms."""
    nx, ny = d.nx, d.ny
    pn, un, vn, tn, dn = p.copy(), u.copy(), v.copy(), t.copy(), den.copy()
    if ss is not None:          # src/main.f:706-727: the time-level-n fields carry the small scales once more
        model, uss, vss, pss, tss = ss
        usn, vsn, tsn = uss.copy(), vss.copy(), tss.copy()
        un += uss; vn += vss; tn += tss
    nmeiter = d.nmeiter if d.thermal else min(d.nmeiter, 1)
    nql = nsor = None
    for l in range(1, nmeiter + 1):
        us, vs, ts = u.copy(), v.copy(), t.copy()
        us[1:ny + 2, 1:nx + 2] = un[1:ny + 2, 1:nx + 2]
        vs[1:ny + 2, 1:nx + 2] = vn[1:ny + 2, 1:nx + 2]
        nql = -1
        for it in range(1, d.mqiter + 1):
            py_velbc(d, us, vs, outflow_only=True)
            dus = np_xmomentum(d, us, vs, un, vn)
            dvs = np_ymomentum(d, us, vs, un, vn, den, dn)
            us[1:ny + 1, 1:nx + 1] += dus[1:ny + 1, 1:nx + 1]
            vs[1:ny + 1, 1:nx + 1] += dvs[1:ny + 1, 1:nx + 1]
            if max(np_dmaxnorm(nx, ny, dus), np_dmaxnorm(nx, ny, dvs)) <= d.qtol:
                nql = it
                break
        py_velbc(d, us, vs, False)
        py_presbc(d, p)
        nsor = np_ppe_general(d, us, vs, p)
        py_presbc(d, p)
        py_project(d, p, us, vs)
        py_velbc(d, us, vs, False)
        py_presbc(d, p)
        if d.thermal:
            np_thermenergy(d, un, vn, us, vs, tn, ts)
        if d.eqstate:
            np_eqstate(d, p, ts, den)
        dif = [np_diffmaxnorm(nx, ny, u, us), np_diffmaxnorm(nx, ny, v, vs), np_diffmaxnorm(nx, ny, t, ts)]
        u[:ny + 2, :nx + 2] = us[:ny + 2, :nx + 2]
        v[:ny + 2, :nx + 2] = vs[:ny + 2, :nx + 2]
        t[:ny + 2, :nx + 2] = ts[:ny + 2, :nx + 2]
        if l > 1 and max(dif) < d.dmeittol:
            break
    if d.thermal and d.nfiltt == 1:
        py_filter_t(d, d.fpt, t)
    if ss is not None:          # src/main.f:896-940: strip the old small scales, advance the model, add the new ones
        W = (slice(0, ny + 2), slice(0, nx + 2))
        u[W] = u[W] - usn[W]; v[W] = v[W] - vsn[W]; t[W] = t[W] - tsn[W]
        model.call(1, u, v, t, uss, vss, pss, tss)
        u[W] = u[W] + uss[W]; v[W] = v[W] + vss[W]; t[W] = t[W] + tss[W]
    py_velbc(d, u, v, False)
    py_presbc(d, p)
    if d.thermal:
        py_tempbc(d, t)
    dif = [np_diffmaxnorm(nx, ny, pn, p), np_diffmaxnorm(nx, ny, un, u), np_diffmaxnorm(nx, ny, vn, v),
           np_diffmaxnorm(nx, ny, tn, t)]
    return nql, nsor, dif


@pytest.mark.parametrize("k", range(2))
def test_thermal_time_steps_second_restatement(orc, k):
    """Three time steps with the thermal energy equation (2-3 momentum-energy iterations per step, buoyancy through
    EqState, heat sources, a fixed-temperature block): u, v, p, t, d bit for bit, counts and the four-column
    PrintDiff tuple identical."""
    import dataclasses
    d = _thermal_decks()[k]
    d = dataclasses.replace(d, msorit=60, sortol=1e-7, sorrel=1.5, mqiter=5, qtol=1e-6, nfiltt=1 if k == 1 else 0, fpt=300.0)
    orc.config(d.mnx, d.mny)
    u, v, p, t, den = (d.new_field() for _ in range(5))
    t[:d.ny + 2, :d.nx + 2] = 0.5
    uo, vo, po, to, do = u.copy(), v.copy(), p.copy(), t.copy(), den.copy()
    for step in range(3):
        nql, nsor, dif = np_step_thermal(d, u, v, p, t, den)
        rc, lg = orc.step(d, uo, vo, po, 1, t=to, d=do)
        assert rc == 0
        assert (nql, nsor) == (lg[0]["nQLiter"], lg[0]["nSorConv"]), (step, nql, nsor)
        for name, a, b in zip("uvptd", (u, v, p, t, den), (uo, vo, po, to, do)):
            assert np.array_equal(a, b), (step, name, np.abs(a - b).max())
        assert dif == list(lg[0]["dif"]), step
    assert np.abs(t[2:d.ny + 1, 2:d.nx + 1] - 0.5).max() > 1e-4


@pytest.mark.parametrize("k", range(2))
def test_atd_time_steps_second_restatement(orc, k):
    """Time steps with the ATD small-scale model inside (src/main.f:643-665, :706-727, :896-940) on a developed
    vortex: cold cavity and thermal 2x2 deck.  u, v, p, t, d, uss, vss, pss, tss bit for bit."""
    import dataclasses
    d = _ss_decks()[k]
    d = dataclasses.replace(d, ss_cu0=0.05, msorit=60, sortol=1e-7, sorrel=1.5, mqiter=5, qtol=1e-6, ss_msorit=60)
    orc.config(d.mnx, d.mny)
    gx, gy = d.node_arrays()
    X, Y = gx / gx.max(), gy / gy.max()
    u, v, p, t, den = (d.new_field() for _ in range(5))
    u[:] = np.sin(np.pi * X) ** 2 * np.sin(2 * np.pi * Y)
    v[:] = -np.sin(2 * np.pi * X) * np.sin(np.pi * Y) ** 2
    py_velbc(d, u, v, False)
    uo, vo, po, to, do = (a.copy() for a in (u, v, p, t, den))
    ss_m = [d.new_field() for _ in range(4)]
    ss_o = [d.new_field() for _ in range(4)]
    model = PySmallScale(d)
    model.call(0, u, v, t, *ss_m)
    orc.atd_init(d, uo, vo, to, *ss_o)
    for a, b in zip(ss_m, ss_o):
        assert np.array_equal(a, b)
    active = False
    for step in range(3):
        nql, nsor, dif = np_step_thermal(d, u, v, p, t, den, ss=(model, *ss_m))
        rc, lg = orc.step_full(d, uo, vo, po, to, do, ss_fields=ss_o, nsteps=1)
        assert rc == 0
        assert (nql, nsor) == (lg[0]["nQLiter"], lg[0]["nSorConv"]), (step, nql, nsor)
        names = "u v p t d uss vss pss tss".split()
        for name, a, b in zip(names, (u, v, p, t, den, *ss_m), (uo, vo, po, to, do, *ss_o)):
            assert np.array_equal(a, b), (step, name, np.abs(a - b).max())
        assert dif == list(lg[0]["dif"]), step
        active = active or np.abs(ss_m[0]).max() > 1e-8
    assert active


# ------------------------------------------------------------------ all six solvers, Cartesian flag on and off
def py_ppe_any_solver(d, solver, cartes, u, v, p, msorit):
    """Ppe on a one-region grid (src/pressure.f:90-246) with solver ids 1..6: Sor :411-446, Slor :704-799, SlorRB
    :873-955, SlorRBP :1026-1133, SorRB :478-541, SorRBP :576-652.  With lCartesGrid false the right-hand side is
    rebuilt from the current p where each routine says so."""
    nx, ny, m = d.nx, d.ny, d.metrics
    rau, rgv, rbu, rbv = m["rau"], m["rgv"], m["rbu"], m["rbv"]
    div = d.new_field()
    np_divergence(nx, ny, 1, m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, div)
    b = np.zeros(d.mnx * d.mny)
    rhs = lambda: np_rhsppe(nx, ny, cartes, d.dk, rbu, rbv, div, p, b)
    ind = lambda i, j: (j - 2) * (nx - 1) + i - 2
    coef = lambda i, j: (rgv[j - 1, i], rau[j, i - 1], -rau[j, i] - rau[j, i - 1] - rgv[j, i] - rgv[j - 1, i], rau[j, i], rgv[j, i])

    def point(i, j):
        a1, a2, a3, a4, a5 = coef(i, j)
        s = b[ind(i, j)] - a1 * p[j - 1, i] - a2 * p[j, i - 1] - a4 * p[j, i + 1] - a5 * p[j + 1, i]
        s = s / a3 - p[j, i]
        p[j, i] = p[j, i] + d.sorrel * s
        return abs(s)

    def line(j):
        al, bl = [], []
        for i in range(2, nx + 1):
            a1, a2, a3, a4, a5 = coef(i, j)
            al.append([a2, a3, a4])
            bl.append(b[ind(i, j)] - (a1 * p[j - 1, i] + a5 * p[j + 1, i]))
        py_alttridlu(al, bl)
        return bl
    if cartes:
        rhs()
    for it in range(1, msorit + 1):
        dif = 0.0
        if solver == 1:
            if not cartes: rhs()
            for j in range(2, ny + 1):
                for i in range(2, nx + 1):
                    dif = max(dif, point(i, j))
        elif solver in (5, 6):
            if not cartes: rhs()
            for par in (0, 1):                  # black: i starts at 2 + mod(j, 2); red: 2 + mod(j + 1, 2)
                for j in range(2, ny + 1):
                    for i in range(2 + (j + par) % 2, nx + 1, 2):
                        dif = max(dif, point(i, j))
        elif solver == 2:
            if not cartes: rhs()
            for j in range(2, ny + 1):
                old = [p[j, i] for i in range(2, nx + 1)]
                bl = line(j)
                for k in range(nx - 1):
                    s = bl[k] - old[k]
                    p[j, 2 + k] = old[k] + d.sorrel * s
                    dif = max(dif, abs(s))
        elif solver == 3:
            for k0 in (2, 3):
                if not cartes: rhs()
                for j in range(k0, ny + 1, 2):
                    old = [p[j, i] for i in range(2, nx + 1)]
                    bl = line(j)
                    for k in range(nx - 1):
                        s = bl[k] - old[k]
                        p[j, 2 + k] = old[k] + d.sorrel * s
                        dif = max(dif, abs(s))
        else:                                   # 4: SlorRBP, relaxation after all lines of a colour are solved
            pn = p.copy()
            for k0 in (2, 3):
                if not cartes: rhs()
                for j in range(k0, ny + 1, 2):
                    bl = line(j)
                    for k in range(nx - 1):
                        p[j, 2 + k] = bl[k]
                for j in range(k0, ny + 1, 2):
                    for i in range(2, nx + 1):
                        p[j, i] = pn[j, i] + d.sorrel * (p[j, i] - pn[j, i])
            dif = np.abs(pn[2:ny + 1, 2:nx + 1] - p[2:ny + 1, 2:nx + 1]).max()
        if it > 1 and dif < d.sortol:
            return it
    return msorit


@pytest.mark.parametrize("solver", [1, 2, 3, 4, 5, 6])
@pytest.mark.parametrize("cartes", [1, 0])
def test_all_ppe_solvers_second_restatement(orc, solver, cartes):
    """Every ppe_solver id through Ppe on a skewed grid, with the Cartesian flag on (right-hand side built once)
    and off (rebuilt from the current iterate, cross-derivative terms included): iterate path, iteration count and
    result bit for bit -- at 1 iteration, at a few, and to convergence."""
    from wolfd2_b200 import deck as dk
    x, y = dk.stretched_grid(22, 18)
    d = dk._mk("skew", 22, 18, dk.RegionTables(22, 18), 100.0, 0.01, x=x, y=y, cartesian=bool(cartes))
    d.sorrel, d.sortol = 1.25, 1e-8
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(20 + solver)
    m, r = d.metrics, d.regions
    assert np.abs(m["rbu"]).max() > 1e-4          # the grid really is non-orthogonal
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for msorit in (1, 4, 600):
        po = p.copy()
        nconv = orc.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cartes, solver, msorit, d.dk, d.sortol, d.sorrel, m["rau"],
                        m["rbu"], m["rbv"], m["rgv"], m["xeu"], m["yeu"], m["xzv"], m["yzv"], u, v, po)
        pn = p.copy()
        n = py_ppe_any_solver(d, solver, cartes, u, v, pn, msorit)
        assert n == nconv, (msorit, n, nconv)
        assert np.array_equal(pn, po), (msorit, np.abs(pn - po).max())


# ------------------------------------------------------------------ SmlSclBC with outlets
def py_smlsclbc(d, u, v, p, t):
    """SmlSclBC (src/bound_cond.f:1257-1649): homogeneous velocity ghosts per face type, as written (east OUTLT1
    assigns u(iE+1,j) to itself; the south outlets write row jS-1), then PresBoundCond and homogeneous temperature
    ghosts."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    O1, O2 = dk.BM_OUTLT1, dk.BM_OUTLT2
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            bd = lambda face: int(r.nMomBdTp[face - 1, jr, ir])
            for face in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH):
                tp = bd(face)
                if tp in (dk.BM_WALL1, dk.BM_WALL2, dk.BM_INLET):
                    sgn = 1.0 if tp == dk.BM_WALL2 else -1.0
                    if face == dk.WEST:
                        u[jS:jN + 1, iW] = 0.0
                        v[jS + 1:jN + 1, iW] = sgn * v[jS + 1:jN + 1, iW + 1]
                    elif face == dk.EAST:
                        u[jS:jN + 1, iE] = 0.0
                        v[jS + 1:jN + 1, iE + 1] = sgn * v[jS + 1:jN + 1, iE]
                    elif face == dk.SOUTH:
                        u[jS, iW + 1:iE + 1] = sgn * u[jS + 1, iW + 1:iE + 1]
                        v[jS, iW:iE + 1] = 0.0
                    else:
                        u[jN + 1, iW + 1:iE + 1] = sgn * u[jN, iW + 1:iE + 1]
                        v[jN, iW:iE + 1] = 0.0
                elif face == dk.WEST and tp == O1:
                    for j in range(jS, jN + 1): u[j, iW - 1] = +u[j, iW]
                    for j in range(jS + 1, jN + 1): v[j, iW] = -v[j, iW + 1]
                elif face == dk.WEST and tp == O2:
                    for j in range(jS, jN + 1): u[j, iW] = u[j, iW + 1] - v[j, iW] + v[j - 1, iW]
                    for j in range(jS + 1, jN + 1):
                        v[j, iW] = -v[j - 1, iW] + 5.0 * (v[j, iW + 1] - v[j - 1, iW + 1]) + 8.0 * (u[j, iW + 1] - u[j, iW])
                elif face == dk.EAST and tp == O1:
                    for j in range(jS + 1, jN + 1): v[j, iE + 1] = -v[j, iE]
                elif face == dk.EAST and tp == O2:
                    for j in range(jS + 1, jN + 1): u[j, iE] = u[j, iE - 1] - (v[j, iE] - v[j - 1, iE])
                    for j in range(jS + 1, jN):
                        v[j, iE + 1] = v[j - 1, iE + 1] + 3.0 * (v[j - 1, iE] - v[j, iE]) - 4.0 * (u[j, iE] - u[j, iE - 1])
                elif face == dk.SOUTH and tp == O1:
                    for i in range(iW + 1, iE + 1): u[jS, i] = -u[jS + 1, i]
                    for i in range(iW, iE + 1): v[jS - 1, i] = +v[jS, i]
                elif face == dk.SOUTH and tp == O2:
                    for i in range(iW + 1, iE + 1):
                        u[jS - 1, i] = u[jS - 1, i - 1] + 5.0 * (u[jS, i] - u[jS, i - 1]) + 8.0 * (v[jS, i] - v[jS - 1, i - 1])
                    for i in range(iW, iE + 1): v[jS - 1, i] = v[jS, i] - u[jS, i] - u[jS, i - 1]
                elif face == dk.NORTH and tp == O1:
                    for i in range(iW + 1, iE + 1): u[jN + 1, i] = -u[jN, i]
                    for i in range(iW, iE + 1): v[jN + 1, i] = +v[jN, i]
                elif face == dk.NORTH and tp == O2:
                    for i in range(iW, iE + 1): v[jN, i] = v[jN - 1, i] - (u[jN, i] - u[jN, i - 1])
                    for i in range(iW + 1, iE):
                        u[jN + 1, i] = u[jN + 1, i - 1] + 3.0 * (u[jN, i - 1] - u[jN, i]) - 4.0 * (v[jN, i] - v[jN - 1, i])
    _smlscl_pt(d, p, t)


def _smlscl_pt(d, p, t):
    """PresBoundCond + the homogeneous temperature ghosts of SmlSclBC (:1480-1649)."""
    from wolfd2_b200 import deck as dk
    r = d.regions
    py_presbc(d, p)
    for jr in range(int(r.nReg[1])):
        for ir in range(int(r.nReg[0])):
            iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
            if int(r.nTRgType[jr, ir]) == dk.RT_TEMPER:
                t[jS + 1:jN + 1, iW + 1:iE + 1] = 0.0
                t[jS + 1:jN + 1, iW + 1] = -t[jS + 1:jN + 1, iW]
                t[jS + 1:jN + 1, iE] = -t[jS + 1:jN + 1, iE + 1]
                t[jS + 1, iW + 1:iE + 1] = -t[jS, iW + 1:iE + 1]
                t[jN, iW + 1:iE + 1] = -t[jN + 1, iW + 1:iE + 1]
                continue
            for face in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH):
                bt = int(r.nTemBdTp[face - 1, jr, ir])
                if bt == 0:
                    continue
                sgn = -1.0 if bt == 1 else 1.0
                if face == dk.WEST: t[jS + 1:jN + 1, iW] = sgn * t[jS + 1:jN + 1, iW + 1]
                elif face == dk.EAST: t[jS + 1:jN + 1, iE + 1] = sgn * t[jS + 1:jN + 1, iE]
                elif face == dk.SOUTH: t[jS, iW + 1:iE + 1] = sgn * t[jS + 1, iW + 1:iE + 1]
                else: t[jN + 1, iW + 1:iE + 1] = sgn * t[jN, iW + 1:iE + 1]


@pytest.mark.parametrize("k", range(6))
def test_smlsclbc_second_restatement(orc, k):
    """SmlSclBC on all six decks (every face type on every side): bit for bit."""
    d = make_test_decks()[k]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(1200 + k)
    r = d.regions
    f = [rand_field(d, rng) for _ in range(4)]
    a, b = [x.copy() for x in f], [x.copy() for x in f]
    py_smlsclbc(d, *a)
    orc.smlsclbc(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.nTRgType, r.nTemBdTp, r.dBCVal, *b)
    for name, x, y in zip("uvpt", a, b):
        assert np.array_equal(x, y), name
    assert not np.array_equal(a[0], f[0])


def test_taveraged_second_restatement(orc):
    """TAveraged (src/utility.f:699-736): node average of t; fixed-temperature regions take dTRgVal (large scales)
    or zero (small scales); later regions overwrite shared border nodes."""
    from wolfd2_b200 import deck as dk
    d = _thermal_decks()[1]
    orc.config(d.mnx, d.mny)
    rng = np.random.default_rng(77)
    r = d.regions
    t = rand_field(d, rng)
    for nscale in (0, 1):
        s = rand_field(d, rng)
        a, b = s.copy(), s.copy()
        for jr in range(int(r.nReg[1])):
            for ir in range(int(r.nReg[0])):
                iW, iE, jS, jN = (int(r.nRegBrd[k - 1, jr, ir]) for k in (dk.WEST, dk.EAST, dk.SOUTH, dk.NORTH))
                if int(r.nTRgType[jr, ir]) == dk.RT_TEMPER:
                    a[jS:jN + 1, iW:iE + 1] = r.dTRgVal[jr, ir] if nscale == 0 else 0.0
                    continue
                W = Rng(iW, iE, jS, jN)
                W.put(a, (W(t) + W(t, 0, 1) + W(t, 1, 1) + W(t, 1, 0)) / 4.0)
        orc.taveraged(d.nx, d.ny, nscale, r.nReg, r.nRegBrd, r.nTRgType, r.dTRgVal, t, b)
        assert np.array_equal(a, b) and not np.array_equal(a, s), nscale


# ------------------------------------------------------------------ the committed fixtures, reproduced without the oracle
@pytest.mark.parametrize("name", ["cavity24x20", "channel22x18_fd", "channel22x18_mc", "bstep26x20", "heated_cavity26x22"])
def test_golden_fixtures_reproduced_by_the_second_restatement(name):
    """tests/golden/*.npz were written by the C oracle (make_golden.py).  The second restatement -- no oracle call
    anywhere in this test -- reproduces them bit for bit: cold start, four time steps, counts and norms.  The fixtures
    the GPU tests compare against are therefore backed by two independent restatements."""
    import os
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_golden
    d = make_golden.cases()[name]
    ref = np.load(os.path.join(here, "golden", name + ".npz"))
    thermal = "t0" in ref
    u, v, p, t, den = (d.new_field() for _ in range(5))
    if thermal:
        t[:d.ny + 2, :d.nx + 2] = 0.5
    py_velbc(d, u, v, False)                      # cold start, src/main.f:606-641
    ncold = np_ppe_general(d, u, v, p)
    py_presbc(d, p)
    py_project(d, p, u, v)
    py_velbc(d, u, v, False)
    assert ncold == int(ref["ncold"])
    for k in range(4):
        nql, nsor, dif = np_step_thermal(d, u, v, p, t, den)
        assert (nql, nsor) == (int(ref["nql"][k]), int(ref["nsor"][k])), k
        assert np.array_equal(u, ref[f"u{k}"]) and np.array_equal(v, ref[f"v{k}"]) and np.array_equal(p, ref[f"p{k}"]), k
        if thermal:
            assert np.array_equal(t, ref[f"t{k}"]) and np.array_equal(den, ref[f"d{k}"]), k
        n = ref["dif"].shape[1]
        assert dif[:n] == list(ref["dif"][k]), k
