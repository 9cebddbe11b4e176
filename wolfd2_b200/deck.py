"""Host-side set-up of a wolfd2 run: grid metrics, region/BC tables, input decks.

This is the L2 ("setup") layer of the reference, which stays on the host:
it produces exactly the arrays `program wolfd2` holds after ReadParam/Grid/SetUpBCs
(src/main.f:376-448) and hands them to the device library.  Array convention: a Fortran
``REAL*8 f(0:mnx,0:mny)`` is a C-contiguous numpy array of shape (mny+1, mnx+1) indexed
``f[j, i]``; region tables keep the Fortran (mgri,mgrj,...) column-major order by using
numpy arrays with reversed axes (``nRegBrd[k-1, jr-1, ir-1]``).

Also writes reference-syntax decks (input.dat, grid, fluidprop.dat)
so a third party with gfortran can run the real reference on the same case.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field

import numpy as np

from ._abi import METRIC_NAMES, Metrics, Params, Regions, SmallScale, Thermal, c_f64p, c_i32p

# include/wolfd2.h:15-59
RM_BLOCKG, RM_INTERN, RM_POROUS = 0, 1, 2
BM_INTERN, BM_WALL1, BM_WALL2, BM_INLET, BM_OUTLT1, BM_OUTLT2 = 0, 1, 2, 3, 4, 5
WEST, EAST, SOUTH, NORTH = 1, 2, 3, 4
_U_, _V_, _P_, _T_ = 1, 2, 3, 4
FACE = {"w": WEST, "e": EAST, "s": SOUTH, "n": NORTH}
RT_NOSRCE, RT_HEATGN, RT_TEMPER = 0, 1, 2
BT_INTERN, BT_TEMPER, BT_HTFLUX = 0, 1, 2
PPE_SOLVERS = {"sor": 1, "lsor": 2, "rb_lsor": 3, "par_rb_lsor": 4, "rb_sor": 5, "par_rb_sor": 6}


def field2d(mnx: int, mny: int) -> np.ndarray:
    """Zero-initialised REAL*8 f(0:mnx,0:mny) (static storage in the reference)."""
    return np.zeros((mny + 1, mnx + 1), dtype=np.float64)


# --------------------------------------------------------------------------- grid.f

def uniform_grid(nx: int, ny: int, lx: float = 1.0, ly: float = 1.0):
    """Node coordinates x(i,j)=(i-1)/(nx-1)*lx, y(i,j)=(j-1)/(ny-1)*ly."""
    xi = np.arange(nx, dtype=np.float64) / float(nx - 1) * lx
    yj = np.arange(ny, dtype=np.float64) / float(ny - 1) * ly
    return np.broadcast_to(xi[None, :], (ny, nx)).copy(), np.broadcast_to(yj[:, None], (ny, nx)).copy()


def stretched_grid(nx: int, ny: int, skew: float = 0.15, stretch: float = 1.5):
    """A smooth non-Cartesian test grid (sheared and stretched) to exercise every metric."""
    s = np.arange(nx, dtype=np.float64) / float(nx - 1)
    t = np.arange(ny, dtype=np.float64) / float(ny - 1)
    S, T = np.meshgrid(s, t)
    xs = np.tanh(stretch * (2 * S - 1)) / np.tanh(stretch) * 0.5 + 0.5
    ys = np.tanh(stretch * (2 * T - 1)) / np.tanh(stretch) * 0.5 + 0.5
    x = xs + skew * np.sin(np.pi * ys) * 0.2
    y = ys + skew * np.sin(np.pi * xs) * 0.2
    return x, y


def metrics_from_grid(x_nodes: np.ndarray, y_nodes: np.ndarray, mnx: int, mny: int,
                      dlref: float = 1.0) -> dict:
    """Grid -> MirrorPts -> FullGrid -> Metric (src/grid.f:98-124, 258-535) in numpy.

    Same expressions and operation order as the reference; arrays are written on
    1..nx,1..ny only and stay zero elsewhere (load-bearing)."""
    ny, nx = x_nodes.shape
    if nx + 1 > mnx or ny + 1 > mny:  # CheckGridSize, src/grid.f:551
        raise ValueError(f"grid {nx}x{ny} needs mnx>={nx + 1}, mny>={ny + 1}")
    gx, gy = field2d(mnx, mny), field2d(mnx, mny)
    gx[1:ny + 1, 1:nx + 1] = x_nodes / dlref
    gy[1:ny + 1, 1:nx + 1] = y_nodes / dlref
    for g in (gx, gy):  # MirrorPts :276-303
        g[0, 1:nx + 1] = 2.0 * g[1, 1:nx + 1] - g[2, 1:nx + 1]
        g[ny + 1, 1:nx + 1] = 2.0 * g[ny, 1:nx + 1] - g[ny - 1, 1:nx + 1]
        g[1:ny + 1, 0] = 2.0 * g[1:ny + 1, 1] - g[1:ny + 1, 2]
        g[1:ny + 1, nx + 1] = 2.0 * g[1:ny + 1, nx] - g[1:ny + 1, nx - 1]
        g[0, 0] = 2.0 * g[1, 1] - g[2, 2]
        g[ny + 1, 0] = 2.0 * g[ny, 1] - g[ny - 1, 2]
        g[0, nx + 1] = 2.0 * g[1, nx] - g[2, nx - 1]
        g[ny + 1, nx + 1] = 2.0 * g[ny, nx] - g[ny - 1, nx - 1]
    xu, yu, xv, yv, xc, yc = (field2d(mnx, mny) for _ in range(6))
    # FullGrid :338-359
    xu[1:ny + 2, 0:nx + 2] = 0.5 * (gx[1:ny + 2, 0:nx + 2] + gx[0:ny + 1, 0:nx + 2])
    yu[1:ny + 2, 0:nx + 2] = 0.5 * (gy[1:ny + 2, 0:nx + 2] + gy[0:ny + 1, 0:nx + 2])
    xv[0:ny + 2, 1:nx + 2] = 0.5 * (gx[0:ny + 2, 1:nx + 2] + gx[0:ny + 2, 0:nx + 1])
    yv[0:ny + 2, 1:nx + 2] = 0.5 * (gy[0:ny + 2, 1:nx + 2] + gy[0:ny + 2, 0:nx + 1])
    xc[1:ny + 2, 1:nx + 2] = 0.5 * (xu[1:ny + 2, 1:nx + 2] + xu[1:ny + 2, 0:nx + 1])
    yc[1:ny + 2, 1:nx + 2] = 0.5 * (yv[1:ny + 2, 1:nx + 2] + yv[0:ny + 1, 1:nx + 2])

    m = {n: field2d(mnx, mny) for n in METRIC_NAMES}
    J, I = slice(1, ny + 1), slice(1, nx + 1)
    Jp, Ip = slice(2, ny + 2), slice(2, nx + 2)
    Jm, Im = slice(0, ny), slice(0, nx)

    def tensor(xz, xe, yz, ye):
        dj = 1.0 / (xz * ye - xe * yz)
        g11 = xz * xz + yz * yz
        g12 = xz * xe + yz * ye
        g22 = xe * xe + ye * ye
        return dj, g11, g12, g22

    # natural grid points (N) :425-448
    m["xzn"][J, I] = xv[J, Ip] - xv[J, I]
    m["xen"][J, I] = xu[Jp, I] - xu[J, I]
    m["yzn"][J, I] = yv[J, Ip] - yv[J, I]
    m["yen"][J, I] = yu[Jp, I] - yu[J, I]
    dj, g11, g12, g22 = tensor(m["xzn"][J, I], m["xen"][J, I], m["yzn"][J, I], m["yen"][J, I])
    m["djn"][J, I] = dj
    m["ran"][J, I] = dj * g22
    m["rbn"][J, I] = -dj * g12 * 0.25
    m["rgn"][J, I] = dj * g11
    # U cells :451-477
    m["xzu"][J, I] = xc[J, Ip] - xc[J, I]
    m["xeu"][J, I] = gx[J, I] - gx[Jm, I]
    m["yzu"][J, I] = yc[J, Ip] - yc[J, I]
    m["yeu"][J, I] = gy[J, I] - gy[Jm, I]
    dj, g11, g12, g22 = tensor(m["xzu"][J, I], m["xeu"][J, I], m["yzu"][J, I], m["yeu"][J, I])
    m["dju"][J, I] = dj
    m["rau"][J, I] = dj * g22
    m["rbu"][J, I] = -dj * g12 * 0.25
    # V cells :480-506
    m["xzv"][J, I] = gx[J, I] - gx[J, Im]
    m["xev"][J, I] = xc[Jp, I] - xc[J, I]
    m["yzv"][J, I] = gy[J, I] - gy[J, Im]
    m["yev"][J, I] = yc[Jp, I] - yc[J, I]
    dj, g11, g12, g22 = tensor(m["xzv"][J, I], m["xev"][J, I], m["yzv"][J, I], m["yev"][J, I])
    m["djv"][J, I] = dj
    m["rbv"][J, I] = -dj * g12 * 0.25
    m["rgv"][J, I] = dj * g11
    # P cells :509-532
    m["xzc"][J, I] = xu[J, I] - xu[J, Im]
    m["xec"][J, I] = xv[J, I] - xv[Jm, I]
    m["yzc"][J, I] = yu[J, I] - yu[J, Im]
    m["yec"][J, I] = yv[J, I] - yv[Jm, I]
    dj, g11, g12, g22 = tensor(m["xzc"][J, I], m["xec"][J, I], m["yzc"][J, I], m["yec"][J, I])
    m["djc"][J, I] = dj
    m["rac"][J, I] = dj * g22
    m["rbc"][J, I] = -dj * g12 * 0.25
    m["rgc"][J, I] = dj * g11
    m["_gx"], m["_gy"] = gx, gy
    return m


# --------------------------------------------------------- region / BC tables

class RegionTables:
    """nReg, nRegBrd, nRegType, nMomBdTp, dBCVal, dPRporos/c1/c2 as SetUpBCs leaves them.

    Mirrors `section boundary_conditions` statements (src/parse.f:1213-1760) and the
    post-processing of SetUpBCs (src/bound_cond.f:165-223)."""

    def __init__(self, nx, ny, ni=1, nj=1, i_borders=(), j_borders=(), mgri=20, mgrj=10,
                 uref=1.0, densref=1.0):
        if ni > mgri or nj > mgrj:
            raise ValueError("too many regions")
        if len(i_borders) != ni - 1 or len(j_borders) != nj - 1:
            raise ValueError("need ni-1 i_borders and nj-1 j_borders")
        self.nx, self.ny, self.mgri, self.mgrj = nx, ny, mgri, mgrj
        self.uref, self.densref = uref, densref
        self.i_borders, self.j_borders = tuple(i_borders), tuple(j_borders)
        self.nReg = np.array([ni, nj], dtype=np.int32)
        self.nRegBrd = np.zeros((4, mgrj, mgri), dtype=np.int32)
        self.nRegType = np.zeros((mgrj, mgri), dtype=np.int32)
        self.nMomBdTp = np.zeros((4, mgrj, mgri), dtype=np.int32)
        self.dBCVal = np.zeros((4, 4, mgrj, mgri), dtype=np.float64)   # [l-1, k-1, jr-1, ir-1]
        self.dPRporos = np.zeros((mgrj, mgri), dtype=np.float64)
        self.dPRporc1 = np.zeros((mgrj, mgri), dtype=np.float64)
        self.dPRporc2 = np.zeros((mgrj, mgri), dtype=np.float64)
        self._statements = []
        # border consistency checks, src/bound_cond.f:124-162
        for b, n, nm in ((self.i_borders, nx, "i"), (self.j_borders, ny, "j")):
            for q, v in enumerate(b):
                if v < 2 or v > n - 2:
                    raise ValueError(f"bad {nm}_border {v}: must be >=2 and <= n-2")
                if q and v - b[q - 1] < 2:
                    raise ValueError(f"{nm}_borders must increase by at least 2")
        for q, v in enumerate(self.i_borders):
            self.nRegBrd[WEST - 1, 0, q + 1] = v
        for q, v in enumerate(self.j_borders):
            self.nRegBrd[SOUTH - 1, q + 1, 0] = v
        # InitBCFlags, src/parse.f:2257-2379
        self.nRegType[:nj, :ni] = RM_INTERN
        self.nMomBdTp[:, :nj, :ni] = BM_INTERN
        self.nMomBdTp[WEST - 1, :nj, 0] = BM_WALL1
        self.nMomBdTp[EAST - 1, :nj, ni - 1] = BM_WALL1
        self.nMomBdTp[SOUTH - 1, 0, :ni] = BM_WALL1
        self.nMomBdTp[NORTH - 1, nj - 1, :ni] = BM_WALL1
        # thermal defaults (parse.f:2340-2376): no heat source, interfaces inside, adiabatic outer walls
        self.nTRgType = np.zeros((mgrj, mgri), dtype=np.int32)
        self.nTemBdTp = np.zeros((4, mgrj, mgri), dtype=np.int32)
        self.dTRgVal = np.zeros((mgrj, mgri), dtype=np.float64)
        self.dHGSTval = np.zeros((mgrj, mgri), dtype=np.float64)
        self.nTemBdTp[WEST - 1, :nj, 0] = BT_HTFLUX
        self.nTemBdTp[EAST - 1, :nj, ni - 1] = BT_HTFLUX
        self.nTemBdTp[SOUTH - 1, 0, :ni] = BT_HTFLUX
        self.nTemBdTp[NORTH - 1, nj - 1, :ni] = BT_HTFLUX
        self._completed = False

    # -- statements ------------------------------------------------------------
    def _k(self, face):
        return FACE[face] if isinstance(face, str) else int(face)

    def wall(self, ir, jr, face, no_stress=False, tangent_vel=None, press=None):
        """`wall ir jr face [no_slip|no_stress] [tangent_vel v] [press p]` (parse.f:1495-1572)."""
        k = self._k(face)
        self.nMomBdTp[k - 1, jr - 1, ir - 1] = BM_WALL2 if no_stress else BM_WALL1
        if tangent_vel is not None:
            var = _V_ if k in (WEST, EAST) else _U_
            self.dBCVal[var - 1, k - 1, jr - 1, ir - 1] = tangent_vel / self.uref
        if press is not None:
            self.dBCVal[_P_ - 1, k - 1, jr - 1, ir - 1] = press / (self.densref * self.uref * self.uref)
        self._statements.append(("wall", ir, jr, face, no_stress, tangent_vel, press))
        return self

    def inlet(self, ir, jr, face, normal_vel=None, tangent_vel=None):
        """`inlet ir jr face normal_vel v [tangent_vel v]` (parse.f:1577-1678)."""
        k = self._k(face)
        self.nMomBdTp[k - 1, jr - 1, ir - 1] = BM_INLET
        if normal_vel is not None:
            var = _U_ if k in (WEST, EAST) else _V_
            self.dBCVal[var - 1, k - 1, jr - 1, ir - 1] = normal_vel / self.uref
        if tangent_vel is not None:
            var = _V_ if k in (WEST, EAST) else _U_
            self.dBCVal[var - 1, k - 1, jr - 1, ir - 1] = tangent_vel / self.uref
        self._statements.append(("inlet", ir, jr, face, normal_vel, tangent_vel))
        return self

    def outlet(self, ir, jr, face, fully_dev=False, press=None):
        """`outlet ir jr face [fully_dev|mass_cons] [press p]` (parse.f:1683-1760);
        the default type is mass_cons (BM_OUTLT2)."""
        k = self._k(face)
        self.nMomBdTp[k - 1, jr - 1, ir - 1] = BM_OUTLT1 if fully_dev else BM_OUTLT2
        if press is not None:
            self.dBCVal[_P_ - 1, k - 1, jr - 1, ir - 1] = press / (self.densref * self.uref * self.uref)
        self._statements.append(("outlet", ir, jr, face, fully_dev, press))
        return self

    def blockage(self, ir, jr):
        """`blockage ir jr` (parse.f:1297-1344)."""
        ni, nj = int(self.nReg[0]), int(self.nReg[1])
        self.nRegType[jr - 1, ir - 1] = RM_BLOCKG
        self.nMomBdTp[:, jr - 1, ir - 1] = BM_WALL1
        if ir > 1:
            self.nMomBdTp[EAST - 1, jr - 1, ir - 2] = BM_WALL1
        if ir < ni:
            self.nMomBdTp[WEST - 1, jr - 1, ir] = BM_WALL1
        if jr > 1:
            self.nMomBdTp[NORTH - 1, jr - 2, ir - 1] = BM_WALL1
        if jr < nj:
            self.nMomBdTp[SOUTH - 1, jr, ir - 1] = BM_WALL1
        self._statements.append(("blockage", ir, jr))
        return self

    def porous(self, ir, jr, poros, c1, c2):
        """Mark a porous region with already non-dimensional coefficients (the reference derives
        them in PorRegCnst, src/bound_cond.f:1883-1986, from permeability models)."""
        self.nRegType[jr - 1, ir - 1] = RM_POROUS
        self._porous = getattr(self, "_porous", {})
        self._porous[(ir, jr)] = (poros, c1, c2)
        return self

    # -- thermal statements (already non-dimensional values; the reference scales them in parse.f) ------
    def wall_temperature(self, ir, jr, face, temp):
        """Fixed-temperature face (BT_TEMPER, e.g. parse.f:1557): ghost = 2*temp - inner."""
        k = self._k(face)
        self.nTemBdTp[k - 1, jr - 1, ir - 1] = BT_TEMPER
        self.dBCVal[_T_ - 1, k - 1, jr - 1, ir - 1] = temp
        return self

    def wall_heat_flux(self, ir, jr, face, flux=0.0):
        """Prescribed-flux face (BT_HTFLUX, parse.f:1924): ghost = flux + inner; 0 = adiabatic."""
        k = self._k(face)
        self.nTemBdTp[k - 1, jr - 1, ir - 1] = BT_HTFLUX
        self.dBCVal[_T_ - 1, k - 1, jr - 1, ir - 1] = flux
        return self

    def heat_generation(self, ir, jr, value):
        """Region with a (dimensionless) volumetric heat source (RT_HEATGN, parse.f:1789)."""
        self.nTRgType[jr - 1, ir - 1] = RT_HEATGN
        self.dHGSTval[jr - 1, ir - 1] = value
        return self

    def fixed_temperature_region(self, ir, jr, temp):
        """Region held at `temp` (RT_TEMPER, parse.f:1826-1845): its faces and the neighbours' facing faces
        become BT_TEMPER with the same value."""
        ni, nj = int(self.nReg[0]), int(self.nReg[1])
        self.nTRgType[jr - 1, ir - 1] = RT_TEMPER
        self.dTRgVal[jr - 1, ir - 1] = temp
        for k in (WEST, EAST, SOUTH, NORTH):
            self.nTemBdTp[k - 1, jr - 1, ir - 1] = BT_TEMPER
            self.dBCVal[_T_ - 1, k - 1, jr - 1, ir - 1] = temp
        for (di, dj, k) in ((-1, 0, EAST), (1, 0, WEST), (0, -1, NORTH), (0, 1, SOUTH)):
            i2, j2 = ir + di, jr + dj
            if 1 <= i2 <= ni and 1 <= j2 <= nj:
                self.nTemBdTp[k - 1, j2 - 1, i2 - 1] = BT_TEMPER
                self.dBCVal[_T_ - 1, k - 1, j2 - 1, i2 - 1] = temp
        return self

    # -- SetUpBCs post-processing, src/bound_cond.f:165-223 ----------------------
    def complete(self):
        nx, ny = self.nx, self.ny
        ni, nj = int(self.nReg[0]), int(self.nReg[1])
        B, M = self.nRegBrd, self.nMomBdTp
        B[WEST - 1, :nj, 0] = 1
        B[EAST - 1, :nj, ni - 1] = nx
        B[SOUTH - 1, 0, :ni] = 1
        B[NORTH - 1, nj - 1, :ni] = ny
        for jr in range(1, nj):
            for ir in range(1, ni):
                B[WEST - 1, jr, ir] = B[WEST - 1, 0, ir]
                B[SOUTH - 1, jr, ir] = B[SOUTH - 1, jr, 0]
        for jr in range(nj):
            for ir in range(ni - 1):
                B[EAST - 1, jr, ir] = B[WEST - 1, 0, ir + 1]
        for jr in range(nj - 1):
            for ir in range(ni):
                B[NORTH - 1, jr, ir] = B[SOUTH - 1, jr + 1, 0]
        for jr in range(nj):
            for ir in range(ni - 1):
                for tp in (BM_INLET, BM_WALL1):
                    if M[EAST - 1, jr, ir] == tp:
                        M[WEST - 1, jr, ir + 1] = tp
                    if M[WEST - 1, jr, ir + 1] == tp:
                        M[EAST - 1, jr, ir] = tp
        for jr in range(nj - 1):
            for ir in range(ni):
                for tp in (BM_INLET, BM_WALL1):
                    if M[SOUTH - 1, jr + 1, ir] == tp:
                        M[NORTH - 1, jr, ir] = tp
                    if M[NORTH - 1, jr, ir] == tp:
                        M[SOUTH - 1, jr + 1, ir] = tp
        # PorRegCnst non-porous defaults, src/bound_cond.f:1971-1978
        self.dPRporos[:nj, :ni] = 1.0
        self.dPRporc1[:nj, :ni] = 0.0
        self.dPRporc2[:nj, :ni] = 0.0
        for (ir, jr), (po, c1, c2) in getattr(self, "_porous", {}).items():
            self.dPRporos[jr - 1, ir - 1] = po
            self.dPRporc1[jr - 1, ir - 1] = c1
            self.dPRporc2[jr - 1, ir - 1] = c2
        self._completed = True
        return self

    def as_struct(self) -> Regions:
        assert self._completed, "call complete() first"
        r = Regions()
        r.nReg = self.nReg.ctypes.data_as(c_i32p)
        r.nRegBrd = self.nRegBrd.ctypes.data_as(c_i32p)
        r.nRegType = self.nRegType.ctypes.data_as(c_i32p)
        r.nMomBdTp = self.nMomBdTp.ctypes.data_as(c_i32p)
        r.dBCVal = self.dBCVal.ctypes.data_as(c_f64p)
        r.dPRporos = self.dPRporos.ctypes.data_as(c_f64p)
        r.dPRporc1 = self.dPRporc1.ctypes.data_as(c_f64p)
        r.dPRporc2 = self.dPRporc2.ctypes.data_as(c_f64p)
        return r


# ------------------------------------------------------------------------- decks

@dataclass
class Deck:
    """Everything `program wolfd2` holds when the time loop starts (src/main.f:687)."""
    name: str
    nx: int
    ny: int
    mnx: int
    mny: int
    regions: RegionTables
    metrics: dict
    dt: float                      # dimensional time_step_size
    re: float
    mqiter: int = 20
    nmeiter: int = 1
    ppe_solver: str = "rb_sor"
    msorit: int = 2000
    qtol: float = 1e-4
    sortol: float = 1e-8
    sorrel: float = 1.0
    cartesian: bool = True
    dlref: float = 1.0
    uref: float = 1.0
    nfiltu: int = 0
    nfiltv: int = 0
    fpu: float = 5e2
    fpv: float = 5e2
    x_nodes: np.ndarray = field(default=None, repr=False)
    y_nodes: np.ndarray = field(default=None, repr=False)
    # thermal energy equation (thermal_energy / eq_state / filter_t of the reference's input deck)
    thermal: bool = False
    eqstate: bool = False
    nfiltt: int = 0
    fpt: float = 5e2
    prandtl: float = 0.71
    dmeittol: float = 1e-6
    densref: float = 1.2
    tmax: float = 310.0
    tref: float = 300.0
    rconst: float = 287.0
    # ATD small-scale model (small_scale / atd_* of the reference's input deck; defaults src/parse.f:144-178)
    smallscale: bool = False
    ss_ppe_solver: str = "sor"
    ss_msorit: int = 2000
    ss_sortol: float = 1e-8
    ss_sorrel: float = 1.0
    ss_filt: tuple = (5e2, 5e2, 0.0, 5e2)
    ss_cu0: float = 0.0
    ss_tscoef: float = 1.0
    ss_hscoef: float = 1.0
    ss_temcoef: float = 1.0
    ss_bncrit: float = 2e2
    ss_rmpmax: float = 0.95
    ss_rmpexp: float = 5.0
    # one rank's share of a multi-GPU run: (rank, world, J0, J1, A0, A1, HG), wolfd2_b200/slab.py.  nx, ny,
    # regions and params stay global; metrics and fields hold rows A0..A1 only (row 0 = global row A0)
    slab: tuple = None

    @property
    def dk(self):  # src/file_manip.f:134
        return self.dt * self.uref / self.dlref

    @property
    def fr(self):  # src/file_manip.f:159
        return self.uref ** 2.0 / (self.dlref * 9.81)

    def params(self) -> Params:
        p = Params()
        p.nx, p.ny = self.nx, self.ny
        p.mqiter, p.nmeiter = self.mqiter, self.nmeiter
        p.nPpeSolver, p.msorit = PPE_SOLVERS[self.ppe_solver], self.msorit
        p.lCartesGrid = 1 if self.cartesian else 0
        p.nfiltu, p.nfiltv = self.nfiltu, self.nfiltv
        p.dk, p.re, p.fr = self.dk, self.re, self.fr
        p.qtol, p.sortol, p.sorrel = self.qtol, self.sortol, self.sorrel
        p.fpu, p.fpv = self.fpu, self.fpv
        return p

    @property
    def pe(self):   # Peclet number, src/file_manip.f
        return self.re * self.prandtl

    def thermal_struct(self) -> Thermal:
        t = Thermal()
        t.nthermen, t.neqstate, t.nfiltt = int(self.thermal), int(self.eqstate), self.nfiltt
        t.pe, t.dmeittol, t.fpt = self.pe, self.dmeittol, self.fpt
        t.uref, t.densref, t.tmax, t.tref, t.rconst = self.uref, self.densref, self.tmax, self.tref, self.rconst
        r = self.regions
        t.nTRgType = r.nTRgType.ctypes.data_as(c_i32p)
        t.nTemBdTp = r.nTemBdTp.ctypes.data_as(c_i32p)
        t.dTRgVal = r.dTRgVal.ctypes.data_as(c_f64p)
        t.dHGSTval = r.dHGSTval.ctypes.data_as(c_f64p)
        return t

    def smallscale_struct(self) -> SmallScale:
        q = SmallScale()
        q.nsmallscl = int(self.smallscale)
        q.nssPpeSlvr, q.mssSorIt = PPE_SOLVERS[self.ss_ppe_solver], self.ss_msorit
        q.dlref, q.uref, q.tref, q.tmax, q.pe = self.dlref, self.uref, self.tref, self.tmax, self.pe
        q.ssSorTol, q.ssSorRel = self.ss_sortol, self.ss_sorrel
        for k in range(4):
            q.ssFiltPar[k] = self.ss_filt[k]
        q.ssCu0, q.ssTsCoef, q.ssHsCoef, q.ssTemCoef = self.ss_cu0, self.ss_tscoef, self.ss_hscoef, self.ss_temcoef
        q.ssBnCrit, q.ssRMpMax, q.ssRMpExp = self.ss_bncrit, self.ss_rmpmax, self.ss_rmpexp
        return q

    def node_arrays(self):
        """Grid nodes as main.f holds them: x(0:mnx,0:mny), y, nodes at 1..nx, 1..ny (src/grid.f)."""
        gx, gy = self.new_field(), self.new_field()
        gx[1:self.ny + 1, 1:self.nx + 1] = self.x_nodes / self.dlref
        gy[1:self.ny + 1, 1:self.nx + 1] = self.y_nodes / self.dlref
        return gx, gy

    def metrics_struct(self) -> Metrics:
        m = Metrics()
        for n in METRIC_NAMES:
            if n in self.metrics:    # an empty dict (lazy_metrics) leaves NULL pointers: nothing is uploaded at create
                setattr(m, n, self.metrics[n].ctypes.data_as(c_f64p))
        return m

    def new_field(self) -> np.ndarray:
        return field2d(self.mnx, self.mny)

    def cells(self) -> int:
        """Pressure unknowns (of this rank's slab, if any)."""
        if self.slab:
            return (self.nx - 1) * (self.slab[3] - self.slab[2] + 1)
        return (self.nx - 1) * (self.ny - 1)

    def window(self, f_global: np.ndarray) -> np.ndarray:
        """This rank's rows of a global field, in the local host layout."""
        if not self.slab:
            return f_global
        a0, a1 = self.slab[4], self.slab[5]
        out = self.new_field()
        out[0:a1 - a0 + 1, :] = f_global[a0:a1 + 1, 0:self.mnx + 1]
        return out

    def to_slab(self, rank: int, world: int) -> "Deck":
        """Cut a global deck into the share of `rank` (tests: slices the global metric arrays)."""
        import dataclasses
        from .slab import slab_layout
        j0, j1, a0, a1, hg = slab_layout(self.nx, self.ny, world, rank)
        d = dataclasses.replace(self, mny=a1 - a0, metrics={}, slab=(rank, world, j0, j1, a0, a1, hg))
        d.metrics = {k: np.ascontiguousarray(v[a0:a1 + 1, :]) for k, v in self.metrics.items()}
        return d

    # ---- reference-syntax files ---------------------------
    def write_reference_files(self, directory, n_time_steps=100):
        import os
        os.makedirs(directory, exist_ok=True)
        r = self.regions
        L = ["# generated by wolfd2_b200.deck -- reference-syntax deck",
             "section input_parameters",
             "grid_file grid.dat" + (" cartesian_grid" if self.cartesian else ""),
             f"n_time_steps {n_time_steps}",
             f"time_step_size {self.dt:.17g}",
             f"max_ql_iter {self.mqiter}",
             f"ql_tolerance {self.qtol:.17g}",
             f"max_me_iter {self.nmeiter}",
             f"ppe_solver {self.ppe_solver}",
             f"max_sor_iter {self.msorit}",
             f"sor_tolerance {self.sortol:.17g}",
             f"sor_relaxation {self.sorrel:.17g}",
             f"ref_length {self.dlref:.17g}",
             f"ref_velocity {self.uref:.17g}",
             "ref_temperature 300.0",
             "fluid_prop_file fluidprop.dat",
             "fluid_prop_table synth",
             "write_restart restart.out",
             "output_prefix run",
             "print_diff_freq 1"]
        if self.nfiltu:
            L.append(f"filter_u {self.fpu:.17g}")
        if self.nfiltv:
            L.append(f"filter_v {self.fpv:.17g}")
        L += ["end section", "",
              "section boundary_conditions",
              f"number_of_regions {int(r.nReg[0])} {int(r.nReg[1])}",
              "i_borders " + " ".join(str(b) for b in r.i_borders),
              "j_borders " + " ".join(str(b) for b in r.j_borders)]
        for st in r._statements:
            if st[0] == "wall":
                _, ir, jr, f, ns, tv, pr = st
                s = f"wall {ir} {jr} {f}" + (" no_stress" if ns else "")
                s += f" tangent_vel {tv:.17g}" if tv is not None else ""
                s += f" press {pr:.17g}" if pr is not None else ""
            elif st[0] == "inlet":
                _, ir, jr, f, nv, tv = st
                s = f"inlet {ir} {jr} {f}"
                s += f" normal_vel {nv:.17g}" if nv is not None else ""
                s += f" tangent_vel {tv:.17g}" if tv is not None else ""
            elif st[0] == "outlet":
                _, ir, jr, f, fd, pr = st
                s = f"outlet {ir} {jr} {f} " + ("fully_dev" if fd else "mass_cons")
                s += f" press {pr:.17g}" if pr is not None else ""
            else:
                s = f"blockage {st[1]} {st[2]}"
            L.append(s)
        L += ["end section", ""]
        with open(os.path.join(directory, "input.dat"), "w") as f:
            f.write("\n".join(L))
        # fluid table: rho=1, mu=1/Re -> re = dlref*uref*rho/mu (src/file_manip.f:143-157)
        mu = self.dlref * self.uref / self.re
        with open(os.path.join(directory, "fluidprop.dat"), "w") as f:
            f.write("fluid synth gasconstant 287.0\n")
            for T in (200.0, 400.0):
                f.write(f"{T:.1f} 1.0 1000.0 {mu:.17g} 0.026 0.71\n")
            f.write("end\n")
        with open(os.path.join(directory, "grid.dat"), "w") as f:
            f.write(f"{self.nx} {self.ny}\n")
            for arr in (self.x_nodes, self.y_nodes):
                flat = np.asarray(arr).reshape(-1)
                for q in range(0, flat.size, 4):
                    f.write(" ".join(f"{v:.17g}" for v in flat[q:q + 4]) + "\n")


def metrics_window(nx: int, ny: int, a0: int, a1: int, mnx: int, dlref: float = 1.0, lx: float = 1.0,
                   ly: float = 1.0) -> dict:
    """Metric arrays of a uniform grid on rows a0..a1 only (a multi-GPU rank never forms the global
    arrays).  Every metric at row j is a formula of the node rows j-2..j+2, so running the same code on the
    node rows a0-2..a1+2 and dropping the two rows at each cut reproduces the global values bit for bit
    (the extrapolated mirror rows are right at the physical edges and discarded at the cuts)."""
    lo, hi = max(1, a0 - 2), min(ny, a1 + 2)
    xi = np.arange(nx, dtype=np.float64) / float(nx - 1) * lx
    yj = np.arange(lo - 1, hi, dtype=np.float64) / float(ny - 1) * ly
    nyw = hi - lo + 1
    xw = np.broadcast_to(xi[None, :], (nyw, nx))
    yw = np.broadcast_to(yj[:, None], (nyw, nx))
    m = metrics_from_grid(xw, yw, mnx, nyw + 1, dlref)    # local row l <-> global row lo - 1 + l
    out = {}
    for k, v in m.items():
        w = np.zeros((a1 - a0 + 1, mnx + 1))
        g_lo, g_hi = max(a0, lo - 1), min(a1, hi + 1)          # rows 0 / ny+1 stay zero like the global arrays
        w[g_lo - a0:g_hi - a0 + 1, :] = v[g_lo - (lo - 1):g_hi - (lo - 1) + 1, :]
        out[k] = w
    return out


def _mk(name, nx, ny, regions, re, dt, x=None, y=None, mnx=None, mny=None, slab=None, **kw) -> Deck:
    mnx = mnx if mnx is not None else nx + 1   # smallest legal size, src/grid.f:551
    # lazy_metrics: leave the metric dict empty; api.Context then builds and uploads the metrics of the uniform grid
    # window by window (a 16384^2 grid would need 64 GB of host arrays otherwise)
    lazy = kw.pop("lazy_metrics", False)
    if lazy and x is not None:
        raise ValueError("lazy_metrics is for uniform grids")
    if slab is not None:   # (rank, world): build this rank's rows directly (uniform grids)
        from .slab import slab_layout
        if x is not None:
            raise ValueError("slab decks are built for uniform grids; cut others with Deck.to_slab")
        rank, world = slab
        j0, j1, a0, a1, hg = slab_layout(nx, ny, world, rank)
        met = {} if lazy else metrics_window(nx, ny, a0, a1, mnx, kw.get("dlref", 1.0))
        return Deck(name=name, nx=nx, ny=ny, mnx=mnx, mny=a1 - a0, regions=regions.complete(), metrics=met,
                    dt=dt, re=re, slab=(rank, world, j0, j1, a0, a1, hg), **kw)
    mny = mny if mny is not None else ny + 1
    if lazy:
        return Deck(name=name, nx=nx, ny=ny, mnx=mnx, mny=mny, regions=regions.complete(), metrics={},
                    dt=dt, re=re, **kw)
    if x is None:
        x, y = uniform_grid(nx, ny)
    met = metrics_from_grid(x, y, mnx, mny, kw.get("dlref", 1.0))
    return Deck(name=name, nx=nx, ny=ny, mnx=mnx, mny=mny, regions=regions.complete(), metrics=met,
                dt=dt, re=re, x_nodes=x, y_nodes=y, **kw)


def cavity(n=64, re=100.0, dt=0.01, ny=None, **kw) -> Deck:
    """Lid-driven cavity: `wall 1 1 n tangent_vel 1.0`, other faces default no-slip walls."""
    nx, ny = n, (ny or n)
    reg = RegionTables(nx, ny).wall(1, 1, "n", tangent_vel=1.0)
    return _mk(f"cavity{nx}x{ny}_re{re:g}", nx, ny, reg, re, dt, **kw)


def channel(n=64, re=100.0, dt=0.01, ny=None, fully_dev=True, **kw) -> Deck:
    """Channel: `inlet 1 1 w normal_vel 1.0`, `outlet 1 1 e fully_dev|mass_cons`, walls N/S."""
    nx, ny = n, (ny or n)
    reg = RegionTables(nx, ny).inlet(1, 1, "w", normal_vel=1.0).outlet(1, 1, "e", fully_dev=fully_dev)
    return _mk(f"channel{nx}x{ny}_re{re:g}", nx, ny, reg, re, dt, **kw)


def backward_step(n=64, re=100.0, dt=0.01, ny=None, fully_dev=True, **kw) -> Deck:
    """Backward-facing step: 2x2 regions, `blockage 1 1`, inlet on (1,2) west, outlet on east."""
    nx, ny = n, (ny or n)
    ib, jb = max(2, nx // 4), max(2, ny // 2)
    reg = RegionTables(nx, ny, 2, 2, (ib,), (jb,))
    reg.blockage(1, 1).inlet(1, 2, "w", normal_vel=1.0)
    reg.outlet(2, 1, "e", fully_dev=fully_dev).outlet(2, 2, "e", fully_dev=fully_dev)
    return _mk(f"bstep{nx}x{ny}_re{re:g}", nx, ny, reg, re, dt, **kw)


def heated_cavity(n=64, re=100.0, dt=0.01, ny=None, t_hot=1.0, t_cold=0.0, **kw) -> Deck:
    """Differentially heated cavity: west wall hot, east wall cold, adiabatic top and bottom, buoyancy through
    the ideal-gas density (EqState) in the v-momentum equation; thermal energy equation on."""
    nx, ny = n, (ny or n)
    reg = RegionTables(nx, ny)
    reg.wall_temperature(1, 1, "w", t_hot).wall_temperature(1, 1, "e", t_cold)
    kw.setdefault("thermal", True)
    kw.setdefault("eqstate", True)
    kw.setdefault("nmeiter", 3)
    return _mk(f"heated_cavity{nx}x{ny}_re{re:g}", nx, ny, reg, re, dt, **kw)
