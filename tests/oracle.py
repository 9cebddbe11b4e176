"""Loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY:
nothing under wolfd2_b200/ may import this module."""
import ctypes as C
import os
import subprocess

import numpy as np

from wolfd2_b200 import _abi

# Routines only the oracle exposes individually (internal to the reference's call tree).
ORACLE_ONLY = {
    # src/pressure.f:384, 457, 548, 673, 819, 976 (they take the assembled matrix a(mn,5))
    "sor_": (None, "ii" "i" "io" "ddd" "DD" "DD" "D" "D"),
    "sorrb_": (None, "ii" "i" "io" "ddd" "DD" "DD" "D" "D"),
    "sorrbp_": (None, "ii" "i" "io" "ddd" "DD" "DD" "D" "D"),
    "slor_": (None, "ii" "i" "i" "io" "ddd" "DD" "DD" "D" "D"),
    "slorrb_": (None, "ii" "i" "i" "io" "ddd" "DD" "DD" "D" "D"),
    "slorrbp_": (None, "ii" "i" "io" "ddd" "DD" "DD" "D" "D"),
}


class Particles(C.Structure):
    """orc_particles of oracle/wolfd2_oracle.c."""
    _fields_ = [("tr", C.POINTER(_abi.Traject)), ("gx", _abi.c_f64p), ("gy", _abi.c_f64p),
                ("cpartx", _abi.c_f64p), ("cparty", _abi.c_f64p), ("repc", _abi.c_f64p),
                ("xp", _abi.c_f64p), ("yp", _abi.c_f64p), ("up", _abi.c_f64p), ("vp", _abi.c_f64p),
                ("nTOutBnd", _abi.c_i32p)]


ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ODIR = os.path.join(ROOT, "oracle")


def _build(target="liboracle.so"):
    path = os.path.join(ODIR, target)
    src = os.path.join(ODIR, "wolfd2_oracle.c")
    inc = os.path.join(ODIR, "wolfd2_oracle_atd.inc")
    if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(inc)):
        subprocess.check_call(["make", "-C", ODIR, target], stdout=subprocess.DEVNULL)
    return path


class Oracle:
    def __init__(self, target="liboracle.so"):
        self.lib = C.CDLL(_build(target))
        table = dict(_abi.SIGNATURES)
        table.update(ORACLE_ONLY)
        self.f = _abi.bind(self.lib, prefix="orc_", table=table)
        self.lib.orc_config.argtypes = [C.c_int32] * 4
        self.lib.orc_get_errflag.restype = C.c_int32
        self.lib.orc_coldstart.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions),
                                           C.POINTER(_abi.Metrics), _abi.c_f64p, _abi.c_f64p,
                                           _abi.c_f64p, _abi.c_i32p]
        self.lib.orc_step.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions),
                                      C.POINTER(_abi.Metrics)] + [_abi.c_f64p] * 5 + \
                                     [C.c_int32, C.POINTER(_abi.StepLog)]
        self.lib.orc_step.restype = C.c_int32
        self.lib.orc_step_thermal.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions), C.POINTER(_abi.Metrics),
                                              C.POINTER(_abi.Thermal)] + [_abi.c_f64p] * 5 + \
                                             [C.c_int32, C.POINTER(_abi.StepLog)]
        self.lib.orc_step_thermal.restype = C.c_int32
        self.lib.orc_atd_init.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions), C.POINTER(_abi.Metrics),
                                          C.POINTER(_abi.Thermal), C.POINTER(_abi.SmallScale)] + [_abi.c_f64p] * 7
        self.lib.orc_step_full.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions), C.POINTER(_abi.Metrics),
                                           C.POINTER(_abi.Thermal), C.POINTER(_abi.SmallScale), C.POINTER(Particles)] + \
                                          [_abi.c_f64p] * 9 + [C.c_int32, C.POINTER(_abi.StepLog)]
        self.lib.orc_step_full.restype = C.c_int32
        self.lib.orc_ss_map.argtypes = [C.c_int32, C.c_int32]
        self.lib.orc_ss_map.restype = _abi.c_f64p
        self.lib.orc_ss_tarea.restype = C.c_double
        self.lib.orc_ppe_matrix.argtypes = [C.c_int32, C.c_int32, _abi.c_i32p, _abi.c_i32p, _abi.c_i32p,
                                            _abi.c_f64p, _abi.c_f64p, _abi.c_f64p, _abi.c_f64p]
        self.lib.orc_grid.argtypes = [C.c_int32, C.c_int32, C.c_double, _abi.c_f64p, _abi.c_f64p,
                                      C.POINTER(_abi.c_f64p)]
        self._dims = None

    def config(self, mnx, mny, mgri=20, mgrj=10):
        if self._dims != (mnx, mny, mgri, mgrj):
            self.lib.orc_config(mnx, mny, mgri, mgrj)
            self._dims = (mnx, mny, mgri, mgrj)

    def __getattr__(self, name):
        try:
            return self.__dict__["f"][name]
        except KeyError:
            raise AttributeError(name)

    # ---- whole-run drivers -------------------------------------------------
    def grid_metrics(self, x_nodes, y_nodes, mnx, mny, dlref=1.0):
        """orc_grid: src/grid.f Grid minus the file reader. Returns dict of 30 arrays."""
        self.config(mnx, mny)
        ny, nx = x_nodes.shape
        gx = np.zeros((mny + 1, mnx + 1)); gy = np.zeros((mny + 1, mnx + 1))
        gx[1:ny + 1, 1:nx + 1] = x_nodes
        gy[1:ny + 1, 1:nx + 1] = y_nodes
        m = {n: np.zeros((mny + 1, mnx + 1)) for n in _abi.METRIC_NAMES}
        ptrs = (_abi.c_f64p * 30)(*[m[n].ctypes.data_as(_abi.c_f64p) for n in _abi.METRIC_NAMES])
        self.lib.orc_grid(nx, ny, dlref, gx.ctypes.data_as(_abi.c_f64p), gy.ctypes.data_as(_abi.c_f64p), ptrs)
        return m

    def coldstart(self, deck, u, v, p):
        self.config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
        par, reg, met = deck.params(), deck.regions.as_struct(), deck.metrics_struct()
        n = C.c_int32(0)
        self.lib.orc_coldstart(C.byref(par), C.byref(reg), C.byref(met),
                               u.ctypes.data_as(_abi.c_f64p), v.ctypes.data_as(_abi.c_f64p),
                               p.ctypes.data_as(_abi.c_f64p), C.byref(n))
        return n.value

    def step(self, deck, u, v, p, nsteps=1, t=None, d=None):
        self.config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
        par, reg, met = deck.params(), deck.regions.as_struct(), deck.metrics_struct()
        t = deck.new_field() if t is None else t
        d = deck.new_field() if d is None else d
        logs = (_abi.StepLog * nsteps)()
        if getattr(deck, "thermal", False) or getattr(deck, "eqstate", False):
            th = deck.thermal_struct()
            rc = self.lib.orc_step_thermal(C.byref(par), C.byref(reg), C.byref(met), C.byref(th),
                                           *[a.ctypes.data_as(_abi.c_f64p) for a in (u, v, p, t, d)], nsteps, logs)
        else:
            rc = self.lib.orc_step(C.byref(par), C.byref(reg), C.byref(met),
                                   *[a.ctypes.data_as(_abi.c_f64p) for a in (u, v, p, t, d)], nsteps, logs)
        return rc, [dict(nQLiter=l.nQLiter, nSorConv=l.nSorConv, dif=list(l.dif)) for l in logs]


    def atd_init(self, deck, u, v, t, uss, vss, pss, tss):
        """src/main.f:643-665: SmallScale(initflg=0)."""
        self.config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
        par, reg, met = deck.params(), deck.regions.as_struct(), deck.metrics_struct()
        th, ss = deck.thermal_struct(), deck.smallscale_struct()
        self.lib.orc_atd_init(C.byref(par), C.byref(reg), C.byref(met), C.byref(th), C.byref(ss),
                              *[a.ctypes.data_as(_abi.c_f64p) for a in (u, v, t, uss, vss, pss, tss)])

    def step_full(self, deck, u, v, p, t, d, ss_fields=None, particles=None, nsteps=1):
        """Step body with the ATD blocks (ss_fields = (uss, vss, pss, tss)) and / or the trajectory block
        (particles = dict(tr, gx, gy, cpartx, cparty, repc, xp, yp, up, vp, out))."""
        self.config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
        par, reg, met = deck.params(), deck.regions.as_struct(), deck.metrics_struct()
        th = deck.thermal_struct()
        ss = deck.smallscale_struct() if ss_fields is not None else None
        pt = None
        if particles is not None:
            pt = Particles()
            pt.tr = C.pointer(particles["tr"])
            for k in ("gx", "gy", "cpartx", "cparty", "repc", "xp", "yp", "up", "vp"):
                setattr(pt, k, particles[k].ctypes.data_as(_abi.c_f64p))
            pt.nTOutBnd = particles["out"].ctypes.data_as(_abi.c_i32p)
        f = list(ss_fields) if ss_fields is not None else [None] * 4
        logs = (_abi.StepLog * nsteps)()
        rc = self.lib.orc_step_full(C.byref(par), C.byref(reg), C.byref(met), C.byref(th),
                                    C.byref(ss) if ss is not None else None, C.byref(pt) if pt is not None else None,
                                    *[a.ctypes.data_as(_abi.c_f64p) for a in (u, v, p, t, d)],
                                    *[a.ctypes.data_as(_abi.c_f64p) if a is not None else None for a in f], nsteps, logs)
        return rc, [dict(nQLiter=l.nQLiter, nSorConv=l.nSorConv, dif=list(l.dif)) for l in logs]

    def timeavg_accumulate(self, deck, npass, u, v, p, t, us, vs, ts, pn, acc):
        """src/main.f:1107-1208 on the 19 accumulators `acc` (list of (0:mnx,0:mny) arrays, SaveTmAvgP3D order + gradients)."""
        self.config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
        r, m = deck.regions, deck.metrics
        pp = (_abi.c_f64p * 19)(*[a.ctypes.data_as(_abi.c_f64p) for a in acc])
        f = self.lib.orc_timeavg_accumulate
        f.restype = None
        f.argtypes = [C.c_int32] * 3 + [_abi.c_i32p] * 4 + [_abi.c_f64p] * 14 + [C.POINTER(_abi.c_f64p)]
        f(npass, deck.nx, deck.ny, r.nReg.ctypes.data_as(_abi.c_i32p), r.nRegBrd.ctypes.data_as(_abi.c_i32p),
          r.nRegType.ctypes.data_as(_abi.c_i32p), r.nTRgType.ctypes.data_as(_abi.c_i32p),
          r.dTRgVal.ctypes.data_as(_abi.c_f64p),
          *[m[k].ctypes.data_as(_abi.c_f64p) for k in ("djn", "xen", "yen", "xzn", "yzn")],
          *[a.ctypes.data_as(_abi.c_f64p) for a in (u, v, p, t, us, vs, ts, pn)], pp)

    def timeavg_finish(self, deck, npass, nts, acc):
        pp = (_abi.c_f64p * 19)(*[a.ctypes.data_as(_abi.c_f64p) for a in acc])
        f = self.lib.orc_timeavg_finish
        f.restype = None
        f.argtypes = [C.c_int32] * 4 + [C.c_double] * 3 + [C.POINTER(_abi.c_f64p)]
        f(npass, deck.nx, deck.ny, nts, deck.uref, deck.dlref, deck.re, pp)

    def ss_map(self, deck, family, plane):
        """A copy of one plane of the oracle's saved map iterates."""
        ptr = self.lib.orc_ss_map(family, plane)
        n = (deck.mny + 1) * (deck.mnx + 1)
        return np.ctypeslib.as_array(ptr, shape=(n,)).reshape(deck.mny + 1, deck.mnx + 1).copy()


_ORACLE = None


def get_oracle():
    global _ORACLE
    if _ORACLE is None:
        _ORACLE = Oracle()
    return _ORACLE
