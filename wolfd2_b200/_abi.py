"""ctypes signature table of the reference's hot-path subroutines (gfortran ABI).

The table binds the literal shims of libwolfd2_b200.so (include/wolfd2_b200.h, section (1)).
`bind` accepts a symbol prefix so that the test suite can bind a second library exposing the same
argument lists (the CPU checker under tests/); nothing in this package loads such a library.

Argument kinds:  i = INTEGER scalar by reference (in),  o = INTEGER scalar (out),
d = REAL*8 scalar by reference,  I = INTEGER array,  D = REAL*8 array.
Argument order follows the Fortran declarations cited next to each entry.
"""
import ctypes as C

import numpy as np

c_i32p = C.POINTER(C.c_int32)
c_f64p = C.POINTER(C.c_double)

M30 = "D" * 30

SIGNATURES = {
    # src/momentum.f:33-48
    "nauxmomentum_": (C.c_int32, "iii" "IIII" "dddd" "DDD" "D" + "D" * 24 + "DD" "DDDD"),
    # src/momentum.f:199-208
    "xmomentum_": (None, "ii" "IIII" "dd" "DDD" "DDDDD" "DDDD" "DDDD" "DDDD" "D"),
    # src/momentum.f:520-530
    "ymomentum_": (None, "ii" "IIII" "ddd" "DDD" "DDDDD" "DDDD" "DDDD" "DD" "DDDD" "D"),
    # src/momentum.f:1307
    "alttridlu_": (None, "iDD"),
    # src/pressure.f:30-38
    "ppe_": (None, "ii" "III" "i" "iio" "ddd" "DDDD" "DDDD" "DDD"),
    # src/pressure.f:265-267
    "divergence_": (None, "iii" "DDDD" "DDD"),
    # src/utility.f:253-259
    "project_": (None, "ii" "IIII" "d" "DD" "DDDD" "DDD"),
    # src/bound_cond.f:511-513
    "velboundcond_": (None, "ii" "III" "D" "DD"),
    # src/bound_cond.f:853-856
    "presboundcond_": (None, "ii" "IIII" "D" "D"),
    # src/bound_cond.f:1656-1659
    "veloutflowbcs_": (None, "ii" "III" "D" "DD"),
    # src/utility.f:33-36
    "filter_": (None, "iii" "IIII" "I" "d" "D"),
    # src/utility.f:446, 479
    "diffmaxnorm_": (C.c_double, "ii" "DD"),
    "dmaxnorm_": (C.c_double, "ii" "D"),
    # routines below XMomentum/YMomentum/Ppe, exported for unit parity:
    # src/momentum.f:864-866, 987, 1015, 1051, 1079, 1115-1118; src/pressure.f:329-330
    "convcoef_": (None, "iiii" "DDDD" "DD" "DD"),
    "dconvu_": (None, "ii" "DDD" "D"),
    "ddiffu_": (None, "ii" "DDDD" "D" "D"),
    "dconvv_": (None, "ii" "DDD" "D"),
    "ddiffv_": (None, "ii" "DDDD" "D" "D"),
    "poroscoef_": (None, "iiii" "III" "DDD" "DD" "D"),
    "rhsppe_": (None, "ii" "i" "d" "DD" "D" "D" "D"),
    # src/bound_cond.f:1030-1034
    "tempboundcond_": (None, "ii" "II" "II" "DD" "D"),
    # src/thermal.f:24-33
    "thermenergy_": (None, "ii" "II" "II" "dd" "DDD" "DDDDD" "DDDD" "DDDD" "DDDDDD"),
    # src/thermal.f:283-285
    "eqstate_": (None, "ii" "ddddd" "DDD"),
    # src/small_scale.f:31-49
    "smallscale_": (None, "iiii" "i" "II" "IIII" "ii" "dddd" "ddd" "dd" "D" "dddd" "ddd" "DD" "DDDD" "DDD"
                          "DDDD" "DDDD" "DDDD" "DDD" "DDDD"),
    # src/bound_cond.f:1209-1213
    "smlsclbc_": (None, "ii" "II" "IIII" "D" "DDDD"),
    # src/utility.f:512, 574
    "ptdavg_": (None, "ii" "III" "DD"),
    "velavg_": (None, "ii" "III" "DDDD"),
    # src/utility.f:668-671
    "taveraged_": (None, "iii" "II" "I" "D" "DD"),
    # src/traject.f:154-164
    "traject_": (None, "ii" "ii" "iii" "I" "ddddd" "DDD" "DD" "DDDD" "DD" "DDDD"),
}

_KIND = {"i": c_i32p, "o": c_i32p, "d": c_f64p, "I": c_i32p, "D": c_f64p}


def _as_array(x, dtype, name):
    if x is None:
        return None
    if not isinstance(x, np.ndarray) or x.dtype != dtype or not x.flags["C_CONTIGUOUS"]:
        raise TypeError(f"{name}: expected C-contiguous numpy array of {dtype}, got {type(x)} "
                        f"{getattr(x, 'dtype', None)}")
    return x


def bind(lib, prefix="", table=None):
    """Return {name_without_underscore: callable} for every routine in `table`.

    Each callable takes Python ints/floats for scalars and numpy arrays for arrays;
    'o' arguments are omitted from the call and returned (after the function result).
    """
    table = SIGNATURES if table is None else table
    out = {}
    for name, (restype, kinds) in table.items():
        fn = getattr(lib, prefix + name)
        fn.restype = restype
        fn.argtypes = [_KIND[k] for k in kinds]

        def call(*args, _fn=fn, _kinds=kinds, _name=name, _restype=restype):
            n_in = sum(1 for k in _kinds if k != "o")
            if len(args) != n_in:
                raise TypeError(f"{_name}: expected {n_in} arguments, got {len(args)}")
            cargs, outs, keep = [], [], []
            it = iter(args)
            for k in _kinds:
                if k == "o":
                    v = C.c_int32(0)
                    outs.append(v)
                    cargs.append(C.byref(v))
                    continue
                a = next(it)
                if k == "i":
                    v = C.c_int32(int(a)); keep.append(v); cargs.append(C.byref(v))
                elif k == "d":
                    v = C.c_double(float(a)); keep.append(v); cargs.append(C.byref(v))
                elif k == "I":
                    arr = _as_array(a, np.int32, _name)
                    cargs.append(arr.ctypes.data_as(c_i32p) if arr is not None else None)
                else:
                    arr = _as_array(a, np.float64, _name)
                    cargs.append(arr.ctypes.data_as(c_f64p) if arr is not None else None)
            r = _fn(*cargs)
            if outs:
                vals = tuple(o.value for o in outs)
                return (r,) + vals if _restype is not None else (vals[0] if len(vals) == 1 else vals)
            return r

        out[name.rstrip("_")] = call
    return out


# ---- plain-data descriptors of include/wolfd2_b200.h ---------------------------------

class Params(C.Structure):
    _fields_ = [("nx", C.c_int32), ("ny", C.c_int32), ("mqiter", C.c_int32),
                ("nmeiter", C.c_int32), ("nPpeSolver", C.c_int32), ("msorit", C.c_int32),
                ("lCartesGrid", C.c_int32), ("nfiltu", C.c_int32), ("nfiltv", C.c_int32),
                ("reserved_", C.c_int32),
                ("dk", C.c_double), ("re", C.c_double), ("fr", C.c_double),
                ("qtol", C.c_double), ("sortol", C.c_double), ("sorrel", C.c_double),
                ("fpu", C.c_double), ("fpv", C.c_double)]


class Regions(C.Structure):
    _fields_ = [("nReg", c_i32p), ("nRegBrd", c_i32p), ("nRegType", c_i32p),
                ("nMomBdTp", c_i32p), ("dBCVal", c_f64p), ("dPRporos", c_f64p),
                ("dPRporc1", c_f64p), ("dPRporc2", c_f64p)]


METRIC_NAMES = ("rau rbu rbv rgv ran rbn rgn rac rbc rgc dju djv djc djn "
                "xen yen xzn yzn xec yec xzc yzc xeu yeu xzv yzv xzu yzu xev yev").split()


class Metrics(C.Structure):
    _fields_ = [(n, c_f64p) for n in METRIC_NAMES]


class Thermal(C.Structure):
    """wolfd2_thermal of include/wolfd2_b200.h."""
    _fields_ = [("nthermen", C.c_int32), ("neqstate", C.c_int32), ("nfiltt", C.c_int32), ("reserved_", C.c_int32),
                ("pe", C.c_double), ("dmeittol", C.c_double), ("fpt", C.c_double),
                ("uref", C.c_double), ("densref", C.c_double), ("tmax", C.c_double), ("tref", C.c_double),
                ("rconst", C.c_double),
                ("nTRgType", c_i32p), ("nTemBdTp", c_i32p), ("dTRgVal", c_f64p), ("dHGSTval", c_f64p)]


class SmallScale(C.Structure):
    """wolfd2_smallscale of include/wolfd2_b200.h."""
    _fields_ = [("nsmallscl", C.c_int32), ("nssPpeSlvr", C.c_int32), ("mssSorIt", C.c_int32), ("reserved_", C.c_int32),
                ("dlref", C.c_double), ("uref", C.c_double), ("tref", C.c_double), ("tmax", C.c_double),
                ("pe", C.c_double), ("ssSorTol", C.c_double), ("ssSorRel", C.c_double),
                ("ssFiltPar", C.c_double * 4),
                ("ssCu0", C.c_double), ("ssTsCoef", C.c_double), ("ssHsCoef", C.c_double), ("ssTemCoef", C.c_double),
                ("ssBnCrit", C.c_double), ("ssRMpMax", C.c_double), ("ssRMpExp", C.c_double)]


class Traject(C.Structure):
    """wolfd2_traject of include/wolfd2_b200.h."""
    _fields_ = [("ntr", C.c_int32), ("ntsubstp", C.c_int32), ("nTrMethod", C.c_int32), ("nTrCdEq", C.c_int32),
                ("mTrHTmit", C.c_int32), ("reserved_", C.c_int32),
                ("densref", C.c_double), ("dTrHTtol", C.c_double), ("dTrHTdel", C.c_double)]


class StepLog(C.Structure):
    _fields_ = [("nQLiter", C.c_int32), ("nSorConv", C.c_int32), ("sor_converged", C.c_int32),
                ("diverged", C.c_int32), ("dif", C.c_double * 4)]
