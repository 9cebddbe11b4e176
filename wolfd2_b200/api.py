"""Python face of libwolfd2_b200.so (ctypes over the C ABI in include/wolfd2_b200.h).

Two layers, as in the header:
  * literal routines with the reference's names and argument lists -- ``nAuxMomentum``, ``XMomentum``,
    ``YMomentum``, ``AltTridLU``, ``Ppe``, ``Divergence``, ``Project``, ``VelBoundCond``,
    ``PresBoundCond``, ``VelOutflowBCs``, ``Filter``, ``DiffMaxNorm``, ``DMaxNorm`` -- taking host
    numpy arrays laid out as REAL*8 f(0:mnx,0:mny);
  * ``Context``: fields resident in HBM across time steps.

There is no CPU path here: importing works anywhere, but every call needs the CUDA library and a
B200; failures are loud (exception with the library's message, or abort() inside the literal shims,
which -- like the reference's `stop` -- have no error channel).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import Metrics, Params, Regions, SmallScale, StepLog, Thermal, Traject, c_f64p, c_i32p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("WOLFD2_B200_LIB") or os.path.join(_HERE, "libwolfd2_b200.so")   # (env: A/B builds of the same source)

F_U, F_V, F_P, F_US, F_VS, F_UN, F_VN, F_PN, F_D, F_DN, F_B, F_T, F_TS, F_TN, F_USS, F_VSS, F_PSS, F_TSS = range(18)

_lib = None
_fn = None


class Wolfd2Error(RuntimeError):
    pass


def lib():
    """Load the CUDA library; raise if it has not been built (no fallback)."""
    global _lib, _fn
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise Wolfd2Error(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). wolfd2_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    L.wolfd2_b200_last_error.restype = C.c_char_p
    L.wolfd2_b200_version.restype = C.c_char_p
    L.wolfd2_b200_config.argtypes = [C.c_int32] * 4
    L.wolfd2_b200_set_device.argtypes = [C.c_int32]
    L.wolfd2_b200_set_option.argtypes = [C.c_char_p, C.c_int32]
    L.wolfd2_b200_create.argtypes = [C.POINTER(C.c_void_p), C.POINTER(Params), C.POINTER(Regions), C.POINTER(Metrics)]
    L.wolfd2_b200_create_slab.argtypes = L.wolfd2_b200_create.argtypes + [C.c_int32, C.c_int32]
    L.wolfd2_b200_slab_layout.argtypes = [C.c_int32] * 4 + [C.POINTER(C.c_int32)]
    L.wolfd2_b200_comm_unique_id.argtypes = [C.POINTER(C.c_ubyte)]
    L.wolfd2_b200_comm_init.argtypes = [C.c_int32, C.c_int32, C.POINTER(C.c_ubyte)]
    L.wolfd2_b200_destroy.argtypes = [C.c_void_p]
    L.wolfd2_b200_destroy.restype = None
    L.wolfd2_b200_set_params.argtypes = [C.c_void_p, C.POINTER(Params)]
    L.wolfd2_b200_set_thermal.argtypes = [C.c_void_p, C.POINTER(Thermal)]
    L.wolfd2_b200_set_smallscale.argtypes = [C.c_void_p, C.POINTER(SmallScale)]
    L.wolfd2_b200_smallscale_init.argtypes = [C.c_void_p]
    L.wolfd2_b200_smallscale_map.argtypes = [C.c_void_p, C.c_int32, C.c_int32, c_f64p, C.c_int32]
    L.wolfd2_b200_set_trajectories.argtypes = [C.c_void_p, C.POINTER(Traject)] + [c_f64p] * 9 + [c_i32p]
    L.wolfd2_b200_get_particles.argtypes = [C.c_void_p] + [c_f64p] * 4 + [c_i32p]
    L.wolfd2_b200_node_averages.argtypes = [C.c_void_p, C.c_int32] + [c_f64p] * 4
    L.wolfd2_b200_timeavg.argtypes = [C.c_void_p, C.c_int32]
    L.wolfd2_b200_timeavg_finish.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_double, C.c_double]
    L.wolfd2_b200_timeavg_get.argtypes = [C.c_void_p, C.c_int32, c_f64p]
    L.wolfd2_b200_set_probes.argtypes = [C.c_void_p, C.c_int32, c_i32p, c_i32p, C.c_int32]
    L.wolfd2_b200_get_probe_records.argtypes = [C.c_void_p, C.c_int32, c_f64p, c_i32p, c_i32p]
    L.wolfd2_b200_upload_field.argtypes = [C.c_void_p, C.c_int32, c_f64p]
    L.wolfd2_b200_download_field.argtypes = [C.c_void_p, C.c_int32, c_f64p]
    L.wolfd2_b200_upload_metric_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, c_f64p]
    L.wolfd2_b200_gather_global.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
    L.wolfd2_b200_compare_global.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_uint64), C.POINTER(C.c_double)]
    L.wolfd2_b200_coldstart.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
    L.wolfd2_b200_step.argtypes = [C.c_void_p, C.c_int32, C.POINTER(StepLog)]
    L.wolfd2_b200_step_host.argtypes = [C.c_void_p, C.c_int32, c_f64p, c_f64p, c_f64p, C.POINTER(StepLog)]
    L.wolfd2_b200_last_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.wolfd2_b200_last_sor_timing.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_int64)]
    L.wolfd2_b200_last_host_syncs.argtypes = [C.c_void_p, C.POINTER(C.c_int64)]
    L.wolfd2_b200_sync.argtypes = [C.c_void_p]
    L.wolfd2_b200_host_alloc.argtypes = [C.c_uint64]
    L.wolfd2_b200_host_alloc.restype = C.c_void_p
    L.wolfd2_b200_host_free.argtypes = [C.c_void_p]
    L.wolfd2_b200_host_free.restype = None
    _lib = L
    _fn = _abi.bind(L)
    return L


def _check(rc, what):
    if rc != 0:
        raise Wolfd2Error(f"{what} failed (code {rc}): {lib().wolfd2_b200_last_error().decode()}")


def config(mnx, mny, mgri=20, mgrj=10):
    """Run-time stand-in for the reference's compile-time include/config.f."""
    _check(lib().wolfd2_b200_config(mnx, mny, mgri, mgrj), "wolfd2_b200_config")


def set_device(dev):
    _check(lib().wolfd2_b200_set_device(dev), "wolfd2_b200_set_device")


def set_option(name, value):
    _check(lib().wolfd2_b200_set_option(name.encode(), int(value)), "wolfd2_b200_set_option")


def _routine(name):
    def f(*args):
        lib()
        return _fn[name](*args)
    f.__name__ = name
    return f


# literal routines, reference spelling (argument order = the Fortran declarations cited in _abi.py)
nAuxMomentum = _routine("nauxmomentum")
XMomentum = _routine("xmomentum")
YMomentum = _routine("ymomentum")
AltTridLU = _routine("alttridlu")
Ppe = _routine("ppe")
Divergence = _routine("divergence")
Project = _routine("project")
VelBoundCond = _routine("velboundcond")
PresBoundCond = _routine("presboundcond")
VelOutflowBCs = _routine("veloutflowbcs")
Filter = _routine("filter")
TempBoundCond = _routine("tempboundcond")
ThermEnergy = _routine("thermenergy")
EqState = _routine("eqstate")
SmallScale_ = _routine("smallscale")     # `SmallScale` is the parameter block (ctypes struct)
SmlSclBC = _routine("smlsclbc")
PTDAvg = _routine("ptdavg")
VelAvg = _routine("velavg")
TAveraged = _routine("taveraged")
Traject_ = _routine("traject")           # `Traject` is the parameter block
DiffMaxNorm = _routine("diffmaxnorm")
DMaxNorm = _routine("dmaxnorm")
# the operators below XMomentum / YMomentum / Ppe, one at a time (unit parity)
ConvCoef = _routine("convcoef")
DConvU = _routine("dconvu")
DDiffU = _routine("ddiffu")
DConvV = _routine("dconvv")
DDiffV = _routine("ddiffv")
PorosCoef = _routine("poroscoef")
RhsPpe = _routine("rhsppe")


class Context:
    """Device-resident run: what `program wolfd2` holds after set-up, living in HBM."""

    def __init__(self, deck, stream_metrics=True):
        L = lib()
        self.deck = deck
        config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
        self._par, self._reg, self._met = deck.params(), deck.regions.as_struct(), deck.metrics_struct()
        h = C.c_void_p()
        if getattr(deck, "slab", None):   # one rank of a multi-GPU run: this deck holds rows A0..A1 only
            rank, world = deck.slab[0], deck.slab[1]
            _check(L.wolfd2_b200_create_slab(C.byref(h), C.byref(self._par), C.byref(self._reg), C.byref(self._met),
                                             rank, world), "wolfd2_b200_create_slab")
        else:
            _check(L.wolfd2_b200_create(C.byref(h), C.byref(self._par), C.byref(self._reg), C.byref(self._met)),
                   "wolfd2_b200_create")
        self._h = h
        if not deck.metrics and stream_metrics:   # uniform grid, metrics built and uploaded window by window
            # (stream_metrics=False: the caller fills them, e.g. Context.gather_global)
            self.stream_uniform_metrics()
        if getattr(deck, "thermal", False) or getattr(deck, "eqstate", False) or getattr(deck, "smallscale", False):
            self._th = deck.thermal_struct()   # cold ATD runs still need the thermal region tables
            _check(L.wolfd2_b200_set_thermal(h, C.byref(self._th)), "wolfd2_b200_set_thermal")
        if getattr(deck, "smallscale", False):
            self._ss = deck.smallscale_struct()
            _check(L.wolfd2_b200_set_smallscale(h, C.byref(self._ss)), "wolfd2_b200_set_smallscale")

    def close(self):
        if getattr(self, "_h", None):
            try:
                lib().wolfd2_b200_destroy(self._h)
            except TypeError:      # interpreter shutdown: the module globals are already gone
                pass
            self._h = None

    __del__ = close

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def set_params(self, **kw):
        for k, v in kw.items():
            setattr(self._par, k, v)
        _check(lib().wolfd2_b200_set_params(self._h, C.byref(self._par)), "wolfd2_b200_set_params")

    def set_thermal(self, **kw):
        """Change fields of the wolfd2_thermal block (e.g. nthermen=0) for the following steps."""
        if not hasattr(self, "_th"):
            self._th = self.deck.thermal_struct()
        for k, v in kw.items():
            setattr(self._th, k, v)
        _check(lib().wolfd2_b200_set_thermal(self._h, C.byref(self._th)), "wolfd2_b200_set_thermal")

    def smallscale_init(self):
        """src/main.f:643-665: SmallScale(initflg=0) on the current u, v, t."""
        _check(lib().wolfd2_b200_smallscale_init(self._h), "wolfd2_b200_smallscale_init")

    def smallscale_map(self, family, plane, arr=None):
        """Download (arr None) or upload one plane of the saved chaotic-map iterates."""
        out = self.deck.new_field() if arr is None else arr
        _check(lib().wolfd2_b200_smallscale_map(self._h, family, plane, out.ctypes.data_as(c_f64p), 0 if arr is None else 1),
               "wolfd2_b200_smallscale_map")
        return out

    def set_trajectories(self, tr, x, y, cpartx, cparty, repc, xp, yp, up, vp, out=None):
        """tr: _abi.Traject; x, y: grid nodes (0:mnx,0:mny); particle arrays of tr.ntr float64 (out: int32 or None)."""
        a = [np.ascontiguousarray(q, dtype=np.float64) for q in (x, y, cpartx, cparty, repc, xp, yp, up, vp)]
        o = None if out is None else np.ascontiguousarray(out, dtype=np.int32)
        _check(lib().wolfd2_b200_set_trajectories(self._h, C.byref(tr), *[q.ctypes.data_as(c_f64p) for q in a],
                                                  o.ctypes.data_as(c_i32p) if o is not None else None),
               "wolfd2_b200_set_trajectories")
        self._ntr = tr.ntr

    def particles(self):
        n = self._ntr
        xp, yp, up, vp = (np.zeros(n) for _ in range(4))
        out = np.zeros(n, dtype=np.int32)
        _check(lib().wolfd2_b200_get_particles(self._h, *[q.ctypes.data_as(c_f64p) for q in (xp, yp, up, vp)],
                                               out.ctypes.data_as(c_i32p)), "wolfd2_b200_get_particles")
        return xp, yp, up, vp, out

    def node_averages(self, small_scale=False, temperature=False):
        """(util, vbar, pav[, tav]): VelAvg / PTDAvg / TAveraged of the resident fields, computed on the device
        (output dumps)."""
        out = [self.deck.new_field() for _ in range(4 if temperature else 3)]
        ptr = [q.ctypes.data_as(c_f64p) for q in out] + ([] if temperature else [None])
        _check(lib().wolfd2_b200_node_averages(self._h, 1 if small_scale else 0, *ptr), "wolfd2_b200_node_averages")
        return tuple(out)

    def stream_uniform_metrics(self, chunk_rows=None, threads=None):
        """Metric arrays of a uniform grid, built by deck.metrics_window in windows of chunk_rows rows (several
        windows at a time on host threads: numpy releases the GIL) and uploaded with wolfd2_b200_upload_metric_rows.
        Bit-identical to the global arrays (tests/test_slab_cpu.py); the host never holds more than a few windows."""
        from concurrent.futures import ThreadPoolExecutor
        from . import deck as dk
        d = self.deck
        a0, a1 = (d.slab[4], d.slab[5]) if d.slab else (0, d.ny + 1)
        if chunk_rows is None:
            chunk_rows = max(16, min(512, (1 << 22) // (d.nx + 2)))     # ~32 MB per array window
        if threads is None:
            threads = max(1, min(8, (os.cpu_count() or 2) // max(1, int(os.environ.get("LOCAL_WORLD_SIZE", "1")))))
        wins = [(j, min(a1, j + chunk_rows - 1)) for j in range(a0, a1 + 1, chunk_rows)]
        L = lib()

        def build(w):
            return w, dk.metrics_window(d.nx, d.ny, w[0], w[1], d.mnx, d.dlref)

        with ThreadPoolExecutor(max_workers=threads) as ex:
            for (j0, j1), m in ex.map(build, wins):
                for k, name in enumerate(_abi.METRIC_NAMES):
                    _check(L.wolfd2_b200_upload_metric_rows(self._h, k, j0, j1 - j0 + 1, m[name].ctypes.data_as(c_f64p)),
                           "wolfd2_b200_upload_metric_rows")

    def gather_global(self, glob, metrics=True, fields=True):
        """Slab context: collect every rank's rows into the one-GPU context `glob` on rank 0 (None elsewhere)."""
        _check(lib().wolfd2_b200_gather_global(self._h, glob._h if glob is not None else None,
                                               (1 if metrics else 0) | (2 if fields else 0)), "wolfd2_b200_gather_global")

    def compare_global(self, glob, which):
        """(cells whose bit patterns differ, max |difference|) of field `which`: slab run vs `glob` (rank 0)."""
        n, m = C.c_uint64(0), C.c_double(0.0)
        _check(lib().wolfd2_b200_compare_global(self._h, glob._h if glob is not None else None, which, C.byref(n), C.byref(m)),
               "wolfd2_b200_compare_global")
        return int(n.value), float(m.value)

    TIMEAVG_NAMES = ("ubar vbar tbar pbar upb vpb tpb upupb vpvpb upvpb uptpb vptpb upxsb upysb vpxsb vpysb "
                     "trbke dssrt dtdyb").split()

    def timeavg(self, op):
        """Time-averaging mode (-D_TIMEAVG_): op 'begin', 1 / 2 (accumulate pass 1 / 2 from now on), 'stop', 'release'."""
        code = {"begin": 0, 1: 1, 2: 2, "stop": 3, "release": 4}[op]
        _check(lib().wolfd2_b200_timeavg(self._h, code), "wolfd2_b200_timeavg")

    def timeavg_finish(self, npass, nts):
        _check(lib().wolfd2_b200_timeavg_finish(self._h, npass, nts, self.deck.uref, self.deck.dlref), "wolfd2_b200_timeavg_finish")

    def timeavg_get(self, name):
        out = self.deck.new_field()
        _check(lib().wolfd2_b200_timeavg_get(self._h, self.TIMEAVG_NAMES.index(name), out.ctypes.data_as(c_f64p)),
               "wolfd2_b200_timeavg_get")
        return out

    def set_probes(self, iTS, jTS, freq=1):
        """Time-series monitor points (SaveTimeSrs): sampled inside every freq-th step from the resident fields."""
        i = np.ascontiguousarray(iTS, dtype=np.int32)
        j = np.ascontiguousarray(jTS, dtype=np.int32)
        self._nprobes = int(i.size)
        _check(lib().wolfd2_b200_set_probes(self._h, i.size, i.ctypes.data_as(c_i32p), j.ctypes.data_as(c_i32p), int(freq)),
               "wolfd2_b200_set_probes")

    def probe_records(self):
        """(steps[nrec], records[nrec][npoints][8]) gathered since the last call; columns u, v, p, t, uss, vss, pss, tss."""
        n = C.c_int32(0)
        _check(lib().wolfd2_b200_get_probe_records(self._h, 0, None, None, C.byref(n)), "wolfd2_b200_get_probe_records")
        out = np.zeros((max(n.value, 1), self._nprobes, 8))
        steps = np.zeros(max(n.value, 1), dtype=np.int32)
        _check(lib().wolfd2_b200_get_probe_records(self._h, n.value, out.ctypes.data_as(c_f64p), steps.ctypes.data_as(c_i32p),
                                                   C.byref(n)), "wolfd2_b200_get_probe_records")
        return steps[:n.value], out[:n.value]

    def upload(self, which, arr):
        assert arr.dtype == np.float64 and arr.flags["C_CONTIGUOUS"]
        assert arr.shape == (self.deck.mny + 1, self.deck.mnx + 1)
        _check(lib().wolfd2_b200_upload_field(self._h, which, arr.ctypes.data_as(c_f64p)), "upload_field")

    def download(self, which, out=None):
        out = self.deck.new_field() if out is None else out
        _check(lib().wolfd2_b200_download_field(self._h, which, out.ctypes.data_as(c_f64p)), "download_field")
        return out

    def coldstart(self):
        n = C.c_int32(0)
        _check(lib().wolfd2_b200_coldstart(self._h, C.byref(n)), "wolfd2_b200_coldstart")
        return n.value

    @staticmethod
    def _logs(raw):
        return [dict(nQLiter=l.nQLiter, nSorConv=l.nSorConv, sor_converged=l.sor_converged,
                     diverged=l.diverged, dif=list(l.dif)) for l in raw]

    def step(self, nsteps=1):
        raw = (StepLog * nsteps)()
        _check(lib().wolfd2_b200_step(self._h, nsteps, raw), "wolfd2_b200_step")
        return self._logs(raw)

    def step_host(self, u, v, p, nsteps=1):
        raw = (StepLog * nsteps)()
        _check(lib().wolfd2_b200_step_host(self._h, nsteps, u.ctypes.data_as(c_f64p), v.ctypes.data_as(c_f64p),
                                           p.ctypes.data_as(c_f64p), raw), "wolfd2_b200_step_host")
        return self._logs(raw)

    def timing(self):
        ms = (C.c_double * 4)()
        ln = (C.c_int64 * 4)()
        _check(lib().wolfd2_b200_last_timing(self._h, ms, ln), "last_timing")
        sm, si = C.c_double(0), C.c_int64(0)
        _check(lib().wolfd2_b200_last_sor_timing(self._h, C.byref(sm), C.byref(si)), "last_sor_timing")
        hs = C.c_int64(0)
        _check(lib().wolfd2_b200_last_host_syncs(self._h, C.byref(hs)), "last_host_syncs")
        return dict(total_ms=ms[0], momentum_ms=ms[1], ppe_ms=ms[2], other_ms=ms[3], host_syncs=hs.value,
                    launches=dict(momentum=ln[1], ppe=ln[2], other=ln[3]),
                    sor_ms=sm.value, sor_iters=si.value)

    def sync(self):
        _check(lib().wolfd2_b200_sync(self._h), "sync")


def slab_layout(nx, ny, world, rank):
    """(J0, J1, A0, A1, HG) as the library computes it (wolfd2_b200_slab_layout)."""
    out = (C.c_int32 * 5)()
    _check(lib().wolfd2_b200_slab_layout(nx, ny, world, rank, out), "wolfd2_b200_slab_layout")
    return tuple(out)


def comm_finalize():
    lib().wolfd2_b200_comm_finalize()


def pinned_field(deck):
    """A zeroed REAL*8 f(0:mnx,0:mny) in page-locked host memory (for the e2e path)."""
    n = (deck.mny + 1) * (deck.mnx + 1)
    ptr = lib().wolfd2_b200_host_alloc(n * 8)
    if not ptr:
        raise Wolfd2Error("pinned allocation failed: " + lib().wolfd2_b200_last_error().decode())
    buf = (C.c_double * n).from_address(ptr)
    arr = np.frombuffer(buf, dtype=np.float64).reshape(deck.mny + 1, deck.mnx + 1)
    return arr, ptr


def pinned_free(ptr):
    lib().wolfd2_b200_host_free(ptr)
