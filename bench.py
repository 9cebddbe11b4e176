#!/usr/bin/env python
"""bench.py -- throughput of the wolfd2 time-step hot path on B200 (contract in the task brief).

A "step" is one pass of the step body src/main.f:690-981 (QL momentum loop, PPE red/black SOR,
projection, ghost fills, norms) over one synthetic grid.

Workloads (BASELINE.json):
  * 1 GPU (default): lid-driven cavity Re=1000 on the metric's 4096 x 4096 grid (deck family of configs[1]).
  * N > 1 GPUs (default): configs[3], the cavity on 16384 x 16384 cut into N row slabs (strong scaling over
    N = 2, 4, 8; the N = 1 line stays the 4096^2 metric grid).  `--scaling weak --n 4096` gives the round-1
    secondary profile (n x n*N grid, n rows per GPU).
  Both run in FIXED-WORK mode (ql_tolerance 0, max_ql_iter 2, sor_tolerance 0, max_sor_iter 100 -- SURVEY.md
  section 8d) so that the algorithmic bytes per step are known a priori: cells*(256*Q + 56 + 64*S + 88 + 16) B.
  * `extra` (1 GPU, default run only): one compact entry each for configs[1] (cavity 1024^2), configs[2]
    (channel / backward step 4096^2, inflow + outflow) and configs[4] (ATD small-scale model + 10^6 particles).

  value : Gcell-updates/s = (nx-1)(ny-1)*K / device time (CUDA events, max over ranks), fields resident in HBM.
  e2e   : same metric through wolfd2_b200_step_host with pinned HOST u,v,p buffers, i.e. H2D of
          the state before and D2H after every step inside the timed region (wall clock between barriers).
  verify: N > 1 only.  After the timed runs every rank's rows (metrics and u, v, p) are gathered into a one-GPU
          context of the SAME global grid on rank 0; both then advance 2 more steps and the fields are compared
          bit for bit on the device (wolfd2_b200_compare_global), together with the per-step log tuples.
  --impl reference : the reference's CPU implementation.  The Fortran reference cannot be built in
          this image (no Fortran front-end), so this times the C restatement in oracle/
          (cpu_baseline.kind = "port"), 1 thread because the reference is serial.  It runs the GPU arm's own
          grid when K+W steps of it fit the time budget (4096^2: about 7 min), else a bounded sample of the
          same deck family; `config.grid_run` always names the grid that actually ran.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def stable_dt(n, re):
    """The reference's split steps are both solved along i, so j-direction diffusion is only
    iterated explicitly by the QL loop: dt must respect dt/(Re h^2) <~ 0.25 as well as a CFL of 0.25."""
    h = 1.0 / (n - 1)
    return min(0.25 * h, 0.2 * re * h * h)


def sor_omega(n):
    """Optimal over-relaxation factor of point SOR for an n-point Laplacian, 2/(1+sin(pi h)): with it the red/black SOR
    reaches sor_tolerance 1e-8 in about 2.5 n iterations (measured with the oracle at 256^2 and 512^2; with 1.9 the
    4096^2 solve needs > 10^5)."""
    import math
    return 2.0 / (1.0 + math.sin(math.pi / (n - 1)))


def resolve_grid(args, world):
    """(n, ny_global, scaling) of the GPU arm.  1 GPU: the metric grid 4096^2.  N > 1: BASELINE configs[3],
    16384^2 in N row slabs (strong), unless --n / --scaling ask for the secondary weak profile."""
    if world == 1:
        n = args.n or 4096
        return n, n, "weak"
    scaling = args.scaling or "strong"
    n = args.n or (16384 if scaling == "strong" else 4096)
    return n, (n if scaling == "strong" else n * world), scaling


def make_deck(workload, n, fixed_work, q_iters, s_iters, ny=None, slab=None, atd=False, lazy=False, msorit=20000,
              outlet="fully_dev"):
    from wolfd2_b200 import deck as dk
    ny = ny or n
    kw = {"ny": ny}
    if atd:   # BASELINE.json configs[4]: ATD small-scale model on; its own Ppe does the same fixed work
        kw.update(smallscale=True, ss_cu0=0.02, ss_bncrit=2.0, ss_ppe_solver="rb_sor",
                  ss_msorit=s_iters if fixed_work else 2000, ss_sortol=0.0 if fixed_work else 1e-8)
    if slab is not None:
        kw["slab"] = slab          # (rank, world): this rank's rows only
    if lazy:
        kw["lazy_metrics"] = True  # metrics built and uploaded window by window (api.Context)
    nmax = max(n, ny)
    fd = outlet == "fully_dev"
    if workload == "cavity":
        d = dk.cavity(n, re=1000.0, dt=stable_dt(nmax, 1000.0), **kw)
    elif workload == "channel":
        d = dk.channel(n, re=100.0, dt=stable_dt(nmax, 100.0), fully_dev=fd, **kw)
    elif workload == "bstep":
        d = dk.backward_step(n, re=100.0, dt=stable_dt(nmax, 100.0), fully_dev=fd, **kw)
    else:
        raise SystemExit(f"unknown workload {workload}")
    d.ppe_solver = "rb_sor"
    if fixed_work:
        d.qtol, d.mqiter, d.sortol, d.msorit = 0.0, q_iters, 0.0, s_iters
    else:   # converged mode: the reference's tolerances, the optimal point-SOR factor of the grid, a cap it does not hit
        d.sorrel, d.msorit = sor_omega(nmax), msorit
    return d


def developed_state(d):
    """A smooth, non-quiescent restart field on the staggered locations of src/grid.f:337-359, so that the timed
    steps do not run on a mostly-zero cold start: one vortex filling the box for the closed cavity; for the
    inflow/outflow decks the plug flow u = 1 the inlet feeds (zero inside a blockage).  Returns u, v, p."""
    nx, ny = d.nx, d.ny
    a0, a1 = (d.slab[4], d.slab[5]) if d.slab else (0, ny + 1)     # rows this rank holds
    u, v, p = d.new_field(), d.new_field(), d.new_field()
    nr = a1 - a0 + 1
    if not d.name.startswith("cavity"):
        u[:nr, :nx + 2] = 1.0
        r = d.regions
        for jr in range(int(r.nReg[1])):
            for ir in range(int(r.nReg[0])):
                if int(r.nRegType[jr, ir]) == 0:       # RM_BLOCKG (include/wolfd2.h)
                    iw, ie, js, jn = (int(r.nRegBrd[k, jr, ir]) for k in range(4))
                    lo, hi = max(js, a0), min(jn + 1, a1)
                    if hi >= lo:
                        u[lo - a0:hi - a0 + 1, iw:ie + 1] = 0.0
        return u, v, p
    xi = (np.arange(nx + 2) - 1.0) / (nx - 1.0)
    yj = (np.arange(a0, a1 + 1) - 1.0) / (ny - 1.0)
    xh, yh = xi - 0.5 / (nx - 1.0), yj - 0.5 / (ny - 1.0)
    A = 0.2
    # psi = A sin^2(pi x) sin^2(pi y);  u = dpsi/dy at (x_i, y_{j-1/2}),  v = -dpsi/dx at (x_{i-1/2}, y_j)
    u[:nr, :nx + 2] = A * np.outer(np.pi * np.sin(2 * np.pi * yh), np.sin(np.pi * xi) ** 2)
    v[:nr, :nx + 2] = -A * np.outer(np.sin(np.pi * yj) ** 2, np.pi * np.sin(2 * np.pi * xh))
    p[:nr, :nx + 2] = 0.05 * np.outer(np.cos(np.pi * yh), np.cos(np.pi * xh))
    return u, v, p


def algorithmic_bytes_per_step(cells, q, s):
    return cells * (256.0 * q + 56.0 + 64.0 * s + 88.0 + 16.0)


class ClockSampler:
    """nvidia-smi clock / throttle sampling during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic():
    """dram bytes per SOR-sweep launch from the committed ncu summary, if any."""
    p = os.path.join(ROOT, "profiles", "sor_traffic.json")
    try:
        return json.load(open(p))
    except Exception:
        return None


# --------------------------------------------------------------------------- CPU arm
def oracle_lib(opt=True):
    """Build (on this host) and load the CPU port.  Only used for cpu_baseline / --impl reference."""
    odir = os.path.join(ROOT, "oracle")
    src = os.path.join(odir, "wolfd2_oracle.c")
    bdir = os.path.join(odir, "_build")
    os.makedirs(bdir, exist_ok=True)
    if opt:
        out = os.path.join(bdir, "liboracle_opt_native.so")
        cmd = ["gcc", "-std=c99", "-O3", "-march=native", "-ffast-math", "-fPIC", "-shared", "-o", out, src, "-lm"]
        kind = "C restatement of the reference (oracle/), gcc -O3 -march=native -ffast-math ~ wolfd2_opt (Makefile:54)"
    else:
        out = os.path.join(bdir, "liboracle_O2.so")
        cmd = ["gcc", "-std=c99", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-o", out, src, "-lm"]
        kind = "C restatement of the reference (oracle/), gcc -O2 ~ reference `make` (Makefile:46)"
    if not os.path.exists(out) or os.path.getmtime(out) < os.path.getmtime(src):
        subprocess.check_call(cmd)
    return C.CDLL(out), kind


def cpu_steps(deck, nsteps, warm=0, opt=True):
    """Time nsteps of orc_step (+coldstart, untimed) on one core; returns (seconds, kind)."""
    from wolfd2_b200 import _abi
    lib, kind = oracle_lib(opt)
    lib.orc_config.argtypes = [C.c_int32] * 4
    lib.orc_step.restype = C.c_int32
    lib.orc_step.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions), C.POINTER(_abi.Metrics)] + \
        [_abi.c_f64p] * 5 + [C.c_int32, C.POINTER(_abi.StepLog)]
    lib.orc_coldstart.argtypes = [C.POINTER(_abi.Params), C.POINTER(_abi.Regions), C.POINTER(_abi.Metrics),
                                  _abi.c_f64p, _abi.c_f64p, _abi.c_f64p, _abi.c_i32p]
    lib.orc_config(deck.mnx, deck.mny, deck.regions.mgri, deck.regions.mgrj)
    par, reg, met = deck.params(), deck.regions.as_struct(), deck.metrics_struct()
    f = list(developed_state(deck)) + [deck.new_field() for _ in range(2)]
    ptr = [a.ctypes.data_as(_abi.c_f64p) for a in f]
    n = C.c_int32(0)
    lib.orc_coldstart(C.byref(par), C.byref(reg), C.byref(met), ptr[0], ptr[1], ptr[2], C.byref(n))
    if warm:
        lib.orc_step(C.byref(par), C.byref(reg), C.byref(met), *ptr, warm, None)
    t0 = time.perf_counter()
    rc = lib.orc_step(C.byref(par), C.byref(reg), C.byref(met), *ptr, nsteps, None)
    dt = time.perf_counter() - t0
    if rc != 0:
        raise RuntimeError("CPU port diverged")
    return dt, kind


def cpu_seconds_per_cell(args):
    """CPU port: seconds per cell per step, measured on a 256^2 grid of the same deck family and mode."""
    d = make_deck(args.workload, 256, args.fixed_work, args.q_iters, args.s_iters)
    t, _ = cpu_steps(d, 1)
    return t / d.cells()


def cpu_sample_size(args, budget_s, nsteps, nmax):
    """Grid size n <= nmax such that nsteps of the CPU port take about budget_s."""
    per_cell = cpu_seconds_per_cell(args)
    n = int((budget_s / max(nsteps, 1) / per_cell) ** 0.5)
    n = max(128, min(nmax, n))
    return (n // 64) * 64 if n >= 256 else n, per_cell


def mode_text(args):
    return ("fixed-work: ql_tolerance 0, max_ql_iter %d, sor_tolerance 0, max_sor_iter %d" % (args.q_iters, args.s_iters)
            if args.fixed_work else "converged: ql_tolerance 1e-4, sor_tolerance 1e-8, sor_relaxation 2/(1+sin(pi h)), max_sor_iter %d" % args.max_sor)


def config_dict(args, world):
    """The GPU arm's workload; the reference arm prints the same dict plus `grid_run`, the grid it really ran."""
    n, nyg, scaling = resolve_grid(args, world)
    cells = (n - 1) * (nyg - 1)
    re = 1000.0 if args.workload == "cavity" else 100.0
    dt = stable_dt(max(n, nyg), re)
    which = ("configs[3]" if (world > 1 and n == 16384 and nyg == 16384) else
             {"channel": "configs[2]", "cavity": "metric grid 4096^2; deck family of configs[1]", "bstep": "configs[2]"}[args.workload])
    return {"workload": f"{args.workload} {n}x{nyg} uniform grid, Re={re:g}, dt={dt:g}, ppe_solver rb_sor (BASELINE.json {which})",
            "mode": mode_text(args),
            "grid": [n, nyg], "cells": cells,
            "l2": "working set (>= 40 arrays x %.0f MB per GPU) exceeds the 126 MB L2; no flush needed" % (cells / world * 8 / 1e6)
            if cells / world * 8 * 4 > 126e6 else "working set fits L2 (latency-bound regime)",
            "parallelism": "1 GPU" if world == 1 else
            f"{world} row slabs of ~{(nyg - 1) // world} rows ({scaling} scaling), one process per GPU; halo rows of us,vs per QL "
            f"iteration (NCCL) and of p per fused SOR pass (peer stores over NVLink), all-reduced max-norms and tridiagonal "
            f"segment records (wolfd2_b200/csrc/w2_dist.cu); results bit-identical to one GPU (see verify)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except Exception:
        pass
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    n_arm, ny_arm, scaling = resolve_grid(args, world)
    nsteps = args.steps + args.warmup
    per_cell = cpu_seconds_per_cell(args)
    est = per_cell * (n_arm - 1) * (ny_arm - 1) * nsteps
    if n_arm == ny_arm and est <= args.ref_budget:
        n, why = n_arm, "the GPU arm's own grid"
    else:
        n, _ = cpu_sample_size(args, budget_s=min(args.ref_budget, 150.0), nsteps=nsteps, nmax=min(n_arm, 4096))
        why = (f"bounded sample: {nsteps} steps of the GPU arm's {n_arm}x{ny_arm} grid would take about {est / 60:.0f} min "
               f"on one core")
    d = make_deck(args.workload, n, args.fixed_work, args.q_iters, args.s_iters)
    t, kind = cpu_steps(d, args.steps, warm=args.warmup)
    value = d.cells() * args.steps / t / 1e9
    sample = (f"{args.workload} {n}x{n}, dt={d.dt:g} ({why}), {args.steps} steps after {args.warmup} warm-up, "
              f"1 thread (the reference is serial)")
    cfg = config_dict(args, world)
    cfg["grid_run"] = [n, n]
    cfg["cells_run"] = d.cells()
    cfg["same_grid_as_gpu_arm"] = bool(n == n_arm and n == ny_arm)
    line = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "build": kind},
        "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------- GPU arm
def timed_steps(ctx, steps, warmup, barrier):
    """W untimed steps, then K steps between barriers: device time (CUDA events inside the library), wall time."""
    ctx.step(warmup)
    l0 = ctx.timing()["launches"]
    barrier()
    t0 = time.perf_counter()
    logs = ctx.step(steps)
    barrier()
    wall = time.perf_counter() - t0
    tm = ctx.timing()
    launches = sum(tm["launches"].values()) - sum(l0.values())
    return logs, tm, wall, launches


def iterations_of(logs, d):
    q = [abs(l["nQLiter"]) if l["nQLiter"] > 0 else d.mqiter for l in logs]
    s = [l["nSorConv"] for l in logs]
    return float(np.mean(q)), float(np.mean(s))


def run_extra(label, d, steps, warmup, particles=0, atd=False):
    """One compact bench entry for a secondary BASELINE config on one GPU (own context, device-resident)."""
    from wolfd2_b200 import api
    out = {"config": label, "grid": [d.nx, d.ny], "cells": d.cells(), "steps": steps, "warmup": warmup}
    try:
        with api.Context(d) as ctx:
            for w, f in zip((api.F_U, api.F_V, api.F_P), developed_state(d)):
                ctx.upload(w, f)
            ctx.coldstart()
            if atd:
                ctx.smallscale_init()
            if particles:
                set_particles(ctx, d, particles)
            logs, tm, wall, launches = timed_steps(ctx, steps, warmup, ctx.sync)
            q, s = iterations_of(logs, d)
            ms = tm["total_ms"] / steps
            out.update({"value": d.cells() / (ms * 1e-3) / 1e9, "unit": "Gcell-updates/s", "ms_per_step": ms,
                        "wall_ms_per_step": wall / steps * 1e3, "steps_per_s": 1e3 / ms,
                        "iterations": {"ql_per_step": q, "sor_per_step": s,
                                       "sor_converged": bool(all(l["sor_converged"] for l in logs))},
                        "sections_ms_per_step": {"momentum": tm["momentum_ms"] / steps, "ppe": tm["ppe_ms"] / steps,
                                                 "other": tm["other_ms"] / steps},
                        "sor_us_per_iteration": tm["sor_ms"] / max(tm["sor_iters"], 1) * 1e3,
                        "dif_last": logs[-1]["dif"][:3], "gpu_launches": int(launches),
                        "host_syncs_per_step": tm["host_syncs"] / steps})
            if particles:
                gxp, gyp, gup, gvp, gout = ctx.particles()
                out["particles"] = {"n": int(gxp.size), "in_bounds": int((gout == 0).sum()),
                                    "finite": bool(np.isfinite(gxp).all() and np.isfinite(gup).all())}
    except Exception as e:      # an extra entry never takes the headline line down
        out["error"] = str(e)[:300]
    return out


def set_particles(ctx, d, particles):
    """configs[4]: particles on a uniform lattice, FwdEuler, Stokes drag."""
    from wolfd2_b200 import _abi
    side = max(1, int(round(particles ** 0.5)))
    gx, gy = d.node_arrays()
    lat = (np.arange(side) + 0.5) / side * 0.8 + 0.1
    xp, yp = [a.ravel().copy() for a in np.meshgrid(lat, lat)]
    npart = xp.size
    tr = _abi.Traject()
    tr.ntr, tr.ntsubstp, tr.nTrMethod, tr.nTrCdEq, tr.mTrHTmit = npart, 1, 2, 1, 1
    tr.densref, tr.dTrHTtol, tr.dTrHTdel = 1.2, 1e-8, 1.0
    ctx.set_trajectories(tr, gx, gy, np.full(npart, 2.0), np.full(npart, 2.0), np.full(npart, 10.0),
                         xp, yp, np.zeros(npart), np.zeros(npart))


def verify_against_one_gpu(ctx, d, args, dist, rank, world, nsteps=2):
    """Multi-GPU correctness, visible to the driver: gather the slab run into a one-GPU context of the same
    global grid on rank 0, advance both nsteps more steps, compare u, v, p bit for bit and the log tuples."""
    import torch
    from wolfd2_b200 import api
    glob, why, light = None, "", ""
    if rank == 0:
        try:
            g = make_deck(args.workload, d.nx, args.fixed_work, args.q_iters, args.s_iters, ny=d.ny, lazy=True)
            glob = api.Context(g, stream_metrics=False)
            # arrays the first step of a one-GPU context allocates on demand (the four colour-split SOR buffers, the
            # coefficient tiles, the second momentum stream's work arrays: about 11 fields) must fit as well
            field_bytes = 8 * (d.nx + 18) * (d.ny + 3)
            free_b, _ = torch.cuda.mem_get_info()
            if free_b < 11.5 * field_bytes + (1 << 30):
                # Not enough for those: the one-GPU run then takes the plain half-sweep SOR kernels (no colour-split
                # buffers, no tiles) and one momentum stream, which need nothing beyond the context itself and give
                # the same bits (tests/test_gpu_fullsize.py) -- the slab run keeps its fused, tiled kernels.
                if free_b > 0.25 * field_bytes + (1 << 30):
                    light = (f"one-GPU run with the plain half-sweep SOR kernels, one momentum stream, no explicit-term cache: {free_b / 1e9:.0f} GB "
                             f"free next to rank 0's slab, {(11.5 * field_bytes + (1 << 30)) / 1e9:.0f} GB needed for the fused path")
                else:
                    why = (f"{free_b / 1e9:.0f} GB free after creating the one-GPU context")
                    glob.close()
                    glob = None
        except Exception as e:
            why = str(e)[:200]
    flag = torch.tensor([1 if (rank != 0 or glob is not None) else 0], device="cuda")
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if int(flag.item()) == 0:
        return {"skipped": f"one-GPU context of {d.nx}x{d.ny} could not be created next to rank 0's slab: {why}"}
    ctx.gather_global(glob, metrics=True, fields=True)
    logs_s = ctx.step(nsteps)
    ms1 = None
    if rank == 0:
        if light:   # none of these changes a bit of the result (tests/test_gpu_fullsize.py, test_gpu_momentum_np.py)
            api.set_option("sor_fused_T", 0)
            api.set_option("mom_two_streams", 0)
            api.set_option("mom_np_cache", 0)      # (its cache of v's explicit terms lives in two colour-split SOR buffers)
        try:
            logs_g = glob.step(nsteps)
        finally:
            if light:
                api.set_option("sor_fused_T", -1)
                api.set_option("mom_two_streams", 1)
                api.set_option("mom_np_cache", 1)
        ms1 = glob.timing()["total_ms"] / nsteps
    res = {}
    for nm, w in (("u", api.F_U), ("v", api.F_V), ("p", api.F_P)):
        nd, mx = ctx.compare_global(glob, w)
        res[nm] = {"cells_differing": nd, "max_abs_diff": mx}
    out = None
    if rank == 0:
        same_logs = all(a["nQLiter"] == b["nQLiter"] and a["nSorConv"] == b["nSorConv"] and a["dif"] == b["dif"]
                        for a, b in zip(logs_s, logs_g))
        out = {"grid": [d.nx, d.ny], "steps": nsteps, "fields": res, "logs_identical": bool(same_logs),
               "identical": bool(same_logs and all(v["cells_differing"] == 0 for v in res.values())),
               "dif_last": logs_s[-1]["dif"][:3], "dif_last_one_gpu": logs_g[-1]["dif"][:3],
               "iterations_last": [logs_s[-1]["nQLiter"], logs_s[-1]["nSorConv"]],
               "one_gpu_ms_per_step": ms1, **({"one_gpu_variant": light} if light else {}),
               "how": "wolfd2_b200_gather_global + wolfd2_b200_compare_global: every cell 0..nx+1 x 0..ny+1 of u, v, p "
                      "compared by bit pattern on the device after both runs advanced the same state"}
        glob.close()
    return out


def run_gpu(args):
    from wolfd2_b200 import api
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    api.set_device(local)
    for o in args.opt:
        k, v = o.split("=")
        api.set_option(k, int(v))
    from wolfd2_b200 import slab
    n, nyg, scaling = resolve_grid(args, world)
    lazy = (n - 1) * (nyg - 1) // world > 20e6      # large slabs: metrics built and uploaded window by window, on host threads
    if world > 1:
        slab.init_comm(dist, local)     # the library's own NCCL communicator; torch only carries the id
        d = make_deck(args.workload, n, args.fixed_work, args.q_iters, args.s_iters, ny=nyg, slab=(rank, world), lazy=lazy,
                      outlet=args.outlet)
    else:
        d = make_deck(args.workload, n, args.fixed_work, args.q_iters, args.s_iters, atd=args.atd, lazy=lazy,
                      outlet=args.outlet, msorit=args.max_sor)
    if world > 1 and (args.atd or args.particles):
        raise SystemExit("bench: --atd / --particles run on one GPU (SURVEY 8f N2/N3 are single-GPU rows)")
    cells = (d.nx - 1) * (d.ny - 1)          # pressure unknowns of the whole grid
    cells_local = (d.nx - 1) * (d.slab[3] - d.slab[2] + 1) if d.slab else cells
    t_setup = time.perf_counter()
    ctx = api.Context(d)
    for w, f in zip((api.F_U, api.F_V, api.F_P), developed_state(d)):
        ctx.upload(w, f)
    ctx.coldstart()
    t_setup = time.perf_counter() - t_setup
    if args.atd:
        ctx.smallscale_init()                # src/main.f:643-665
    if args.particles:
        set_particles(ctx, d, args.particles)

    def barrier():
        ctx.sync()
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    # ---- device-resident timing --------------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()          # samples cover warm-up + timed steps (all under the same load)
        time.sleep(0.3)
    logs, tm, wall_s, launches = timed_steps(ctx, args.steps, args.warmup, barrier)
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = tm["total_ms"]
    dev_ms = slab.max_over_ranks(dev_ms, dist, "cuda" if dist is not None else None)
    wall_s = slab.max_over_ranks(wall_s, dist, "cuda" if dist is not None else None)
    value = cells * args.steps / (dev_ms * 1e-3) / 1e9
    sor_iters = tm["sor_iters"]
    fused_T = int(os.environ.get("W2_SOR_T", "2"))
    iters_per_launch = fused_T if fused_T > 0 else 0.5          # T=0: one launch per colour half-sweep
    sor_launch_ms = tm["sor_ms"] / max(sor_iters / iters_per_launch, 1)
    q_mean, s_mean = iterations_of(logs, d)

    # ---- end-to-end through the host-buffer ABI --------------------------------------------
    hu, pu = api.pinned_field(d)
    hv, pv = api.pinned_field(d)
    hp, pp = api.pinned_field(d)
    ctx.download(api.F_U, hu); ctx.download(api.F_V, hv); ctx.download(api.F_P, hp)
    ctx.step_host(hu, hv, hp, 1)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ctx.step_host(hu, hv, hp, 1)
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_s = slab.max_over_ranks(e2e_s, dist, "cuda" if dist is not None else None)
    e2e = cells * args.steps / e2e_s / 1e9
    part_info = None
    if args.particles:
        gxp, gyp, gup, gvp, gout = ctx.particles()
        part_info = {"particles": int(gxp.size), "in_bounds": int((gout == 0).sum()),
                     "finite": bool(np.isfinite(gxp).all() and np.isfinite(gup).all()),
                     "mean_speed": float(np.hypot(gup, gvp)[gout == 0].mean()) if (gout == 0).any() else 0.0}
    for nm, f in (("u", hu), ("v", hv), ("p", hp)):   # the timed run must have produced a sane flow
        if not np.isfinite(f).all() or np.abs(f).max() > 1.0e3:
            raise SystemExit(f"bench: field {nm} is not finite/bounded after the run (max {np.abs(f).max()})")
    rows_held = (d.slab[5] - d.slab[4] + 1) if d.slab else d.ny + 2
    copy_bytes = 3 * (d.nx + 2) * rows_held * 8 * world     # all ranks (their slabs are equal to within a row)
    for q in (pu, pv, pp):
        api.pinned_free(q)
    del hu, hv, hp
    verify = None
    if world > 1 and not args.no_verify:
        verify = verify_against_one_gpu(ctx, d, args, dist, rank, world)
    ctx.close()
    if world > 1:
        api.comm_finalize()
        dist.barrier()
        dist.destroy_process_group()

    if rank != 0:
        return
    peak, peak_src = measured_peak()
    # dominant kernel: the red/black SOR pass; one R/B iteration = 64 algorithmic B/cell,
    # a launch covers iters_per_launch iterations
    alg_launch = cells_local * 64.0 * iters_per_launch       # per GPU
    achieved = alg_launch / (sor_launch_ms * 1e-3) / 1e9 if sor_iters else 0.0
    kname = (f"sor_rb_fused_kernel<{fused_T}> ({fused_T} red+black iteration(s) per launch)" if fused_T > 0
             else "sor_rb_sweep (one colour half-sweep per launch)")
    tr = ncu_traffic()
    traffic = ((tr or {}).get("dram_bytes_per_launch")
               if (tr and tr.get("cells") == cells and tr.get("iterations_per_launch") == iters_per_launch) else None)
    step_bytes = algorithmic_bytes_per_step(cells, q_mean, s_mean)
    cfg = config_dict(args, world)
    cfg["setup_s"] = round(t_setup, 1)
    line = {
        "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "wall_ms_per_step": wall_s / args.steps * 1e3,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (uniform grid, analytic one-vortex restart field)",
        "config": cfg,
        "steps_per_s": args.steps / (dev_ms * 1e-3),
        "iterations": {"ql_per_step": q_mean, "sor_per_step": s_mean},
        "step_roofline": {"algorithmic_GB_per_step": step_bytes / 1e9,
                          "achieved_GBs": step_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / world,
                          "frac_of_measured_peak": step_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / peak / world,
                          "note": "per GPU: algorithmic bytes of the whole step / step time / N"},
        "sections_ms_per_step": {"momentum": tm["momentum_ms"] / args.steps, "ppe": tm["ppe_ms"] / args.steps,
                                 "other": tm["other_ms"] / args.steps},
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_launch, "avg_launch_ms": sor_launch_ms,
                     "traffic": traffic,
                     "traffic_GBs": traffic / (sor_launch_ms * 1e-3) / 1e9 if traffic else None,
                     "traffic_frac": traffic / (sor_launch_ms * 1e-3) / 1e9 / peak if traffic else None,
                     "note": "achieved counts ALGORITHMIC bytes (64 B/cell/iteration, SURVEY 8d); the temporally blocked "
                             "kernel moves fewer (traffic, ncu dram bytes per launch), so frac can exceed 1; traffic_frac "
                             "is the physical DRAM fraction"
                             + ("; per GPU, and at N>1 avg_launch_ms includes the peer halo stores and the cross-GPU "
                                "barrier that follow every pass" if world > 1 else "")},
        "e2e": {"value": e2e, "unit": "Gcell-updates/s", "h2d_bytes_per_step": copy_bytes,
                "d2h_bytes_per_step": copy_bytes, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(launches),
        "host_syncs_per_step": tm["host_syncs"] / args.steps,
        "clocks": clocks,
    }
    if verify is not None:
        line["verify"] = verify
    if args.atd or args.particles:   # BASELINE.json configs[4]; a parity-test configuration, not the headline line
        line["config"]["optional_paths"] = {
            "small_scale": bool(args.atd), "trajectories": part_info,
            "note": "SmallScale (own Ppe of the same fixed work) and Traject run inside every step on the device and "
                    "inside the timed region; step_roofline counts the large-scale step's algorithmic bytes only"}
    headline = (world == 1 and args.workload == "cavity" and n == 4096 and args.fixed_work and not (args.atd or args.particles))
    if headline and not args.no_extras:
        ex = []
        fw = (True, args.q_iters, args.s_iters)
        ex.append(run_extra("configs[1]: cavity Re=1000 1024x1024, fixed work", make_deck("cavity", 1024, *fw), 100, 10))
        ex.append(run_extra("configs[1]: cavity Re=1000 1024x1024, converged (ql 1e-4, sor 1e-8, omega 2/(1+sin(pi h)), cap 20000)",
                            make_deck("cavity", 1024, False, 0, 0), 10, 3))
        ex.append(run_extra("configs[2]: channel 4096x4096, inlet W / fully_dev outlet E, fixed work",
                            make_deck("channel", 4096, *fw), 5, 3))
        ex.append(run_extra("configs[2]: channel 4096x4096, inlet W / mass_cons outlet E, fixed work",
                            make_deck("channel", 4096, *fw, outlet="mass_cons"), 5, 3))
        ex.append(run_extra("configs[2]: backward step 4096x4096 (2x2 regions, blockage), fully_dev outlets, fixed work",
                            make_deck("bstep", 4096, *fw), 5, 3))
        ex.append(run_extra("configs[2]: channel 4096x4096, fully_dev outlet, converged (ql 1e-4, sor 1e-8, omega 2/(1+sin(pi h)), cap 40000)",
                            make_deck("channel", 4096, False, 0, 0, msorit=40000), 2, 1))
        ex.append(run_extra("configs[4]: cavity 4096x4096 + ATD small-scale model + 10^6 particles, fixed work",
                            make_deck("cavity", 4096, *fw, atd=True), 5, 3, particles=1000000, atd=True))
        line["extra"] = ex
    if world == 1 and not args.no_cpu and not (args.atd or args.particles):
        try:
            ncpu, _ = cpu_sample_size(args, budget_s=20.0, nsteps=1, nmax=n)
            dc = make_deck(args.workload, ncpu, args.fixed_work, args.q_iters, args.s_iters)
            t, kind = cpu_steps(dc, 1)
            line["cpu_baseline"] = {"value": dc.cells() / t / 1e9, "unit": "Gcell-updates/s", "cores": 1, "kind": "port",
                                    "sample": f"1 step of {args.workload} {ncpu}x{ncpu} in the same mode (about {t:.0f} s of CPU)",
                                    "build": kind}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "Gcell-updates/s", "cores": 1, "kind": "port",
                                    "sample": f"failed: {e}"}
    print(json.dumps(line))
    if verify is not None and verify.get("identical") is False:
        sys.exit(3)      # a slab run that differs from one GPU is a bug, not a number


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cavity", choices=["channel", "cavity", "bstep"])
    ap.add_argument("--outlet", default="fully_dev", choices=["fully_dev", "mass_cons"])
    # (--grid-n: under `python -m torch.distributed.run` a bare --n is claimed by the launcher's own abbreviations)
    ap.add_argument("--n", "--grid-n", dest="n", type=int, default=None,
                    help="grid points per side (default: 4096 on 1 GPU, 16384 on N > 1)")
    ap.add_argument("--mode", default="fixed", choices=["fixed", "converged"])
    ap.add_argument("--q-iters", type=int, default=2)
    ap.add_argument("--s-iters", type=int, default=100)
    ap.add_argument("--max-sor", type=int, default=20000, help="max_sor_iter of the converged mode")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the configs[1]/[2]/[4] entries of the default 1-GPU run")
    ap.add_argument("--no-verify", action="store_true", help="N > 1: skip the comparison with a one-GPU run")
    ap.add_argument("--ref-budget", type=float, default=600.0,
                    help="--impl reference: run the GPU arm's own grid if K+W steps fit this many seconds")
    ap.add_argument("--atd", action="store_true", help="ATD small-scale model on (BASELINE.json configs[4])")
    ap.add_argument("--particles", type=int, default=0, help="Lagrangian particles (configs[4]: 1000000)")
    ap.add_argument("--scaling", default=None, choices=["weak", "strong"],
                    help="N>1: strong (default) = the n x n grid cut into N slabs; weak = n x (n*N) grid, n rows per GPU")
    ap.add_argument("--opt", action="append", default=[], metavar="NAME=VALUE",
                    help="library option (wolfd2_b200_set_option), e.g. mom_np_cache=0, sor_fused_T=1")
    args = ap.parse_args()
    args.fixed_work = args.mode == "fixed"
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
