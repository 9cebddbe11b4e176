#!/usr/bin/env python
"""bench.py -- throughput of the wolfd2 time-step hot path on B200 (contract in the task brief).

A "step" is one pass of the step body src/main.f:690-981 (QL momentum loop, PPE red/black SOR,
projection, ghost fills, norms) over one synthetic grid.  Default workload: BASELINE.json
configs[2], channel inflow/outflow on a 4096 x 4096 uniform grid, in FIXED-WORK mode
(ql_tolerance 0, max_ql_iter 2, sor_tolerance 0, max_sor_iter 100 -- SURVEY.md §8d) so that the
algorithmic bytes per step are known a priori:  cells*(256*Q + 56 + 64*S + 88 + 16) B.

  value : Gcell-updates/s = (nx-1)(ny-1)*K / device time, fields resident in HBM.
  e2e   : same metric through wolfd2_b200_step_host with pinned HOST u,v,p buffers, i.e. H2D of
          the state before and D2H after every step inside the timed region.
  --impl reference : the reference's CPU implementation.  The Fortran reference cannot be built in
          this image (no Fortran front-end), so this times the C restatement in oracle/
          (cpu_baseline.kind = "port"), 1 thread because the reference is serial.
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


def global_ny(args, world):
    """Weak scaling: every GPU keeps an n-row slab of an n x (n*world) grid; strong: the grid stays n x n."""
    return args.n * world if args.scaling == "weak" else args.n


def make_deck(workload, n, fixed_work, q_iters, s_iters, ny=None, slab=None, atd=False):
    from wolfd2_b200 import deck as dk
    ny = ny or n
    kw = {"ny": ny}
    if atd:   # BASELINE.json configs[4]: ATD small-scale model on; its own Ppe does the same fixed work
        kw.update(smallscale=True, ss_cu0=0.02, ss_bncrit=2.0, ss_ppe_solver="rb_sor",
                  ss_msorit=s_iters if fixed_work else 2000, ss_sortol=0.0 if fixed_work else 1e-8)
    if slab is not None:
        kw["slab"] = slab          # (rank, world): this rank's rows only
    nmax = max(n, ny)
    if workload == "cavity":
        d = dk.cavity(n, re=1000.0, dt=stable_dt(nmax, 1000.0), **kw)
    elif workload == "channel":
        d = dk.channel(n, re=100.0, dt=stable_dt(nmax, 100.0), fully_dev=True, **kw)
    elif workload == "bstep":
        d = dk.backward_step(n, re=100.0, dt=stable_dt(nmax, 100.0), fully_dev=True, **kw)
    else:
        raise SystemExit(f"unknown workload {workload}")
    d.ppe_solver = "rb_sor"
    if fixed_work:
        d.qtol, d.mqiter, d.sortol, d.msorit = 0.0, q_iters, 0.0, s_iters
    else:
        d.sorrel = 1.9
    return d


def developed_state(d):
    """A smooth, non-quiescent restart field (one vortex filling the box) on the staggered locations of
    src/grid.f:337-359, so that the timed steps do not run on a mostly-zero cold start.  Returns u, v, p."""
    nx, ny = d.nx, d.ny
    a0, a1 = (d.slab[4], d.slab[5]) if d.slab else (0, ny + 1)     # rows this rank holds
    u, v, p = d.new_field(), d.new_field(), d.new_field()
    xi = (np.arange(nx + 2) - 1.0) / (nx - 1.0)
    yj = (np.arange(a0, a1 + 1) - 1.0) / (ny - 1.0)
    xh, yh = xi - 0.5 / (nx - 1.0), yj - 0.5 / (ny - 1.0)
    A = 0.2
    nr = a1 - a0 + 1
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


def cpu_sample_size(args, budget_s, nsteps):
    """Grid size n such that nsteps of the CPU port take about budget_s."""
    d = make_deck(args.workload, 256, args.fixed_work, args.q_iters, args.s_iters)
    t, _ = cpu_steps(d, 1)
    per_cell = t / d.cells()
    n = int((budget_s / max(nsteps, 1) / per_cell) ** 0.5)
    n = max(128, min(args.n, n))
    return (n // 64) * 64 if n >= 256 else n, per_cell


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    try:
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})
    except Exception:
        pass
    n, _ = cpu_sample_size(args, budget_s=120.0, nsteps=args.steps + args.warmup)
    d = make_deck(args.workload, n, args.fixed_work, args.q_iters, args.s_iters)
    t, kind = cpu_steps(d, args.steps, warm=args.warmup)
    value = d.cells() * args.steps / t / 1e9
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    nyg = global_ny(args, world)
    sample = (f"{args.workload} {n}x{n} (same deck family as the GPU arm's {args.n}x{nyg}; size bounded so the run "
              f"ends in minutes), {args.steps} steps after {args.warmup} warm-up, 1 thread (the reference is serial)")
    line = {
        "impl": "reference", "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t / args.steps * 1e3,
        "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, make_deck(args.workload, 256, args.fixed_work, args.q_iters, args.s_iters), world),
        "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": 1, "kind": "port", "sample": sample,
                         "build": kind},
        "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def config_dict(args, d, world):
    nyg = global_ny(args, world)
    cells = (args.n - 1) * (nyg - 1)
    re = d.re
    dt = stable_dt(max(args.n, nyg), re)
    return {"workload": f"{args.workload} {args.n}x{nyg} uniform grid, Re={re:g}, dt={dt:g}, ppe_solver rb_sor "
                        f"(BASELINE.json metric grid 4096^2; deck family of configs[{ {'channel': 2, 'cavity': 1, 'bstep': 2}[args.workload] }])",
            "mode": ("fixed-work: ql_tolerance 0, max_ql_iter %d, sor_tolerance 0, max_sor_iter %d" % (args.q_iters, args.s_iters))
            if args.fixed_work else "converged: ql_tolerance 1e-4, sor_tolerance 1e-8, max_sor_iter 2000",
            "grid": [args.n, nyg], "cells": cells,
            "l2": "working set (>= 40 arrays x %.0f MB per GPU) exceeds the 126 MB L2; no flush needed" % (cells / world * 8 / 1e6)
            if cells / world * 8 * 4 > 126e6 else "working set fits L2 (latency-bound regime)",
            "parallelism": "1 GPU" if world == 1 else
            f"{world} row slabs of ~{(nyg - 1) // world} rows, one process per GPU; NCCL halo exchange of us,vs per QL "
            f"iteration and of p per fused SOR pass, all-reduced max-norms and tridiagonal segment records "
            f"(wolfd2_b200/csrc/w2_dist.cu); results bit-identical to one GPU"}


# --------------------------------------------------------------------------- GPU arm
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
    from wolfd2_b200 import slab
    nyg = global_ny(args, world)
    if world > 1:
        slab.init_comm(dist, local)     # the library's own NCCL communicator; torch only carries the id
        d = make_deck(args.workload, args.n, args.fixed_work, args.q_iters, args.s_iters, ny=nyg, slab=(rank, world))
    else:
        d = make_deck(args.workload, args.n, args.fixed_work, args.q_iters, args.s_iters, atd=args.atd)
    if world > 1 and (args.atd or args.particles):
        raise SystemExit("bench: --atd / --particles run on one GPU (SURVEY 8f N2/N3 are single-GPU rows)")
    cells = (d.nx - 1) * (d.ny - 1)          # pressure unknowns of the whole grid
    cells_local = (d.nx - 1) * (d.slab[3] - d.slab[2] + 1) if d.slab else cells
    ctx = api.Context(d)
    for w, f in zip((api.F_U, api.F_V, api.F_P), developed_state(d)):
        ctx.upload(w, f)
    ctx.coldstart()
    if args.atd:
        ctx.smallscale_init()                # src/main.f:643-665
    if args.particles:                       # configs[4]: particles on a uniform lattice, FwdEuler, Stokes drag
        from wolfd2_b200 import _abi
        side = max(1, int(round(args.particles ** 0.5)))
        gx, gy = d.node_arrays()
        lat = (np.arange(side) + 0.5) / side * 0.8 + 0.1
        xp, yp = [a.ravel().copy() for a in np.meshgrid(lat, lat)]
        npart = xp.size
        tr = _abi.Traject()
        tr.ntr, tr.ntsubstp, tr.nTrMethod, tr.nTrCdEq, tr.mTrHTmit = npart, 1, 2, 1, 1
        tr.densref, tr.dTrHTtol, tr.dTrHTdel = 1.2, 1e-8, 1.0
        ctx.set_trajectories(tr, gx, gy, np.full(npart, 2.0), np.full(npart, 2.0), np.full(npart, 10.0),
                             xp, yp, np.zeros(npart), np.zeros(npart))

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
    ctx.step(args.warmup)
    l0 = ctx.timing()["launches"]
    barrier()
    logs = ctx.step(args.steps)
    barrier()
    tm = ctx.timing()
    clocks = sampler.stop() if rank == 0 else None
    dev_ms = tm["total_ms"]
    launches = sum(tm["launches"].values()) - sum(l0.values())
    dev_ms = slab.max_over_ranks(dev_ms, dist, "cuda" if dist is not None else None)
    value = cells * args.steps / (dev_ms * 1e-3) / 1e9
    sor_iters = tm["sor_iters"]
    fused_T = int(os.environ.get("W2_SOR_T", "2"))
    iters_per_launch = fused_T if fused_T > 0 else 0.5          # T=0: one launch per colour half-sweep
    sor_launch_ms = tm["sor_ms"] / max(sor_iters / iters_per_launch, 1)
    q_done = [abs(l["nQLiter"]) if l["nQLiter"] > 0 else d.mqiter for l in logs]
    s_done = [l["nSorConv"] for l in logs]

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
    step_bytes = algorithmic_bytes_per_step(cells, float(np.mean(q_done)), float(np.mean(s_done)))
    line = {
        "metric": "Gcell-updates/s", "value": value, "unit": "Gcell-updates/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
        "higher_is_better": True, "scaling": args.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic (uniform grid, analytic one-vortex restart field)",
        "config": config_dict(args, d, world),
        "steps_per_s": args.steps / (dev_ms * 1e-3),
        "iterations": {"ql_per_step": float(np.mean(q_done)), "sor_per_step": float(np.mean(s_done))},
        "step_roofline": {"algorithmic_GB_per_step": step_bytes / 1e9,
                          "achieved_GBs": step_bytes / (dev_ms / args.steps * 1e-3) / 1e9,
                          "frac_of_measured_peak": step_bytes / (dev_ms / args.steps * 1e-3) / 1e9 / peak},
        "sections_ms_per_step": {"momentum": tm["momentum_ms"] / args.steps, "ppe": tm["ppe_ms"] / args.steps,
                                 "other": tm["other_ms"] / args.steps},
        "roofline": {"bound": "hbm", "kernel": kname, "achieved": achieved,
                     "peak": peak, "unit": "GB/s", "frac": achieved / peak, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_launch, "avg_launch_ms": sor_launch_ms,
                     "traffic": (tr or {}).get("dram_bytes_per_launch")
                     if (tr and tr.get("cells") == cells and tr.get("iterations_per_launch") == iters_per_launch) else None,
                     "note": "achieved counts ALGORITHMIC bytes; the fused kernel moves fewer (see traffic), "
                             "so frac can exceed what a copy kernel reaches"
                             + ("; per GPU, and at N>1 avg_launch_ms includes the NCCL all-reduce and halo exchange "
                                "that follow every pass" if world > 1 else "")},
        "e2e": {"value": e2e, "unit": "Gcell-updates/s", "h2d_bytes_per_step": copy_bytes,
                "d2h_bytes_per_step": copy_bytes, "ms_per_step": e2e_s / args.steps * 1e3},
        "gpu_launches": int(launches),
        "clocks": clocks,
    }
    if args.atd or args.particles:   # BASELINE.json configs[4]; a parity-test configuration, not the headline line
        line["config"]["optional_paths"] = {
            "small_scale": bool(args.atd), "trajectories": part_info,
            "note": "SmallScale (own Ppe of the same fixed work) and/or Traject run inside every step on the device; "
                    "step_roofline counts the large-scale step's algorithmic bytes only"}
    if world == 1 and not args.no_cpu and not (args.atd or args.particles):
        try:
            n, _ = cpu_sample_size(args, budget_s=20.0, nsteps=1)
            dc = make_deck(args.workload, n, args.fixed_work, args.q_iters, args.s_iters)
            t, kind = cpu_steps(dc, 1)
            line["cpu_baseline"] = {"value": dc.cells() / t / 1e9, "unit": "Gcell-updates/s", "cores": 1, "kind": "port",
                                    "sample": f"1 step of {args.workload} {n}x{n} in the same mode (about {t:.0f} s of CPU)",
                                    "build": kind}
        except Exception as e:  # the baseline is reported, never required
            line["cpu_baseline"] = {"value": None, "unit": "Gcell-updates/s", "cores": 1, "kind": "port",
                                    "sample": f"failed: {e}"}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cavity", choices=["channel", "cavity", "bstep"])
    ap.add_argument("--n", type=int, default=4096)
    ap.add_argument("--mode", default="fixed", choices=["fixed", "converged"])
    ap.add_argument("--q-iters", type=int, default=2)
    ap.add_argument("--s-iters", type=int, default=100)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--atd", action="store_true", help="ATD small-scale model on (BASELINE.json configs[4])")
    ap.add_argument("--particles", type=int, default=0, help="Lagrangian particles (configs[4]: 1000000)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = n x (n*N) grid, n rows per GPU; strong = the n x n grid cut into N slabs")
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
