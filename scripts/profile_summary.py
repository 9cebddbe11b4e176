"""Summaries committed under profiles/: (1) per-kernel totals of an ncu launch list
(`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python bench.py ...`),
(2) selected raw metrics of `ncu --set full` reports.  Usage:
    python scripts/profile_summary.py launches gpurun_out/launches.csv [skip_first_n]
    python scripts/profile_summary.py raw gpurun_out/x.ncu-rep [...]"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict

METRICS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
           "launch__registers_per_thread", "launch__waves_per_multiprocessor",
           "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
           "smsp__issue_active.avg.pct_of_peak_sustained_active",
           "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
           "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
           "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
           "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]


def short(name):
    name = re.sub(r"^void ", "", name)
    name = re.sub(r"\(.*$", "", name)
    return name.replace("(int)", "").replace("(bool)", "")


def launches(path, skip=0):
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ik, iv, im = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name")
    agg = OrderedDict()
    seen = 0
    for r in rows[1:]:
        if r[im] != "gpu__time_duration.sum":
            continue
        seen += 1
        if seen <= skip:
            continue
        a = agg.setdefault(short(r[ik]), [0, 0.0])
        a[0] += 1
        a[1] += float(r[iv]) / 1e3
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':46s}{'n':>6s}{'avg_us':>10s}{'share%':>8s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:46s}{n:6d}{t / n:10.1f}{100 * t / tot:8.1f}")
    grp = {"SOR (fused pass, pack, reset)": ("sor", "sorf"), "momentum (assemble+reduce, upper levels, finalize, QL update)":
           ("mom_", "tri_", "ql_"), "other (div/rhs, project, ghost fills, norms, masks)": ()}
    used = set()
    for g, keys in grp.items():
        if keys:
            ks = [k for k in agg if k.startswith(keys)]
            used.update(ks)
        else:
            ks = [k for k in agg if k not in used]
        print(f"# {g}: {100 * sum(agg[k][1] for k in ks) / tot:.1f} %")
    print(f"# total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")


def raw(paths):
    for p in paths:
        out = subprocess.run(["ncu", "-i", p, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(out.splitlines()))
        h = rows[0]
        print(f"## report: {p.split('/')[-1]}")
        for m in METRICS:
            if m in h:
                i = h.index(m)
                print(f"{m:90s} {[short(r[i]) if m == 'Kernel Name' else r[i] for r in rows[1:]]}")
        print()


def dram(path, peak=6451.2, traffic_json=None, cells=16769025):
    """Per-kernel time and DRAM bytes of a launch list taken with
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum; optionally writes the dominant SOR
    kernel's bytes per launch to profiles/sor_traffic.json (read by bench.py)."""
    rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
    h = rows[0]
    ik, iv, im, ii = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Name"), h.index("ID")
    per = OrderedDict()
    for r in rows[1:]:
        per.setdefault((r[ii], short(r[ik])), {})[r[im]] = float(r[iv].replace(",", ""))
    agg = OrderedDict()
    for (_, k), m in per.items():
        a = agg.setdefault(k, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += m.get("gpu__time_duration.sum", 0.0) / 1e3
        a[2] += m.get("dram__bytes_read.sum", 0.0)
        a[3] += m.get("dram__bytes_write.sum", 0.0)
    tot = sum(v[1] for v in agg.values())
    print(f"{'kernel':40s}{'n':>5s}{'avg_us':>9s}{'share%':>8s}{'rd_MB':>9s}{'wr_MB':>9s}{'GB/s':>8s}{'%peak':>7s}")
    for k, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        gbs = (rd + wr) / (t * 1e-6) / 1e9 if t > 0 else 0.0
        print(f"{k:40s}{n:5d}{t / n:9.1f}{100 * t / tot:8.1f}{rd / n / 1e6:9.1f}{wr / n / 1e6:9.1f}{gbs:8.0f}{100 * gbs / peak:7.1f}")
    groups = (("SOR (fused pass, tiles, pack/unpack, reset)", ("sor", "sorf")),
              ("momentum (assemble+reduce, upper levels, finalize, QL update/decide)", ("mom_", "tri_", "ql_")))
    used = set()
    for g, keys in groups:
        ks = [k for k in agg if k.startswith(keys)]
        used.update(ks)
        print(f"# {g}: {100 * sum(agg[k][1] for k in ks) / tot:.1f} %")
    print(f"# other: {100 * sum(v[1] for k, v in agg.items() if k not in used) / tot:.1f} %")
    print(f"# total {tot / 1e3:.2f} ms over {sum(v[0] for v in agg.values())} launches")
    if traffic_json:
        import json
        k = next((k for k in agg if k.startswith("sor_rb_fused_kernel<2>")), None)
        if k:
            n, t, rd, wr = agg[k]
            json.dump({"kernel": k, "cells": cells, "iterations_per_launch": 2, "dram_bytes_per_launch": (rd + wr) / n,
                       "dram_bytes_read": rd / n, "dram_bytes_write": wr / n, "launches_averaged": n,
                       "source": "ncu launch list of bench.py (cavity 4096^2), mean over the launches: " + path.split("/")[-1]},
                      open(traffic_json, "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "dram":
        dram(sys.argv[2], traffic_json=sys.argv[3] if len(sys.argv) > 3 else None)
    elif sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]) if len(sys.argv) > 3 else 0)
    else:
        raw(sys.argv[2:])
