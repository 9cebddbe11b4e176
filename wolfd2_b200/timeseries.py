"""Time-series files of the reference: SaveTimeSrs (src/file_manip.f:686-854).

The device samples the monitor points inside every step (csrc/w2_probes.cu, `Context.set_probes`); this module
keeps the file side: one `<prefix>NNN.ts` per point with the reference's header (case 0, :741-800) and one
`(17(e14.6))` row per sample (case 1, :829-830): the dimensional time `dtime*dlref/uref` (main.f:985) followed
by u, v, P[, T][, u*, v*, P*[, T*]] -- the columns depend on thermal_energy / small_scale exactly as nLamV and
nVars do (:746-756)."""
from __future__ import annotations

import numpy as np

from .plot3d import _e14_6


class TimeSeriesWriter:
    def __init__(self, prefix, iTS, jTS, thermal=False, smallscale=False):
        self.prefix = prefix
        self.iTS, self.jTS = [int(i) for i in iTS], [int(j) for j in jTS]
        if len(self.iTS) != len(self.jTS) or len(self.iTS) > 999:      # character*3 TSNumber (:728)
            raise ValueError("need as many i as j indices, at most 999 points")
        self.thermal, self.smallscale = bool(thermal), bool(smallscale)
        # positions in a device record (u, v, p, t, uss, vss, pss, tss) of the columns written (:810-826)
        lam = [0, 1, 2] + ([3] if thermal else [])
        self.cols = lam + ([4 + k for k in lam] if smallscale else [])
        self.files = []

    def filenames(self):
        return [f"{self.prefix}{k + 1:03d}.ts" for k in range(len(self.iTS))]      # '(i3.3)' (:773-775)

    def open(self):
        """case 0: create the files and write their headers."""
        names = ["u", "v", "P"] + (["T"] if self.thermal else [])
        if self.smallscale:
            names += [n + "*" for n in names]
        for k, fn in enumerate(self.filenames()):
            f = open(fn, "w")
            f.write("#\n")
            f.write(f"# Time-series No. {k + 1:4d}\n")
            f.write(f"# Location: {self.iTS[k]:4d}{self.jTS[k]:4d}\n")
            f.write("#\n")
            f.write("#    Time" + "".join(f"{n:>14s}" for n in names) + "\n")      # '(a,8a14)'
            self.files.append(f)
        return self

    def write(self, rdimtime, record):
        """case 1: one row per point; record[point][8] as wolfd2_b200_get_probe_records returns it."""
        rec = np.asarray(record, dtype=np.float64).reshape(len(self.iTS), 8)
        for f, r in zip(self.files, rec):
            f.write(_e14_6(rdimtime) + "".join(_e14_6(r[c]) for c in self.cols) + "\n")      # '(17(e14.6))'

    def close(self):
        """case 2."""
        for f in self.files:
            f.close()
        self.files = []

    def __enter__(self):
        return self.open()

    def __exit__(self, *a):
        self.close()
