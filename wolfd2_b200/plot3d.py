"""PLOT3D output in the reference's own layout: the `.qqq` function file, the `.xyz` grid
file and the `.nam` name file that `SaveStdVarsP3D` / `SaveGrid2DP3D` write (src/file_manip.f:1024-1173, :319-357),
so that a post-processor set up for wolfd2 reads a run computed here.

The arrays passed in are the NODE-AVERAGED fields main.f hands to SaveStdVarsP3D (src/main.f:1090-1095, :1337-1342):
`VelAvg` / `PTDAvg` of the staggered fields (src/utility.f:513-647) -- on the device these are the `velavg_` /
`ptdavg_` entry points or `Context.node_averages`; this module only does the host-side file I/O.

    unformatted (`output_format plot3d unformatted`):
        write(15) nx, ny, nVars                                  ! 3 x INTEGER*4
        write(15) (((f(i,j,l), i=1,nx), j=1,ny), l=1,nVars)      ! REAL*8 (FLOAT = real*8, include/wolfd2.h:6)
    grid, unformatted:
        write(22) nx, ny
        write(22) ((sngl(x(i,j)),i=1,nx),j=1,ny), ((sngl(y(i,j)),i=1,nx),j=1,ny)     ! REAL*4
    Variable order: p, u, v, w(=0) [, t] [, ps, us, vs, ws(=0) [, ts]].

Unformatted files are byte-exact gfortran sequential records (4-byte markers, sub-records above 2 GiB as in
restart.py).  The formatted variants are list-directed in the reference (`write(15,*)`), whose column layout is
compiler-specific; here they are written one value per token with 17 significant digits (the grid with the
reference's own `5(e14.6)` edit descriptor), which every free-format PLOT3D reader -- and a Fortran list-directed
read -- accepts."""
from __future__ import annotations

import struct

import numpy as np

from .restart import _read_record, _write_record

FT_FORMATTED, FT_UNFORMATTED = 0, 1   # the nForm flags of include/wolfd2.h:82-83


def _nodes(a, nx, ny):
    """(i = 1..nx, j = 1..ny) window of a (0:mnx,0:mny) array, i fastest."""
    return np.ascontiguousarray(a[1:ny + 1, 1:nx + 1], dtype=np.float64)


def std_vars(nx, ny, u, v, p, t=None, us=None, vs=None, ps=None, ts=None):
    """The function planes of SaveStdVarsP3D in file order (src/file_manip.f:1075-1118): thermal energy adds t,
    the small-scale model doubles the set."""
    zero = np.zeros((ny, nx))
    planes = [_nodes(p, nx, ny), _nodes(u, nx, ny), _nodes(v, nx, ny), zero]
    if t is not None:
        planes.append(_nodes(t, nx, ny))
    if us is not None:
        planes += [_nodes(ps, nx, ny), _nodes(us, nx, ny), _nodes(vs, nx, ny), zero]
        if t is not None:
            planes.append(_nodes(ts, nx, ny))
    return planes


def var_names(thermal=False, smallscale=False):
    """Lines of the `.nam` file (src/file_manip.f:1146-1162)."""
    names = ["Complete Pressure", "Complete U ; Complete Velocity", "Complete V", "Complete W"]
    if thermal:
        names.append("Complete Temperature")
    if smallscale:
        names += ["Small-Scale Pressure", "Small-Scale U ; Small-Scale Velocity", "Small-Scale V", "Small-Scale W"]
        if thermal:
            names.append("Small-Scale Temperature")
    return names


def save_grid_p3d(path, nx, ny, x, y, form=FT_UNFORMATTED):
    """SaveGrid2DP3D (src/file_manip.f:319-357)."""
    gx, gy = _nodes(x, nx, ny), _nodes(y, nx, ny)
    if form == FT_UNFORMATTED:
        with open(path, "wb") as f:
            _write_record(f, struct.pack("<ii", nx, ny))
            _write_record(f, gx.astype("<f4").tobytes() + gy.astype("<f4").tobytes())
    elif form == FT_FORMATTED:
        vals = np.concatenate([gx.ravel(), gy.ravel()])
        with open(path, "w") as f:
            f.write(f" {nx:11d} {ny:11d}\n")
            for k in range(0, vals.size, 5):          # format (5(e14.6)): 0.dddddde+xx, width 14
                f.write("".join(_e14_6(v) for v in vals[k:k + 5]) + "\n")
    else:
        raise ValueError("* Wrong nForm flag passed to SaveGrid2DP3D")


def _e14_6(v):
    """Fortran E14.6: 0.ddddddE+ee right-justified in 14 columns."""
    if v == 0.0:
        return "  0.000000E+00"
    s = f"{abs(v):.5e}"                     # d.dddddE+ee
    mant, exp = s.split("e")
    digits = mant.replace(".", "")
    e = int(exp) + 1
    txt = f"{'-' if v < 0 else ''}0.{digits}E{'+' if e >= 0 else '-'}{abs(e):02d}"
    return txt.rjust(14)


def save_std_vars_p3d(prefix, nx, ny, u, v, p, t=None, us=None, vs=None, ps=None, ts=None, form=FT_UNFORMATTED,
                      grid=None, names=True):
    """SaveStdVarsP3D (src/file_manip.f:1024-1173): `prefix.qqq`, and when asked `prefix.xyz` (grid = (x, y) node
    arrays) and `prefix.nam`.  Returns the number of variables written."""
    planes = std_vars(nx, ny, u, v, p, t, us, vs, ps, ts)
    nvars = len(planes)
    if grid is not None:
        save_grid_p3d(prefix + ".xyz", nx, ny, grid[0], grid[1], form)
    if form == FT_UNFORMATTED:
        with open(prefix + ".qqq", "wb") as f:
            _write_record(f, struct.pack("<iii", nx, ny, nvars))
            _write_record(f, b"".join(pl.astype("<f8").tobytes() for pl in planes))
    elif form == FT_FORMATTED:
        with open(prefix + ".qqq", "w") as f:
            f.write(f" {nx:11d} {ny:11d} {nvars:11d}\n")
            flat = np.concatenate([pl.ravel() for pl in planes])
            for k in range(0, flat.size, 3):
                f.write(" " + " ".join(f"{x:24.16E}" for x in flat[k:k + 3]) + "\n")
    else:
        raise ValueError("* Wrong nForm flag passed to SaveStdVarsP3D")
    if names:
        with open(prefix + ".nam", "w") as f:
            for line in var_names(t is not None, us is not None):
                f.write(" " + line + "\n")
    return nvars


TMAVG_ORDER = "ubar vbar - tbar pbar upb vpb tpb upupb vpvpb upvpb uptpb vptpb trbke dssrt dtdyb".split()
TMAVG_NAMES = ["Avgd. U-vel ; Avgd. Velocity", "Avgd. V-vel", "Avgd. W-vel (null)", "Avgd. T", "Avgd. P", "Avgd. up", "Avgd. vp",
               "Avgd. tp", "Avgd. up * up", "Avgd. vp * vp", "Avgd. up * vp", "Avgd. up * tp", "Avgd. vp * tp",
               "Avgd. Turb. Kinetic En.", "Avgd. Dissipation Rate ", "Avgd. dT*/dy*"]


def save_tmavg_p3d(prefix, nx, ny, arrays, form=FT_UNFORMATTED):
    """SaveTmAvgP3D (src/file_manip.f:899-1012): `prefix.qqq` with 16 SINGLE-precision planes (sngl(...), the third a
    null W-velocity) and `prefix.nam`.  arrays: dict name -> (0:mnx,0:mny) array (Context.timeavg_get)."""
    zero = np.zeros((ny, nx), dtype=np.float32)
    planes = [zero if k == "-" else _nodes(arrays[k], nx, ny).astype(np.float32) for k in TMAVG_ORDER]
    if form == FT_UNFORMATTED:
        with open(prefix + ".qqq", "wb") as f:
            _write_record(f, struct.pack("<iii", nx, ny, len(planes)))
            _write_record(f, b"".join(pl.astype("<f4").tobytes() for pl in planes))
    elif form == FT_FORMATTED:
        with open(prefix + ".qqq", "w") as f:
            f.write(f" {nx:11d} {ny:11d} {len(planes):11d}\n")
            flat = np.concatenate([pl.ravel() for pl in planes])
            for k in range(0, flat.size, 5):        # list-directed REAL*4: five per line in gfortran's layout
                f.write(" " + " ".join(f"{x:15.8E}" for x in flat[k:k + 5]) + "\n")
    else:
        raise ValueError("* Wrong nForm flag passed to SaveTmAvgP3D")
    with open(prefix + ".nam", "w") as f:
        for line in TMAVG_NAMES:
            f.write(" " + line + "\n")
    return len(planes)


def read_std_vars_p3d(path, form=FT_UNFORMATTED):
    """Read a `.qqq` file back: (nx, ny, [planes (ny, nx)])."""
    if form == FT_UNFORMATTED:
        with open(path, "rb") as f:
            nx, ny, nvars = struct.unpack("<iii", _read_record(f))
            data = np.frombuffer(_read_record(f), dtype="<f8")
    else:
        toks = open(path).read().split()
        nx, ny, nvars = (int(tk) for tk in toks[:3])
        data = np.array([float(tk) for tk in toks[3:]])
    if data.size != nx * ny * nvars:
        raise ValueError("size mismatch in " + path)
    return nx, ny, [data[k * nx * ny:(k + 1) * nx * ny].reshape(ny, nx) for k in range(nvars)]
