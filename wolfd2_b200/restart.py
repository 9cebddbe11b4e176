"""Restart files in the reference's own format, so a gfortran build of wolfd2 can resume from a state
computed here and vice versa (SURVEY §5 "Checkpoint / resume", §8f N4).

SaveRestart (src/file_manip.f:364-401) writes three Fortran unformatted sequential records:
    write(12) k, dtime                      ! INTEGER*4, REAL*8
    write(12) nx, ny                        ! 2 x INTEGER*4
    write(12) p, u, v, t, pss, uss, vss, tss   each ((f(i,j), i=0,nx+1), j=0,ny+1), REAL*8
gfortran frames every record with 4-byte little-endian length markers; records longer than
2147483639 bytes are split into subrecords whose leading marker is negative when another subrecord
follows and whose trailing marker is negative when one precedes (gfortran's default record-marker
convention; the multi-subrecord path could not be checked against a real gfortran here).
Host-side I/O only -- nothing in this module touches the device."""
from __future__ import annotations

import struct

import numpy as np

_MAXREC = 2147483639


def _write_record(f, payload: bytes):
    n = len(payload)
    if n <= _MAXREC:
        f.write(struct.pack("<i", n)); f.write(payload); f.write(struct.pack("<i", n))
        return
    pos, first = 0, True
    while pos < n:
        m = min(_MAXREC, n - pos)
        last = pos + m >= n
        f.write(struct.pack("<i", m if last else -m))
        f.write(payload[pos:pos + m])
        f.write(struct.pack("<i", m if first else -m))
        pos += m
        first = False


def _read_record(f) -> bytes:
    out = []
    while True:
        (head,) = struct.unpack("<i", f.read(4))
        m = abs(head)
        out.append(f.read(m))
        (tail,) = struct.unpack("<i", f.read(4))
        if abs(tail) != m:
            raise ValueError("corrupt record markers")
        if head >= 0:
            return b"".join(out)


def _window(a, nx, ny):
    return np.ascontiguousarray(a[0:ny + 2, 0:nx + 2], dtype="<f8").tobytes()


def save_restart(path, deck, k, dtime, u, v, p, t=None, uss=None, vss=None, pss=None, tss=None):
    """SaveRestart, src/file_manip.f:387-398.  Fields are (0:mnx,0:mny) arrays; missing ones are zero
    (t and the small-scale fields are identically zero on the cold-flow path)."""
    nx, ny = deck.nx, deck.ny
    z = np.zeros_like(u)
    order = [p, u, v, t, pss, uss, vss, tss]
    with open(path, "wb") as f:
        _write_record(f, struct.pack("<id", int(k), float(dtime)))
        _write_record(f, struct.pack("<ii", nx, ny))
        _write_record(f, b"".join(_window(z if a is None else a, nx, ny) for a in order))


def read_restart(path, deck):
    """ReadRestart, src/file_manip.f:442-460 (including its size check).  Returns
    (k, dtime, dict(p,u,v,t,pss,uss,vss,tss)) with fields in the deck's (0:mnx,0:mny) layout."""
    nx, ny = deck.nx, deck.ny
    with open(path, "rb") as f:
        k, dtime = struct.unpack("<id", _read_record(f))
        nxf, nyf = struct.unpack("<ii", _read_record(f))
        if (nxf, nyf) != (nx, ny):
            raise ValueError(f"Error: Index mismatch in restart file: {nx} {ny} {nxf} {nyf}")
        raw = np.frombuffer(_read_record(f), dtype="<f8")
    n = (nx + 2) * (ny + 2)
    if raw.size != 8 * n:
        raise ValueError("restart record has the wrong length")
    out = {}
    for q, name in enumerate(("p", "u", "v", "t", "pss", "uss", "vss", "tss")):
        a = deck.new_field()
        a[0:ny + 2, 0:nx + 2] = raw[q * n:(q + 1) * n].reshape(ny + 2, nx + 2)
        out[name] = a
    return k, dtime, out
