"""Host-side helpers for running the hot path on several GPUs of one node (one process per GPU).

Row-slab decomposition in j: no tridiagonal line crosses a slab because both momentum split
steps run along i.  The device side is csrc/w2_dist.cu; this module mirrors its partition
arithmetic (checked against the library in tests) and carries the launcher-side helpers."""
from __future__ import annotations


def slab_rows(ny: int, world: int, rank: int):
    """Unknown pressure rows j = 2..ny split into `world` contiguous slabs; returns (j0, j1) inclusive.
    Earlier ranks take the remainder rows, every slab has at least 4 rows."""
    nrows = ny - 1
    if world < 1 or rank < 0 or rank >= world:
        raise ValueError("bad rank/world")
    if nrows < 4 * world:
        raise ValueError(f"{nrows} rows cannot be split into {world} slabs of >= 4 rows")
    base, rem = divmod(nrows, world)
    j0 = 2 + rank * base + min(rank, rem)
    j1 = j0 + base + (1 if rank < rem else 0) - 1
    return j0, j1


def halo_depth(nx: int, world: int) -> int:
    """Halo rows on each side of a slab (w2_slab_layout in csrc/w2_context.cu): 2T = 4 rows for the fused
    SOR pass, the rows a straddling 2048-unknown momentum segment reaches into, plus the stencil row."""
    if world == 1:
        return 0
    return max(5, 4 + 2047 // (nx - 1))


def slab_layout(nx: int, ny: int, world: int, rank: int):
    """(J0, J1, A0, A1, HG): owned unknown rows, rows held in memory, halo depth."""
    j0, j1 = slab_rows(ny, world, rank) if world > 1 else (2, ny)
    hg = halo_depth(nx, world)
    if world > 1 and (ny - 1) // world < 2 * hg:
        raise ValueError(f"{ny - 1} rows cannot be split into {world} slabs of >= {2 * hg} rows")
    a0 = 0 if rank == 0 else max(0, j0 - hg)
    a1 = ny + 1 if rank == world - 1 else min(ny + 1, j1 + hg)
    return j0, j1, a0, a1, hg


def init_comm(dist, device_index: int):
    """Create the library's NCCL communicator over the ranks of an initialised torch.distributed group:
    rank 0 makes the id, the group broadcasts it (plumbing only), every rank joins."""
    import ctypes as C
    import torch
    from . import api
    L = api.lib()
    rank, world = dist.get_rank(), dist.get_world_size()
    api.set_device(device_index)
    buf = (C.c_ubyte * 128)()
    if rank == 0 and world > 1:
        api._check(L.wolfd2_b200_comm_unique_id(buf), "comm_unique_id")
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if world > 1:
        if dist.get_backend() == "nccl":
            t = t.cuda(device_index)
        dist.broadcast(t, 0)
    ident = (C.c_ubyte * 128)(*[int(x) for x in t.cpu().tolist()])
    api._check(L.wolfd2_b200_comm_init(rank, world, ident), "comm_init")
    return rank, world


def halo_rows(j0: int, j1: int, ny: int, depth: int):
    """Rows a slab must receive from its south / north neighbour for a fused SOR pass of `depth` = 2T
    half-sweeps (None at a physical boundary)."""
    south = None if j0 == 2 else (j0 - depth, j0 - 1)
    north = None if j1 == ny else (j1 + 1, j1 + depth)
    return south, north


def max_over_ranks(value: float, dist=None, device=None) -> float:
    """Device-time reduction used by bench.py: the slowest rank defines the step time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(value)
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device or "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
