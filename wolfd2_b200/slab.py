"""Host-side helpers for running the hot path on several GPUs of one node (one process per GPU).

Row-slab decomposition in j: no tridiagonal line crosses a slab because both momentum split
steps run along i.  Round 1 ships the partition arithmetic and the timing reduction; the
device-side halo exchange is not implemented yet, so bench.py runs independent replicas at N > 1."""
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
