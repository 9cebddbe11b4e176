"""Host-side multi-process logic on CPU (gloo, world_size 2): slab partition and the max-over-ranks
timing reduction that bench.py uses at N > 1."""
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from wolfd2_b200 import slab


def test_slab_rows_partition_is_exact():
    for ny in (17, 64, 4096, 16384):
        for world in (1, 2, 3, 4, 8):
            if ny - 1 < 4 * world:
                continue
            rows = [slab.slab_rows(ny, world, r) for r in range(world)]
            assert rows[0][0] == 2 and rows[-1][1] == ny
            for a, b in zip(rows, rows[1:]):
                assert b[0] == a[1] + 1
            sizes = [j1 - j0 + 1 for j0, j1 in rows]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        slab.slab_rows(9, 4, 0)


def test_halo_rows():
    assert slab.halo_rows(2, 100, 400, 4) == (None, (101, 104))
    assert slab.halo_rows(101, 200, 400, 4) == ((97, 100), (201, 204))
    assert slab.halo_rows(301, 400, 400, 4) == ((297, 300), None)


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        t = slab.max_over_ranks(10.0 + rank, dist)
        j0, j1 = slab.slab_rows(257, world, rank)
        # neighbours agree on the rows they exchange
        south, north = slab.halo_rows(j0, j1, 257, 4)
        msg = torch.tensor([j0, j1], dtype=torch.int64)
        gathered = [torch.zeros(2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(gathered, msg)
        ok = True
        if north is not None:
            ok &= int(gathered[rank + 1][0]) == north[0]
        if south is not None:
            ok &= int(gathered[rank - 1][1]) == south[1]
        out[rank] = (t, ok)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_timing_reduction_and_neighbour_agreement():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert out[0][0] == out[1][0] == 11.0
    assert out[0][1] and out[1][1]
