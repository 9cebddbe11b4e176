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


# ---------------------------------------------------------------------------------------------------------
# slab layout / slab decks (the host side of csrc/w2_dist.cu)

def test_slab_layout_matches_library():
    """wolfd2_b200/slab.py and the library compute the same partition (pure arithmetic, no GPU needed)."""
    from wolfd2_b200 import api
    for nx, ny, world in [(300, 120, 2), (512, 2048, 2), (4096, 4096, 8), (4096, 32768, 8), (260, 400, 3), (64, 64, 1)]:
        for r in range(world):
            assert api.slab_layout(nx, ny, world, r) == slab.slab_layout(nx, ny, world, r)
    j0, j1, a0, a1, hg = slab.slab_layout(4096, 4096, 8, 3)
    assert hg == 5 and a0 == j0 - 5 and a1 == j1 + 5
    assert slab.slab_layout(300, 120, 2, 0)[4] == 10      # a 2048-unknown segment spans 7 rows of 299
    with pytest.raises(ValueError):
        slab.slab_layout(300, 40, 2, 0)                   # slabs must hold at least two halo depths
    with pytest.raises(api.Wolfd2Error):
        api.slab_layout(300, 40, 2, 0)


def test_slab_deck_metrics_are_the_global_ones():
    """A rank builds its metric rows from a node window; they must be bit-identical to the global arrays."""
    import numpy as np
    from wolfd2_b200 import deck
    for nx, ny, world in [(300, 200, 2), (260, 97, 3), (300, 400, 4)]:
        g = deck.cavity(nx, ny=ny, re=100.0)
        held = np.zeros(ny + 2, dtype=int)
        for r in range(world):
            d = deck.cavity(nx, ny=ny, re=100.0, slab=(r, world))
            cut = g.to_slab(r, world)
            assert d.slab == cut.slab and d.mny == cut.mny
            _, _, j0, j1, a0, a1, hg = d.slab
            held[j0:j1 + 1] += 1
            for k in d.metrics:
                if not k.startswith("_"):
                    assert np.array_equal(d.metrics[k], cut.metrics[k]), (nx, ny, world, r, k)
            f = g.new_field()
            f[:] = np.arange(f.size).reshape(f.shape)
            assert np.array_equal(d.window(f)[:a1 - a0 + 1], f[a0:a1 + 1, :d.mnx + 1])
        assert (held[2:ny + 1] == 1).all()                # every unknown row has exactly one owner


def _halo_worker(rank, world, port, out):
    """Replays w2_halo_exchange's row arithmetic with gloo send/recv on slab windows of one global field."""
    import numpy as np
    from wolfd2_b200 import deck
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = deck.cavity(300, ny=160, re=100.0)
        d = g.to_slab(rank, world)
        _, _, j0, j1, a0, a1, hg = d.slab
        rng = np.random.default_rng(5)
        f = g.new_field()
        f[:] = rng.standard_normal(f.shape)
        w = d.window(f).copy()
        w[:j0 - a0] = 0.0                                   # wipe the halos, then refill them from the owners
        w[j1 - a0 + 1:] = 0.0 if rank < world - 1 else w[j1 - a0 + 1:]
        if rank == 0:
            w[:2] = d.window(f)[:2]                         # rows 0,1 belong to rank 0
        reqs = []
        t = torch.from_numpy(w)
        if rank + 1 < world:
            reqs.append(dist.isend(t[j1 - hg + 1 - a0:j1 + 1 - a0].clone(), rank + 1))
            up = torch.zeros(hg, t.shape[1], dtype=torch.float64)
            reqs.append(dist.irecv(up, rank + 1))
        if rank > 0:
            reqs.append(dist.isend(t[j0 - a0:j0 + hg - a0].clone(), rank - 1))
            dn = torch.zeros(hg, t.shape[1], dtype=torch.float64)
            reqs.append(dist.irecv(dn, rank - 1))
        for q in reqs:
            q.wait()
        if rank + 1 < world:
            t[j1 + 1 - a0:j1 + 1 + hg - a0] = up
        if rank > 0:
            t[j0 - hg - a0:j0 - a0] = dn
        out[rank] = bool(np.array_equal(w[:a1 - a0 + 1], d.window(f)[:a1 - a0 + 1]))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_halo_rows_reassemble_the_global_field():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    port = 29100 + (os.getpid() % 500)
    mp.spawn(_halo_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)


def test_metric_windows_of_any_size_are_the_global_rows():
    """api.Context.stream_uniform_metrics uploads deck.metrics_window chunk by chunk: every window, down to a single
    row and including the physical edge rows 0 and ny+1, must reproduce the global arrays bit for bit."""
    import numpy as np
    from wolfd2_b200 import deck
    from wolfd2_b200._abi import METRIC_NAMES
    nx, ny = 41, 53
    g = deck.cavity(nx, ny=ny, re=100.0)
    for chunk in (1, 2, 5, 16, 60):
        for j0 in range(0, ny + 2, chunk):
            j1 = min(ny + 1, j0 + chunk - 1)
            w = deck.metrics_window(nx, ny, j0, j1, g.mnx)
            for k in METRIC_NAMES:
                assert np.array_equal(w[k], g.metrics[k][j0:j1 + 1, :]), (chunk, j0, k)
    lazy = deck.cavity(nx, ny=ny, re=100.0, lazy_metrics=True)
    assert lazy.metrics == {} and lazy.nx == nx and lazy.mny == g.mny
    m = lazy.metrics_struct()
    assert all(not getattr(m, k) for k in METRIC_NAMES)       # NULL pointers: nothing uploaded at create
