"""Worker of tests/test_gpu_multi.py: launched by torchrun, one rank per GPU.

Each rank steps its slab of a deck through the slab context and, on the same GPU, the whole deck through a
one-GPU context; the slab result must equal the one-GPU result bit for bit on every row the rank holds
(owned rows and halo rows), with identical QL / SOR iteration counts and max-norms."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from wolfd2_b200 import api, deck, slab  # noqa: E402


def stable_dt(nx, ny, re):
    """The scheme has no implicit j-coupling: dt/(Re h^2) must stay below ~0.25."""
    h = 1.0 / (max(nx, ny) - 1)
    return min(0.25 * h, 0.2 * re * h * h)


def cases():
    # nx >= 254 (fused SOR pipeline); slabs of >= 2*HG rows
    yield "cavity300x120", deck.cavity(300, ny=120, re=100.0, dt=stable_dt(300, 120, 100.0), sortol=1e-6, msorit=400,
                                       sorrel=1.5), 3
    # loose tolerance: the SOR converges, at odd and even iterations (mid-pass repeat of the fused pipeline)
    yield "cavity300x120_conv", deck.cavity(300, ny=120, re=100.0, dt=stable_dt(300, 120, 100.0), sortol=2e-3,
                                            msorit=400, sorrel=1.5), 6
    yield "channel320x96", deck.channel(320, ny=96, re=50.0, dt=stable_dt(320, 96, 50.0), fully_dev=True, sortol=1e-5,
                                        msorit=300, sorrel=1.3), 2
    yield "bstep400x128", deck.backward_step(400, ny=128, re=50.0, dt=stable_dt(400, 128, 50.0), fully_dev=True,
                                             sortol=1e-5, msorit=300, sorrel=1.3), 2
    # mass-conserving outlets (OUTLT2): the ghost fill of an east / west face is a recurrence along the whole face,
    # i.e. across the slabs (w2_bc.cu bc_scan_slab)
    yield "channel320x96_mass_cons", deck.channel(320, ny=96, re=50.0, dt=stable_dt(320, 96, 50.0), fully_dev=False,
                                                  sortol=1e-5, msorit=300, sorrel=1.3), 3
    yield "bstep400x128_mass_cons", deck.backward_step(400, ny=128, re=50.0, dt=stable_dt(400, 128, 50.0), fully_dev=False,
                                                       sortol=1e-5, msorit=300, sorrel=1.3), 2
    yield "cavity512x2048", deck.cavity(512, ny=2048, re=400.0, dt=stable_dt(512, 2048, 400.0), sortol=1e-7, msorit=60,
                                        sorrel=1.7), 2
    yield "cavity512x2048_filter", deck.cavity(512, ny=2048, re=400.0, dt=stable_dt(512, 2048, 400.0), sortol=1e-7,
                                               msorit=60, sorrel=1.7, nfiltu=1, nfiltv=1), 2


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    local = int(os.environ.get("LOCAL_RANK", rank))
    slab.init_comm(dist, local)
    for kv in os.environ.get("W2_OPTS", "").split(","):   # e.g. W2_OPTS=sor_slab_inpass=1 (A/B of a library option)
        if kv:
            k, v = kv.split("=")
            api.set_option(k, int(v))
    failures = []
    for name, g, nsteps in cases():
        try:
            slab.slab_layout(g.nx, g.ny, world, 0)
        except ValueError as e:   # too few rows for this many ranks: the same verdict on every rank
            if rank == 0:
                print(f"[rank 0] {name}: skipped ({e})", flush=True)
            continue
        rng = np.random.default_rng(7)
        u0, v0, p0 = g.new_field(), g.new_field(), g.new_field()
        for f in (u0, v0, p0):   # smooth-ish random start, identical on every rank
            f[0:g.ny + 2, 0:g.nx + 2] = 0.01 * rng.standard_normal((g.ny + 2, g.nx + 2))
        # one-GPU answer
        with api.Context(g) as c1:
            c1.upload(api.F_U, u0); c1.upload(api.F_V, v0); c1.upload(api.F_P, p0)
            ns1 = c1.coldstart()
            logs1 = c1.step(nsteps)
            ref = [c1.download(w) for w in (api.F_U, api.F_V, api.F_P)]
        d = g.to_slab(rank, world)
        with api.Context(d) as cs:
            cs.upload(api.F_U, d.window(u0)); cs.upload(api.F_V, d.window(v0)); cs.upload(api.F_P, d.window(p0))
            nss = cs.coldstart()
            logss = cs.step(nsteps)
            got = [cs.download(w) for w in (api.F_U, api.F_V, api.F_P)]
        _, _, j0, j1, a0, a1, hg = d.slab
        ok = ns1 == nss
        for l1, ls in zip(logs1, logss):
            same = l1["nQLiter"] == ls["nQLiter"] and l1["nSorConv"] == ls["nSorConv"] and l1["dif"] == ls["dif"]
            if not same:
                print(f"[rank {rank}] {name}: logs differ {l1} vs {ls}", flush=True)
            ok = ok and same
        worst = 0.0
        # owned rows (incl. the physical ghost rows a boundary rank owns) must match exactly; halo rows too,
        # except the outermost one (spare)
        lo = a0 if rank == 0 else a0 + 1
        hi = a1 if rank == world - 1 else a1 - 1
        for k, (r, s) in enumerate(zip(ref, got)):
            if k == 2:   # p: the SOR passes refresh 2T = 4 halo rows; deeper halo rows are never read
                lo_k, hi_k = max(lo, j0 - 4) if rank else a0, min(hi, j1 + 4) if rank < world - 1 else a1
            else:
                lo_k, hi_k = lo, hi
            a, b = r[lo_k:hi_k + 1, :], s[lo_k - a0:hi_k - a0 + 1, :]
            if not np.isfinite(a).all():
                ok = False
                print(f"[rank {rank}] {name}: the one-GPU field {'uvp'[k]} is not finite (bad test deck)", flush=True)
            if not np.array_equal(a, b):
                ok = False
                bad = np.argwhere(a != b)
                worst = max(worst, float(np.nanmax(np.abs(a - b))))
                print(f"[rank {rank}] {name}: field {'uvp'[k]} differs at {len(bad)} cells, rows "
                      f"{sorted(set((bad[:, 0] + lo_k).tolist()))[:12]} cols {sorted(set(bad[:, 1].tolist()))[:12]} "
                      f"nan {int(np.isnan(a).sum())}/{int(np.isnan(b).sum())}", flush=True)
        print(f"[rank {rank}] {name}: rows {a0}..{a1} (owned {j0}..{j1}) coldstart {ns1}/{nss} "
              f"QL {[l['nQLiter'] for l in logss]} SOR {[l['nSorConv'] for l in logss]} "
              f"{'OK' if ok else 'MISMATCH maxabs=%g' % worst}", flush=True)
        if not ok:
            failures.append(name)
    t = torch.tensor([len(failures)], dtype=torch.int64)
    dist.all_reduce(t)
    api.comm_finalize()
    dist.destroy_process_group()
    sys.exit(1 if int(t.item()) else 0)


if __name__ == "__main__":
    main()
