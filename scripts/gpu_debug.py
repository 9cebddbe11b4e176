"""Prints every parity figure GPU-vs-oracle without asserting (development aid)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from wolfd2_b200 import api, deck as dk
from oracle import get_oracle
from util import rel_l2, rand_field, make_test_decks, region_args

o = get_oracle()
rng = np.random.default_rng(12345)
print(api.lib().wolfd2_b200_version().decode())

def cmp(tag, a, b):
    print(f"  {tag:28s} bitwise={np.array_equal(a, b)}  maxabs={np.max(np.abs(a-b)):.3e} rel_l2={rel_l2(a,b):.3e}")

for d in make_test_decks():
    print("==", d.name, d.nx, d.ny)
    api.config(d.mnx, d.mny); o.config(d.mnx, d.mny)
    r = d.regions
    u, v, p = rand_field(d, rng), rand_field(d, rng), rand_field(d, rng)
    for name, gf, of in (("VelBoundCond", api.VelBoundCond, o.velboundcond), ("VelOutflowBCs", api.VelOutflowBCs, o.veloutflowbcs)):
        ug, vg, uo, vo = u.copy(), v.copy(), u.copy(), v.copy()
        gf(d.nx, d.ny, r.nReg, r.nRegBrd, r.nMomBdTp, r.dBCVal, ug, vg)
        of(d.nx, d.ny, r.nReg, r.nRegBrd, r.nMomBdTp, r.dBCVal, uo, vo)
        cmp(name + " u", ug, uo); cmp(name + " v", vg, vo)
    pg, po = p.copy(), p.copy()
    api.PresBoundCond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.dBCVal, pg)
    o.presboundcond(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, r.dBCVal, po)
    cmp("PresBoundCond", pg, po)
    m = d.metrics
    for nloc, names in ((1, "xeu yeu xzv yzv"), (2, "xev yev xzu yzu")):
        dg, do = d.new_field(), d.new_field()
        ms = [m[n] for n in names.split()]
        api.Divergence(d.nx, d.ny, nloc, *ms, u, v, dg)
        o.divergence(d.nx, d.ny, nloc, *ms, u, v, do)
        cmp(f"Divergence nloc={nloc}", dg, do)
    ug, vg, uo, vo = u.copy(), v.copy(), u.copy(), v.copy()
    pm = [m[n] for n in "dju djv yeu xzv yzu xev".split()]
    api.Project(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, *pm, p, ug, vg)
    o.project(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, d.dk, *pm, p, uo, vo)
    cmp("Project u", ug, uo); cmp("Project v", vg, vo)
    print("  DiffMaxNorm", api.DiffMaxNorm(d.nx, d.ny, u, v), o.diffmaxnorm(d.nx, d.ny, u, v),
          " DMaxNorm", api.DMaxNorm(d.nx, d.ny, u), o.dmaxnorm(d.nx, d.ny, u))
    for comp in (1, 2):
        qg, qo = u.copy(), u.copy()
        api.Filter(d.nx, d.ny, comp, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, np.zeros(200, np.int32), 5.0, qg)
        o.filter(d.nx, d.ny, comp, r.nReg, r.nRegBrd, r.nRegType, r.nMomBdTp, np.zeros(200, np.int32), 5.0, qo)
        cmp(f"Filter comp={comp}", qg, qo)
    # momentum
    us, vs, un, vn = (rand_field(d, rng, -0.5, 0.5) for _ in range(4))
    xm = [m[n] for n in "rbn rgn rac rbc dju xec yec xzn yzn xeu yeu xzu yzu".split()]
    dg, do = d.new_field(), d.new_field()
    api.XMomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2, *xm, us, vs, un, vn, dg)
    o.xmomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, r.dPRporos, r.dPRporc1, r.dPRporc2, *xm, us, vs, un, vn, do)
    cmp("XMomentum dus", dg, do)
    ym = [m[n] for n in "ran rbn rbc rgc djv xen yen xzc yzc xev yev xzv yzv".split()]
    z = d.new_field()
    dg, do = d.new_field(), d.new_field()
    api.YMomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, z, z, us, vs, un, vn, dg)
    o.ymomentum(d.nx, d.ny, *region_args(d), d.dk, d.re, d.fr, r.dPRporos, r.dPRporc1, r.dPRporc2, *ym, z, z, us, vs, un, vn, do)
    cmp("YMomentum dvs", dg, do)
    # Ppe
    for cart in (1, 0):
        pg, po = p.copy() * 0.01, p.copy() * 0.01
        pm8 = [m[n] for n in "rau rbu rbv rgv xeu yeu xzv yzv".split()]
        ng = api.Ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, 5, 300, d.dk, 1e-8, 1.3, *pm8, u * 0.01, v * 0.01, pg)
        no = o.ppe(d.nx, d.ny, r.nReg, r.nRegBrd, r.nRegType, cart, 5, 300, d.dk, 1e-8, 1.3, *pm8, u * 0.01, v * 0.01, po)
        print(f"  Ppe cart={cart} nSorConv gpu={ng} oracle={no}", end="")
        cmp("", pg, po)

# tridiagonal
for n in (7, 100, 4096, 4097, 70000, 1000003):
    api.config(2000, 2000); o.config(2000, 2000)
    a = np.zeros((n, 3)); a[:, 0] = rng.uniform(-1, 1, n); a[:, 2] = rng.uniform(-1, 1, n)
    a[:, 1] = 2.5 + rng.uniform(0, 1, n)
    b = rng.uniform(-1, 1, n)
    ag, bg, ao, bo = a.copy(), b.copy(), a.copy(), b.copy()
    api.AltTridLU(n, ag.reshape(-1), bg)
    o.alttridlu(n, ao.reshape(-1), bo)
    print(f"AltTridLU n={n}: maxabs={np.max(np.abs(bg-bo)):.3e} rel={rel_l2(bg,bo):.3e}")

# whole steps
for d in [dk.cavity(64, re=100., dt=0.01), dk.channel(48, re=100., dt=0.005, ny=40), dk.backward_step(64, re=100., dt=0.005, ny=48),
          dk.channel(48, re=100., dt=0.005, ny=40, fully_dev=False)]:
    d.msorit = 400
    print("== steps", d.name)
    o.config(d.mnx, d.mny)
    uo, vo, po = d.new_field(), d.new_field(), d.new_field()
    nso = o.coldstart(d, uo, vo, po)
    with api.Context(d) as ctx:
        z = d.new_field()
        for w in (api.F_U, api.F_V, api.F_P): ctx.upload(w, z)
        nsg = ctx.coldstart()
        print("  coldstart nSor", nsg, nso)
        for k in range(6):
            t0 = time.time(); lg = ctx.step(1)[0]; tg = time.time() - t0
            rc, lo = o.step(d, uo, vo, po, 1)
            ug, vg, pg = ctx.download(api.F_U), ctx.download(api.F_V), ctx.download(api.F_P)
            print(f"  step {k}: QL {lg['nQLiter']}/{lo[0]['nQLiter']} SOR {lg['nSorConv']}/{lo[0]['nSorConv']} "
                  f"relL2 u={rel_l2(ug,uo):.2e} v={rel_l2(vg,vo):.2e} p={rel_l2(pg,po):.2e} dif_g={lg['dif'][:3]} t={tg*1e3:.1f}ms")
        print("  timing", ctx.timing())
