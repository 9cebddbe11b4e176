"""How fast do the host<->device copies of one field go?  (1-D pinned copy vs the pitched 2-D copy of the library)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from wolfd2_b200 import api, deck as dk
import bench

d = bench.make_deck("cavity", 4096, True, 2, 100)
ctx = api.Context(d)
h, ptr = api.pinned_field(d)
nb = h.nbytes
for name, fn in (("upload 2-D", lambda: ctx.upload(api.F_U, h)), ("download 2-D", lambda: ctx.download(api.F_U, h))):
    fn(); ctx.sync()
    t = time.perf_counter()
    for _ in range(10):
        fn()
    ctx.sync()
    dt = (time.perf_counter() - t) / 10
    print(f"{name}: {dt * 1e3:.2f} ms  {nb / dt / 1e9:.1f} GB/s")
x = torch.empty(nb // 8, dtype=torch.float64).pin_memory()
y = torch.empty(nb // 8, dtype=torch.float64, device="cuda")
for name, fn in (("h2d 1-D", lambda: y.copy_(x, non_blocking=True)), ("d2h 1-D", lambda: x.copy_(y, non_blocking=True))):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 10
    print(f"{name}: {dt * 1e3:.2f} ms  {nb / dt / 1e9:.1f} GB/s")
