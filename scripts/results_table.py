"""Round results for BASELINE.md / profiles/README.md: CPU port at 64^2 (both builds), run on the GPU box."""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from wolfd2_b200 import deck as dk

out = {}
d = dk.cavity(64, re=100.0, dt=0.01)          # configs[0]: reference defaults, ppe_solver sor, tol 1e-8, 2000 its
d.ppe_solver = "sor"
for opt in (False, True):
    t, kind = bench.cpu_steps(d, 100, warm=5, opt=opt)
    out["cpu_cavity64_sor_" + ("O3fast" if opt else "O2")] = {"steps_per_s": 100 / t, "gcell_per_s": d.cells() * 100 / t / 1e9, "build": kind}
d.ppe_solver = "rb_sor"
t, kind = bench.cpu_steps(d, 100, warm=5, opt=True)
out["cpu_cavity64_rbsor_O3fast"] = {"steps_per_s": 100 / t, "gcell_per_s": d.cells() * 100 / t / 1e9}
import platform
try:
    out["cpu_model"] = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
except Exception:
    out["cpu_model"] = platform.processor()
out["nproc"] = os.cpu_count()
print(json.dumps(out, indent=1))
