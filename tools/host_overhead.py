"""Where the host time of a Gibbs step goes (measurement aid): PF_HOST_PROF=1 python tools/host_overhead.py B"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from pedfac_b200 import synth, Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
n, S, soft, _, _ = bench.WORKLOADS[wl]
d = synth.simulate(n, S, soft=soft, seed=46)
tr = bench.Trace(d, 48)
e = Engine(S, tr.cap_i, tr.cap_m)
synth.upload(e, d)
e.sweep(tr.full_view)
for rep in range(3):
    t0 = time.perf_counter()
    for v, k, p in tr.steps:
        e.sweep_eval(v, k, p)
    dt = time.perf_counter() - t0
print(f"wall per Gibbs step through python: {dt / len(tr.steps) * 1e6:.1f} us")
e.profile(True)
for v, k, p in tr.steps:
    e.sweep_eval(v, k, p)
ms, cnt = e.profile_read()
print("device us per step:", {k: round(ms[k] * 1e3 / len(tr.steps), 1) for k in ms})
e.close()
