"""Run a few Gibbs-step passes of a workload (measurement aid for ncu): python tools/profile_eval.py C 6"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench
from pedfac_b200 import synth, Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "C"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
n, S, soft, _, _ = bench.WORKLOADS[wl]
d = synth.simulate(n, S, soft=soft, seed=46)
tr = bench.Trace(d, nsteps)
e = Engine(S, tr.cap_i, tr.cap_m)
synth.upload(e, d)
e.sweep(tr.full_view)
e.profile(True)
for rep in range(3):
    for v, k, p in tr.steps:
        e.sweep(v); e.eval(k, p)
ms, cnt = e.profile_read()
sb, eb = tr.algorithmic_bytes(S)
print("per launch us:", {k: round(ms[k] * 1e3 / max(cnt[k], 1), 1) for k in ms}, "launches", cnt)
print("eval GB/s", eb * 3 / (ms["eval"] * 1e-3) / 1e9, "sweep GB/s", sb * 3 / (ms["sweep"] * 1e-3) / 1e9)
