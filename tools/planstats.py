"""Print the shape and device time of every sweep of the bench trace (measurement aid)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from pedfac_b200 import synth, Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 48
n, S, soft, _, _ = bench.WORKLOADS[wl]
if len(sys.argv) > 3: S = int(sys.argv[3])
quiet = len(sys.argv) > 4
d = synth.simulate(n, S, soft=soft, seed=46)
tr = bench.Trace(d, nsteps)
e = Engine(S, tr.cap_i, tr.cap_m)
synth.upload(e, d)
e.sweep(tr.full_view); print('full', e.plan_stats())
for rep in range(2):
    e.profile(True)
    rows = []
    for v, k, p in tr.steps:
        e.sweep(v); st = e.plan_stats(); ms, _ = e.profile_read(); st['us'] = round(ms['sweep'] * 1e3, 1); st['ni'] = len(v.indiv)
        e.eval(k, p); ms, _ = e.profile_read(); st['eval_us'] = round(ms['eval'] * 1e3, 1); st['np'] = len(p)
        rows.append(st)
for r in ([] if quiet else rows):
    print({k: r[k] for k in ('levels', 'groups', 'tasks', 'rows', 'rows_smem', 'ni', 'us', 'np', 'eval_us')})
import collections
print('S', S, 'groups histogram', sorted(collections.Counter(r['groups'] for r in rows).items()), 'mean levels', np.mean([r['levels'] for r in rows]))
print('mean sweep us', np.mean([r['us'] for r in rows]), 'mean eval us', np.mean([r['eval_us'] for r in rows]))
