"""Wall-clock latency of tiny engine calls through the C ABI (measurement aid)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from pedfac_b200 import Engine, Pedigree, PF_PROP_TRIO, make_proposals

S = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
e = Engine(S, 64, 32)
e.set_locus_params(np.full(S, 0.02), np.full(S, 0.3))
rng = np.random.default_rng(0)
for i in range(8):
    e.set_hard(i, rng.integers(0, 3, size=S).astype(np.uint8))
ped = Pedigree(64, 32)
for i in range(8):
    ped.add_individual(i, i < 2)
view, _ = ped.view()
ped.add_marriage(0, 0, 1, [2, 3])
view2, _ = ped.view()
props = make_proposals([(PF_PROP_TRIO, 0, 1, -1)] * 300)
for name, fn in [("sweep 8 singletons", lambda: e.sweep(view)), ("sweep family+4", lambda: e.sweep(view2)), ("eval 300 proposals", lambda: e.eval(4, props))]:
    for _ in range(50): fn()
    t = time.perf_counter()
    for _ in range(500): fn()
    print(f"{name}: {(time.perf_counter() - t) / 500 * 1e6:.1f} us per call")
e.profile(True)
for _ in range(100): e.sweep(view2); e.eval(4, props)
ms, cnt = e.profile_read()
print({k: round(ms[k] * 1e3 / max(cnt[k], 1), 1) for k in ms}, "us per kernel launch")
