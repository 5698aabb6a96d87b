"""Start-up prefilter (_RefineParentOffPair) at workload scale: every observed kid against every observed
individual one generation up, top 15 per kid. Prints device time, pair-locus evaluations per second and
the algorithmic bandwidth (measurement aid): python tools/prefilter_bench.py B"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import bench
from pedfac_b200 import synth, Engine

wl = sys.argv[1] if len(sys.argv) > 1 else "B"
n, S, soft, _, _ = bench.WORKLOADS[wl]
d = synth.simulate(n, S, soft=soft, seed=46)
cap_i, cap_m = 3 * n, 2 * n
e = Engine(S, cap_i, cap_m)
synth.upload(e, d)
ped = synth.singletons_pedigree(d, cap_i, cap_m)
view, glob = ped.view()
ll = e.sweep(view)
obs = np.flatnonzero(d.observed)
tot_pairs = 0; tot_ms = 0.0; wall = 0.0
for g in sorted(set(d.gen[obs])):
    kids = obs[d.gen[obs] == g]
    cands = obs[d.gen[obs] == g + 1]
    if len(kids) == 0 or len(cands) == 0:
        continue
    role = (d.sex[kids] == 1).astype(np.uint8)
    bias = -ll[[int(np.where(glob == ped.components()[0][c])[0][0]) if False else 0 for c in cands]] if False else np.zeros(len(cands))
    for rep in range(3):
        e.profile(True)
        t0 = time.perf_counter()
        slot, score = e.prefilter(kids, role, cands, bias, top_k=15)
        dt = time.perf_counter() - t0
        ms, cnt = e.profile_read()
    tot_pairs += len(kids) * len(cands); tot_ms += ms["prefilter"]; wall += dt
    print(f"generation {g}: {len(kids)} kids x {len(cands)} candidates x {S} loci: device {ms['prefilter']:.3f} ms ({cnt['prefilter']} launches), call {dt * 1e3:.3f} ms")
pl = tot_pairs * S
print(f"total: {tot_pairs} pairs, {pl / 1e9:.2f} G pair-locus evaluations in {tot_ms:.3f} ms device = {pl / tot_ms / 1e9 * 1e3:.1f} G pair-loci/s; "
      f"fp64 ~{pl * 30 / tot_ms / 1e9 * 1e3 / 1e3:.2f} TFLOP/s at 30 flops per pair-locus; call wall {wall * 1e3:.1f} ms")
