"""Trace-parity soak (measurement aid): random data sets and option mixes, the CUDA sampler against the
oracle-backed sampler (same source, same seed); ped.txt must be identical and ll.txt equal to its
printed 6 decimals (allowing the last digit)."""
import os, shutil, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pedfac_b200 import synth, frontend
GPU, REF = ROOT / "pedfac_b200/bin/pedigraph", ROOT / "oracle/pedigraph_ref"
n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 12
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 2026)
bad = 0
for case in range(n_cases):
    n = int(rng.integers(80, 420)); S = int(rng.integers(40, 260)); soft = bool(rng.integers(2))
    selfing = bool(rng.random() < 0.25); seeded = bool(rng.random() < 0.3); unobs = int(rng.integers(1, 3))
    obs_frac = float(rng.choice([0.5, 0.8, 1.0])); max_age = float(rng.choice([1.0, 2.0])); iters = int(rng.integers(2, 5))
    n_gen = int(rng.choice([2, 3, 3, 4, 5]))
    d = synth.simulate(n, S, soft=soft, seed=int(rng.integers(1 << 30)), observe_frac=float(rng.choice([0.6, 0.8, 0.95])), n_gen=n_gen)
    seeds = (int(rng.integers(1, 1 << 30)), int(rng.integers(1, 1 << 30)))
    outs = []
    for tag, b in ((("ref", REF), ("ref", REF)) if os.environ.get("SOAK_REF_ONLY") else (("gpu", GPU), ("ref", REF))):
        rd = Path(f"/tmp/soak_{case}_{tag}")
        shutil.rmtree(rd, ignore_errors=True)
        frontend.write_run_dir(d, rd, seeds=seeds, max_unobs=unobs, max_age=max_age, observe_frac=obs_frac, seed_marriages=seeded)
        t = time.time()
        r = subprocess.run([str(b), "-d", str(rd), "-n", str(iters), "-r", str(rd / "rand")] + (["-s"] if selfing else []),
                           capture_output=True, text=True, env=dict(os.environ, PEDFAC_CHECK="1", PEDFAC_STATS="1"))
        outs.append((r.returncode, (rd / "out/ped.txt").read_text() if r.returncode == 0 else r.stderr[-300:],
                     (rd / "out/ll.txt").read_text() if r.returncode == 0 else "", time.time() - t))
    ok = outs[0][0] == 0 and outs[1][0] == 0 and outs[0][1] == outs[1][1]
    if ok:
        a = np.array([float(x.split()[2]) for x in outs[0][2].splitlines()]); b2 = np.array([float(x.split()[2]) for x in outs[1][2].splitlines()])
        ok = len(a) == len(b2) and bool(np.all(np.abs(a - b2) <= 2e-6 + 1e-9 * np.abs(b2)))
    bad += not ok
    print(f"case {case}: n={n} S={S} gens={n_gen} soft={soft} selfing={selfing} seeded={seeded} unobs={unobs} obs_frac={obs_frac} max_age={max_age} iters={iters} "
          f"rows={len(outs[0][1].splitlines()) if outs[0][0] == 0 else -1} gpu {outs[0][3]:.1f}s ref {outs[1][3]:.1f}s -> {'OK' if ok else 'MISMATCH ' + str((outs[0][0], outs[1][0]))}")
    if not ok and outs[0][0] != 0: print(outs[0][1])
    if not ok and outs[1][0] != 0: print(outs[1][1])
print("mismatches:", bad, "of", n_cases)
sys.exit(1 if bad else 0)
