"""GPU-only chain soak at config-B scale (measurement aid): several data sets and option mixes, every run under
PEDFAC_CHECK=1 (structure + likelihood of every picked move verified against a re-propagation)."""
import os, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pedfac_b200 import synth, frontend
GPU = ROOT / "pedfac_b200/bin/pedigraph"
rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 1)
bad = 0
for case in range(int(sys.argv[2]) if len(sys.argv) > 2 else 6):
    n = int(rng.choice([2000, 5000])); S = int(rng.choice([500, 2000])); soft = bool(rng.integers(2))
    selfing = bool(rng.random() < 0.3); seeded = bool(rng.random() < 0.3); unobs = int(rng.integers(1, 3)); max_age = float(rng.choice([1.0, 2.0]))
    n_gen = int(rng.choice([3, 4, 5]))
    d = synth.simulate(n, S, soft=soft, seed=int(rng.integers(1 << 30)), observe_frac=float(rng.choice([0.6, 0.8])), n_gen=n_gen)
    rd = Path(f"/tmp/gsoak_{case}")
    import shutil; shutil.rmtree(rd, ignore_errors=True)
    frontend.write_run_dir(d, rd, seeds=(int(rng.integers(1, 1 << 30)), int(rng.integers(1, 1 << 30))), max_unobs=unobs, max_age=max_age, seed_marriages=seeded)
    t = time.time()
    r = subprocess.run([str(GPU), "-d", str(rd), "-n", "4", "-r", str(rd / "rand")] + (["-s"] if selfing else []), capture_output=True, text=True,
                       env=dict(os.environ, PEDFAC_CHECK="1", PEDFAC_STATS="1"))
    last = [l for l in r.stderr.splitlines() if l.startswith("pedigraph-stats")][-1:] or [""]
    bad += r.returncode != 0
    print(f"case {case}: n={n} S={S} gens={n_gen} soft={soft} selfing={selfing} seeded={seeded} unobs={unobs} max_age={max_age} rc={r.returncode} {time.time() - t:.1f}s {last[0]}")
    if r.returncode: print(r.stderr[-400:])
print("failures:", bad)
