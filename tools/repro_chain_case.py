"""Debug aid: re-run one case of tools/gpu_chain_soak.py (same RNG stream) with PEDFAC_CHECK and state dumps.
CASE_ITERS=4 python tools/repro_chain_case.py SOAK_SEED CASE [binary]  ->  gpurun_out/case_{pre,post,err}.txt"""
import os, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pedfac_b200 import synth, frontend
seed, want = int(sys.argv[1]), int(sys.argv[2])
binary = sys.argv[3] if len(sys.argv) > 3 else str(ROOT / "pedfac_b200/bin/pedigraph")
rng = np.random.default_rng(seed)
for case in range(want + 1):
    n = int(rng.choice([2000, 5000])); S = int(rng.choice([500, 2000])); soft = bool(rng.integers(2))
    selfing = bool(rng.random() < 0.3); seeded = bool(rng.random() < 0.3); unobs = int(rng.integers(1, 3)); max_age = float(rng.choice([1.0, 2.0]))
    n_gen = int(rng.choice([3, 4, 5]))
    dseed = int(rng.integers(1 << 30)); of = float(rng.choice([0.6, 0.8]))
    seeds = (int(rng.integers(1, 1 << 30)), int(rng.integers(1, 1 << 30)))
print("case", case, dict(n=n, S=S, soft=soft, selfing=selfing, seeded=seeded, unobs=unobs, max_age=max_age, n_gen=n_gen, dseed=dseed, observe_frac=of, seeds=seeds))
d = synth.simulate(n, S, soft=soft, seed=dseed, observe_frac=of, n_gen=n_gen)
rd = Path(os.environ.get("CASE_DIR", "/tmp/chain_case")); shutil.rmtree(rd, ignore_errors=True)
frontend.write_run_dir(d, rd, seeds=seeds, max_unobs=unobs, max_age=max_age, seed_marriages=seeded)
if os.environ.get("CASE_ONLY_WRITE"): sys.exit(0)
out = ROOT / "gpurun_out"; out.mkdir(exist_ok=True)
args = [binary, "-d", str(rd), "-n", os.environ.get("CASE_ITERS", "1"), "-r", str(rd / "rand")] + (["-s"] if selfing else [])
env = dict(os.environ, PEDFAC_CHECK="1", PEDFAC_DEBUG="1", PEDFAC_DUMP=str(out / "case_post.txt"))
r = subprocess.run(args, capture_output=True, text=True, env=env)
(out / "case_err.txt").write_text(r.stderr[-30000:])
print(r.returncode, r.stderr[-400:])
bad = [l for l in r.stderr.splitlines() if l.startswith("PEDFAC_CHECK: iter")]
if bad:   # second run: dump the state the failing step was scored on
    kid = bad[0].split()[4]
    env.update(PEDFAC_DUMP_PRE=str(out / "case_pre.txt"), PEDFAC_DUMP_KID=kid)
    subprocess.run(args, capture_output=True, text=True, env=env)
