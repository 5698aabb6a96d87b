"""Where a Gibbs step of the real sampler spends its time (measurement aid): runs pedfac_b200/bin/pedigraph
on a workload with PEDFAC_STATS=1 PF_HOST_PROF=1 and prints the per-iteration lines.
    python tools/chain_profile.py [B] [n_iter]"""
import os, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from pedfac_b200 import frontend, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "B"
n_iter = sys.argv[2] if len(sys.argv) > 2 else "3"
n, S, soft, _, _ = bench.WORKLOADS[wl]
d = synth.simulate(n, S, soft=soft, seed=46)
with tempfile.TemporaryDirectory() as tmp:
    rd = Path(tmp) / "run"
    frontend.write_run_dir(d, rd)
    r = subprocess.run([str(ROOT / "pedfac_b200/bin/pedigraph"), "-d", str(rd), "-n", n_iter, "-r", str(rd / "rand")],
                       capture_output=True, text=True, env=dict(os.environ, PEDFAC_STATS="1", PF_HOST_PROF="1"))
    print(r.stderr[-3000:])
