"""The real sampler on a workload, on 1 or N GPUs of this box (measurement aid):
    python tools/chain_scale.py C 2 [gpus]     # workload, iterations, PEDFAC_GPUS
prints the per-iteration stats lines of rank 0 and the wall time."""
import os, subprocess, sys, tempfile, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from pedfac_b200 import frontend, synth
wl = sys.argv[1] if len(sys.argv) > 1 else "C"
n_iter = sys.argv[2] if len(sys.argv) > 2 else "2"
gpus = sys.argv[3] if len(sys.argv) > 3 else "1"
n, S, soft, _, _ = bench.WORKLOADS[wl]
t0 = time.time()
d = synth.simulate(n, S, soft=soft, seed=46)
with tempfile.TemporaryDirectory() as tmp:
    rd = Path(tmp) / "run"
    frontend.write_run_dir(d, rd)
    print(f"data + files: {time.time() - t0:.1f} s, geno.txt {os.path.getsize(rd / 'geno.txt') / 1e6:.0f} MB", flush=True)
    env = dict(os.environ, PEDFAC_STATS="1")
    if int(gpus) > 1:
        env["PEDFAC_GPUS"] = gpus
    t0 = time.time()
    r = subprocess.run([str(ROOT / "pedfac_b200/bin/pedigraph"), "-d", str(rd), "-n", n_iter, "-r", str(rd / "rand")], capture_output=True, text=True, env=env)
    print(f"rc {r.returncode}, wall {time.time() - t0:.1f} s")
    print("\n".join(l for l in r.stderr.splitlines() if "chain 0" in l or "warning" in l or "rror" in l)[-3000:])
    print(r.stdout[-300:])
