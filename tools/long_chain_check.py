"""Stability run (measurement aid): config-B-sized chain on the GPU for N iterations with PEDFAC_CHECK=1
(every picked move is verified against a re-propagation), plus a selfing-enabled (-s) and a
marriage.txt-seeded trace-parity run against the oracle-backed sampler on a small data set."""
import os, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pedfac_b200 import synth, frontend
GPU, REF = ROOT / "pedfac_b200/bin/pedigraph", ROOT / "oracle/pedigraph_ref"
n_iter = sys.argv[1] if len(sys.argv) > 1 else "10"

def run(binary, rd, n, extra=(), check=True):
    env = dict(os.environ, PEDFAC_STATS="1")
    if check:
        env["PEDFAC_CHECK"] = "1"
    t = time.time()
    r = subprocess.run([str(binary), "-d", str(rd), "-n", str(n), "-r", str(rd / "rand"), *extra], capture_output=True, text=True, env=env)
    return r, time.time() - t

d = synth.simulate(5000, 2000, soft=True, seed=46)
rd = Path("/tmp/lc_B"); frontend.write_run_dir(d, rd)
r, dt = run(GPU, rd, n_iter)
stats = [l for l in r.stderr.splitlines() if l.startswith("pedigraph-stats")]
print("config B", n_iter, "iterations rc", r.returncode, f"{dt:.1f}s"); print("\n".join(stats[-3:])); print(r.stderr[-300:] if r.returncode else "")

for name, extra, soft in (("selfing", ("-s",), False), ("selfing-soft", ("-s",), True)):
    d2 = synth.simulate(300, 160, soft=soft, seed=5)
    outs = []
    for tag, b in (("gpu", GPU), ("ref", REF)):
        rd2 = Path(f"/tmp/lc_{name}_{tag}"); frontend.write_run_dir(d2, rd2, seeds=(99, 1001))
        r2, dt2 = run(b, rd2, 4, extra)
        outs.append((r2.returncode, (rd2 / "out/ped.txt").read_text(), (rd2 / "out/ll.txt").read_text()))
    print(name, "rc", outs[0][0], outs[1][0], "ped identical:", outs[0][1] == outs[1][1], "ll identical:", outs[0][2] == outs[1][2], "rows", len(outs[0][1].splitlines()))
