import sys, subprocess, os, time
sys.path.insert(0, '/root/repo')
from pathlib import Path
from pedfac_b200 import frontend, synth
ROOT = Path('/root/repo')
for soft in (False, True):
    d = synth.simulate(1000, 500, soft=soft, seed=11)
    outs = []
    for name, binary in (("gpu", ROOT / "pedfac_b200/bin/pedigraph"), ("ref", ROOT / "oracle/pedigraph_ref")):
        rd = Path(f"/tmp/bt_{name}_{int(soft)}")
        frontend.write_run_dir(d, rd, seeds=(4711, 1234))
        t = time.time()
        r = subprocess.run([str(binary), "-d", str(rd), "-n", "3", "-r", str(rd / "rand")], capture_output=True, text=True,
                           env=dict(os.environ, PEDFAC_STATS="1", PEDFAC_CHECK="1"))
        print(name, "soft" if soft else "hard", "rc", r.returncode, f"{time.time() - t:.1f}s", [l for l in r.stderr.splitlines() if l.startswith("pedigraph-stats")][-1])
        outs.append(((rd / "out/ped.txt").read_text(), (rd / "out/ll.txt").read_text()))
    print("ped identical:", outs[0][0] == outs[1][0], "ll lines:", len(outs[0][1].splitlines()))
    import numpy as np
    a = np.array([float(x.split()[2]) for x in outs[0][1].splitlines()]); b = np.array([float(x.split()[2]) for x in outs[1][1].splitlines()])
    print("max |ll diff|:", np.abs(a - b).max() if len(a) == len(b) else "length differs")
