"""Trace parity against the REFERENCE ITSELF (measurement aid): random data sets and option mixes run
through the reference's own Mach-O `pedigraph` (oracle/_ref/pedigraph_macho, see oracle/macho_run.c)
and through one of our samplers (default: oracle/pedigraph_ref = the product's host chain logic on
the CPU oracle; `--gpu` = pedfac_b200/bin/pedigraph). Reports the first differing ped.txt / ll.txt row.

    python tools/ref_compare.py [n_cases] [seed] [--gpu] [--keep] [--small]
"""
import os, shutil, subprocess, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import numpy as np
from pedfac_b200 import synth, frontend
MACHO = ROOT / "oracle/_ref/pedigraph_macho"
OURS = ROOT / ("pedfac_b200/bin/pedigraph" if "--gpu" in sys.argv else "oracle/pedigraph_ref")
args = [a for a in sys.argv[1:] if not a.startswith("--")]
n_cases = int(args[0]) if args else 12
rng = np.random.default_rng(int(args[1]) if len(args) > 1 else 2026)
small = "--small" in sys.argv
bad = 0
died = 0


def first_diff(a, b):
    la, lb = a.splitlines(), b.splitlines()
    for k, (x, y) in enumerate(zip(la, lb)):
        if x != y:
            return k, x, y
    return (min(len(la), len(lb)), "<end>", "<end>") if len(la) != len(lb) else None


for case in range(n_cases):
    n = int(rng.integers(12, 60) if small else rng.integers(60, 320)); S = int(rng.integers(8, 60) if small else rng.integers(40, 200))
    soft = bool(rng.integers(2)); selfing = bool(rng.random() < 0.25); seeded = bool(rng.random() < 0.3)
    noswap = bool(rng.random() < 0.3) or selfing; unobs = int(rng.integers(1, 3))
    obs_frac = float(rng.choice([0.5, 0.8, 1.0])); max_age = float(rng.choice([1.0, 2.0])); iters = int(rng.integers(2, 5))
    n_gen = int(rng.choice([2, 3, 3, 4, 5]))
    d = synth.simulate(n, S, soft=soft, seed=int(rng.integers(1 << 30)), observe_frac=float(rng.choice([0.6, 0.8, 0.95])), n_gen=n_gen)
    seeds = (int(rng.integers(1, 1 << 30)), int(rng.integers(1, 1 << 30)))
    outs = []
    for tag, b in (("macho", MACHO), ("ours", OURS)):
        rd = Path(f"/tmp/refcmp_{case}_{tag}")
        shutil.rmtree(rd, ignore_errors=True)
        frontend.write_run_dir(d, rd, seeds=seeds, max_unobs=unobs, max_age=max_age, observe_frac=obs_frac, seed_marriages=seeded, seed_one_kid=True)
        t = time.time()
        try:
            r = subprocess.run([str(b), "-d", str(rd), "-n", str(iters), "-r", str(rd / "rand")] + (["-s"] if selfing else []) + (["-w"] if noswap else []),
                               capture_output=True, text=True, timeout=900, env=dict(os.environ, PEDFAC_STATS="1"))
            rc, so, se = r.returncode, r.stdout, r.stderr
        except subprocess.TimeoutExpired:
            rc, so, se = -999, "", "timeout"
        prior = [l for l in so.splitlines() if l.startswith("prior:")]
        outs.append(dict(rc=rc, ped=(rd / "out/ped.txt").read_text() if (rd / "out/ped.txt").exists() else "",
                         ll=(rd / "out/ll.txt").read_text() if (rd / "out/ll.txt").exists() else "",
                         prior=prior[0].strip() if prior else "", err=se[-300:], t=time.time() - t))
    a, b = outs
    dp, dl = first_diff(a["ped"], b["ped"]), first_diff(a["ll"], b["ll"])
    ok = a["rc"] == 0 and b["rc"] == 0 and dp is None and dl is None and a["prior"] == b["prior"]
    # the reference aborting on its own defects is not a mismatch as
    # long as everything it printed before dying is a prefix of ours (the last row may be the bad step)
    if not ok and a["rc"] != 0 and b["rc"] == 0 and a["prior"] == b["prior"]:
        la, lb = a["ped"].splitlines(), b["ped"].splitlines()
        if la[:-1] == lb[:max(0, len(la) - 1)]:
            print(f"case {case}: REF_DIED rc={a['rc']} after {len(la)} rows ({a['err'].strip()[-90:]}); ours agrees on the first {max(0, len(la) - 1)} rows"
                  f" (last row {'same' if la and la[-1] == lb[len(la) - 1] else 'differs'}) [n={n} S={S} max_age={max_age} unobs={unobs} selfing={selfing}]", flush=True)
            died += 1
            continue
    bad += not ok
    print(f"case {case}: n={n} S={S} gens={n_gen} soft={soft} selfing={selfing} seeded={seeded} noswap={noswap} unobs={unobs} "
          f"obs_frac={obs_frac} max_age={max_age} iters={iters} rows={len(a['ped'].splitlines())} macho {a['t']:.1f}s ours {b['t']:.1f}s "
          f"-> {'OK' if ok else 'MISMATCH'}", flush=True)
    if not ok:
        print(f"    rc {a['rc']} / {b['rc']}; prior '{a['prior']}' / '{b['prior']}'; first ped diff {dp}; first ll diff {dl}")
        if a["rc"] != 0: print("    macho stderr:", a["err"])
        if b["rc"] != 0: print("    ours stderr:", b["err"])
    elif "--keep" not in sys.argv:
        for tag in ("macho", "ours"):
            shutil.rmtree(f"/tmp/refcmp_{case}_{tag}", ignore_errors=True)
print("mismatches:", bad, "of", n_cases, "; reference died on its own defects in", died)
sys.exit(1 if bad else 0)
