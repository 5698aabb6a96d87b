"""Generates tests/golden/ref_traces.json from RUNS OF THE REFERENCE ITSELF.

Runs only where /root/reference exists (the build container): `make -C oracle ref` embeds the
reference's Mach-O `exec/pedigraph` into the user-space loader oracle/_ref/pedigraph_macho
(oracle/macho_run.c), and this script drives that binary over a fixed list of small run directories.
Per case the fixture keeps the input files (so the tests do not depend on the synthetic generator
staying bit-stable), the command-line flags, and what the reference produced: its `prior:` line,
out/ped.txt, out/ll.txt, and — through the loader's PEDFAC_REF_DUMP_LL trap at
`_AddObsFracPrior@0x100019800` — the complete proposal list of the first Gibbs steps with every
log-likelihood at %.17g.

    python tools/make_reference_goldens.py        # rewrites tests/golden/ref_traces.json
"""
import json, os, shutil, subprocess, sys, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pedfac_b200 import frontend, synth
from pedfac_b200 import runpedfac as rp

MACHO = ROOT / "oracle/_ref/pedigraph_macho"
MAX_STEPS_WITH_PROPOSALS = 30

# name, simulate kwargs, write_run_dir kwargs, flags, iterations
CASES = [
    ("hard_default", dict(n_total=36, n_loci=24, soft=False, seed=46), dict(), [], 3),
    ("soft_default", dict(n_total=32, n_loci=16, soft=True, seed=47), dict(), [], 3),
    ("hard_noswap", dict(n_total=40, n_loci=20, soft=False, seed=48), dict(), ["-w"], 3),
    ("soft_seeded_marriages", dict(n_total=34, n_loci=12, soft=True, seed=49), dict(seed_marriages=True, seed_one_kid=True), [], 2),
    ("hard_two_unobs_layers", dict(n_total=30, n_loci=20, soft=False, seed=50, n_gen=4), dict(max_unobs=2, observe_frac=0.5), [], 3),
    ("hard_selfing_noswap", dict(n_total=24, n_loci=20, soft=False, seed=51, n_gen=2), dict(), ["-s", "-w"], 2),
    ("hard_sparse_observed", dict(n_total=60, n_loci=30, soft=False, seed=52, observe_frac=0.6), dict(observe_frac=0.5), [], 4),
    ("hard_swaps_many_iters", dict(n_total=48, n_loci=16, soft=False, seed=53, n_gen=4), dict(), [], 8),
]


def run_reference(run_dir, flags, n_iter):
    r = subprocess.run([str(MACHO), "-d", str(run_dir), "-n", str(n_iter), "-r", str(run_dir / "rand"), *flags], capture_output=True, text=True,
                       env=dict(os.environ, PEDFAC_REF_DUMP_LL="1", PEDFAC_REF_TRACE_FN="0x100019800"))
    if r.returncode != 0:
        raise RuntimeError(f"the reference exited with {r.returncode} on {run_dir}: {r.stderr[-400:]}")
    steps = []
    for line in r.stderr.splitlines():
        t = line.split()
        if line.startswith("choices kid"):
            steps.append({"kid": int(t[2]), "ids": [], "ll": []})
        elif line.startswith("c ") and steps:
            steps[-1]["ids"].append([int(x) for x in t[3:7]]); steps[-1]["ll"].append(t[7])
    prior = [l for l in r.stdout.splitlines() if l.startswith("prior:")][0].strip()
    return {"prior": prior, "ped": (run_dir / "out/ped.txt").read_text(), "ll": (run_dir / "out/ll.txt").read_text(),
            "n_steps": len(steps), "proposals": steps[:MAX_STEPS_WITH_PROPOSALS]}


def main():
    if not MACHO.exists():
        sys.exit("oracle/_ref/pedigraph_macho is missing: run `make -C oracle ref` where /root/reference exists")
    out = {"_made_by": "tools/make_reference_goldens.py", "_reference": "ngthomas/pedFac exec/pedigraph (Mach-O x86_64) run under oracle/macho_run.c on glibc",
           "cases": {}}
    with tempfile.TemporaryDirectory() as tmp:
        tmp = Path(tmp)
        for name, sim, wr, flags, n_iter in CASES:
            d = synth.simulate(**sim)
            rd = tmp / name
            frontend.write_run_dir(d, rd, **wr)
            files = {f.name: f.read_text() for f in sorted(rd.iterdir()) if f.is_file()}
            out["cases"][name] = {"files": files, "flags": flags, "n_iter": n_iter, **run_reference(rd, flags, n_iter)}
            print(name, out["cases"][name]["prior"], "rows", len(out["cases"][name]["ped"].splitlines()), "steps", out["cases"][name]["n_steps"])
        # BASELINE.json configs[0]: the reference's own example through the front-end, -n 5
        geno = Path("/root/reference/example/case_0/genotype.txt")
        (tmp / "c0").mkdir()
        rp.run_pedfac(geno, output_path=tmp / "c0", n_iter=5, binary=MACHO)
        rd = tmp / "c0" / "run"
        files = {k: (rd / k).read_text() for k in ("geno.txt", "prior.txt", "rand")}
        shutil.rmtree(rd / "out")
        out["cases"]["example_case_0_n5"] = {"files": files, "flags": [], "n_iter": 5, "front_end_input": "example/case_0/genotype.txt (run_pedfac defaults, random_seed=26)",
                                             **run_reference(rd, [], 5)}
        print("example_case_0_n5", out["cases"]["example_case_0_n5"]["prior"])
    dst = ROOT / "tests/golden/ref_traces.json"
    dst.write_text(json.dumps(out, separators=(",", ":")))
    print("wrote", dst, dst.stat().st_size, "bytes")


if __name__ == "__main__":
    main()
