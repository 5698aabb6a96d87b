"""Per-proposal log-likelihood parity against the REFERENCE ITSELF (measurement aid): runs the
reference's Mach-O `pedigraph` under oracle/macho_run.c with PEDFAC_REF_DUMP_LL (every proposal list of
every Gibbs step, %.17g, before the observation-fraction prior) and one of our samplers with
PEDFAC_DEBUG_LL, on the same run directory, and reports the largest relative difference.

    python tools/ref_ll_compare.py RUN_DIR N_ITER [--gpu] [pedigraph flags...]
"""
import os, shutil, subprocess, sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
MACHO = ROOT / "oracle/_ref/pedigraph_macho"


def proposal_lists(binary, run_dir, n_iter, flags, env):
    rd = Path(str(run_dir) + "_lltmp")
    shutil.rmtree(rd, ignore_errors=True); shutil.copytree(run_dir, rd); shutil.rmtree(rd / "out", ignore_errors=True)
    r = subprocess.run([str(binary), "-d", str(rd), "-n", str(n_iter), "-r", str(rd / "rand"), *flags],
                       capture_output=True, text=True, env=dict(os.environ, **env))
    steps = []
    for line in r.stderr.splitlines():
        t = line.split()
        if line.startswith("choices kid"):
            steps.append((int(t[2]), []))
        elif line.startswith("c ") and steps:
            steps[-1][1].append((tuple(int(x) for x in t[3:7]), float(t[7])))
    shutil.rmtree(rd, ignore_errors=True)
    return r.returncode, steps


def compare(run_dir, n_iter, flags=(), gpu=False):
    ours = ROOT / ("pedfac_b200/bin/pedigraph" if gpu else "oracle/pedigraph_ref")
    rc_a, a = proposal_lists(MACHO, run_dir, n_iter, flags, {"PEDFAC_REF_DUMP_LL": "1", "PEDFAC_REF_TRACE_FN": "0x100019800"})
    rc_b, b = proposal_lists(ours, run_dir, n_iter, flags, {"PEDFAC_DEBUG_LL": "1"})
    worst, n_prop, structural = 0.0, 0, 0
    for (ka, la), (kb, lb) in zip(a, b):
        if ka != kb or len(la) != len(lb):
            structural += 1
            continue
        for (ia, xa), (ib, xb) in zip(la, lb):
            n_prop += 1
            if ia != ib:
                structural += 1
            elif xa != xb:
                worst = max(worst, abs(xa - xb) / max(1.0, abs(xa)))
    return dict(rc_ref=rc_a, rc_ours=rc_b, steps_ref=len(a), steps_ours=len(b), proposals=n_prop, structural=structural, max_rel=worst)


if __name__ == "__main__":
    args = [x for x in sys.argv[1:] if x != "--gpu"]
    print(compare(args[0], int(args[1]), args[2:], gpu="--gpu" in sys.argv))
