"""CPU tests of bench.py's host side: the Gibbs-step trace it replays, its byte accounting, the CPU
sample sizing and the reference arm's JSON line (no GPU needed: the reference arm runs the reference's
own binary inside oracle/_ref/pedigraph_macho where that was built, else the C port)."""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench
from oracle.binding import OracleEngine
from pedfac_b200 import synth


def small_trace():
    d = synth.simulate(300, 128, soft=False, seed=46)
    return d, bench.Trace(d, 12)


def test_trace_steps_are_replayable_and_restore_the_start_state():
    d, tr = small_trace()
    o = OracleEngine(d.n_loci, tr.cap_i, tr.cap_m)
    synth.upload(o, d)
    ll0 = o.sweep(tr.full_view).sum()
    stub0 = o.get_stub(int(tr.steps[0][1])).copy()
    for rep in range(2):                              # the trace must be repeatable: same numbers on every replay
        tot = []
        for view, kid, props in tr.steps:
            ll = o.sweep(view)
            ls = o.eval(kid, props)
            assert np.isfinite(ll).all() and len(ls) == len(props) and (ls <= 1e-9).all()
            tot.append((float(ll.sum()), float(ls[0])))
        o.sweep(tr.final_view)
        np.testing.assert_allclose(o.get_stub(int(tr.steps[0][1])), stub0, rtol=1e-12)   # re-attached at the end of the sibship: same value, another product order
        if rep == 0:
            first = tot
        else:
            assert tot == first
    assert np.isfinite(ll0)
    # the reference schedule has two sweeps per Gibbs step over the same proposals
    assert len(tr.ref_steps) == len(tr.steps)
    assert all(len(a[2]) == len(b[2]) for a, b in zip(tr.ref_steps, tr.steps))


def test_algorithmic_bytes_follow_the_design_formulas():
    d, tr = small_trace()
    sweep_b, eval_b = tr.algorithmic_bytes(d.n_loci)
    assert sweep_b > 0 and eval_b > 0
    # per sweep: at least one 24-byte stub per listed individual and locus
    n_ind = sum(len(v.indiv) for v, _, _ in tr.steps) + len(tr.final_view.indiv)
    assert sweep_b >= 24 * n_ind * d.n_loci
    # per eval: at most (kid + two rows per proposal) x 24 bytes per locus
    n_prop = sum(len(p) for _, _, p in tr.steps)
    assert eval_b <= 24 * d.n_loci * (2 * n_prop + len(tr.steps)) + 8 * n_prop


def test_cpu_sample_is_a_whole_number_of_64_locus_blocks():
    d, tr = small_trace()
    assert bench.cpu_sample_loci(d, tr, 1e9, 1) == d.n_loci          # unlimited budget: all loci, no extrapolation
    s = bench.cpu_sample_loci(d, tr, 1e-6, 1)
    assert s == 64                                                  # no budget: the minimum block


def test_reference_arm_prints_the_contract_line():
    r = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--workload", "A", "--steps", "1", "--warmup", "0",
                        "--trace-steps", "6", "--cpu-budget", "2"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-800:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "mcmc_proposals_per_sec" and line["unit"] == "proposals/s"
    assert line["higher_is_better"] is True and line["value"] > 0 and line["n_gpus"] == 1
    kind = "reference" if (ROOT / "oracle" / "_ref" / "pedigraph_macho").exists() else "port"
    assert line["cpu_baseline"]["kind"] == kind and line["cpu_baseline"]["cores"] == 1 and line["cpu_baseline"]["value"] == line["value"]
    assert line["e2e"] == {"value": line["value"], "unit": "proposals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # same `config` object as the CUDA arm prints for this workload and GPU count
    assert line["config"] == bench.workload_config("A", 1, trace_steps=6)
