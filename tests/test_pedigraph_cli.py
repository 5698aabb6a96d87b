"""The `pedigraph` sampler executable: CLI + file contract and chain behaviour.
CPU tests drive oracle/pedigraph_ref (the product's host chain logic linked against the CPU
oracle); the -m gpu test drives pedfac_b200/bin/pedigraph (linked against libpedfac_cuda) and
requires the two to sample the identical pedigree trace from the same inputs and seed."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from pedfac_b200 import frontend, synth

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "pedigraph_ref"
GPU = ROOT / "pedfac_b200" / "bin" / "pedigraph"


def run(binary, run_dir, n_iter=2, extra=(), env=None):
    import os
    return subprocess.run([str(binary), "-d", str(run_dir), "-n", str(n_iter), "-r", str(Path(run_dir) / "rand"), *extra],
                          capture_output=True, text=True, timeout=900, env={**os.environ, **(env or {})})


def swap_counts(stderr):
    """(proposed, picked) swap moves summed over the iterations (PEDFAC_STATS lines)."""
    prop = pick = 0
    for line in stderr.splitlines():
        if line.startswith("pedigraph-stats"):
            t = line.split()
            kv = dict(zip(t[1::2], t[2::2]))
            prop += int(kv["swaps"]); pick += int(kv["swaps_picked"])
    return prop, pick


def make_run(tmp_path, name, n=80, S=60, soft=False, seed=46, seeds=(2590, 7579545), seed_marriages=False):
    d = synth.simulate(n, S, soft=soft, seed=seed)
    order = frontend.write_run_dir(d, tmp_path / name, seeds=seeds, seed_marriages=seed_marriages)
    return d, order, tmp_path / name


def test_outputs_follow_the_file_contract(tmp_path):
    d, order, rd = make_run(tmp_path, "r")
    r = run(REF, rd, n_iter=3)
    assert r.returncode == 0, r.stderr
    out = r.stdout
    for needle in ["SEEDFILE: Using contents of", "Allow selfing: 0", "Allow loopy: 0", "propagating message",
                   "prior: ", "finish propagating message", "sample for 3 iterations: --", "..0", "..2", ".. DONE"]:
        assert needle in out
    ped = (rd / "out" / "ped.txt").read_text().splitlines()
    ll = (rd / "out" / "ll.txt").read_text().splitlines()
    assert len(ped) >= len(ll) > 0          # a picked swap prints one row per kid of both families
    assert all(re.fullmatch(r"\d+ \d+ \d+ \d+", x) for x in ped)          # "%d %d %d %d\n"
    assert all(re.fullmatch(r"\d+ \d+ -?\d+\.\d{6}", x) for x in ll)       # "%d %d %f\n"
    assert (rd / "out" / "geno.txt").read_text() == ""                     # opened, never written
    assert (rd / "out" / "historyNode.csv").read_text().startswith("t,action,id,label,level,group\n")
    assert (rd / "out" / "historyEdge.csv").read_text().startswith("t,action,id,from,to\n")
    rows = np.array([list(map(int, x.split())) for x in ped])
    assert set(rows[:, 0]) == {0, 1, 2}
    n_obs = len(order)
    # observed ids are 1..n; unobserved parents are numbered maxID + slot = (n + 1) + slot, slot >= n
    assert rows[:, 1].min() >= 1 and (rows[:, 2:] >= 1).all()
    assert ((rows[:, 2:] <= n_obs) | (rows[:, 2:] >= 2 * n_obs + 1)).all()
    # the oldest generation is never sampled (gen >= nGen - 1)
    top_ids = {k + 1 for k, i in enumerate(order) if d.gen[i] == d.gen[order].max()}
    assert not (set(rows[:, 1]) & top_ids)


def test_same_seed_same_trace_and_recovers_true_parents(tmp_path):
    d, order, rd = make_run(tmp_path, "a", n=200, S=150)
    _, _, rd2 = make_run(tmp_path, "b", n=200, S=150)
    _, _, rd3 = make_run(tmp_path, "c", n=200, S=150, seeds=(11, 22))
    for x in (rd, rd2, rd3):
        assert run(REF, x, n_iter=3).returncode == 0
    assert (rd / "out" / "ped.txt").read_text() == (rd2 / "out" / "ped.txt").read_text()
    assert (rd / "out" / "ll.txt").read_text() == (rd2 / "out" / "ll.txt").read_text()
    assert (rd / "out" / "ped.txt").read_text() != (rd3 / "out" / "ped.txt").read_text()
    rows = np.loadtxt(rd / "out" / "ped.txt", dtype=int)
    last = rows[rows[:, 0] == rows[:, 0].max()]
    uid = {int(i): k + 1 for k, i in enumerate(order)}
    inv = {v: k for k, v in uid.items()}
    ok = tot = 0
    for _, kid, pa, ma in last:
        k = inv.get(int(kid))
        if k is None or d.pa[k] < 0 or not (d.observed[d.pa[k]] and d.observed[d.ma[k]]):
            continue
        tot += 1
        ok += {int(pa), int(ma)} == {uid[int(d.pa[k])], uid[int(d.ma[k])]}
    assert tot > 20 and ok / tot > 0.9


def test_soft_calls_and_error_paths(tmp_path):
    d, order, rd = make_run(tmp_path, "s", n=60, S=40, soft=True)
    assert run(REF, rd, n_iter=1).returncode == 0
    bad = tmp_path / "bad"
    frontend.write_run_dir(d, bad)
    (bad / "prior.txt").write_text((bad / "prior.txt").read_text() + "\nnonsense 3")
    r = run(REF, bad)
    assert r.returncode != 0 and "cannot recognize this keyword: nonsense" in r.stderr
    # the two non-default modes, as observed on the reference binary (tests/test_reference_parity.py)
    r = subprocess.run([str(REF), "-d", str(rd), "-r", str(rd / "rand"), "-l"], capture_output=True, text=True)
    assert r.returncode == 255 and r.stdout.endswith("propagating message\npropagation step exceeds the acyclic upper limit")
    assert "exceeding the upper limits of sum-product recursion counter" in r.stderr
    ped0 = (rd / "out" / "ped.txt").read_text()
    assert run(REF, rd, n_iter=1).returncode == 0
    ped0 = (rd / "out" / "ped.txt").read_text()
    r = subprocess.run([str(REF), "-d", str(rd), "-n", "1", "-r", str(rd / "rand"), "-c", "1"], capture_output=True, text=True)
    assert r.returncode == 0 and "runs as -c 0" in r.stderr and (rd / "out" / "ped.txt").read_text() == ped0


def test_swap_moves_are_scored_consistently_and_can_be_disabled(tmp_path):
    """Default mode proposes swap moves (_ProposeSwapping); PEDFAC_CHECK re-propagates after every
    picked move (attach, prong and swap alike) and exits 3 unless the pedigree's total
    log-likelihood equals the score the move was picked with. -w disables swaps."""
    d, order, rd = make_run(tmp_path, "w1", n=200, S=150)
    _, _, rd2 = make_run(tmp_path, "w2", n=200, S=150)
    r = run(REF, rd, n_iter=4, env={"PEDFAC_CHECK": "1", "PEDFAC_STATS": "1"})
    assert r.returncode == 0, r.stderr[-2000:]
    proposed, picked = swap_counts(r.stderr)
    assert proposed > 50 and picked > 0
    ped = (rd / "out" / "ped.txt").read_text().splitlines()
    ll = (rd / "out" / "ll.txt").read_text().splitlines()
    assert len(ped) > len(ll)                         # the picked swaps printed whole families
    assert all(re.fullmatch(r"\d+ \d+ \d+ \d+", x) for x in ped)
    r2 = run(REF, rd2, n_iter=4, extra=("-w",), env={"PEDFAC_CHECK": "1", "PEDFAC_STATS": "1"})
    assert r2.returncode == 0, r2.stderr[-2000:]
    assert swap_counts(r2.stderr) == (0, 0)
    assert len((rd2 / "out" / "ped.txt").read_text().splitlines()) == len((rd2 / "out" / "ll.txt").read_text().splitlines())
    assert (rd / "out" / "ped.txt").read_text() != (rd2 / "out" / "ped.txt").read_text()


def test_marriage_file_seeds_the_chain(tmp_path):
    """Optional marriage.txt (`kid pa ma` rows by user id, nMar couples in prior.txt;
    _ReadPedfile@0x10000b7b0): the chain starts from that pedigree instead of all singletons."""
    d, order, rd = make_run(tmp_path, "m1", n=200, S=120, seed_marriages=True)
    _, _, rd0 = make_run(tmp_path, "m0", n=200, S=120)
    assert (rd / "marriage.txt").exists() and not (rd0 / "marriage.txt").exists()
    env = {"PEDFAC_STATS": "1", "PEDFAC_CHECK": "1"}
    r, r0 = run(REF, rd, n_iter=2, env=env), run(REF, rd0, n_iter=2, env=env)
    assert r.returncode == 0 and r0.returncode == 0, r.stderr[-1000:]
    assert "optional marriage.txt is not found" in r0.stdout and "optional marriage.txt is not found" not in r.stdout
    # the seeded start has families already: another initial log-likelihood (kids without parents carry
    # no prior at all, so it is lower); the unseeded chain has no families, hence no swap proposals, in iteration 0
    prior = lambda out: float(re.search(r"prior: (-?[\d.]+)", out).group(1))
    assert prior(r.stdout) < prior(r0.stdout)
    first = lambda err: dict(zip(*[iter(err.splitlines()[[l.startswith("pedigraph-stats") for l in err.splitlines()].index(True)].split()[1:])] * 2))
    assert swap_counts(r.stderr)[0] > 0 and int(first(r0.stderr)["swaps"]) == 0
    bad = tmp_path / "mbad"
    frontend.write_run_dir(d, bad, seed_marriages=True)
    (bad / "prior.txt").write_text((bad / "prior.txt").read_text().replace("nMar ", "nMar 1"))
    rb = run(REF, bad, n_iter=1)
    assert rb.returncode != 0 and "doesn't match up with nMar" in rb.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("soft", [False, True])
def test_gpu_sampler_reproduces_the_cpu_trace_from_a_seeded_pedigree(tmp_path, soft):
    d, order, rd = make_run(tmp_path, "gm", n=220, S=150, soft=soft, seed_marriages=True)
    _, _, rd2 = make_run(tmp_path, "hm", n=220, S=150, soft=soft, seed_marriages=True)
    env = {"PEDFAC_STATS": "1", "PEDFAC_CHECK": "1"}
    rg, rc = run(GPU, rd, n_iter=3, env=env), run(REF, rd2, n_iter=3, env=env)
    assert rg.returncode == 0, rg.stderr
    assert rc.returncode == 0, rc.stderr
    assert (rd / "out" / "ped.txt").read_text() == (rd2 / "out" / "ped.txt").read_text()
    a = np.loadtxt(rd / "out" / "ll.txt"); b = np.loadtxt(rd2 / "out" / "ll.txt")
    np.testing.assert_allclose(a[:, 2], b[:, 2], rtol=1e-9, atol=2e-6)


@pytest.mark.gpu
@pytest.mark.parametrize("soft", [False, True])
def test_gpu_sampler_reproduces_the_cpu_trace(tmp_path, soft):
    """north_star: sampled pedigree traces identical for the same inputs and seed; log-likelihoods
    of the picked choices agree to 1e-9 relative (ll.txt prints 6 decimals)."""
    d, order, rd = make_run(tmp_path, "g", n=260, S=180, soft=soft)
    _, _, rd2 = make_run(tmp_path, "h", n=260, S=180, soft=soft)
    env = {"PEDFAC_STATS": "1", "PEDFAC_CHECK": "1"}
    rg, rc = run(GPU, rd, n_iter=4, env=env), run(REF, rd2, n_iter=4, env=env)
    assert rg.returncode == 0, rg.stderr
    assert rc.returncode == 0, rc.stderr
    assert swap_counts(rg.stderr) == swap_counts(rc.stderr) and swap_counts(rg.stderr)[1] > 0   # swap moves are part of the trace
    assert (rd / "out" / "ped.txt").read_text() == (rd2 / "out" / "ped.txt").read_text()
    a = np.loadtxt(rd / "out" / "ll.txt"); b = np.loadtxt(rd2 / "out" / "ll.txt")
    np.testing.assert_array_equal(a[:, :2], b[:, :2])
    np.testing.assert_allclose(a[:, 2], b[:, 2], rtol=1e-9, atol=2e-6)
    pg = float(re.search(r"prior: (-?[\d.]+)", rg.stdout).group(1)); pc = float(re.search(r"prior: (-?[\d.]+)", rc.stdout).group(1))
    assert abs(pg - pc) <= 1e-9 * abs(pc) + 2e-6


@pytest.mark.gpu
def test_gpu_sampler_trace_parity_with_other_priors(tmp_path):
    """Two unobserved layers, a two-generation age window and a lower observed fraction: more
    unobserved parents are created and trimmed, more candidates per step."""
    d = synth.simulate(240, 140, soft=True, seed=9, observe_frac=0.6)
    outs = []
    for name, binary in (("g", GPU), ("c", REF)):
        rd = tmp_path / name
        frontend.write_run_dir(d, rd, seeds=(31, 77), max_unobs=2, min_age=1.0, max_age=2.0, observe_frac=0.5)
        r = run(binary, rd, n_iter=3, env={"PEDFAC_STATS": "1", "PEDFAC_CHECK": "1"})
        assert r.returncode == 0, r.stderr[-1500:]
        outs.append(((rd / "out" / "ped.txt").read_text(), np.loadtxt(rd / "out" / "ll.txt")))
    assert outs[0][0] == outs[1][0]
    np.testing.assert_allclose(outs[0][1][:, 2], outs[1][1][:, 2], rtol=1e-9, atol=2e-6)


def test_other_priors_run_consistently_on_the_cpu(tmp_path):
    d = synth.simulate(160, 90, soft=False, seed=9, observe_frac=0.6)
    rd = tmp_path / "p"
    frontend.write_run_dir(d, rd, seeds=(31, 77), max_unobs=2, min_age=1.0, max_age=2.0, observe_frac=0.5)
    r = run(REF, rd, n_iter=3, env={"PEDFAC_STATS": "1", "PEDFAC_CHECK": "1"})
    assert r.returncode == 0, r.stderr[-1500:]
    rows = np.loadtxt(rd / "out" / "ped.txt", dtype=int)
    assert (rows[:, 1] > len(np.flatnonzero(d.observed))).any()        # unobserved individuals get sampled too


def test_geno_parser_accepts_equivalent_spellings(tmp_path):
    """CRLF line ends, repeated blanks, trailing zeros and short decimals spell the same data; five
    significant decimals force the exact-double path. Same chain in every case."""
    d, order, rd = make_run(tmp_path, "f0", n=70, S=50, soft=True)
    base = (rd / "geno.txt").read_text()
    variants = {
        "crlf": base.replace("\n", "\r\n"),
        "blanks": base.replace(" ", "  "),
        "zeros": re.sub(r"(\d\.\d{4})(?=[, \n])", r"\g<1>00", base),
    }
    want = None
    for name, text in {"plain": base, **variants}.items():
        vd = tmp_path / name
        frontend.write_run_dir(d, vd)
        (vd / "geno.txt").write_text(text)
        r = run(REF, vd, n_iter=2)
        assert r.returncode == 0, (name, r.stderr[-500:])
        got = (vd / "out" / "ped.txt").read_text(), (vd / "out" / "ll.txt").read_text()
        want = want or got
        assert got == want, name
    # a fifth significant decimal cannot be a 16-bit numerator: the row goes through the f64 path and still runs
    odd = tmp_path / "odd"
    frontend.write_run_dir(d, odd)
    (odd / "geno.txt").write_text(re.sub(r"0\.(\d{4}),", r"0.\g<1>1,", base, count=3))
    assert run(REF, odd, n_iter=1).returncode == 0
    short = tmp_path / "short"
    frontend.write_run_dir(d, short)
    lines = base.splitlines()
    lines[3] = " ".join(lines[3].split(" ")[:-2])
    (short / "geno.txt").write_text("\n".join(lines) + "\n")
    r = run(REF, short, n_iter=1)
    assert r.returncode != 0 and "fewer SNPs than reported" in r.stderr and "obs: 48 , exp: 50" in r.stderr


TRIM_CASES = [
    # n, S, soft, data seed, rng seeds, max_unobs, observe_frac (prior), iterations, generations
    (245, 58, False, 500392788, (614661122, 956123187), 2, 0.8, 3, 3),
    # the chosen ancestor descends from a selfing: its stub is not the founder prior, so replacing it
    # with a new unobserved parent (an earlier fix) left a pedigree 1.07 log units off the scored one
    (362, 60, True, 138171697, (594499943, 43070780), 1, 0.5, 2, 5),
]


@pytest.mark.parametrize("case", TRIM_CASES, ids=["two-layers", "selfed-ancestor"])
def test_choice_survives_the_same_steps_trimming(tmp_path, case):
    """Found by tools/trace_soak.py: a kid can pick an unobserved ancestor of its own ex-parent (or
    the marriage of two); trimming the ex-parent then cascades to that ancestor in the same step.
    The reference links the kid to the freed slot; here the chosen individuals and marriage are
    kept through the trimming, and PEDFAC_CHECK (structure + scored == resulting likelihood) must
    hold through the run."""
    n, S, soft, dseed, seeds, unobs, ofrac, iters, gens = case
    d = synth.simulate(n, S, soft=soft, seed=dseed, observe_frac=0.8, n_gen=gens)
    rd = tmp_path / "t"
    frontend.write_run_dir(d, rd, seeds=seeds, max_unobs=unobs, max_age=2.0, observe_frac=ofrac)
    r = run(REF, rd, n_iter=iters, extra=("-s",), env={"PEDFAC_CHECK": "1", "PEDFAC_STATS": "1"})
    assert r.returncode == 0, r.stderr[-800:]


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 3])
def test_gpu_chains_batched_in_one_process_equal_single_chain_runs(tmp_path, threads):
    """PEDFAC_CHAINS=C (BASELINE.json's last configuration at test size): C chains on one GPU, their
    likelihood calls batched into one launch per round. Chain c must sample exactly what a
    single-chain run samples with the seeds (a + c, b + c); chain 0 keeps the unchanged file layout.
    PEDFAC_THREADS deals the chains to scheduler threads with an engine each (here 3 threads for 4
    chains: groups of 1, 1 and 2): same traces."""
    import os
    C = 4
    d = synth.simulate(160, 120, soft=True, seed=46)
    rd = tmp_path / "multi"
    frontend.write_run_dir(d, rd, seeds=(2590, 7579545))
    r = run(GPU, rd, n_iter=3, env={"PEDFAC_CHAINS": str(C), "PEDFAC_STATS": "1", "PEDFAC_THREADS": str(threads)})
    assert r.returncode == 0, r.stderr[-1500:]
    assert r.stdout.count(".. DONE") == 1 and "..2" in r.stdout          # one progress log: chain 0's
    assert sum(1 for l in r.stderr.splitlines() if l.startswith("pedigraph-stats")) == 3 * C
    for c in range(C):
        one = tmp_path / f"single{c}"
        frontend.write_run_dir(d, one, seeds=(2590 + c, 7579545 + c))
        r1 = run(GPU, one, n_iter=3)
        assert r1.returncode == 0, r1.stderr[-1500:]
        got = rd / "out" if c == 0 else rd / "out" / f"chain{c}"
        for name in ("ped.txt", "ll.txt", "historyNode.csv", "historyEdge.csv"):
            assert (got / name).read_text() == (one / "out" / name).read_text(), (c, name)
    assert (rd / "out" / "ped.txt").read_text() != (rd / "out" / "chain1" / "ped.txt").read_text()
