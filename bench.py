#!/usr/bin/env python
"""bench.py — the pedFac hot path on B200: MCMC proposals/s (and locus-trio evals/s).

One "step" = one replay of a fixed trace of T Gibbs-step passes of the hot path over synthetic
data of the named BASELINE.json shape: for every focal kid t the dirty components are
re-propagated (sum-product sweep) and the whole proposal list is evaluated; the trace ends with the
sweep that restores the starting graph, so a step is repeatable. See DESIGN.md §Measurement.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload A|B|C|D] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU): loci are sharded across ranks, every Gibbs
step's per-proposal / per-component partial sums are all-reduced with NCCL.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

from pedfac_b200 import PF_PROP_PRONG, PF_PROP_TRIO, synth  # noqa: E402

WORKLOADS = {
    # name: (individuals, loci, soft calls, chains per GPU, description)
    "A": (1000, 500, False, 1, "synthetic 1k individuals x 500 biallelic SNPs, hard calls"),
    "B": (5000, 2000, True, 1, "synthetic 5k individuals x 2k SNPs with 0,1,2 genotype likelihoods"),
    "C": (20000, 10000, False, 1, "synthetic 20k individuals x 10k SNPs, hard calls, loci sharded across GPUs"),
    "D": (5000, 2000, True, 64, "synthetic 5k x 2k soft, 64 independent chains batched per GPU"),
}
from pedfac_b200.sharding import shard_range  # noqa: E402


# ---------------------------------------------------------------------------------------------
# trace of Gibbs-step passes

class Trace:
    """Host-side description of T Gibbs-step passes on the truth pedigree of `data`."""

    def __init__(self, data, T, seed=2590):
        self.data = data
        rng = np.random.default_rng(seed)
        cap_i, cap_m = data.n_total + 64, data.n_total + 64
        self.cap_i, self.cap_m = cap_i, cap_m
        ped = synth.truth_pedigree(data, cap_i, cap_m)
        self.ped = ped
        labels0, _ = ped.components()
        self.full_view, _ = ped.view(labels0)
        with_parents = np.flatnonzero((data.pa >= 0) & data.observed)
        focal = rng.choice(with_parents, size=min(T, len(with_parents)), replace=False)
        self.steps = []          # fused schedule: (view, kid, props)
        self.ref_steps = []      # reference schedule: (view_prune, kid, props, view_pick)
        prev_kid = None
        for kt in focal:
            kid = int(data.slot[kt])
            pa, ma, m = ped.detach_kid(kid)
            labels, _ = ped.components()
            dirty_a = {int(labels[kid]), int(labels[pa]), int(labels[ma])}
            dirty = set(dirty_a)
            if prev_kid is not None:
                dirty.add(int(labels[prev_kid]))
            view, _ = ped.view(labels, only=dirty)
            view_a, _ = ped.view(labels, only=dirty_a)
            kid_slot, props = synth.gibbs_step_proposals(data, ped, labels, int(kt), rng)
            # restore ("pick" = the true parents)
            if m is None:
                m = max(ped.marriages) + 1
                ped.add_marriage(m, pa, ma, [kid])
            else:
                ped.marriages[m].kids.append(kid)
            labels_b, _ = ped.components()
            view_b, _ = ped.view(labels_b, only=[int(labels_b[kid])])
            self.steps.append((view, kid_slot, props))
            self.ref_steps.append((view_a, kid_slot, props, view_b))
            prev_kid = kid
        self.final_view = self.ref_steps[-1][3]
        self.n_props = int(sum(len(p) for _, _, p in self.steps))

    def algorithmic_bytes(self, S_local):
        """sweep S*(b_in*n_obs + 24*n_i + 24*n_m) + 8*n_cc; eval S*(24*D + 24*R) + 8*P."""
        d = self.data
        b_in = 6.0 if d.soft_num is not None else 0.25      # dec16 soft rows / 2-bit hard calls
        sweep = evalb = 0.0
        views = [v for v, _, _ in self.steps] + [self.final_view]
        for v in views:
            n_obs = int((v.indiv < d.n_obs).sum())
            sweep += S_local * (b_in * n_obs + 24.0 * len(v.indiv) + 24.0 * len(v.marr)) + 8.0 * v.n_cc
        for _, kid, props in self.steps:
            tr = props[props["kind"] != PF_PROP_PRONG]
            rows = set(tr["pa"].tolist()) | set(tr["ma"].tolist()) | {kid}
            rows.discard(-1)
            R = len(set(props["marr"][props["kind"] == PF_PROP_PRONG].tolist()))
            evalb += S_local * (24.0 * len(rows) + 24.0 * R) + 8.0 * len(props)
        return sweep, evalb


# ---------------------------------------------------------------------------------------------

class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.idx), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for k, n in enumerate(names) if any(len(r) > 2 + k and r[2 + k].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def measured_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture of this command
    (profiles/r02_traffic.json; config B), or None."""
    p = ROOT / "profiles" / "r02_traffic.json"
    try:
        return float(json.loads(p.read_text())[kernel]["dram_bytes_per_launch"])
    except (OSError, KeyError, ValueError):
        return None


def cpu_count():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference schedule (test infrastructure used as the timed baseline)

def cpu_sample_loci(data, trace, budget_s, passes):
    """How many loci the CPU arm can afford: one Gibbs-step pass is timed on 64 loci, then the
    largest multiple of 64 (at most all loci) whose `passes` full passes fit in `budget_s`."""
    from oracle.binding import OracleEngine
    S = data.n_loci
    S0 = min(S, 64)
    o = OracleEngine(S, trace.cap_i, trace.cap_m, locus_range=(0, S0))
    synth.upload(o, data)
    o.sweep(trace.full_view)
    t0 = time.perf_counter()
    k = min(6, len(trace.ref_steps))
    for view_a, kid, props, view_b in trace.ref_steps[:k]:
        o.sweep(view_a); o.eval(kid, props); o.sweep(view_b)
    per_locus_pass = (time.perf_counter() - t0) / k * len(trace.ref_steps) / S0
    fit = int(budget_s / max(passes, 1) / max(per_locus_pass, 1e-9))
    return int(min(S, max(64, fit // 64 * 64)))


def run_cpu_reference(data, trace, S_sample, budget_s, steps, warmup, opt="O2"):
    """Times the reference's own schedule — per Gibbs step: Propagating_msg over the dirty
    components, every proposal re-contracted over all loci, Propagating_msg again — on the first
    `S_sample` loci (cost is exactly linear in S) with the single thread the
    reference has. Returns proposals/s scaled to the full locus count, and a description."""
    from oracle.binding import OracleEngine
    S = data.n_loci
    o = OracleEngine(S, trace.cap_i, trace.cap_m, locus_range=(0, S_sample), opt=opt)
    synth.upload(o, data)
    o.sweep(trace.full_view)

    def one_pass(max_steps, deadline):
        n_p = n_s = 0
        for view_a, kid, props, view_b in trace.ref_steps[:max_steps]:
            o.sweep(view_a)
            o.eval(kid, props)
            o.sweep(view_b)
            n_p += len(props); n_s += 1
            if time.perf_counter() > deadline:
                break
        return n_p, n_s

    # size the sample: how many Gibbs steps fit in the budget
    t0 = time.perf_counter()
    one_pass(1, t0 + 1e9)
    t_one = max(time.perf_counter() - t0, 1e-6)
    per_pass = max(1, min(len(trace.ref_steps), int(budget_s / max(steps + warmup, 1) / t_one)))
    for _ in range(warmup):
        one_pass(per_pass, time.perf_counter() + 1e9)
    t0 = time.perf_counter()
    tot_p = tot_s = 0
    for _ in range(steps):
        a, b = one_pass(per_pass, time.perf_counter() + 1e9)
        tot_p += a; tot_s += b
    dt = time.perf_counter() - t0
    value = tot_p / dt * (S_sample / S)
    sample = (f"{tot_s} Gibbs-step passes ({tot_p} proposals) of the same trace on the first {S_sample} of {S} loci, "
              f"{dt:.1f} s, 1 thread, gcc -{opt} oracle port; proposals/s scaled by {S_sample}/{S} (cost is linear in loci)")
    return value, dt / max(steps, 1) * 1e3, sample


# ---------------------------------------------------------------------------------------------
# the reference itself: ngthomas/pedFac's own `exec/pedigraph` (Mach-O) inside oracle/_ref/pedigraph_macho

MACHO = ROOT / "oracle" / "_ref" / "pedigraph_macho"
REF_SAMPLE_INDIVIDUALS = 250


def reference_sample(wl, tmp, also_gpu_device=None, all_cores=False):
    """The reference's own sampler on a bounded sample of workload `wl`: the same generator, seed,
    encoding and ALL loci, but REF_SAMPLE_INDIVIDUALS individuals (the binary needs ~25 s for them at
    config B; its start-up prefilter alone is quadratic in individuals), one MCMC iteration from the
    all-singletons start through the unchanged CLI + files. The cost of a proposal is linear in loci and
    independent of the number of individuals, so proposals/s of the sample stands for the workload.
    Timed inside the loader (PEDFAC_REF_STATS: first to last Gibbs step of the iteration)."""
    from pedfac_b200 import frontend
    n_ind, S, soft, _, desc = WORKLOADS[wl]
    d = synth.simulate(REF_SAMPLE_INDIVIDUALS, S, soft=soft, seed=46)
    rd = Path(tmp) / "ref_sample"
    frontend.write_run_dir(d, rd)
    use_ref = MACHO.exists()
    binary = MACHO if use_ref else ROOT / "oracle" / "pedigraph_ref"
    env = dict(os.environ, PEDFAC_REF_STATS="1", PEDFAC_STATS="1")
    t0 = time.perf_counter()
    r = subprocess.run(["taskset", "-c", "0", str(binary), "-d", str(rd), "-n", "1", "-r", str(rd / "rand")], capture_output=True, text=True, env=env)
    wall = time.perf_counter() - t0
    it = None
    for line in r.stderr.splitlines():
        t = line.split()
        if line.startswith("ref-stats") or line.startswith("pedigraph-stats"):
            kv = dict(zip(t[1::2], t[2::2]))
            it = {"steps": int(kv["steps"]), "proposals": int(kv["proposals"]), "seconds": float(kv["seconds"])}
            break
    if it is None or it["seconds"] <= 0:
        return {"error": f"the reference sample run failed (rc {r.returncode}): {r.stderr[-200:]}"}
    out = {"value": it["proposals"] / it["seconds"], "unit": "proposals/s", "cores": 1, "kind": "reference" if use_ref else "port",
           "sample": (f"{'ngthomas/pedFac exec/pedigraph (the Mach-O binary itself, run by oracle/macho_run.c on glibc)' if use_ref else 'oracle/pedigraph_ref (C port; oracle/_ref/pedigraph_macho missing)'}"
                      f", 1 thread pinned with taskset -c 0: first MCMC iteration ({it['steps']} Gibbs steps, {it['proposals']} proposals, {it['seconds']:.1f} s) "
                      f"from the all-singletons start on {d.n_obs} observed of {d.n_total} synthetic individuals x all {S} loci of workload {wl} "
                      f"(same generator and seed; {n_ind} individuals would take the reference hours: its start-up prefilter is quadratic)"),
           "seconds": it["seconds"], "gibbs_steps": it["steps"], "proposals": it["proposals"], "process_wall_seconds": wall,
           "host_cores_available": cpu_count()}
    if all_cores and use_ref:
        # the best the CPU box could do: one independent chain of the reference per host core (different
        # seeds, same sample), all at once
        import shutil
        n = cpu_count()
        procs, t0 = [], time.perf_counter()
        for c in range(n):
            rc_dir = Path(tmp) / f"ref_sample_{c}"
            shutil.copytree(rd, rc_dir, ignore=shutil.ignore_patterns("out"))
            (rc_dir / "rand").write_text(f"{2590 + c} {7579545 + c} ")
            procs.append(subprocess.Popen(["taskset", "-c", str(c), str(binary), "-d", str(rc_dir), "-n", "1", "-r", str(rc_dir / "rand")],
                                          stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, env=env))
        tot_p, worst = 0, 0.0
        for pr in procs:
            _, err = pr.communicate()
            for line in err.splitlines():
                if line.startswith("ref-stats"):
                    kv = dict(zip(line.split()[1::2], line.split()[2::2]))
                    tot_p += int(kv["proposals"]); worst = max(worst, float(kv["seconds"]))
                    break
        if worst > 0:
            out["all_cores"] = {"cores": n, "value": tot_p / worst, "unit": "proposals/s",
                                "what": f"{n} independent chains of the reference binary at once, one per host core (seeds 2590+c, 7579545+c): "
                                        f"{tot_p} proposals in {worst:.1f} s (slowest chain), wall {time.perf_counter() - t0:.1f} s"}
    if also_gpu_device is not None:
        # the CUDA sampler on the very same run directory: same seed, so the same trace is required
        ref_ped = (rd / "out" / "ped.txt").read_text()
        import shutil
        shutil.rmtree(rd / "out")
        env2 = dict(os.environ, PEDFAC_STATS="1", PEDFAC_DEVICE=str(also_gpu_device))
        g = subprocess.run([str(ROOT / "pedfac_b200" / "bin" / "pedigraph"), "-d", str(rd), "-n", "1", "-r", str(rd / "rand")], capture_output=True, text=True, env=env2)
        st = parse_stats(g.stderr) if g.returncode == 0 else []
        if st:
            out["same_run_on_gpu"] = {"seconds": st[0]["seconds"], "proposals": st[0]["proposals"], "gibbs_steps": st[0]["steps"],
                                      "identical_trace": (rd / "out" / "ped.txt").read_text() == ref_ped,
                                      "speedup": it["seconds"] / st[0]["seconds"]}
    return out


# ---------------------------------------------------------------------------------------------
# the real sampler: pedfac_b200/bin/pedigraph (CUDA) vs oracle/pedigraph_ref (CPU) through the CLI

def parse_stats(stderr):
    out = []
    for line in stderr.splitlines():
        if line.startswith("pedigraph-stats"):
            t = line.split()
            kv = dict(zip(t[1::2], t[2::2]))
            out.append({"iter": int(kv["iter"]), "steps": int(kv["steps"]), "proposals": int(kv["proposals"]), "sweeps": int(kv["sweeps"]),
                        "swaps": int(kv.get("swaps", 0)), "swaps_picked": int(kv.get("swaps_picked", 0)), "seconds": float(kv["seconds"]),
                        "h2d_bytes": int(kv.get("h2d_bytes", 0)), "d2h_bytes": int(kv.get("d2h_bytes", 0)), "chain": int(kv.get("chain", 0))})
    return out


def run_chain(data, S, device, chains=1):
    """BASELINE.md's definition of the metric: one full MCMC iteration (every observed individual
    gets a Gibbs step, plus the unobserved parents it creates) timed after one warm-up iteration
    from the all-singletons start, through the unchanged CLI + file contract. chains > 1: that many
    independent chains in the one process (PEDFAC_CHAINS), batched per round; proposals are summed
    over the chains, the time is the wall time of the iteration (the chains run in lock step)."""
    import tempfile
    from pedfac_b200 import frontend
    n_iter = 3 if chains == 1 else 2
    with tempfile.TemporaryDirectory() as tmp:
        rd = Path(tmp) / "run"
        frontend.write_run_dir(data, rd)
        env = dict(os.environ, PEDFAC_STATS="1", PEDFAC_DEVICE=str(device))
        if chains > 1:
            env["PEDFAC_CHAINS"] = str(chains)
            env.setdefault("PEDFAC_THREADS", str(max(1, min(12, cpu_count() - 4, chains // 4))))
        t0 = time.perf_counter()
        r = subprocess.run([str(ROOT / "pedfac_b200" / "bin" / "pedigraph"), "-d", str(rd), "-n", str(n_iter), "-r", str(rd / "rand")],
                           capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": r.stderr[-300:]}
        allst = parse_stats(r.stderr)
        st = []
        for it in sorted({x["iter"] for x in allst}):
            rows = [x for x in allst if x["iter"] == it]
            agg = dict(rows[0])
            for k in ("steps", "proposals", "sweeps", "swaps", "swaps_picked"):
                agg[k] = sum(x[k] for x in rows)
            agg["seconds"] = max(x["seconds"] for x in rows)
            agg["chains"] = len(rows)
            st.append(agg)
        it = min(st[1:], key=lambda x: x["seconds"] / max(x["proposals"], 1))
        return {"what": f"pedigraph -n {n_iter} via CLI + files" + (f", {chains} chains in one process (PEDFAC_CHAINS) on {env.get('PEDFAC_THREADS')} scheduler threads (PEDFAC_THREADS)" if chains > 1 else "") +
                        "; iteration 1 = warm-up from all singletons, the fastest later iteration reported (all listed in timed_iterations)",
                "timed_iterations": st[1:], "swap_proposals": it["swaps"], "swaps_picked": it["swaps_picked"], "chains": chains,
                "gibbs_steps": it["steps"], "proposals": it["proposals"], "sweeps": it["sweeps"], "seconds": it["seconds"],
                "proposals_per_sec": it["proposals"] / it["seconds"], "gibbs_steps_per_sec": it["steps"] / it["seconds"],
                "trio_evals_per_sec": it["proposals"] * S / it["seconds"], "first_iteration": st[0],
                "h2d_bytes_per_gibbs_step": it["h2d_bytes"] / max(it["steps"], 1) * (chains if chains > 1 else 1), "d2h_bytes_per_gibbs_step": it["d2h_bytes"] / max(it["steps"], 1) * (chains if chains > 1 else 1),
                "process_wall_seconds": wall}


def workload_config(wl, world, S_local=None, n_individuals=None, n_observed=None, trace_steps=48, proposals_per_step=None):
    """The `config` object of the JSON line: identical for both arms of one (workload, GPU count)."""
    n_ind, S, soft, chains, desc = WORKLOADS[wl]
    sharded = world > 1 and wl != "D"
    return {"workload": f"{wl}: {desc}", "individuals": n_ind, "loci": S, "encoding": "soft dec16" if soft else "hard 2-bit",
            "loci_per_gpu": S if not sharded else -(-S // world), "chains_per_gpu": chains, "trace_steps": trace_steps,
            "l2": "256 MiB flush buffer written between timed steps",
            "parallelism": "locus-sharded + all-reduce per Gibbs step" if sharded else ("independent replicas" if (wl == "D" and world > 1) else "single GPU")}


# ---------------------------------------------------------------------------------------------

def replay(wl, args, rank, world, local_rank, trace_steps, steps, warmup, with_clocks):
    """Times the trace replay of workload `wl` on this rank's GPU (sharded over the ranks when
    world > 1): device-resident (`value`), through host buffers (`replay_e2e`), per-kernel event times."""
    import torch
    import torch.distributed as dist
    from pedfac_b200 import Engine
    n_ind, S, soft, chains, desc = WORKLOADS[wl]
    s0, s1 = shard_range(S, rank, world) if wl in ("C", "A", "B") and world > 1 else (0, S)
    if wl == "D" and world > 1:
        s0, s1 = 0, S            # replicas: every GPU runs its own 64 chains on all loci

    data = synth.simulate(n_ind, S, soft=soft, seed=46, locus_range=(s0, s1))
    trace = Trace(data, trace_steps)
    sharded = world > 1 and (s1 - s0) != S
    eng = Engine(S, trace.cap_i, trace.cap_m, locus_range=(s0, s1), n_chains=chains, device=local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    if sharded:
        eng.comm_init_from_torch(dist, rank, world)          # the library reduces over the shards itself from here on
        eng.comm_set_async(True)                             # device-resident replay: the sums overlap the next sweep
    synth.upload(eng, data)
    S_local = s1 - s0
    for c in range(chains):
        eng.sweep(trace.full_view, chain=c)

    # device output buffer: per Gibbs step [ll_cc | lsum] contiguous
    offs, tot = [], 0
    for view, kid, props in trace.steps:
        offs.append(tot); tot += view.n_cc + len(props)
    off_final = tot; tot += trace.final_view.n_cc
    d_out = torch.zeros(max(tot, 1) * chains, dtype=torch.float64, device="cuda")
    base_ptr = d_out.data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

    # multi-chain batching (config D): one sweep launch and one eval launch serve all chains
    if chains > 1:
        from pedfac_b200.abi import PfGraphView
        chain_ids = np.arange(chains, dtype=np.int32)
        batch_views = [(PfGraphView * chains)(*[v.struct for _ in range(chains)]) for v, _, _ in trace.steps]
        final_views = (PfGraphView * chains)(*[trace.final_view.struct for _ in range(chains)])
        batch_kids = [np.full(chains, k, dtype=np.int32) for _, k, _ in trace.steps]
        batch_pb = [np.arange(chains + 1, dtype=np.int32) * len(p) for _, _, p in trace.steps]
        batch_props = [np.ascontiguousarray(np.tile(p, chains)) for _, _, p in trace.steps]

    def view_bytes(v):
        return int(v.indiv.nbytes + v.indiv_cc.nbytes + v.indiv_founder.nbytes + v.marr.nbytes + v.marr_pa.nbytes +
                   v.marr_ma.nbytes + v.marr_selfing.nbytes + v.marr_kid_begin.nbytes + v.marr_kids.nbytes)

    def step_device_batched():
        for t, ((view, kid, props), o) in enumerate(zip(trace.steps, offs)):
            eng.sweep_batch_dev(chain_ids, batch_views[t], chains, base_ptr + 8 * chains * o)
            eng.eval_batch_dev(chain_ids, batch_kids[t], batch_pb[t], batch_props[t], base_ptr + 8 * chains * (o + view.n_cc))
        eng.sweep_batch_dev(chain_ids, final_views, chains, base_ptr + 8 * chains * off_final)

    def step_e2e_batched():
        nbytes_in = nbytes_out = 0
        for t, (view, kid, props) in enumerate(trace.steps):
            eng.sweep_batch(chain_ids, [view] * chains)
            eng.eval_batch(chain_ids, batch_kids[t], batch_pb[t], batch_props[t])
            nbytes_in += chains * view_bytes(view) + batch_props[t].nbytes
            nbytes_out += 8 * chains * (view.n_cc + len(props))
        eng.sweep_batch(chain_ids, [trace.final_view] * chains)
        return nbytes_in + chains * view_bytes(trace.final_view), nbytes_out + 8 * chains * trace.final_view.n_cc

    def step_device():
        """T Gibbs-step passes, results stay on the device. On a sharded engine every _dev call leaves
        the sum over all shards in its output (the library's own all-reduce, on the engine's stream)."""
        for c in range(chains):
            cb = c * tot
            for t, ((view, kid, props), o) in enumerate(zip(trace.steps, offs)):
                eng.sweep_eval_dev(view, kid, props, base_ptr + 8 * (cb + o), chain=c)
            eng.sweep_dev(trace.final_view, base_ptr + 8 * (cb + off_final), chain=c)
        if sharded:
            eng.comm_fence()                                 # the step ends when every cross-GPU sum has landed

    def step_e2e():
        """The call sequence pedigraph makes: host buffers in, host results out, every call."""
        nbytes_in = nbytes_out = 0
        for c in range(chains):
            for (view, kid, props), o in zip(trace.steps, offs):
                eng.sweep_eval(view, kid, props, chain=c)          # pf_sweep_eval: what pedigraph calls per Gibbs step
                nbytes_in += view_bytes(view) + props.nbytes
                nbytes_out += 8 * (view.n_cc + len(props))
            eng.sweep(trace.final_view, chain=c)
            nbytes_in += view_bytes(trace.final_view); nbytes_out += 8 * trace.final_view.n_cc
        return nbytes_in, nbytes_out

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, wall):
        """W warm-ups, then K timed steps; L2 flushed between steps (outside the timed spans)."""
        for _ in range(warmup):
            fn()
        barrier()
        total_ms, extra = 0.0, None
        for _ in range(steps):
            flush.fill_(1)
            barrier()
            if wall:
                t0 = time.perf_counter()
                extra = fn()
                torch.cuda.synchronize()
                total_ms += (time.perf_counter() - t0) * 1e3
            else:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                fn()
                e1.record(stream)
                torch.cuda.synchronize()
                total_ms += e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([total_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item()), extra

    clocks = ClockSampler(local_rank) if with_clocks else None
    if clocks:
        clocks.start()
    l0 = eng.launch_count()
    dev_fn, e2e_fn = (step_device_batched, step_e2e_batched) if chains > 1 else (step_device, step_e2e)
    dev_ms, _ = timed(dev_fn, wall=False)
    launches = (eng.launch_count() - l0) * steps // (steps + warmup)
    clk = clocks.stop() if clocks else None
    e2e_ms, io = timed(e2e_fn, wall=True)

    # per-kernel device time of the same step (separate pass: events around every launch)
    eng.profile(True)
    for _ in range(2):
        flush.fill_(1)
        dev_fn()
    k_ms, k_n = eng.profile_read()
    eng.profile(False)

    n_props_step = trace.n_props * chains
    # sharded: every rank evaluates every proposal on its loci -> the job's proposal count is the same
    replicas = world if (wl == "D" and world > 1) else 1
    value = n_props_step * replicas * steps / (dev_ms * 1e-3)
    e2e_value = n_props_step * replicas * steps / (e2e_ms * 1e-3)

    sweep_b, eval_b = trace.algorithmic_bytes(S_local)
    sweep_b *= chains; eval_b *= chains
    dom = "eval" if k_ms["eval"] >= k_ms["sweep"] else "sweep"
    alg = (eval_b if dom == "eval" else sweep_b) * 2          # two profiled steps
    peak, peak_src = measured_peaks()
    achieved = alg / (k_ms[dom] * 1e-3) / 1e9 if k_ms[dom] > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": f"{dom}_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ncu_traffic(f"{dom}_kernel") if wl == "B" else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg / max(k_n[dom], 1),
                "avg_launch_us": k_ms[dom] * 1e3 / max(k_n[dom], 1),
                "kernel_ms_per_step": {k: v / 2 for k, v in k_ms.items()},
                "kernel_launches_per_step": {k: v // 2 for k, v in k_n.items()},
                # both hot kernels against the same HBM peak (the other one for context)
                "per_kernel": {name: {"achieved_gbs": b * 2 / (k_ms[name] * 1e-3) / 1e9 if k_ms[name] > 0 else 0.0,
                                      "frac": (b * 2 / (k_ms[name] * 1e-3) / 1e9 / peak) if k_ms[name] > 0 else 0.0,
                                      "algorithmic_bytes_per_launch": b * 2 / max(k_n[name], 1),
                                      "avg_launch_us": k_ms[name] * 1e3 / max(k_n[name], 1)}
                               for name, b in (("sweep", sweep_b), ("eval", eval_b))},
                "note": "the sweep is bound by fp64 issue and instruction fetch of one block per (32-locus tile, component), not by HBM: "
                        "ncu (profiles/r02_sweep2_ncu_summary.txt) fp64 pipe 10.8 %, issue slots 16.5 %, DRAM < 1 % of peak"}
    del eng
    torch.cuda.empty_cache()
    return {"value": value, "ms_per_step": dev_ms / steps, "e2e_value": e2e_value, "e2e_ms_per_step": e2e_ms / steps, "io": io,
            "launches": int(launches), "clocks": clk, "roofline": roofline, "data": data, "trace": trace, "S": S, "S_local": S_local,
            "sharded": sharded, "replicas": replicas, "n_props_step": n_props_step, "shard": (s0, s1)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pedfac_b200", choices=["pedfac_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "A", "B", "C", "D"])
    ap.add_argument("--trace-steps", type=int, default=48, help="Gibbs-step passes per bench step")
    ap.add_argument("--cpu-budget", type=float, default=10.0, help="seconds of CPU work for the C-port baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="skip the real-sampler (pedigraph CLI) measurement")
    ap.add_argument("--no-anchor", action="store_true", help="skip the config-C line that anchors the multi-GPU scaling curve")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload if args.workload != "auto" else ("B" if args.gpus == 1 else "C")
    n_ind, S, soft, chains, desc = WORKLOADS[wl]
    metric, unit = "mcmc_proposals_per_sec", "proposals/s"
    config = workload_config(wl, max(world, args.gpus), trace_steps=args.trace_steps)

    if args.impl == "reference":
        if rank != 0:
            return 0
        import tempfile
        with tempfile.TemporaryDirectory() as tmp:
            ref = reference_sample(wl, tmp)
        if "error" in ref:
            print(json.dumps({"impl": "reference", "unavailable": ref["error"]}))
            return 0
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": ref["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ref["seconds"] * 1e3, "higher_is_better": True,
            "scaling": "weak" if wl == "D" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "note": "one run of the reference's own binary; a 'step' of this arm is its whole timed MCMC iteration on the bounded sample "
                    "(the binary cannot replay a trace, and repeating it K times would only repeat the same deterministic run)",
            "cpu_baseline": {k: ref[k] for k in ("value", "unit", "cores", "kind", "sample", "host_cores_available")},
            "e2e": {"value": ref["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "trio_evals_per_sec": ref["value"] * S,
        }))
        return 0

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pedfac_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    R = replay(wl, args, rank, world, local_rank, args.trace_steps, args.steps, args.warmup, with_clocks=True)

    if rank == 0:
        import tempfile
        data, trace = R["data"], R["trace"]
        out = {
            "metric": metric, "value": R["value"], "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": R["ms_per_step"], "higher_is_better": True,
            "scaling": "weak" if wl == "D" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config,
            "workload_detail": {"individuals": data.n_total, "observed": data.n_obs, "loci_on_this_gpu": R["S_local"],
                                "proposals_per_step": R["n_props_step"], "trace_steps": len(trace.steps),
                                "value_is": "replay of a fixed trace of Gibbs-step passes, inputs resident in HBM, CUDA events"},
            "trio_evals_per_sec": R["value"] * S,
            "replay_e2e": {"value": R["e2e_value"], "unit": unit, "h2d_bytes_per_step": R["io"][0], "d2h_bytes_per_step": R["io"][1],
                           "ms_per_step": R["e2e_ms_per_step"], "what": "the same trace through the host-buffer C ABI (pf_sweep_eval), wall clock"},
            "gpu_launches": R["launches"], "clocks": R["clocks"], "roofline": R["roofline"],
        }
        # e2e: the product's public entry point is the `pedigraph` executable (CLI + files, every Gibbs
        # step through the host-buffer C ABI) — BASELINE.md's timing window. Where that cannot run
        # (sharded / batched workloads) e2e is the trace replay through the host-buffer C ABI.
        out["e2e"] = dict(out["replay_e2e"])
        if world == 1 and not args.no_chain and wl in ("A", "B", "D"):
            full = data if R["shard"] == (0, S) else synth.simulate(n_ind, S, soft=soft, seed=46)
            ch = run_chain(full, S, local_rank, chains=chains)
            out["chain"] = ch
            if "error" not in ch:
                out["e2e"] = {"value": ch["proposals_per_sec"], "unit": unit, "h2d_bytes_per_step": ch["h2d_bytes_per_gibbs_step"],
                              "d2h_bytes_per_step": ch["d2h_bytes_per_gibbs_step"], "ms_per_step": ch["seconds"] * 1e3,
                              "what": "one full MCMC iteration of the real sampler (pedfac_b200/bin/pedigraph through CLI + files) after a warm-up iteration; "
                                      "a 'step' here is that iteration; bytes are per Gibbs step, counted by the library (pf_io_bytes)",
                              "trio_evals_per_sec": ch["trio_evals_per_sec"]}
        if world == 1 and not args.no_cpu_baseline:
            with tempfile.TemporaryDirectory() as tmp:
                ref = reference_sample(wl, tmp, also_gpu_device=local_rank if chains == 1 else None, all_cores=True)
            out["cpu_baseline"] = ref if "error" not in ref else None
            # the C port of the algorithm beside it (what round 1 used as the baseline): -O2 and -O0
            try:
                v, _, sample = run_cpu_reference(data, trace, cpu_sample_loci(data, trace, args.cpu_budget, 1), args.cpu_budget, 1, 0)
                out["cpu_port"] = {"value": v, "unit": unit, "cores": 1, "kind": "port", "sample": sample}
                # the shipped reference binary is an -O0 build: the same port at -O0, on a 256-locus sample
                v0, _, sample0 = run_cpu_reference(data, trace, min(S, 256), args.cpu_budget / 2, 1, 0, opt="O0")
                out["cpu_port_O0"] = {"value": v0, "unit": unit, "cores": 1, "kind": "port", "sample": sample0}
            except Exception as ex:                     # reported, not swallowed
                out["cpu_port"] = {"error": repr(ex)}
        else:
            out["cpu_baseline"] = None
    anchor = world == 1 and not args.no_anchor and wl == "B"
    if rank == 0 and not anchor:
        print(json.dumps(out), flush=True)
    if anchor:
        # the scaling curve's one-GPU anchor: config C (what --gpus 2/4/8 run) on this single GPU
        A = replay("C", args, rank, world, local_rank, min(args.trace_steps, 24), max(2, args.steps // 2), 2, with_clocks=False)
        out["scale_anchor"] = {"workload": WORKLOADS["C"][4], "config": workload_config("C", 1, trace_steps=min(args.trace_steps, 24)),
                               "value": A["value"], "e2e": A["e2e_value"], "ms_per_step": A["ms_per_step"], "unit": unit,
                               "proposals_per_step": A["n_props_step"], "roofline_per_kernel": A["roofline"]["per_kernel"]}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
