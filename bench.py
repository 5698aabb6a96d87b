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
    (profiles/r01_traffic.json; config B), or None."""
    p = ROOT / "profiles" / "r01_traffic.json"
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


def run_cpu_reference(data, trace, S_sample, budget_s, steps, warmup):
    """Times the reference's own schedule — per Gibbs step: Propagating_msg over the dirty
    components, every proposal re-contracted over all loci, Propagating_msg again — on the first
    `S_sample` loci (cost is exactly linear in S) with the single thread the
    reference has. Returns proposals/s scaled to the full locus count, and a description."""
    from oracle.binding import OracleEngine
    S = data.n_loci
    o = OracleEngine(S, trace.cap_i, trace.cap_m, locus_range=(0, S_sample))
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
              f"{dt:.1f} s, 1 thread, gcc -O2 oracle port; proposals/s scaled by {S_sample}/{S} (cost is linear in loci)")
    return value, dt / max(steps, 1) * 1e3, sample


# ---------------------------------------------------------------------------------------------
# the real sampler: pedfac_b200/bin/pedigraph (CUDA) vs oracle/pedigraph_ref (CPU) through the CLI

def parse_stats(stderr):
    out = []
    for line in stderr.splitlines():
        if line.startswith("pedigraph-stats"):
            t = line.split()
            kv = dict(zip(t[1::2], t[2::2]))
            out.append({"iter": int(kv["iter"]), "steps": int(kv["steps"]), "proposals": int(kv["proposals"]), "sweeps": int(kv["sweeps"]),
                        "swaps": int(kv.get("swaps", 0)), "swaps_picked": int(kv.get("swaps_picked", 0)), "seconds": float(kv["seconds"])})
    return out


def run_chain(data, S, device, cpu_budget_s, with_cpu=True):
    """BASELINE.md's definition of the metric: one full MCMC iteration (every observed individual
    gets a Gibbs step, plus the unobserved parents it creates) timed after one warm-up iteration
    from the all-singletons start, through the unchanged CLI + file contract. The CPU arm runs the
    same chain logic on the oracle for the first K Gibbs steps on a 64-locus subset (linear in loci)."""
    import tempfile
    from pedfac_b200 import frontend
    res = {}
    with tempfile.TemporaryDirectory() as tmp:
        rd = Path(tmp) / "run"
        frontend.write_run_dir(data, rd)
        env = dict(os.environ, PEDFAC_STATS="1", PEDFAC_DEVICE=str(device))
        t0 = time.perf_counter()
        r = subprocess.run([str(ROOT / "pedfac_b200" / "bin" / "pedigraph"), "-d", str(rd), "-n", "3", "-r", str(rd / "rand")],
                           capture_output=True, text=True, env=env)
        wall = time.perf_counter() - t0
        if r.returncode != 0:
            return {"error": r.stderr[-300:]}
        st = parse_stats(r.stderr)
        it = min(st[1:], key=lambda x: x["seconds"] / max(x["proposals"], 1))
        res = {"what": "pedigraph -n 3 via CLI + files; iteration 1 = warm-up from all singletons, the faster of iterations 2 and 3 reported (both listed in timed_iterations)",
               "timed_iterations": st[1:], "swap_proposals": it["swaps"], "swaps_picked": it["swaps_picked"],
               "gibbs_steps": it["steps"], "proposals": it["proposals"], "sweeps": it["sweeps"], "seconds": it["seconds"],
               "proposals_per_sec": it["proposals"] / it["seconds"], "gibbs_steps_per_sec": it["steps"] / it["seconds"],
               "trio_evals_per_sec": it["proposals"] * S / it["seconds"], "first_iteration": st[0],
               "process_wall_seconds": wall}
        if with_cpu:
            # CPU arm: same sampler source on the oracle, first iteration in full (warm-up) then the
            # first K Gibbs steps of iteration 2, on a 64-locus subset; cost is linear in loci
            Ss = min(S, 64)
            rc_dir = Path(tmp) / "run_cpu"
            frontend.write_run_dir(data, rc_dir, loci=Ss)
            K = 120
            env2 = dict(os.environ, PEDFAC_STATS="1", PEDFAC_MAX_STEPS=str(st[0]["steps"] + K))
            try:
                r2 = subprocess.run([str(ROOT / "oracle" / "pedigraph_ref"), "-d", str(rc_dir), "-n", "2", "-r", str(rc_dir / "rand")],
                                    capture_output=True, text=True, env=env2, timeout=max(120.0, cpu_budget_s * 10))
                cs = parse_stats(r2.stderr) if r2.returncode == 0 else []
            except subprocess.TimeoutExpired:
                cs = []
            if len(cs) >= 2 and cs[1]["steps"] > 0:
                c = cs[1]
                res["cpu_iteration2_sample"] = {
                    "what": f"oracle-backed pedigraph_ref (same chain source, CPU likelihoods, 1 thread): iteration 1 in full, then the "
                            f"first {c['steps']} Gibbs steps of iteration 2, on the first {Ss} of {S} loci; proposals/s scaled by {Ss}/{S}",
                    "proposals_per_sec": c["proposals"] / c["seconds"] * (Ss / S), "seconds": c["seconds"],
                    "proposals": c["proposals"], "gibbs_steps": c["steps"],
                    "first_iteration_proposals_per_sec": cs[0]["proposals"] / cs[0]["seconds"] * (Ss / S)}
                res["speedup_vs_cpu_iteration2_sample"] = res["proposals_per_sec"] / res["cpu_iteration2_sample"]["proposals_per_sec"]
    return res


# ---------------------------------------------------------------------------------------------

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="pedfac_b200", choices=["pedfac_b200", "reference"])
    ap.add_argument("--workload", default="auto", choices=["auto", "A", "B", "C", "D"])
    ap.add_argument("--trace-steps", type=int, default=48, help="Gibbs-step passes per bench step")
    ap.add_argument("--cpu-budget", type=float, default=20.0, help="seconds of CPU work for the baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-chain", action="store_true", help="skip the real-sampler (pedigraph CLI) measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    wl = args.workload if args.workload != "auto" else ("B" if args.gpus == 1 else "C")
    n_ind, S, soft, chains, desc = WORKLOADS[wl]
    metric, unit = "mcmc_proposals_per_sec", "proposals/s"

    if args.impl == "reference":
        if rank != 0:
            return 0
        data = synth.simulate(n_ind, S, soft=soft, seed=46)
        trace = Trace(data, args.trace_steps)
        S_sample = cpu_sample_loci(data, trace, args.cpu_budget * 3, args.steps + args.warmup)
        value, ms, sample = run_cpu_reference(data, trace, S_sample, args.cpu_budget * 3, args.steps, args.warmup)
        print(json.dumps({
            "impl": "reference", "metric": metric, "value": value, "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak" if wl == "D" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl}: {desc}", "trace_steps": args.trace_steps,
                       "note": "ngthomas/pedFac ships pedigraph only as a macOS binary (no source): this arm times the C "
                               "restatement of its algorithm (oracle/pedfac_oracle.c) with the reference's own schedule"},
            "cpu_baseline": {"value": value, "unit": unit, "cores": 1, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "trio_evals_per_sec": value * S,
        }))
        return 0

    import torch
    import torch.distributed as dist
    from pedfac_b200 import Engine

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: pedfac_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    s0, s1 = shard_range(S, rank, world) if wl in ("C", "A", "B") and world > 1 else (0, S)
    if wl == "D" and world > 1:
        s0, s1 = 0, S            # replicas: every GPU runs its own 64 chains on all loci

    data = synth.simulate(n_ind, S, soft=soft, seed=46, locus_range=(s0, s1))
    trace = Trace(data, args.trace_steps)
    eng = Engine(S, trace.cap_i, trace.cap_m, locus_range=(s0, s1), n_chains=chains, device=local_rank)
    stream = torch.cuda.current_stream()
    eng.set_stream(stream.cuda_stream)
    synth.upload(eng, data)
    S_local = s1 - s0
    for c in range(chains):
        eng.sweep(trace.full_view, chain=c)

    # device output buffer: per Gibbs step [ll_cc | lsum] contiguous so one all-reduce carries both
    offs, tot = [], 0
    for view, kid, props in trace.steps:
        offs.append(tot); tot += view.n_cc + len(props)
    off_final = tot; tot += trace.final_view.n_cc
    d_out = torch.zeros(max(tot, 1) * chains, dtype=torch.float64, device="cuda")
    base_ptr = d_out.data_ptr()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sharded = world > 1 and (s1 - s0) != S

    # multi-chain batching (config D): one sweep launch and one eval launch serve all chains
    if chains > 1:
        import ctypes as C
        from pedfac_b200.abi import PfGraphView
        chain_ids = np.arange(chains, dtype=np.int32)
        batch_views = [(PfGraphView * chains)(*[v.struct for _ in range(chains)]) for v, _, _ in trace.steps]
        final_views = (PfGraphView * chains)(*[trace.final_view.struct for _ in range(chains)])
        batch_kids = [np.full(chains, k, dtype=np.int32) for _, k, _ in trace.steps]
        batch_pb = [np.arange(chains + 1, dtype=np.int32) * len(p) for _, _, p in trace.steps]
        batch_props = [np.ascontiguousarray(np.tile(p, chains)) for _, _, p in trace.steps]

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

    # views of the output vector that each Gibbs step all-reduces (sliced once, outside the timed loop)
    segs = [[d_out[c * tot + o: c * tot + o + view.n_cc + len(props)] for (view, kid, props), o in zip(trace.steps, offs)]
            for c in range(chains)] if sharded else None

    def step_device():
        """T Gibbs-step passes, results stay on the device (plus the NCCL all-reduce when sharded)."""
        for c in range(chains):
            cb = c * tot
            for t, ((view, kid, props), o) in enumerate(zip(trace.steps, offs)):
                eng.sweep_dev(view, base_ptr + 8 * (cb + o), chain=c)
                eng.eval_dev(kid, props, base_ptr + 8 * (cb + o + view.n_cc), chain=c)
                if sharded:
                    dist.all_reduce(segs[c][t])
            eng.sweep_dev(trace.final_view, base_ptr + 8 * (cb + off_final), chain=c)

    h_pinned = torch.empty(max(tot, 1), dtype=torch.float64).pin_memory()

    def step_e2e():
        """The call sequence pedigraph makes: host buffers in, host results out, every call."""
        nbytes_in = nbytes_out = 0
        for c in range(chains):
            for (view, kid, props), o in zip(trace.steps, offs):
                if sharded:
                    eng.sweep_dev(view, base_ptr + 8 * o, chain=c)
                    eng.eval_dev(kid, props, base_ptr + 8 * (o + view.n_cc), chain=c)
                    seg = d_out[o: o + view.n_cc + len(props)]
                    dist.all_reduce(seg)
                    h_pinned[o: o + view.n_cc + len(props)].copy_(seg, non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                else:
                    eng.sweep_eval(view, kid, props, chain=c)      # pf_sweep_eval: what pedigraph calls per Gibbs step
                nbytes_in += view_bytes(view) + props.nbytes
                nbytes_out += 8 * (view.n_cc + len(props))
            eng.sweep(trace.final_view, chain=c)
            nbytes_in += view_bytes(trace.final_view); nbytes_out += 8 * trace.final_view.n_cc
        return nbytes_in, nbytes_out

    def view_bytes(v):
        return int(v.indiv.nbytes + v.indiv_cc.nbytes + v.indiv_founder.nbytes + v.marr.nbytes + v.marr_pa.nbytes +
                   v.marr_ma.nbytes + v.marr_selfing.nbytes + v.marr_kid_begin.nbytes + v.marr_kids.nbytes)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, wall):
        """W warm-ups, then K timed steps; L2 flushed between steps (outside the timed spans)."""
        for _ in range(args.warmup):
            fn()
        barrier()
        total_ms, extra = 0.0, None
        for _ in range(args.steps):
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

    clocks = ClockSampler(local_rank)
    clocks.start()
    l0 = eng.launch_count()
    dev_fn, e2e_fn = (step_device_batched, step_e2e_batched) if chains > 1 else (step_device, step_e2e)
    dev_ms, _ = timed(dev_fn, wall=False)
    launches = (eng.launch_count() - l0) * args.steps // (args.steps + args.warmup)
    clk = clocks.stop()
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
    value = n_props_step * replicas * args.steps / (dev_ms * 1e-3)
    e2e_value = n_props_step * replicas * args.steps / (e2e_ms * 1e-3)

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
                               for name, b in (("sweep", sweep_b), ("eval", eval_b))}}

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            v, _, sample = run_cpu_reference(data, trace, cpu_sample_loci(data, trace, args.cpu_budget, 1), args.cpu_budget, 1, 0)
            cpu = {"value": v, "unit": unit, "cores": 1, "kind": "port", "sample": sample, "host_cores_available": cpu_count()}
        out = {
            "metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak" if wl == "D" else "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"{wl}: {desc}", "individuals": data.n_total, "observed": data.n_obs, "loci": S,
                       "loci_per_gpu": S_local, "chains_per_gpu": chains, "trace_steps": len(trace.steps),
                       "proposals_per_step": n_props_step, "l2": "256 MiB flush buffer written between timed steps",
                       "parallelism": "locus-sharded + NCCL all-reduce per Gibbs step" if sharded else
                                      ("independent replicas" if replicas > 1 else "single GPU")},
            "trio_evals_per_sec": value * S,
            "e2e": {"value": e2e_value, "unit": unit, "h2d_bytes_per_step": io[0], "d2h_bytes_per_step": io[1],
                    "ms_per_step": e2e_ms / args.steps, "trio_evals_per_sec": e2e_value * S},
            "gpu_launches": int(launches),
            "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
        }
        if world == 1 and chains == 1 and not args.no_chain and wl in ("A", "B"):
            del eng
            torch.cuda.empty_cache()
            full = data if (s0, s1) == (0, S) else synth.simulate(n_ind, S, soft=soft, seed=46)
            out["chain"] = run_chain(full, S, local_rank, args.cpu_budget, with_cpu=not args.no_cpu_baseline)
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
