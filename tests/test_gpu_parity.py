"""GPU parity tests (-m gpu): libpedfac_cuda through its C ABI against the CPU oracle on the same
seeded inputs, against exhaustive enumeration, against the golden known answers, and — at the
BASELINE.json sizes — through size-independent properties.

Tolerances (fp64 path; north_star: per-proposal log-likelihoods within 1e-9 relative):
  * stubs / prongs (raw message products):        1e-12 relative
  * data terms lsum and component LLs:            1e-12 relative (+1e-11 absolute near 0)
The GPU multiplies messages in a different association order and takes one log per chunk of
loci instead of one per locus, so bit equality is not expected; 1e-12 is ~4 decimal digits
tighter than the bar."""
import json
from pathlib import Path

import numpy as np
import pytest

import checks
from oracle.binding import OracleEngine
from pedfac_b200 import (PF_PROP_PRONG, PF_PROP_SELFING, PF_PROP_TRIO, Engine, GraphView, PedfacError,
                         Pedigree, compose_ll, synth)
from scenarios import random_forest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def test_golden_known_answers_on_gpu():
    k = json.loads((GOLDEN / "survey_appendix_e.json").read_text())
    e = Engine(1, 8, 4)
    e.set_locus_params([0.02], [0.3])
    ped = Pedigree(8, 4)
    for slot, (code, founder) in enumerate([(0, True), (2, True), (1, False)]):
        ped.add_individual(slot, founder)
        e.set_hard(slot, [code])
    view, _ = ped.view()
    llcc = e.sweep(view)
    np.testing.assert_allclose(llcc, k["kat1"]["ll_cc"], rtol=1e-13, atol=1e-15)
    lltot = float(np.sum(llcc))
    lsum = e.eval(2, [(PF_PROP_TRIO, 0, 1, -1), (PF_PROP_TRIO, -1, -1, -1), (PF_PROP_TRIO, 0, -1, -1)])
    ll = [compose_ll(lltot, lsum[0], llcc[2], True, llcc[0], True, llcc[1]),
          compose_ll(lltot, lsum[1], llcc[2], False, 0.0, False, 0.0),
          compose_ll(lltot, lsum[2], llcc[2], True, llcc[0], False, 0.0)]
    np.testing.assert_allclose(ll, [k["kat1"]["ll_pa_ma"], k["kat1"]["ll_new_new"], k["kat1"]["ll_pa_new"]], rtol=1e-13)
    # KAT 2: family + joining kid
    e = Engine(1, 8, 4)
    e.set_locus_params([0.02], [0.3])
    ped = Pedigree(8, 4)
    ped.add_individual(0, True); e.set_hard(0, [0])
    ped.add_individual(1, True); e.set_hard(1, [2])
    ped.add_individual(2, False); e.set_hard(2, [1])
    ped.add_individual(3, False); e.set_soft(3, [[0.1, 0.7, 0.2]], "f64")
    ped.add_individual(4, False); e.set_hard(4, [2])
    ped.add_marriage(0, 0, 1, [2, 3])
    view, _ = ped.view()
    llcc = e.sweep(view)
    np.testing.assert_allclose(llcc[0], k["kat2"]["ll_cc_family"], rtol=1e-13)
    for slot, name in [(0, "pa"), (1, "ma"), (2, "k1"), (3, "k2")]:
        np.testing.assert_allclose(e.get_stub(slot)[0], k["kat2"]["stub_" + name], rtol=1e-12)
    np.testing.assert_allclose(e.get_prong(0)[0], k["kat2"]["prong"], rtol=1e-12)
    np.testing.assert_allclose(e.eval(4, [(PF_PROP_PRONG, -1, -1, 0)])[0], k["kat2"]["contraction_prong_k3"], rtol=1e-13)


@pytest.mark.parametrize("seed", range(6))
def test_gpu_matches_bruteforce(seed):
    rng = np.random.default_rng(300 + seed)
    sc = random_forest(rng, n_indiv=int(rng.integers(3, 12)), S=3, max_kids=3,
                       p_selfing=0.2 if seed % 3 == 0 else 0.0, max_comp=8)
    e = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(e)
    checks.check_sweep_against_bruteforce(e, sc)
    sc2 = random_forest(rng, n_indiv=10, S=2, max_kids=2, n_marr_target=3, p_selfing=0.25, max_comp=3)
    e2 = Engine(sc2.S, sc2.ped.cap_indiv, sc2.ped.cap_marr)
    sc2.upload(e2)
    assert checks.check_eval_against_bruteforce(e2, sc2, rng) > 0


@pytest.mark.parametrize("seed,n_indiv,S,max_kids,p_self", [
    (0, 1, 1, 1, 0.0), (1, 2, 5, 1, 0.0), (2, 12, 31, 3, 0.0), (3, 40, 33, 4, 0.1), (4, 120, 64, 6, 0.05),
    (5, 300, 65, 5, 0.0), (6, 64, 257, 12, 0.0), (7, 500, 100, 3, 0.02)])
def test_gpu_matches_oracle_random_forests(seed, n_indiv, S, max_kids, p_self):
    """Ragged sizes: S around the 32-locus tile and 64-locus padding edges, deep and wide families."""
    rng = np.random.default_rng(400 + seed)
    sc = random_forest(rng, n_indiv=n_indiv, S=S, max_kids=max_kids, p_selfing=p_self)
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    o = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g); sc.upload(o)
    checks.check_engines_agree(g, o, sc, rng)


def test_gpu_incremental_sweeps_follow_graph_edits():
    """Prune a kid, sweep only the dirty components, re-attach elsewhere: the sequence of calls a
    Gibbs step makes (_MCMC_sampling@0x100019ec0) must track the oracle step by step."""
    rng = np.random.default_rng(77)
    sc = random_forest(rng, n_indiv=60, S=48, max_kids=4)
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    o = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g); sc.upload(o)
    checks.check_engines_agree(g, o, sc, rng)
    kids = sorted({k for mm in sc.ped.marriages.values() for k in mm.kids})
    for kid in kids[:12]:
        old_label = sc.ped.components()[0][kid]
        pa, ma, m = sc.ped.detach_kid(kid)
        labels, _ = sc.ped.components()
        dirty = {int(labels[kid]), int(labels[pa]), int(labels[ma])}
        checks.check_engines_agree(g, o, sc, rng, dirty=dirty, n_kids=3)
        if m is None:
            m = max(list(sc.ped.marriages) + [-1]) + 1
            sc.ped.add_marriage(m, pa, ma, [kid], selfing=(pa == ma))
        else:
            sc.ped.marriages[m].kids.append(kid)
        labels, _ = sc.ped.components()
        checks.check_engines_agree(g, o, sc, rng, dirty={int(labels[kid])}, n_kids=2)


def test_gpu_locus_shards_add_up_to_the_unsharded_engine():
    rng = np.random.default_rng(9)
    sc = random_forest(rng, n_indiv=50, S=150, max_kids=4)
    full = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    parts = [Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr, locus_range=r) for r in [(0, 37), (37, 101), (101, 150)]]
    view, _ = sc.ped.view()
    for e in [full] + parts:
        sc.upload(e)
    want = full.sweep(view)
    np.testing.assert_allclose(sum(p.sweep(view) for p in parts), want, rtol=1e-12)
    props = [(PF_PROP_TRIO, 1, 2, -1), (PF_PROP_TRIO, -1, 3, -1), (PF_PROP_SELFING, 4, -1, -1)]
    np.testing.assert_allclose(sum(p.eval(0, props) for p in parts), full.eval(0, props), rtol=1e-12)
    np.testing.assert_array_equal(np.concatenate([p.get_stub(5) for p in parts]), full.get_stub(5))


def test_gpu_multi_chain_batches_match_single_chain_oracles():
    rng = np.random.default_rng(11)
    scs = [random_forest(np.random.default_rng(500 + c), n_indiv=30, S=40, max_kids=3) for c in range(3)]
    # same genotype data for all chains, different graphs: reuse scenario 0's data
    base = scs[0]
    cap_i = max(s.ped.cap_indiv for s in scs); cap_m = max(s.ped.cap_marr for s in scs)
    g = Engine(base.S, cap_i, cap_m, n_chains=3)
    base.upload(g)
    os_ = []
    for sc in scs:
        o = OracleEngine(base.S, cap_i, cap_m)
        base.upload(o)
        os_.append(o)
    views = [sc.ped.view()[0] for sc in scs]
    got = g.sweep_batch([0, 1, 2], views)
    want = np.concatenate([o.sweep(v) for o, v in zip(os_, views)])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    kids, pb, props = [], [0], []
    for c in range(3):
        kids.append(int(rng.integers(0, 30)))
        for _ in range(int(rng.integers(1, 9))):
            props.append((PF_PROP_TRIO, int(rng.integers(-1, 30)), int(rng.integers(-1, 30)), -1))
        pb.append(len(props))
    got = g.eval_batch([0, 1, 2], kids, pb, props)
    want = np.concatenate([os_[c].eval(kids[c], props[pb[c]:pb[c + 1]]) for c in range(3)])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-11)
    for c in range(3):
        np.testing.assert_allclose(g.get_stub(7, chain=c), os_[c].get_stub(7), rtol=1e-12)


def test_gpu_zero_likelihoods_give_minus_infinity_like_the_reference():
    """eps = 0 makes Mendelian-impossible trios exactly 0: log(0) = -inf in the reference; the
    product-accumulating kernels must not turn that into a finite number or nan."""
    S = 40
    g = Engine(S, 4, 2); o = OracleEngine(S, 4, 2)
    for e in (g, o):
        e.set_locus_params(np.zeros(S), np.full(S, 0.25))
        e.set_hard(0, np.zeros(S, np.uint8)); e.set_hard(1, np.zeros(S, np.uint8)); e.set_hard(2, np.full(S, 2, np.uint8))
    ped = Pedigree(4, 2)
    for i in range(3):
        ped.add_individual(i, i < 2)
    view, _ = ped.view()
    np.testing.assert_allclose(g.sweep(view), o.sweep(view), rtol=1e-12)
    props = [(PF_PROP_TRIO, 0, 1, -1), (PF_PROP_TRIO, 0, -1, -1), (PF_PROP_TRIO, -1, -1, -1)]
    a, b = g.eval(2, props), o.eval(2, props)
    assert a[0] == -np.inf and b[0] == -np.inf and a[1] == -np.inf and b[1] == -np.inf
    np.testing.assert_allclose(a[2], b[2], rtol=1e-12)


def test_gpu_tiny_values_do_not_underflow_in_the_log_product():
    """Unnormalised stubs shrink geometrically with component size: feed values
    around 1e-290 and check the log path agrees with the oracle."""
    S = 70
    g = Engine(S, 4, 2); o = OracleEngine(S, 4, 2)
    lik = np.full((S, 3), 1e-290); lik[:, 1] = 3e-291
    for e in (g, o):
        e.set_locus_params(np.full(S, 0.01), np.full(S, 0.3))
        e.set_soft(0, lik, "f64"); e.set_hard(1, np.ones(S, np.uint8))
    ped = Pedigree(4, 2)
    ped.add_individual(0, True); ped.add_individual(1, False)
    view, _ = ped.view()
    np.testing.assert_allclose(g.sweep(view), o.sweep(view), rtol=1e-13)
    np.testing.assert_allclose(g.eval(1, [(PF_PROP_TRIO, 0, -1, -1)]), o.eval(1, [(PF_PROP_TRIO, 0, -1, -1)]), rtol=1e-13)


def test_gpu_rejects_bad_views_like_the_oracle():
    e = Engine(4, 6, 4)
    e.set_locus_params(np.full(4, 0.02), np.full(4, 0.3))
    ped = Pedigree(6, 4)
    for i in range(4):
        ped.add_individual(i, i < 2)
    ped.add_marriage(0, 0, 1, [2, 3])
    ped.add_marriage(1, 2, 3, [])
    view, _ = ped.view(labels=np.zeros(6, np.int32))
    with pytest.raises(PedfacError) as ei:
        e.sweep(view)
    assert ei.value.code == -5
    open_view = GraphView([0, 1], [0, 0], [1, 1], [0], [0], [1], [0], [0, 1], [2], 1)
    with pytest.raises(PedfacError) as ei:
        e.sweep(open_view)
    assert ei.value.code == -1
    with pytest.raises(PedfacError):
        e.set_hard(0, np.array([0, 1, 2, 7], np.uint8))
    empty = GraphView([], [], [], [], [], [], [], [0], [], 0)
    assert len(e.sweep(empty)) == 0 and len(e.eval(0, [])) == 0


def test_gpu_prefilter_matches_oracle():
    rng = np.random.default_rng(5)
    sc = random_forest(rng, n_indiv=90, S=130, n_marr_target=0, kinds=("hard", "dec16", "f32"))
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr); o = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g); sc.upload(o)
    view, _ = sc.ped.view()
    llcc = o.sweep(view); g.sweep(view)
    kids = list(range(0, 70)); cands = list(range(20, 90))
    role = rng.integers(0, 2, size=len(kids))
    bias = -llcc[cands]
    gs, gv = g.prefilter(kids, role, cands, bias, top_k=15)
    os_, ov = o.prefilter(kids, role, cands, bias, top_k=15)
    np.testing.assert_allclose(gv, ov, rtol=1e-12)
    # ranks may only differ where the oracle's scores tie to 1e-12
    diff = gs != os_
    assert diff.mean() < 0.02
    for k, t in zip(*np.nonzero(diff)):
        assert abs(ov[k, t] - gv[k, t]) <= 1e-9 * abs(ov[k, t])


@pytest.mark.parametrize("name,n,S,soft", [("A_1k_x_500_hard", 1000, 500, False), ("B_5k_x_2k_soft", 5000, 2000, True),
                                           ("C_20k_x_10k_hard", 20000, 10000, False)])
def test_gpu_full_size_properties(name, n, S, soft):
    """BASELINE.json sizes, where the oracle is too slow to be the checker: size-independent
    properties. (1) additivity over locus shards (a checksum of checksums); (2) the joined-family
    identity: attaching a kid to its true parents changes the total LL by exactly
    lsum - LLcc[kid] - LLcc[pa] - LLcc[ma], verified by re-sweeping the edited graph; (3) an
    oracle spot-check on a bounded sub-sample of individuals and loci."""
    d = synth.simulate(n, S, soft=soft, seed=46)
    cap_i, cap_m = d.n_total + 16, d.n_total
    full = Engine(S, cap_i, cap_m)
    synth.upload(full, d)
    ped = synth.truth_pedigree(d, cap_i, cap_m)
    labels, ncc = ped.components()
    view, glob = ped.view(labels)
    ll = full.sweep(view)
    assert np.isfinite(ll).all()
    halves = [Engine(S, cap_i, cap_m, locus_range=r) for r in [(0, S // 3), (S // 3, S)]]
    for h in halves:
        synth.upload(h, d)
    np.testing.assert_allclose(sum(h.sweep(view) for h in halves), ll, rtol=1e-12, atol=1e-9)
    # (2) detach 20 kids one at a time, evaluate "re-attach to the same marriage" and compare with a re-sweep
    rng = np.random.default_rng(3)
    kids_truth = rng.choice(np.flatnonzero(d.pa >= 0), size=20, replace=False)
    lab_pos = {int(g): j for j, g in enumerate(glob)}
    for kt in kids_truth:
        kid = int(d.slot[kt])
        before = ll[lab_pos[int(labels[kid])]]
        pa, ma, m = ped.detach_kid(kid)
        lab2, _ = ped.components()
        dirty = sorted({int(lab2[kid]), int(lab2[pa]), int(lab2[ma])})
        v2, g2 = ped.view(lab2, only=dirty)
        l2 = dict(zip([int(x) for x in g2], full.sweep(v2)))
        if m is None:
            prop = (PF_PROP_TRIO, pa, ma, -1)
        else:
            prop = (PF_PROP_PRONG, pa, ma, m)
        lsum = full.eval(kid, [prop])[0]
        parts = {int(lab2[kid]), int(lab2[pa]), int(lab2[ma])} if m is None else {int(lab2[kid]), int(lab2[pa])}
        # new component LL = lsum (the data term is the log partition value of the joined component)
        # restore the graph and re-sweep the same component
        if m is None:
            mnew = max(ped.marriages) + 1 if ped.marriages else 0
            ped.add_marriage(mnew, pa, ma, [kid])
        else:
            ped.marriages[m].kids.append(kid)
        lab3, _ = ped.components()
        v3, g3 = ped.view(lab3, only=[int(lab3[kid])])
        after = full.sweep(v3)[0]
        np.testing.assert_allclose(after, before, rtol=1e-12)
        np.testing.assert_allclose(lsum, after, rtol=1e-11)
        assert len(parts) >= 2
    # (3) oracle spot check: first 64 loci, the 40 individuals of a few small components
    Ss = 64
    o = OracleEngine(S, cap_i, cap_m, locus_range=(0, Ss))
    gsub = Engine(S, cap_i, cap_m, locus_range=(0, Ss))
    labels, _ = ped.components()
    sizes = np.bincount(labels[labels >= 0])
    pick = [int(c) for c in np.argsort(sizes)[::-1][:3]]
    vs, gl = ped.view(labels, only=pick)
    for e in (o, gsub):
        e.set_locus_params(d.eps, d.afreq)
    for i in np.flatnonzero(d.observed):
        s = int(d.slot[i])
        if s in set(vs.indiv.tolist()):
            for e in (o, gsub):
                if soft:
                    e.set_soft(s, d.soft_num[i], "dec16", 10000)
                else:
                    e.set_hard(s, d.codes[i])
    np.testing.assert_allclose(gsub.sweep(vs), o.sweep(vs), rtol=1e-12)
    for x in vs.indiv[:10]:
        np.testing.assert_allclose(gsub.get_stub(int(x)), o.get_stub(int(x)), rtol=1e-12)


def test_gpu_sweep_eval_equals_sweep_then_eval():
    """pf_sweep_eval (one submission, one host wait) returns exactly what pf_sweep followed by
    pf_eval returns, and agrees with the oracle."""
    rng = np.random.default_rng(4242)
    sc = random_forest(rng, n_indiv=150, S=97, max_kids=5, max_comp=40)
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    o = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g); sc.upload(o)
    labels, _ = sc.ped.components()
    view, glob = sc.ped.view(labels)
    has_parents = {k for mm in sc.ped.marriages.values() for k in mm.kids}
    kid = next(int(i) for i in np.flatnonzero(sc.ped.active) if int(i) not in has_parents)
    others = [int(i) for i in np.flatnonzero(sc.ped.active) if labels[i] != labels[kid]]
    props = [(PF_PROP_TRIO, -1, -1, -1)] + [(PF_PROP_TRIO, c, -1, -1) for c in others[:37]] + [(PF_PROP_SELFING, others[0], -1, -1)]
    ll1 = g.sweep(view); l1 = g.eval(kid, props)
    ll2, l2 = g.sweep_eval(view, kid, props)
    np.testing.assert_array_equal(ll1, ll2)
    np.testing.assert_array_equal(l1, l2)
    ll3, l3 = o.sweep_eval(view, kid, props)
    np.testing.assert_allclose(ll2, ll3, rtol=1e-12, atol=1e-11)
    np.testing.assert_allclose(l2, l3, rtol=1e-12, atol=1e-11)
    ll4, l4 = g.sweep_eval(view, kid, [])
    np.testing.assert_array_equal(ll4, ll1)
    assert len(l4) == 0


def test_gpu_throughput_shape_matches_oracle():
    """Many locus tiles x heavy groups with small genotype tiles (hard calls) make the plan builder
    pick the 8-warp / 2-blocks-per-SM sweep shape (98 KB of shared memory, more spilled rows): same
    results as the oracle, and as the latency shape on a shard that is small enough for it."""
    rng = np.random.default_rng(515)
    sc = random_forest(rng, n_indiv=420, S=6000, max_kids=6, kinds=("hard", "none"), max_comp=200)
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    o = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g); sc.upload(o)
    labels, _ = sc.ped.components()
    view, glob = sc.ped.view(labels)
    ll = g.sweep(view)
    st = g.plan_stats()
    assert st["smem_bytes"] <= 98 * 1024 and st["tiles"] == 188, st
    np.testing.assert_allclose(ll, o.sweep(view), rtol=1e-12, atol=1e-9)
    for i in np.flatnonzero(sc.ped.active)[::17]:
        np.testing.assert_allclose(g.get_stub(int(i)), o.get_stub(int(i)), rtol=1e-12, atol=0)
    # a 1 000-locus shard of the same data runs the latency shape; its log-likelihoods are the partial sums
    g2 = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr, locus_range=(0, 1000))
    sc.upload(g2)
    ll2 = g2.sweep(view)
    assert g2.plan_stats()["smem_bytes"] > 98 * 1024 or g2.plan_stats()["rows"] == g2.plan_stats()["rows_smem"]
    o2 = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr, locus_range=(0, 1000))
    sc.upload(o2)
    np.testing.assert_allclose(ll2, o2.sweep(view), rtol=1e-12, atol=1e-9)


def test_gpu_results_without_the_mailbox_are_identical(monkeypatch):
    """PEDFAC_NO_MAILBOX=1 delivers results by device-to-host copy + stream synchronise instead of
    the mapped pinned mailbox: bit-identical values."""
    rng = np.random.default_rng(616)
    sc = random_forest(rng, n_indiv=90, S=130, max_kids=4, max_comp=30)
    labels, _ = sc.ped.components()
    view, _ = sc.ped.view(labels)
    a = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    monkeypatch.setenv("PEDFAC_NO_MAILBOX", "1")
    b = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    monkeypatch.delenv("PEDFAC_NO_MAILBOX")
    sc.upload(a); sc.upload(b)
    np.testing.assert_array_equal(a.sweep(view), b.sweep(view))
    has_parents = {k for mm in sc.ped.marriages.values() for k in mm.kids}
    kid = next(int(i) for i in np.flatnonzero(sc.ped.active) if int(i) not in has_parents)
    others = [int(i) for i in np.flatnonzero(sc.ped.active) if labels[i] != labels[kid]]
    props = [(PF_PROP_TRIO, c, -1, -1) for c in others[:20]]
    np.testing.assert_array_equal(a.eval(kid, props), b.eval(kid, props))


def test_gpu_small_results_after_growing_large_ones():
    """Results beyond the mailbox size (4096 doubles) go through a pinned buffer that grows on
    demand. A stray free of the mailbox in that growth path once left later small results reading
    freed memory, and a second growth failed with "invalid argument" (found by
    tools/gpu_chain_soak.py in a degenerate chain state with huge proposal lists)."""
    rng = np.random.default_rng(4097)
    sc = random_forest(rng, n_indiv=90, S=70, max_kids=4, max_comp=30)
    labels, _ = sc.ped.components()
    view, _ = sc.ped.view(labels)
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    o = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g); sc.upload(o)
    np.testing.assert_allclose(g.sweep(view), o.sweep(view), rtol=1e-12, atol=1e-9)
    has_parents = {k for mm in sc.ped.marriages.values() for k in mm.kids}
    kid = next(int(i) for i in np.flatnonzero(sc.ped.active) if int(i) not in has_parents)
    others = [int(i) for i in np.flatnonzero(sc.ped.active) if labels[i] != labels[kid]]
    small = [(PF_PROP_TRIO, c, -1, -1) for c in others[:20]]
    want_small = o.eval(kid, small)
    for n_big in (5000, 12000, 30000):
        big = [(PF_PROP_TRIO, others[j % len(others)], -1, -1) for j in range(n_big)]
        got = g.eval(kid, big)
        np.testing.assert_allclose(got[:len(others)], o.eval(kid, big[:len(others)]), rtol=1e-12, atol=1e-9)
        np.testing.assert_array_equal(got[len(others):2 * len(others)], got[:len(others)])
        np.testing.assert_allclose(g.eval(kid, small), want_small, rtol=1e-12, atol=1e-9)
        np.testing.assert_allclose(g.sweep(view), o.sweep(view), rtol=1e-12, atol=1e-9)


def test_gpu_large_chain_batch_is_planned_in_parallel_and_matches_the_oracle():
    """24 chains with different graphs in one sweep_batch call: the views are planned by several
    host threads and the partial plans concatenated; every chain must match its own oracle, and a
    bad view anywhere in the batch must still be reported."""
    n_ch = 24
    scs = [random_forest(np.random.default_rng(700 + c), n_indiv=40, S=70, max_kids=4) for c in range(n_ch)]
    base = scs[0]
    cap_i = max(s.ped.cap_indiv for s in scs); cap_m = max(s.ped.cap_marr for s in scs)
    g = Engine(base.S, cap_i, cap_m, n_chains=n_ch)
    base.upload(g)
    o = OracleEngine(base.S, cap_i, cap_m)
    base.upload(o)
    views = [sc.ped.view()[0] for sc in scs]
    got = g.sweep_batch(list(range(n_ch)), views)
    want = np.concatenate([o.sweep(v) for v in views])
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-11)
    for c in (0, 11, 23):
        o.sweep(views[c])
        for i in np.flatnonzero(scs[c].ped.active)[::5]:
            np.testing.assert_allclose(g.get_stub(int(i), chain=c), o.get_stub(int(i)), rtol=1e-12)
    bad = scs[17].ped.copy()
    m0 = next(iter(bad.marriages)); mm = bad.marriages[m0]
    bad.add_marriage(max(bad.marriages) + 1, mm.pa, mm.ma, [mm.kids[0]])      # a second marriage through the same three: a loop
    views[17] = bad.view()[0]
    with pytest.raises(PedfacError):
        g.sweep_batch(list(range(n_ch)), views)


def test_gpu_prefilter_in_two_halves_over_locus_shards():
    """pf_prefilter_scores_dev on three locus shards, summed, then pf_prefilter_topk_dev = pf_prefilter
    on the unsharded engine (same slots; scores to 1e-12)."""
    import torch
    rng = np.random.default_rng(808)
    sc = random_forest(rng, n_indiv=70, S=300, max_kids=3, kinds=("hard", "dec16", "none"))
    full = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(full)
    view, _ = sc.ped.view()
    full.sweep(view)
    kids = np.arange(0, 40, dtype=np.int32); cands = np.arange(10, 70, dtype=np.int32)
    role = (kids % 2).astype(np.uint8); bias = rng.normal(size=len(cands))
    want_slot, want_score = full.prefilter(kids, role, cands, bias, top_k=15)
    total = torch.zeros(len(kids) * len(cands), dtype=torch.float64, device="cuda")
    shards = []
    for k, (a, b) in enumerate([(0, 100), (100, 150), (150, 300)]):
        e = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr, locus_range=(a, b))
        sc.upload(e); e.sweep(view)
        part = torch.zeros_like(total)
        e.set_stream(torch.cuda.current_stream().cuda_stream)
        e.prefilter_scores_dev(kids, role, cands, bias if k == 0 else None, part.data_ptr())
        torch.cuda.synchronize()
        total += part
        shards.append(e)
    got_slot, got_score = shards[1].prefilter_topk_dev(kids, cands, total.data_ptr(), top_k=15)
    np.testing.assert_array_equal(got_slot, want_slot)
    np.testing.assert_allclose(got_score, want_score, rtol=1e-12, atol=1e-10)
    with pytest.raises(PedfacError):
        shards[0].prefilter(kids, role, cands, bias, top_k=15)        # the one-call form needs all loci


def test_gpu_repeated_sweeps_are_bitwise_identical():
    """The sweep's transient rows are reused by level parity and its tasks run on whichever warp
    takes them: a hazard between tasks would show up as run-to-run differences. 40 repetitions of
    a deep, wide view (and of its swap evaluations) must agree to the bit."""
    rng = np.random.default_rng(919)
    sc = random_forest(rng, n_indiv=600, S=100, max_kids=8, kinds=("hard", "dec16", "none"), max_comp=400)
    g = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(g)
    labels, _ = sc.ped.components()
    view, _ = sc.ped.view(labels)
    ref_ll = g.sweep(view)
    probe = [int(i) for i in np.flatnonzero(sc.ped.active)[::7]]
    ref_stubs = [g.get_stub(i) for i in probe]
    marrs = sorted(sc.ped.marriages)[::5]
    ref_prongs = [g.get_prong(m) for m in marrs]
    assert g.plan_stats()["levels"] >= 8
    for _ in range(40):
        np.testing.assert_array_equal(g.sweep(view), ref_ll)
    for i, s in zip(probe, ref_stubs):
        np.testing.assert_array_equal(g.get_stub(i), s)
    for m, s in zip(marrs, ref_prongs):
        np.testing.assert_array_equal(g.get_prong(m), s)


def test_gpu_config_d_scale_batch_equals_single_chain():
    """BASELINE.json's last configuration at full size: 64 chains batched on one GPU over the config-B
    data (5 000 x 2 000 soft calls). All chains are given the same views and proposals, so every chain
    of every batched call must return, bit for bit, what chain 0 returns when it is called alone
    (the chains share the genotype arrays and differ only in their row block; proposal sums to 1e-12, their
    locus chunking depends on the batch size), through a few Gibbs-step
    passes of the bench trace; stubs and prongs of the last chain are compared with the first's."""
    import bench
    n_chains = 64
    d = synth.simulate(5000, 2000, soft=True, seed=46)
    tr = bench.Trace(d, 4)
    g = Engine(d.n_loci, tr.cap_i, tr.cap_m, n_chains=n_chains)
    synth.upload(g, d)
    chains = np.arange(n_chains, dtype=np.int32)
    ll_all = g.sweep_batch(chains, [tr.full_view] * n_chains).reshape(n_chains, -1)
    ll_one = g.sweep(tr.full_view, chain=0)
    assert np.isfinite(ll_one).all()
    assert (ll_all == ll_one[None, :]).all()
    for view, kid, props in tr.steps:
        got_cc = g.sweep_batch(chains, [view] * n_chains).reshape(n_chains, -1)
        pb = np.arange(n_chains + 1, dtype=np.int32) * len(props)
        got_ls = g.eval_batch(chains, np.full(n_chains, kid, dtype=np.int32), pb, np.tile(props, n_chains)).reshape(n_chains, -1)
        one_cc, one_ls = g.sweep_eval(view, kid, props, chain=0)
        assert (got_cc == one_cc[None, :]).all()
        assert (got_ls == got_ls[0][None, :]).all()            # every chain of the batch: the same bits
        # batch vs single call: the locus chunking of the eval grid depends on the batch size, so the sums
        # are associated differently
        np.testing.assert_allclose(got_ls[0], one_ls, rtol=1e-12, atol=1e-11)
        assert np.isfinite(one_ls).all()
    slot = int(tr.steps[-1][1])
    assert (g.get_stub(slot, chain=n_chains - 1) == g.get_stub(slot, chain=0)).all()
    g.sweep_batch(chains, [tr.final_view] * n_chains)


@pytest.mark.parametrize("fmt", ["hard", "dec16", "mixed"])
def test_gpu_sweep2_general_records_match_oracle(fmt):
    """The record shapes outside sweep2_kernel's common case, in each genotype format it is compiled
    for: an individual with seven marriages (operands with more than three rows), sibships of 5, 9 and
    14 kids (chunk pairs: KFOLD2 / KIDCHUNK2), a selfing marriage, kids with families of their own inside
    a large sibship. Log-likelihoods, every stub and every prong against the oracle."""
    rng = np.random.default_rng(5)
    S, n = 70, 120
    ped = Pedigree(n + 4, n)
    g, o = Engine(S, n + 4, n), OracleEngine(S, n + 4, n)
    eps, af = np.round(rng.uniform(0.005, 0.05, S), 3), np.round(rng.uniform(0.05, 0.5, S), 3)
    for e in (g, o):
        e.set_locus_params(eps, af)
    for i in range(n):
        ped.add_individual(i, founder=False)
        kind = fmt if fmt != "mixed" else ["hard", "dec16", "f32", "none"][i % 4]
        if rng.random() < 0.15:
            continue                                        # unobserved
        for e in (g, o):
            if kind == "hard":
                e.set_hard(i, np.random.default_rng(100 + i).integers(0, 4, S).astype(np.uint8))
            elif kind == "dec16":
                num = np.random.default_rng(100 + i).integers(0, 10001, (S, 3)).astype(np.uint16)
                e.set_soft(i, num, "dec16", 10000)
            elif kind == "f32":
                e.set_soft(i, np.random.default_rng(100 + i).random((S, 3)).astype(np.float32), "f32")
    nxt = iter(range(1, n))
    take = lambda k: [next(nxt) for _ in range(k)]
    m = 0
    hub = 0                                                  # one father, seven wives
    for k in (1, 2, 5, 1, 9, 3, 14):
        wife = take(1)[0]
        ped.add_marriage(m, hub, wife, take(k)); m += 1
    # two kids of the 9-sibship get families of their own; one marriage is a selfing
    kids9 = ped.marriages[4].kids
    ped.add_marriage(m, kids9[6], take(1)[0], take(5)); m += 1
    ped.add_marriage(m, take(1)[0], kids9[8], take(2)); m += 1
    selfer = take(1)[0]
    ped.add_marriage(m, selfer, selfer, take(3), selfing=True); m += 1
    with_parents = {k for mm in ped.marriages.values() for k in mm.kids}
    for i in range(n):
        ped.founder[i] = ped.active[i] and i not in with_parents
    view, _ = ped.view()
    lg, lo = g.sweep(view), o.sweep(view)
    np.testing.assert_allclose(lg, lo, rtol=1e-12, atol=1e-11)
    for i in view.indiv:
        np.testing.assert_allclose(g.get_stub(int(i)), o.get_stub(int(i)), rtol=1e-12, atol=1e-300)
    for mm in view.marr:
        np.testing.assert_allclose(g.get_prong(int(mm)), o.get_prong(int(mm)), rtol=1e-12, atol=1e-300)
    st = g.plan_stats()
    assert st["rows_smem"] == st["rows"]                    # planned by the second-generation builder, nothing spilled
