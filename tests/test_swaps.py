"""Swap proposals (`pf_eval_swaps` ≙ _ProposeSwapping@0x1000223d0). The oracle restates the
reference's holder-message arithmetic literally; the checks pin it to what it must mean — the log
partition value of the component that the rewiring of a picked swap (_PickAndUpdate@0x100017f5a)
produces — by exhaustive enumeration (small) and by a full sweep of the rewired pedigree (larger).
The CUDA engine computes exactly that quantity with an upward-only sweep and is compared with the
oracle's literal restatement."""
import numpy as np
import pytest

import bruteforce as bf
from checks import comp_members, local_factors
from scenarios import Scenario, random_forest, random_genotypes
from oracle.binding import OracleEngine
from pedfac_b200 import Pedigree


def make_oracle(S, ci, cm):
    return OracleEngine(S, ci, cm)


def make_gpu(S, ci, cm):
    from pedfac_b200 import Engine
    return Engine(S, ci, cm)

SWAP_KINDS = [(1, 0), (1, 1), (2, 0), (2, 1), (3, 1)]


def tiny_scenario(rng, S=4, sibs_x=1, sibs_z=1, extra_kid=True, only_child=False):
    """P,Q -> x (+ sibs); paz,maz -> z (+ sibs); x,z -> k (+ a second marriage of z)."""
    n = 0
    ped = Pedigree(24 + sibs_x + sibs_z, 12)
    ids = {}
    for name, founder in [("P", True), ("Q", True), ("paz", True), ("maz", False), ("x", False), ("z", False), ("k", False)]:
        ids[name] = n; ped.add_individual(n, founder); n += 1
    sx = []
    for _ in range(0 if only_child else sibs_x):
        sx.append(n); ped.add_individual(n, False); n += 1
    sz = []
    for _ in range(sibs_z):
        sz.append(n); ped.add_individual(n, False); n += 1
    ped.add_marriage(0, ids["P"], ids["Q"], [ids["x"]] + sx)
    ped.add_marriage(1, ids["paz"], ids["maz"], sz + [ids["z"]])
    ped.add_marriage(2, ids["x"], ids["z"], [ids["k"]])
    if extra_kid:   # something hanging off maz's other side, so that its message into marr_z is not just local
        g = n; ped.add_individual(n, True); n += 1
        gg = n; ped.add_individual(n, True); n += 1
        ped.add_marriage(3, g, gg, [ids["maz"]])
    eps = np.round(rng.uniform(0.005, 0.05, size=S), 3)
    afreq = np.round(rng.uniform(0.05, 0.5, size=S), 3)
    kinds = ("hard", "dec16", "f64", "none")
    geno = {i: random_genotypes(rng, S, kinds[rng.integers(len(kinds))]) for i in range(n)}
    return Scenario(ped, S, eps, afreq, geno), ids


def pruned_cases(sc):
    """For every swap site: (pruned scenario, the swap descriptors, the view of the three components)."""
    for x, z, mx, mz, m_of_x in sc.ped.swap_sites():
        ped = sc.ped.copy()
        pa, ma, prev = ped.detach_kid(x)
        ped.founder[x] = False                      # _PruneIndiv@0x1000141f0 clears isFounder
        labels, _ = ped.components()
        if labels[pa] == labels[x] or labels[ma] == labels[x]:
            continue
        swaps = [(k, mk, x, z, mz, mx, pa, ma, -1 if prev is None else prev) for k, mk in SWAP_KINDS]
        only = {int(labels[x]), int(labels[pa]), int(labels[ma])}
        yield Scenario(ped, sc.S, sc.eps, sc.afreq, sc.geno), swaps, only, labels


def rewired_logz(engine, psc, sw):
    """log Z (summed over loci) of the merged component after applying the swap, by a full sweep."""
    k, mk, x, z, mz, mx, pa, ma, prev = sw
    ped = psc.ped.copy()
    free = next(m for m in range(ped.cap_marr) if m not in ped.marriages)
    ped.apply_swap(k, bool(mk), x, z, mz, pa, ma, None if prev == -1 else prev, new_marr=free)
    labels, _ = ped.components()
    view, glob = ped.view(labels, only=[int(labels[x])])
    return float(engine.sweep(view)[0]), ped, labels


def check_swaps_vs_rewired(engine, sc, sweeper=None, rtol=1e-10, brute_max=0):
    """engine.eval_swaps on the pruned state == log Z of the rewired component (computed by
    `sweeper`, default the same engine, and by enumeration when the component is small)."""
    n = 0
    for psc, swaps, only, labels in pruned_cases(sc):
        psc.upload(engine)
        view_all, _ = psc.ped.view(labels)
        engine.sweep(view_all)
        view3, _ = psc.ped.view(labels, only=only)
        got = engine.eval_swaps(view3, swaps)
        for sw, g in zip(swaps, got):
            sw_engine = sweeper or engine
            if sweeper is not None:
                psc.upload(sweeper)
            want, rped, rl = rewired_logz(sw_engine, psc, sw)
            np.testing.assert_allclose(g, want, rtol=rtol, atol=1e-9, err_msg=f"swap {sw}")
            members = [int(i) for i in np.flatnonzero(rped.active) if rl[i] == rl[sw[2]]]
            if len(members) <= brute_max:
                rsc = Scenario(rped, psc.S, psc.eps, psc.afreq, psc.geno)
                mem, marrs = comp_members(rsc, rl, rl[sw[2]])
                Z = bf.component_quantities(mem, local_factors(rsc), [(mm.pa, mm.ma, mm.kids, mm.selfing) for _, mm in marrs])["Z"]
                np.testing.assert_allclose(g, np.log(Z).sum(), rtol=rtol, atol=1e-9, err_msg=f"swap {sw} (enumeration)")
            n += 1
        # re-sweeping the untouched pedigree afterwards must still give the same stubs: nothing was modified
    return n


@pytest.mark.parametrize("seed", range(6))
def test_oracle_swaps_match_enumeration(seed):
    rng = np.random.default_rng(900 + seed)
    sc, _ = tiny_scenario(rng, sibs_x=seed % 2, sibs_z=(seed // 2) % 2, extra_kid=seed % 3 != 0, only_child=seed % 2 == 0)
    eng = make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    assert check_swaps_vs_rewired(eng, sc, brute_max=10) >= 5


@pytest.mark.parametrize("seed", range(4))
def test_oracle_swaps_match_rewired_sweep(seed):
    rng = np.random.default_rng(950 + seed)
    sc = random_forest(rng, 60, 9, n_marr_target=34, max_kids=3, kinds=("hard", "dec16", "none"), max_comp=30)
    eng = make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    assert check_swaps_vs_rewired(eng, sc) >= 5


def test_swaps_with_very_large_sibships():
    """70 siblings next to z and 66 next to x: no fixed per-marriage edge limit anywhere."""
    rng = np.random.default_rng(31)
    sc, _ = tiny_scenario(rng, S=3, sibs_x=66, sibs_z=70)
    eng = make_oracle(sc.S, sc.ped.cap_indiv + 160, sc.ped.cap_marr)
    assert check_swaps_vs_rewired(eng, sc, rtol=1e-9) >= 5


def test_swap_argument_errors():
    rng = np.random.default_rng(5)
    sc, ids = tiny_scenario(rng)
    eng = make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    psc, swaps, only, labels = next(pruned_cases(sc))
    psc.upload(eng)
    eng.sweep(psc.ped.view(labels)[0])
    view3, _ = psc.ped.view(labels, only=only)
    from pedfac_b200.abi import PedfacError
    bad = list(swaps[0]); bad[0] = 4
    with pytest.raises(PedfacError):
        eng.eval_swaps(view3, [tuple(bad)])
    bad = list(swaps[0]); bad[4] = 2            # marr_z that does not hold z as a kid
    with pytest.raises(PedfacError):
        eng.eval_swaps(view3, [tuple(bad)])


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(4))
def test_gpu_swaps_match_oracle_and_enumeration(seed):
    rng = np.random.default_rng(900 + seed)
    sc, _ = tiny_scenario(rng, S=70, sibs_x=seed % 2, sibs_z=(seed // 2) % 2, only_child=seed % 2 == 0)
    gpu = make_gpu(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    ora = make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    assert check_swaps_vs_rewired(gpu, sc, brute_max=11, rtol=1e-11) >= 5
    for psc, swaps, only, labels in pruned_cases(sc):
        for e in (gpu, ora):
            psc.upload(e)
            e.sweep(psc.ped.view(labels)[0])
        view3, _ = psc.ped.view(labels, only=only)
        np.testing.assert_allclose(gpu.eval_swaps(view3, swaps), ora.eval_swaps(view3, swaps), rtol=1e-12, atol=1e-10)


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(3))
def test_gpu_swaps_match_oracle_on_forests(seed):
    rng = np.random.default_rng(950 + seed)
    sc = random_forest(rng, 200, 130, n_marr_target=120, max_kids=4, kinds=("hard", "dec16", "f32", "none"), max_comp=90)
    gpu = make_gpu(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    ora = make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    n = 0
    for psc, swaps, only, labels in pruned_cases(sc):
        for e in (gpu, ora):
            psc.upload(e)
            e.sweep(psc.ped.view(labels)[0])
        view3, _ = psc.ped.view(labels, only=only)
        before = gpu.get_stub(swaps[0][2])
        np.testing.assert_allclose(gpu.eval_swaps(view3, swaps), ora.eval_swaps(view3, swaps), rtol=1e-12, atol=1e-9)
        np.testing.assert_array_equal(gpu.get_stub(swaps[0][2]), before)      # nothing persisted was touched
        n += len(swaps)
        if n > 60:
            break
    assert n >= 5


@pytest.mark.gpu
def test_gpu_swaps_with_very_large_sibships():
    rng = np.random.default_rng(31)
    sc, _ = tiny_scenario(rng, S=40, sibs_x=66, sibs_z=70)
    gpu = make_gpu(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    ora = make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    for psc, swaps, only, labels in pruned_cases(sc):
        for e in (gpu, ora):
            psc.upload(e)
            e.sweep(psc.ped.view(labels)[0])
        view3, _ = psc.ped.view(labels, only=only)
        np.testing.assert_allclose(gpu.eval_swaps(view3, swaps), ora.eval_swaps(view3, swaps), rtol=1e-11, atol=1e-9)
        for i in np.flatnonzero(psc.ped.active)[::9]:
            np.testing.assert_allclose(gpu.get_stub(int(i)), ora.get_stub(int(i)), rtol=1e-12)


def check_deferred_swaps(eng, sc, max_cases=6):
    """pf_sweep_eval_swaps returns (ll_cc, lsum) of pf_sweep_eval at once and hands the swap values over
    later — with the next call's results, or on request — identical to pf_eval_swaps."""
    from pedfac_b200 import PF_PROP_TRIO
    n = 0
    for psc, swaps, only, labels in pruned_cases(sc):
        psc.upload(eng)
        view_all, _ = psc.ped.view(labels)
        view3, _ = psc.ped.view(labels, only=only)
        x = swaps[0][2]
        others = [int(i) for i in np.flatnonzero(psc.ped.active) if labels[i] != labels[x]]
        props = [(PF_PROP_TRIO, -1, -1, -1)] + [(PF_PROP_TRIO, c, -1, -1) for c in others[:9]]
        pairs = [(PF_PROP_TRIO, a, b, -1) for a, b in zip(others[:5], others[5:10])]
        ll0 = eng.sweep(view_all).copy()
        l0 = eng.eval(x, props).copy()
        want = eng.eval_swaps(view3, swaps).copy()
        # (a) the next call carries the deferred values
        ll1, l1 = eng.sweep_eval_swaps(view_all, x, props, view3, swaps)
        np.testing.assert_array_equal(ll1, ll0); np.testing.assert_array_equal(l1, l0)
        if pairs:
            lp = eng.eval(x, pairs)
            np.testing.assert_array_equal(lp, eng.eval(x, pairs))
        np.testing.assert_array_equal(eng.deferred_results(len(swaps)), want)
        # (b) nothing in between: delivered on request; no proposals in the first submission either
        ll2, l2 = eng.sweep_eval_swaps(view_all, x, [], view3, swaps)
        np.testing.assert_array_equal(ll2, ll0); assert len(l2) == 0
        np.testing.assert_array_equal(eng.deferred_results(len(swaps)), want)
        # (c) one set at a time, and the count has to match
        eng.sweep_eval_swaps(view_all, x, props, view3, swaps)
        with pytest.raises(RuntimeError):
            eng.sweep_eval_swaps(view_all, x, props, view3, swaps)
        with pytest.raises(RuntimeError):
            eng.deferred_results(len(swaps) + 1)
        np.testing.assert_array_equal(eng.deferred_results(len(swaps)), want)
        # (d) no swaps: plain pf_sweep_eval
        ll3, l3 = eng.sweep_eval_swaps(view_all, x, props, view3, [])
        np.testing.assert_array_equal(l3, l0)
        assert len(eng.deferred_results(0)) == 0
        n += 1
        if n >= max_cases:
            break
    return n


def test_oracle_deferred_swap_values():
    rng = np.random.default_rng(950)
    sc = random_forest(rng, 60, 9, n_marr_target=34, max_kids=3, kinds=("hard", "dec16", "none"), max_comp=30)
    assert check_deferred_swaps(make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr), sc) >= 2


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(2))
def test_gpu_deferred_swap_values(seed):
    rng = np.random.default_rng(960 + seed)
    sc = random_forest(rng, 200, 130, n_marr_target=120, max_kids=4, kinds=("hard", "dec16", "f32", "none"), max_comp=90)
    assert check_deferred_swaps(make_gpu(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr), sc) >= 2


def check_swap_batch(eng, sc, max_sets=5):
    """pf_eval_swaps_batch over several (view, swap list) sets == pf_eval_swaps set by set."""
    sets = []
    for psc, swaps, only, labels in pruned_cases(sc):
        sets.append((psc, swaps, only, labels))
        if len(sets) >= max_sets:
            break
    # all sets must be evaluated on one engine state: use the first pruned scenario's state for every set that
    # is valid on it (same pedigree object => same prune), else singletons of each
    psc, swaps, only, labels = sets[0]
    psc.upload(eng)
    eng.sweep(psc.ped.view(labels)[0])
    view3, _ = psc.ped.view(labels, only=only)
    one = eng.eval_swaps(view3, swaps)
    got = eng.eval_swaps_batch([0, 0, 0], [view3, view3, view3], [swaps, [], swaps[:2]])
    np.testing.assert_array_equal(got, np.concatenate([one, one[:2]]))
    assert len(eng.eval_swaps_batch([], [], [])) == 0
    return len(sets)


def test_oracle_swap_batch():
    rng = np.random.default_rng(950)
    sc = random_forest(rng, 60, 9, n_marr_target=34, max_kids=3, kinds=("hard", "dec16", "none"), max_comp=30)
    assert check_swap_batch(make_oracle(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr), sc) >= 1


@pytest.mark.gpu
def test_gpu_swap_batch():
    rng = np.random.default_rng(951)
    sc = random_forest(rng, 200, 130, n_marr_target=120, max_kids=4, kinds=("hard", "dec16", "f32", "none"), max_comp=90)
    assert check_swap_batch(make_gpu(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr), sc) >= 1
