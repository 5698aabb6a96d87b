"""CPU tests: the oracle against (1) the _TPROB bytes extracted from the reference binary,
(2) the SURVEY.md Appendix E known answers, (3) exhaustive enumeration on random pedigrees."""
import json
from pathlib import Path

import numpy as np
import pytest

import bruteforce as bf
import checks
from oracle.binding import OracleEngine, tprob
from pedfac_b200 import PF_PROP_PRONG, PF_PROP_TRIO, GraphView, Pedigree, compose_ll
from scenarios import random_forest

GOLDEN = Path(__file__).parent / "golden"


def test_tprob_matches_reference_binary_bytes():
    g = json.loads((GOLDEN / "tprob.json").read_text())
    want = np.frombuffer(bytes.fromhex(g["hex"]), dtype="<f8")
    assert want.shape == (27,)
    np.testing.assert_array_equal(np.array(tprob()), want)
    np.testing.assert_array_equal(bf.T.ravel(), want)
    # rows of a conditional distribution: sum over the kid's state is 1 for every (gm, gp)
    np.testing.assert_array_equal(want.reshape(3, 3, 3).sum(axis=0), np.ones((3, 3)))


def _kat_engine():
    e = OracleEngine(1, 8, 4)
    e.set_locus_params([0.02], [0.3])
    return e


def test_survey_appendix_e_kat1():
    k = json.loads((GOLDEN / "survey_appendix_e.json").read_text())["kat1"]
    e = _kat_engine()
    ped = Pedigree(8, 4)
    for slot, (code, founder) in enumerate([(0, True), (2, True), (1, False)]):
        ped.add_individual(slot, founder)
        e.set_hard(slot, [code])
    view, glob = ped.view()
    llcc = e.sweep(view)
    np.testing.assert_allclose(llcc, k["ll_cc"], rtol=1e-13)
    lltot = float(np.sum(llcc))
    np.testing.assert_allclose(lltot, k["ll_tot"], rtol=1e-13)
    props = [(PF_PROP_TRIO, 0, 1, -1), (PF_PROP_TRIO, -1, -1, -1), (PF_PROP_TRIO, 0, -1, -1)]
    lsum = e.eval(2, props)
    ll = [compose_ll(lltot, lsum[0], llcc[2], True, llcc[0], True, llcc[1]),
          compose_ll(lltot, lsum[1], llcc[2], False, 0.0, False, 0.0),
          compose_ll(lltot, lsum[2], llcc[2], True, llcc[0], False, 0.0)]
    np.testing.assert_allclose(ll, [k["ll_pa_ma"], k["ll_new_new"], k["ll_pa_new"]], rtol=1e-13)
    np.testing.assert_allclose(lsum[1], k["contraction_new_new"], rtol=1e-13)


def test_survey_appendix_e_kat2():
    k = json.loads((GOLDEN / "survey_appendix_e.json").read_text())["kat2"]
    e = _kat_engine()
    ped = Pedigree(8, 4)
    ped.add_individual(0, True); e.set_hard(0, [0])
    ped.add_individual(1, True); e.set_hard(1, [2])
    ped.add_individual(2, False); e.set_hard(2, [1])
    ped.add_individual(3, False); e.set_soft(3, [[0.1, 0.7, 0.2]], "f64")
    ped.add_individual(4, False); e.set_hard(4, [2])
    ped.add_marriage(0, 0, 1, [2, 3])
    view, glob = ped.view()
    llcc = e.sweep(view)
    np.testing.assert_allclose(llcc[0], k["ll_cc_family"], rtol=1e-13)
    for slot, name in [(0, "pa"), (1, "ma"), (2, "k1"), (3, "k2")]:
        np.testing.assert_allclose(e.get_stub(slot)[0], k["stub_" + name], rtol=1e-12)
    np.testing.assert_allclose(e.get_prong(0)[0], k["prong"], rtol=1e-12)
    lsum = e.eval(4, [(PF_PROP_PRONG, -1, -1, 0)])
    np.testing.assert_allclose(lsum[0], k["contraction_prong_k3"], rtol=1e-13)


@pytest.mark.parametrize("seed", range(12))
def test_oracle_sweep_matches_bruteforce(seed):
    rng = np.random.default_rng(100 + seed)
    sc = random_forest(rng, n_indiv=int(rng.integers(3, 12)), S=3, max_kids=3,
                       p_selfing=0.2 if seed % 3 == 0 else 0.0, max_comp=8)
    e = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(e)
    checks.check_sweep_against_bruteforce(e, sc)


@pytest.mark.parametrize("seed", range(8))
def test_oracle_eval_matches_bruteforce(seed):
    rng = np.random.default_rng(200 + seed)
    sc = random_forest(rng, n_indiv=10, S=2, max_kids=2, n_marr_target=3,
                       p_selfing=0.25 if seed % 2 else 0.0, max_comp=3)
    e = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(e)
    assert checks.check_eval_against_bruteforce(e, sc, rng) > 0


def test_oracle_locus_shards_add_up():
    rng = np.random.default_rng(7)
    sc = random_forest(rng, n_indiv=14, S=11, max_kids=3)
    full = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    parts = [OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr, locus_range=r) for r in [(0, 4), (4, 11)]]
    view, _ = sc.ped.view()
    for e in [full] + parts:
        sc.upload(e)
    want = full.sweep(view)
    got = sum(p.sweep(view) for p in parts)
    np.testing.assert_allclose(got, want, rtol=1e-13)
    props = [(PF_PROP_TRIO, 1, 2, -1), (PF_PROP_TRIO, -1, 3, -1)]
    np.testing.assert_allclose(sum(p.eval(0, props) for p in parts), full.eval(0, props), rtol=1e-13)
    np.testing.assert_array_equal(np.concatenate([p.get_stub(5) for p in parts]), full.get_stub(5))


def test_oracle_rejects_loops_and_open_views():
    e = OracleEngine(1, 6, 4)
    e.set_locus_params([0.02], [0.3])
    ped = Pedigree(6, 4)
    for i in range(4):
        ped.add_individual(i, i < 2)
    ped.add_marriage(0, 0, 1, [2, 3])
    ped.add_marriage(1, 2, 3, [])     # sibs marry with no kid: edges 2-m0-3-m1-2 close a loop
    view, _ = ped.view(labels=np.zeros(6, np.int32))
    with pytest.raises(Exception) as ei:
        e.sweep(view)
    assert ei.value.code == -5
    open_view = GraphView([0, 1], [0, 0], [1, 1], [0], [0], [1], [0], [0, 1], [2], 1)
    with pytest.raises(Exception) as ei:
        e.sweep(open_view)
    assert ei.value.code == -1


def test_oracle_prefilter_orders_by_score():
    rng = np.random.default_rng(5)
    sc = random_forest(rng, n_indiv=12, S=6, n_marr_target=0, kinds=("hard", "dec16"))
    e = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(e)
    view, _ = sc.ped.view()
    llcc = e.sweep(view)
    kids, cands = [0, 1, 2], list(range(12))
    bias = -llcc[[int(c) for c in cands]]
    slots, scores = e.prefilter(kids, [1, 0, 1], cands, bias, top_k=5)
    for k, kid in enumerate(kids):
        allv = []
        for c in cands:
            if c == kid:
                continue
            pr = (PF_PROP_TRIO, c, -1, -1) if [1, 0, 1][k] else (PF_PROP_TRIO, -1, c, -1)
            allv.append((e.eval(kid, [pr])[0] + bias[c], c))
        allv.sort(key=lambda t: (-t[0], t[1]))
        np.testing.assert_array_equal(slots[k], [c for _, c in allv[:5]])
        np.testing.assert_allclose(scores[k], [v for v, _ in allv[:5]], rtol=1e-14)


def test_many_individuals_few_loci():
    """The per-sweep arena must be sized by the pedigree as well as by the loci (a 5000-individual
    sweep over 16 loci once ran out of adjacency space and crashed)."""
    rng = np.random.default_rng(77)
    sc = random_forest(rng, 400, 2, n_marr_target=150, max_kids=5, kinds=("hard", "none"), max_comp=60)
    eng = OracleEngine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(eng)
    view, glob = sc.ped.view()
    ll = eng.sweep(view)
    assert np.isfinite(ll).all() and len(ll) == len(glob)
