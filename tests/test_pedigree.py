import numpy as np

from pedfac_b200 import Pedigree
from pedfac_b200 import synth


def test_components_and_views():
    ped = Pedigree(10, 5)
    for i in range(8):
        ped.add_individual(i, i < 2)
    ped.add_marriage(0, 0, 1, [2, 3])
    ped.add_marriage(1, 2, 4, [5])
    labels, n = ped.components()
    assert n == 3
    assert len({labels[i] for i in [0, 1, 2, 3, 4, 5]}) == 1
    assert labels[6] != labels[7] and labels[8] == -1
    view, glob = ped.view(labels, only=[labels[0]])
    assert sorted(view.indiv.tolist()) == [0, 1, 2, 3, 4, 5] and view.n_cc == 1 and list(view.marr) == [0, 1]
    assert ped.detach_kid(5) == (2, 4, None) and 1 not in ped.marriages
    assert ped.detach_kid(3) == (0, 1, 0) and ped.marriages[0].kids == [2]
    labels, n = ped.components()
    assert n == 6


def test_synthetic_data_shapes_and_frequencies():
    d = synth.simulate(400, 64, soft=False, seed=46)
    assert d.codes.shape == (d.n_total, 64) and d.codes.max() <= 3
    assert (d.afreq <= 0.5).all() and (d.afreq >= 0.001).all()
    assert 0.7 < d.n_obs / d.n_total < 0.9
    kids = np.flatnonzero(d.pa >= 0)
    assert (d.gen[d.pa[kids]] == d.gen[kids] + 1).all() and (d.sex[d.pa[kids]] == 1).all() and (d.sex[d.ma[kids]] == 2).all()
    # Mendelian consistency of the simulated hard calls is high (errors + missing only)
    s = synth.simulate(300, 32, soft=True, seed=1)
    assert s.soft_num.shape == (s.n_total, 32, 3) and abs(s.soft_num.sum(axis=2).astype(int) - 10000).max() <= 3
    ped = synth.truth_pedigree(d, d.n_total + 8, d.n_total)
    labels, n = ped.components()
    assert n >= 1 and (labels[: d.n_total] >= 0).all()


def test_synthetic_truth_is_a_forest_for_any_depth():
    """The sampler's model is acyclic (loops abort: _CheckPed), so the synthetic truth must be a forest
    whatever the number of generations — small deep pedigrees once got a loop-closing fallback couple."""
    from pedfac_b200 import synth
    for n_gen in (2, 3, 4, 5, 6):
        for seed in range(6):
            d = synth.simulate(120, 8, soft=False, seed=seed, n_gen=n_gen)
            ped = synth.truth_pedigree(d, 4000, 4000)
            _, n_cc = ped.components()
            nodes = int(ped.active.sum()) + len(ped.marriages)
            edges = sum(2 + len(m.kids) for m in ped.marriages.values())
            assert edges == nodes - n_cc, (n_gen, seed)
