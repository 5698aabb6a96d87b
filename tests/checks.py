"""Engine-agnostic checks used against the CPU oracle (always) and the CUDA engine (-m gpu)."""
from __future__ import annotations

import numpy as np

import bruteforce as bf
from pedfac_b200 import PF_PROP_PRONG, PF_PROP_SELFING, PF_PROP_TRIO


def local_factors(sc):
    prior = bf.hwe_prior(sc.afreq)
    return {int(i): sc.emission(int(i)) * (prior if sc.ped.founder[i] else 1.0)
            for i in np.flatnonzero(sc.ped.active)}


def comp_members(sc, labels, g):
    members = [int(i) for i in np.flatnonzero(sc.ped.active) if labels[i] == g]
    marrs = [(m, mm) for m, mm in sorted(sc.ped.marriages.items()) if labels[mm.pa] == g]
    return members, marrs


def check_sweep_against_bruteforce(engine, sc, rtol=1e-11):
    labels, _ = sc.ped.components()
    view, glob = sc.ped.view(labels)
    llcc = engine.sweep(view)
    local = local_factors(sc)
    for j, g in enumerate(glob):
        members, marrs = comp_members(sc, labels, g)
        q = bf.component_quantities(members, local, [(mm.pa, mm.ma, mm.kids, mm.selfing) for _, mm in marrs])
        np.testing.assert_allclose(llcc[j], np.log(q["Z"]).sum(), rtol=rtol, atol=1e-12)
        for x in members:
            np.testing.assert_allclose(engine.get_stub(x), q["stub"][x], rtol=rtol, atol=0)
        for mi, (m, _) in enumerate(marrs):
            np.testing.assert_allclose(engine.get_prong(m), q["prong"][mi], rtol=rtol, atol=0)
    return llcc, glob, labels


def check_eval_against_bruteforce(engine, sc, rng, n_kids=4, rtol=1e-11):
    """The data term of a proposal is the log partition value of the joined component (so that
    ll = LLtot + lsum - LLcc[parts] is the new total): lsum must equal sum_s log Z_joined."""
    labels, _ = sc.ped.components()
    view, glob = sc.ped.view(labels)
    llcc = dict(zip([int(g) for g in glob], engine.sweep(view)))
    local = local_factors(sc)
    prior = bf.hwe_prior(sc.afreq)
    has_parents = {k for mm in sc.ped.marriages.values() for k in mm.kids}
    free = [int(i) for i in np.flatnonzero(sc.ped.active) if int(i) not in has_parents]
    rng.shuffle(free)
    checked = 0
    for kid in free[:n_kids]:
        others = [int(i) for i in np.flatnonzero(sc.ped.active) if labels[i] != labels[kid]]
        pa = int(rng.choice(others)) if others and rng.random() < 0.8 else -1
        others2 = [i for i in others if pa == -1 or labels[i] != labels[pa]]
        ma = int(rng.choice(others2)) if others2 and rng.random() < 0.8 else -1
        props = [(PF_PROP_TRIO, pa, ma, -1), (PF_PROP_SELFING, pa, -1, -1)]
        marr_other = [m for m, mm in sc.ped.marriages.items() if labels[mm.pa] != labels[kid]]
        if marr_other:
            props.append((PF_PROP_PRONG, -1, -1, int(rng.choice(marr_other))))
        lsum = engine.eval(kid, props)
        for (kind, p, q, m), got in zip(props, lsum):
            comps = {int(labels[kid])}
            loc = dict(local)
            new_id = 10_000
            extra = []
            if kind == PF_PROP_PRONG:
                mm = sc.ped.marriages[m]
                comps.add(int(labels[mm.pa]))
                replace = (m, (mm.pa, mm.ma, mm.kids + [kid], mm.selfing))
            else:
                selfing = kind == PF_PROP_SELFING
                ids = []
                for par in ([p] if selfing else [p, q]):
                    if par == -1:
                        loc[new_id] = np.ones_like(prior) * prior
                        extra.append(new_id); ids.append(new_id); new_id += 1
                    else:
                        comps.add(int(labels[par])); ids.append(par)
                replace = (None, (ids[0], ids[0] if selfing else ids[1], [kid], selfing))
            members, marrs = [], []
            for g in sorted(comps):
                mem, ms = comp_members(sc, labels, g)
                members += mem
                marrs += ms
            members += extra
            mlist = [(mm.pa, mm.ma, mm.kids, mm.selfing) for mi, mm in marrs if mi != replace[0]] + [replace[1]]
            if len(members) > 10:
                continue
            Z = bf.component_quantities(members, loc, mlist)["Z"]
            want = np.log(Z).sum()
            np.testing.assert_allclose(got, want, rtol=rtol, atol=1e-10)
            checked += 1
    return checked


def check_engines_agree(a, b, sc, rng, rtol=1e-12, n_kids=8, dirty=None):
    """Sweep (optionally only the `dirty` labels) + a proposal batch on two engines."""
    labels, _ = sc.ped.components()
    view, glob = sc.ped.view(labels, only=dirty)
    la, lb = a.sweep(view), b.sweep(view)
    np.testing.assert_allclose(la, lb, rtol=rtol, atol=1e-12)
    for x in view.indiv:
        sa, sb = a.get_stub(int(x)), b.get_stub(int(x))
        np.testing.assert_allclose(sa, sb, rtol=rtol, atol=0)
    for m in view.marr:
        np.testing.assert_allclose(a.get_prong(int(m)), b.get_prong(int(m)), rtol=rtol, atol=0)
    act = [int(i) for i in np.flatnonzero(sc.ped.active)]
    marr = list(sc.ped.marriages)
    for kid in rng.choice(act, size=min(n_kids, len(act)), replace=False):
        props = [(PF_PROP_TRIO, -1, -1, -1)]
        for _ in range(int(rng.integers(1, 40))):
            r = rng.random()
            if r < 0.6:
                props.append((PF_PROP_TRIO, int(rng.choice(act + [-1])), int(rng.choice(act + [-1])), -1))
            elif r < 0.75:
                props.append((PF_PROP_SELFING, int(rng.choice(act + [-1])), -1, -1))
            elif marr:
                props.append((PF_PROP_PRONG, -1, -1, int(rng.choice(marr))))
        ea, eb = a.eval(int(kid), props), b.eval(int(kid), props)
        np.testing.assert_allclose(ea, eb, rtol=rtol, atol=1e-11)
    return la
