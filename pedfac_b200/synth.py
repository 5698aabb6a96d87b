"""Synthetic pedigree + genotype generator for the BASELINE.json workloads,
modelled on the reference's ad-hoc simulator `simGeno` (R/renderGeno.R:257-372): Beta(10,10)
allele frequencies folded to the minor allele, gene-dropping through a mating table, per-allele
flip errors at rate `geno_err`, and the front-end's recoding (R/renderGeno.R:68-240): "0" is the
major allele, `aFreq` is the observed "1"-allele frequency rounded to 3 decimals, soft calls are
normalised triples rounded to 4 decimals.

Everything is numpy with a fixed seed (data seed 46 = the `simGeno` default); no I/O.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .abi import PF_PROP_PRONG, PF_PROP_TRIO, PROPOSAL_DTYPE
from .pedigree import Pedigree


GENO_BLOCK = 50


@dataclass
class SynthData:
    n_total: int               # individuals in the true pedigree (observed + unobserved)
    n_obs: int
    n_loci: int
    gen: np.ndarray            # [n_total] generation, 0 = youngest
    sex: np.ndarray            # [n_total] 1 male, 2 female
    pa: np.ndarray             # [n_total] true father (index into the truth arrays) or -1
    ma: np.ndarray             # [n_total] true mother or -1
    observed: np.ndarray       # [n_total] bool
    slot: np.ndarray           # [n_total] engine slot: observed first (0..n_obs-1), then unobserved
    codes: np.ndarray          # [n_total][S] uint8 hard calls 0/1/2, 3 = missing (error model applied)
    soft_num: np.ndarray | None  # [n_total][S][3] uint16 numerators over 10000 (soft configs)
    afreq: np.ndarray          # [S] observed "1"-allele frequency, 3 decimals
    eps: np.ndarray            # [S]
    n_gen: int


def simulate(n_total: int, n_loci: int, *, soft: bool, seed: int = 46, n_gen: int = 3,
             mean_kids: float = 2.5, geno_err: float = 0.02, observe_frac: float = 0.8,
             missing_frac: float = 0.02, remate_frac: float = 0.1,
             max_component: int = 150, locus_range=None) -> SynthData:
    rng = np.random.default_rng(seed)
    # generation sizes: each couple has Poisson(mean_kids) kids, so sizes grow by mean_kids/2
    growth = mean_kids / 2.0
    w = np.array([growth ** k for k in range(n_gen)])           # oldest .. youngest
    sizes = np.maximum(2, np.round(n_total * w / w.sum()).astype(int))
    sizes[-1] = max(2, n_total - sizes[:-1].sum())
    gen = np.concatenate([np.full(sz, n_gen - 1 - k) for k, sz in enumerate(sizes)]).astype(np.int32)
    n_total = len(gen)
    sex = rng.integers(1, 3, size=n_total).astype(np.int32)
    pa = np.full(n_total, -1, np.int32)
    ma = np.full(n_total, -1, np.int32)
    start = np.concatenate([[0], np.cumsum(sizes)])
    # The sampler's state is always a forest (default -c 0 rejects loop-closing proposals,
    # _checkMatchingCC@0x1000241b0) and its unnormalised messages underflow fp64 beyond a few
    # hundred individuals per component, so the synthetic truth is generated
    # loop-free with bounded components: a couple is formed only across two different components
    # whose joint size stays below `max_component`.
    uf = np.arange(n_total)
    csize = np.ones(n_total, dtype=np.int64)

    def find(x):
        while uf[x] != x:
            uf[x] = uf[uf[x]]
            x = uf[x]
        return x

    def union(a, b):
        ra, rb = find(a), find(b)
        if ra != rb:
            uf[rb] = ra
            csize[ra] += csize[rb]
        return ra

    for k in range(n_gen - 1):                                   # parents in block k, kids in block k+1
        par = np.arange(start[k], start[k + 1])
        kids = np.arange(start[k + 1], start[k + 2])
        males, females = par[sex[par] == 1], par[sex[par] == 2]
        if len(males) == 0 or len(females) == 0:
            sex[par[0]], sex[par[-1]] = 1, 2
            males, females = par[sex[par] == 1], par[sex[par] == 2]
        n_couples = max(1, int(round(len(kids) / mean_kids)))
        # monogamous-ish: walk a permutation of the males; a fraction of them re-mates
        order = list(rng.permutation(males))
        order += [int(x) for x in rng.choice(males, size=int(remate_frac * len(males)) + 1)]
        fpool = list(rng.permutation(females))
        couples = []
        budget = int(max_component - mean_kids - 1)
        for m_ in order:
            if len(couples) >= n_couples or not fpool:
                break
            for tries in range(min(8, len(fpool))):
                f_ = fpool[-1 - tries]
                rm, rf = find(int(m_)), find(int(f_))
                if rm != rf and csize[rm] + csize[rf] <= budget:
                    fpool.pop(-1 - tries)
                    if rng.random() < remate_frac:
                        fpool.insert(int(rng.integers(0, len(fpool) + 1)), f_)
                    couples.append((int(m_), int(f_)))
                    union(int(m_), int(f_))
                    break
        if not couples:
            # nothing fits the component budget: take any couple from two different components; if the
            # whole generation is one component already, its kids stay parentless (never close a loop)
            for m_ in males:
                f_ok = [int(f_) for f_ in females if find(int(f_)) != find(int(m_))]
                if f_ok:
                    couples = [(int(m_), f_ok[0])]
                    union(int(m_), f_ok[0])
                    break
        if not couples:
            continue
        couple_of_kid = np.sort(rng.integers(0, len(couples), size=len(kids)))
        carr = np.array(couples, dtype=np.int32)
        pa[kids], ma[kids] = carr[couple_of_kid, 0], carr[couple_of_kid, 1]
        for kk, c in zip(kids, couple_of_kid):
            union(couples[c][0], int(kk))

    # genotypes are generated in blocks of GENO_BLOCK loci, each from its own seeded stream, so
    # a rank of a locus-sharded run can generate just its shard and still see the same data
    observed = rng.random(n_total) < observe_frac
    if not observed.any():
        observed[:] = True
    l0, l1 = (0, n_loci) if locus_range is None else locus_range
    codes = np.zeros((n_total, n_loci), dtype=np.uint8)
    soft_num = np.zeros((n_total, n_loci, 3), dtype=np.uint16) if soft else None
    afreq = np.full(n_loci, 0.25)
    founders = np.flatnonzero(pa < 0)
    for b0 in range(l0 - l0 % GENO_BLOCK, l1, GENO_BLOCK):
        b1 = min(b0 + GENO_BLOCK, n_loci)
        nb = b1 - b0
        brng = np.random.default_rng([seed, 7919, b0 // GENO_BLOCK])
        af = brng.beta(10.0, 10.0, size=nb)
        maf = np.minimum(af, 1.0 - af)
        hap = np.zeros((2, n_total, nb), dtype=np.uint8)
        hap[:, founders, :] = (brng.random((2, len(founders), nb)) < maf).astype(np.uint8)
        for k in range(1, n_gen):                                 # gene drop, unlinked loci
            kids = np.arange(start[k], start[k + 1])
            pick_p = brng.integers(0, 2, size=(len(kids), nb), dtype=np.uint8)
            pick_m = brng.integers(0, 2, size=(len(kids), nb), dtype=np.uint8)
            hp, hm = hap[:, pa[kids], :], hap[:, ma[kids], :]
            hap[0, kids, :] = np.where(pick_p == 0, hp[0], hp[1])
            hap[1, kids, :] = np.where(pick_m == 0, hm[0], hm[1])
        flip = brng.random((2, n_total, nb)) < geno_err          # per-allele flip errors
        hap ^= flip.astype(np.uint8)
        cb = (hap[0] + hap[1]).astype(np.uint8)
        # front-end recoding: "0" = major allele among observed calls; aFreq = freq of "1", 3 decimals
        f1 = cb[observed].mean(axis=0) / 2.0
        swap = f1 > 0.5
        cb[:, swap] = 2 - cb[:, swap]
        f1 = np.where(swap, 1.0 - f1, f1)
        afreq[b0:b1] = np.clip(np.round(f1, 3), 0.001, 0.5)
        if soft:
            base = np.full((n_total, nb, 3), geno_err / 2.0)
            np.put_along_axis(base, cb[..., None].astype(np.int64), 1.0 - geno_err, axis=2)
            lik = base * brng.gamma(4.0, 0.25, size=base.shape)
            lik /= lik.sum(axis=2, keepdims=True)
            soft_num[:, b0:b1, :] = np.clip(np.round(lik * 10000.0), 1, 65535).astype(np.uint16)
        else:
            cb = np.where(brng.random(cb.shape) < missing_frac, 3, cb).astype(np.uint8)
        codes[:, b0:b1] = cb

    slot = np.empty(n_total, np.int32)
    slot[observed] = np.arange(observed.sum())
    slot[~observed] = observed.sum() + np.arange((~observed).sum())
    return SynthData(n_total, int(observed.sum()), n_loci, gen, sex, pa, ma, observed, slot, codes,
                     soft_num, afreq, np.full(n_loci, geno_err), n_gen)


def upload(engine, data: SynthData, soft_fmt: str = "dec16"):
    """Send locus parameters and the observed individuals' genotype rows to an engine
    (CUDA or oracle: both expose the same methods). Unobserved slots keep emission 1."""
    engine.set_locus_params(data.eps, data.afreq)
    for i in np.flatnonzero(data.observed):
        s = int(data.slot[i])
        if data.soft_num is not None:
            if soft_fmt == "dec16":
                engine.set_soft(s, data.soft_num[i], "dec16", 10000)
            elif soft_fmt == "f64":
                engine.set_soft(s, data.soft_num[i].astype(np.float64) / 10000.0, "f64")
            else:
                engine.set_soft(s, (data.soft_num[i].astype(np.float64) / 10000.0).astype(np.float32), "f32")
        else:
            engine.set_hard(s, data.codes[i])


def truth_pedigree(data: SynthData, cap_indiv: int, cap_marr: int) -> Pedigree:
    """The true pedigree as the sampler would hold it: every individual active (unobserved true
    parents occupy unobserved slots), one marriage per distinct (father, mother) couple,
    the oldest generation are founders."""
    ped = Pedigree(cap_indiv, cap_marr)
    for i in range(data.n_total):
        ped.add_individual(int(data.slot[i]), founder=bool(data.pa[i] < 0))
    couples: dict[tuple[int, int], list[int]] = {}
    for i in np.flatnonzero(data.pa >= 0):
        couples.setdefault((int(data.slot[data.pa[i]]), int(data.slot[data.ma[i]])), []).append(int(data.slot[i]))
    for m, ((p, q), kids) in enumerate(sorted(couples.items())):
        ped.add_marriage(m, p, q, kids)
    return ped


def singletons_pedigree(data: SynthData, cap_indiv: int, cap_marr: int) -> Pedigree:
    """Chain start without marriage.txt: observed individuals only, all singletons; the oldest
    observed generation are founders (R/renderGeno.R:84-86)."""
    ped = Pedigree(cap_indiv, cap_marr)
    top = data.gen[data.observed].max()
    for i in np.flatnonzero(data.observed):
        ped.add_individual(int(data.slot[i]), founder=bool(data.gen[i] == top))
    return ped


def gibbs_step_proposals(data: SynthData, ped: Pedigree, labels: np.ndarray, kid_truth: int,
                         rng: np.random.Generator, n_prefilter: int = 15, top_pairs: int = 10):
    """The proposal list `_EvalMoves@0x100015010` would build for one focal kid on graph `ped`
    (kid already pruned): two new founders; every prefilter candidate and every active unobserved
    individual of the parent generation as a single given parent (skipping the kid's own
    component); father x mother pairs from the first `top_pairs` of each sex — an existing
    marriage of the pair becomes a prong proposal, pairs inside one component are dropped.
    The prefilter list is emulated by the true parents plus random parent-generation individuals.
    Returns a `pf_proposal` array."""
    slot = data.slot
    kid = int(slot[kid_truth])
    pg = data.gen[kid_truth] + 1
    pool = np.flatnonzero((data.gen == pg) & data.observed)
    cand = [int(x) for x in (data.pa[kid_truth], data.ma[kid_truth]) if x >= 0 and data.observed[x]]
    extra = rng.choice(pool, size=min(len(pool), n_prefilter), replace=False) if len(pool) else []
    for x in extra:
        if len(cand) >= n_prefilter:
            break
        if int(x) not in cand:
            cand.append(int(x))
    unobs = [int(x) for x in np.flatnonzero((data.gen == pg) & ~data.observed) if ped.active[slot[x]]]
    props = [(PF_PROP_TRIO, -1, -1, -1)]
    males, females = [], []
    for x in cand + unobs:
        sx = int(slot[x])
        if labels[sx] == labels[kid]:
            continue
        if data.sex[x] == 1:
            props.append((PF_PROP_TRIO, sx, -1, -1)); males.append(sx)
        else:
            props.append((PF_PROP_TRIO, -1, sx, -1)); females.append(sx)
    married = {(mm.pa, mm.ma): m for m, mm in ped.marriages.items()}
    for p in males[:top_pairs]:
        for q in females[:top_pairs]:
            if (p, q) in married:
                props.append((PF_PROP_PRONG, p, q, married[(p, q)]))
            elif labels[p] != labels[q]:
                props.append((PF_PROP_TRIO, p, q, -1))
    return kid, np.array(props, dtype=PROPOSAL_DTYPE)
