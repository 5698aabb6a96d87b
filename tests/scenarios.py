"""Random small pedigrees + genotype data shared by the CPU (oracle vs brute force) and GPU
(CUDA vs oracle) tests."""
from __future__ import annotations

import numpy as np

from pedfac_b200 import Pedigree


class Scenario:
    def __init__(self, ped, S, eps, afreq, geno):
        self.ped, self.S, self.eps, self.afreq = ped, S, eps, afreq
        self.geno = geno  # slot -> ("hard", codes) | ("f32"|"f64"|"dec16", array) | ("none",)

    def upload(self, engine):
        engine.set_locus_params(self.eps, self.afreq)
        for slot, g in self.geno.items():
            if g[0] == "hard":
                engine.set_hard(slot, g[1])
            elif g[0] == "none":
                engine.set_unobserved(slot)
            else:
                engine.set_soft(slot, g[1], g[0])

    def emission(self, slot):
        """fp64 emission [S][3] exactly as the engines decode it."""
        g = self.geno[slot]
        if g[0] == "none":
            return np.ones((self.S, 3))
        if g[0] == "hard":
            e = np.repeat((self.eps / 2.0)[:, None], 3, axis=1)
            for s, c in enumerate(g[1]):
                if c == 3:
                    e[s] = 1.0
                else:
                    e[s, c] = 1.0 - self.eps[s]
            return e
        if g[0] == "f32":
            return np.asarray(g[1], dtype=np.float32).astype(np.float64)
        if g[0] == "f64":
            return np.asarray(g[1], dtype=np.float64)
        return np.asarray(g[1], dtype=np.float64) / 10000.0


def random_genotypes(rng, S, kind):
    if kind == "hard":
        return ("hard", rng.choice([0, 1, 2, 3], size=S, p=[.45, .35, .15, .05]).astype(np.uint8))
    if kind == "none":
        return ("none",)
    lik = rng.dirichlet([0.6, 0.6, 0.6], size=S)
    num = np.clip(np.round(lik * 10000), 1, 65535).astype(np.uint16)
    if kind == "dec16":
        return ("dec16", num)
    if kind == "f64":
        return ("f64", num.astype(np.float64) / 10000.0)
    return ("f32", (num.astype(np.float64) / 10000.0).astype(np.float32))


def random_forest(rng, n_indiv, S, *, n_marr_target=None, max_kids=4, p_selfing=0.0,
                  kinds=("hard", "dec16", "f32", "f64", "none"), cap_extra=4, max_comp=None):
    """A random acyclic pedigree graph on slots 0..n_indiv-1 (union-find keeps it a forest)."""
    cap_i, cap_m = n_indiv + cap_extra, max(4, n_indiv)
    ped = Pedigree(cap_i, cap_m)
    eps = np.round(rng.uniform(0.005, 0.05, size=S), 3)
    afreq = np.round(rng.uniform(0.05, 0.5, size=S), 3)
    geno = {}
    for i in range(n_indiv):
        ped.add_individual(i, founder=bool(rng.random() < 0.4))
        geno[i] = random_genotypes(rng, S, kinds[rng.integers(len(kinds))])
    parent = list(range(n_indiv))
    size = [1] * n_indiv

    def find(x):
        while parent[x] != x:
            parent[x] = parent[parent[x]]
            x = parent[x]
        return x

    has_parents = set()
    n_marr_target = n_indiv // 2 if n_marr_target is None else n_marr_target
    m = 0
    for _ in range(n_marr_target * 6):
        if m >= n_marr_target:
            break
        selfing = rng.random() < p_selfing
        k = int(rng.integers(1, max_kids + 1))
        need = (1 if selfing else 2) + k
        pick = rng.permutation(n_indiv)[:need + 3]
        members, roots = [], set()
        for x in pick:
            r = find(int(x))
            if r in roots:
                continue
            is_kid = len(members) >= (1 if selfing else 2)
            if is_kid and int(x) in has_parents:
                continue
            members.append(int(x)); roots.add(r)
            if len(members) == need:
                break
        if len(members) < (1 if selfing else 2) + 1:
            continue
        if max_comp is not None and sum(size[r] for r in roots) > max_comp:
            continue
        pa = members[0]
        ma = pa if selfing else members[1]
        kids = members[(1 if selfing else 2):]
        ped.add_marriage(m, pa, ma, kids, selfing)
        has_parents.update(kids)
        r0 = find(members[0])
        for x in members[1:]:
            rx = find(x)
            if rx != r0:
                parent[rx] = r0
                size[r0] += size[rx]
        m += 1
    return Scenario(ped, S, eps, afreq, geno)
