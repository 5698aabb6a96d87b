"""Exhaustive-enumeration ground truth for small pedigrees — independent of both the CPU oracle
and the CUDA engine. The reference ships no expected outputs, so this is
what pins the equations of SURVEY.md §8(a): every quantity is a literal sum over all 3^n joint
genotype assignments of a connected component.

    weight(g) = prod_i local_i[g_i] * prod_marriages prod_kids T[g_kid][g_ma][g_pa]
    local_i   = emission_i * (HWE prior if founder else 1)
"""
from __future__ import annotations

import itertools

import numpy as np

# pedigraph@0x100035b10 _TPROB, [gk][gm][gp] (also tests/golden/tprob.json)
T = np.array([1, .5, 0, .5, .25, 0, 0, 0, 0,
              0, .5, 1, .5, .5, .5, 1, .5, 0,
              0, 0, 0, 0, .25, .5, 0, .5, 1], dtype=np.float64).reshape(3, 3, 3)


def hwe_prior(afreq):
    a = np.asarray(afreq, dtype=np.float64)
    q = 1.0 - a
    return np.stack([q * q, (2.0 * (1.0 - q)) * q, (1.0 - q) * (1.0 - q)], axis=-1)


def emission_hard(codes, eps):
    codes = np.asarray(codes)
    eps = np.asarray(eps, dtype=np.float64)
    e = np.repeat((eps / 2.0)[:, None], 3, axis=1)
    for s, c in enumerate(codes):
        if c == 3:
            e[s] = 1.0
        else:
            e[s, c] = 1.0 - eps[s]
    return e


def component_joint(members, local, marriages):
    """members: list of individual ids; local: dict id -> [S][3]; marriages: list of
    (pa, ma, kids, selfing) fully inside the component. Returns array [S, 3, ..., 3] (one axis
    per member, in `members` order) of joint weights."""
    n = len(members)
    S = next(iter(local.values())).shape[0]
    pos = {x: k for k, x in enumerate(members)}
    W = np.ones((S,) + (3,) * n)
    for x in members:
        shape = [S] + [1] * n
        shape[1 + pos[x]] = 3
        W = W * local[x].reshape(shape)
    for (pa, ma, kids, selfing) in marriages:
        for k in kids:
            # factor T[gk][gm][gp] broadcast onto axes (k, ma, pa)
            F = np.ones((1,) + (3,) * n)
            it = np.zeros((3,) * n)
            for g in itertools.product(range(3), repeat=n):
                gp = g[pos[pa]]
                gm = gp if selfing else g[pos[ma]]
                it[g] = T[g[pos[k]], gm, gp]
            W = W * it[None]
    return W


def component_quantities(members, local, marriages):
    """Returns dict with Z [S], stubs {id: [S][3]}, prongs {marriage index: [S][3]}."""
    n = len(members)
    pos = {x: k for k, x in enumerate(members)}
    W = component_joint(members, local, marriages)
    axes = tuple(range(1, n + 1))
    out = {"Z": W.sum(axis=axes), "stub": {}, "prong": {}}
    for x in members:
        other = tuple(a for a in axes if a != 1 + pos[x])
        out["stub"][x] = W.sum(axis=other)
    for mi, (pa, ma, kids, selfing) in enumerate(marriages):
        pr = np.zeros((W.shape[0], 3))
        for g in itertools.product(range(3), repeat=n):
            gp = g[pos[pa]]
            gm = gp if selfing else g[pos[ma]]
            w = W[(slice(None),) + g]
            for gk in range(3):
                pr[:, gk] += T[gk, gm, gp] * w
        out["prong"][mi] = pr
    return out
