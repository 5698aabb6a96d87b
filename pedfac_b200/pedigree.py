"""Host-side mirror of the pedigree factor graph that the reference keeps in `ped->indiv[]` /
`ped->mnode[]`: which individual slots are active, which are founders, and
the marriages (father, mother, kids, selfing flag). It labels connected components
(`_CheckPed@…`/`_UpdateConComp@0x100013…` in the reference) and emits the `pf_graph_view` the
likelihood engine consumes (the worklists of `_TurnOnSpotlight@0x10001bb90`).

Pure integer bookkeeping; no likelihood arithmetic lives here.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .abi import GraphView


@dataclass
class Marriage:
    pa: int
    ma: int
    kids: list = field(default_factory=list)
    selfing: bool = False


class Pedigree:
    def __init__(self, cap_indiv: int, cap_marr: int):
        self.cap_indiv, self.cap_marr = cap_indiv, cap_marr
        self.active = np.zeros(cap_indiv, dtype=bool)
        self.founder = np.zeros(cap_indiv, dtype=bool)
        self.marriages: dict[int, Marriage] = {}

    # -- edits ---------------------------------------------------------------------------------
    def add_individual(self, slot: int, founder: bool):
        assert 0 <= slot < self.cap_indiv
        self.active[slot] = True
        self.founder[slot] = founder

    def add_marriage(self, m: int, pa: int, ma: int, kids, selfing: bool = False):
        assert 0 <= m < self.cap_marr and m not in self.marriages
        self.marriages[m] = Marriage(pa, pa if selfing else ma, list(kids), selfing)

    def remove_marriage(self, m: int):
        del self.marriages[m]

    def parents_marriage_of(self, kid: int):
        for m, mm in self.marriages.items():
            if kid in mm.kids:
                return m
        return None

    def detach_kid(self, kid: int):
        """`_PruneIndiv@0x1000141f0`: remove `kid` from its parental marriage; an only child's
        marriage disappears. Returns (pa, ma, marriage slot or None if it was deleted)."""
        m = self.parents_marriage_of(kid)
        if m is None:
            return None
        mm = self.marriages[m]
        mm.kids.remove(kid)
        if not mm.kids:
            del self.marriages[m]
            return mm.pa, mm.ma, None
        return mm.pa, mm.ma, m

    def apply_swap(self, kind: int, move_kids: bool, x: int, z: int, marr_z: int, prev_pa: int,
                   prev_ma: int, prev_marr, new_marr=None):
        """The rewiring of a picked swap choice (`_PickAndUpdate@0x100016cf5` re-attaches the
        pruned x to its previous parents, then `@0x100017f5a-0x100019281` exchanges the fathers
        (kind 1), mothers (kind 2) or both (kind 3) of that family and of `marr_z`, and with
        `move_kids` the two sibships as well). `self` is the pedigree after `detach_kid(x)`;
        `prev_marr` is None when x was an only child (then `new_marr` names the free slot)."""
        if prev_marr is None:
            prev_marr = new_marr
            self.add_marriage(prev_marr, prev_pa, prev_ma, [x])
        else:
            self.marriages[prev_marr].kids.append(x)
        m0, mz = self.marriages[prev_marr], self.marriages[marr_z]
        assert z in mz.kids and not m0.selfing and not mz.selfing
        if kind in (1, 3):
            m0.pa, mz.pa = mz.pa, m0.pa
        if kind in (2, 3):
            m0.ma, mz.ma = mz.ma, m0.ma
        if move_kids:
            k0 = [k for k in m0.kids if k != x]
            kz = [k for k in mz.kids if k != z]
            m0.kids, mz.kids = [x] + kz, [z] + k0
        return prev_marr

    def swap_sites(self):
        """Every (x, z, marr_x, marr_z, marr_of_x) such that x is a kid of marr_of_x, a parent in
        marr_x with spouse z, and z is a kid of marr_z: the situations in which
        `_EvalMoves@0x100015989` proposes swaps for x (selfing marriages excluded)."""
        kid_of = {k: m for m, mm in self.marriages.items() for k in mm.kids}
        out = []
        for mx, mm in sorted(self.marriages.items()):
            if mm.selfing:
                continue
            for x, z in ((mm.pa, mm.ma), (mm.ma, mm.pa)):
                if x in kid_of and z in kid_of and kid_of[x] != kid_of[z]:
                    if not self.marriages[kid_of[x]].selfing and not self.marriages[kid_of[z]].selfing:
                        out.append((x, z, mx, kid_of[z], kid_of[x]))
        return out

    def copy(self):
        q = Pedigree(self.cap_indiv, self.cap_marr)
        q.active, q.founder = self.active.copy(), self.founder.copy()
        q.marriages = {m: Marriage(mm.pa, mm.ma, list(mm.kids), mm.selfing) for m, mm in self.marriages.items()}
        return q

    # -- components ----------------------------------------------------------------------------
    def components(self):
        """Returns (label[cap_indiv] with -1 for inactive slots, n_cc). Labels are assigned in
        order of the lowest slot of each component."""
        parent = np.arange(self.cap_indiv)

        def find(x):
            while parent[x] != x:
                parent[x] = parent[parent[x]]
                x = parent[x]
            return x

        for mm in self.marriages.values():
            members = [mm.pa] + ([] if mm.selfing else [mm.ma]) + mm.kids
            r = find(members[0])
            for y in members[1:]:
                ry = find(y)
                if ry != r:
                    parent[max(r, ry)] = min(r, ry)
                    r = min(r, ry)
        label = np.full(self.cap_indiv, -1, dtype=np.int32)
        n = 0
        for i in np.flatnonzero(self.active):
            r = find(i)
            if label[r] < 0:
                label[r] = n
                n += 1
            label[i] = label[r]
        return label, n

    def view(self, labels=None, only=None):
        """Graph view of the components in `only` (iterable of global labels; None = everything).
        Returns (GraphView, global_labels) where view label j corresponds to global_labels[j]."""
        if labels is None:
            labels, _ = self.components()
        act = np.flatnonzero(self.active)
        if only is not None:
            only = sorted(set(int(x) for x in only))
            keep = np.isin(labels[act], only)
            act = act[keep]
            glob = np.array(only, dtype=np.int32)
        else:
            glob = np.unique(labels[act])
        remap = {int(g): j for j, g in enumerate(glob)}
        indiv_cc = np.array([remap[int(labels[i])] for i in act], dtype=np.int32)
        ms, pas, mas, selfs, kb, kids = [], [], [], [], [0], []
        for m in sorted(self.marriages):
            mm = self.marriages[m]
            if int(labels[mm.pa]) not in remap:
                continue
            ms.append(m); pas.append(mm.pa); mas.append(mm.ma); selfs.append(1 if mm.selfing else 0)
            kids.extend(mm.kids); kb.append(len(kids))
        gv = GraphView(act, indiv_cc, self.founder[act].astype(np.uint8), ms, pas, mas, selfs, kb,
                       kids if kids else np.zeros(0, np.int32), len(glob))
        return gv, glob
