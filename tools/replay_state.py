"""Debug aid: rebuild a pedigree state dumped by the sampler (PEDFAC_DUMP=path with PEDFAC_CHECK=1) on the CPU
oracle and recompute its log-likelihood:  python tools/replay_state.py RUN_DIR STATE_FILE [kid]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT)]
import numpy as np
from oracle.binding import OracleEngine
from pedfac_b200 import Pedigree, PF_PROP_PRONG, PF_PROP_SELFING, PF_PROP_TRIO

run, state = Path(sys.argv[1]), Path(sys.argv[2])
prior = {l.split()[0]: l.split()[1:] for l in (run / "prior.txt").read_text().splitlines() if l.strip()}
S = int(prior["nSNP"][0]); afreq = np.array(prior["aFreq"], float); eps = np.array(prior["epsilon"], float)
inds, marrs, llcc, tot = [], [], {}, None
for l in state.read_text().splitlines():
    t = l.split()
    if t[0] == "T": tot = float(t[1])
    elif t[0] == "I": inds.append(tuple(int(x) for x in t[1:]))
    elif t[0] == "M": marrs.append((int(t[1]), int(t[2]), int(t[3]), int(t[4]), [int(x) for x in t[5:]]))
    elif t[0] == "C": llcc[int(t[1])] = float(t[2])
cap_i = max(i[0] for i in inds) + 2; cap_m = max(m[0] for m in marrs) + 2
e = OracleEngine(S, cap_i, cap_m)
e.set_locus_params(eps, afreq)
# observed rows fill slots 0.. in file order
slot = 0
for l in (run / "geno.txt").read_text().splitlines():
    t = l.split(" ")
    if len(t) < 5 or t[1] == "0": continue
    toks = t[5:5 + S]
    if "," in toks[0]:
        e.set_soft(slot, np.array([[float(x) for x in tk.split(",")] for tk in toks]), "f64")
    else:
        e.set_hard(slot, np.array([int(x) for x in toks], dtype=np.uint8))
    slot += 1
ped = Pedigree(cap_i, cap_m)
for i, uid, obs, founder, cc, gen, sex in inds:
    ped.add_individual(i, bool(founder))
for m, pa, ma, selfing, kids in marrs:
    ped.add_marriage(m, pa, ma, kids, bool(selfing))
labels, n = ped.components()
view, glob = ped.view(labels)
ll = e.sweep(view)
print("components", n, "oracle total", repr(float(np.sum(ll))), "sampler total", repr(tot), "diff", float(np.sum(ll)) - tot)
if len(sys.argv) > 3:
    kid = int(sys.argv[3])
    m = ped.parents_marriage_of(kid)
    mm = ped.marriages[m]
    print("kid", kid, "is in marriage", m, "pa", mm.pa, "ma", mm.ma, "kids", mm.kids, "founder flag", bool(ped.founder[kid]))
    after = float(np.sum(ll))
    was_selfing = mm.selfing
    pa, ma, prev = ped.detach_kid(kid)
    labels2, n2 = ped.components()
    view2, glob2 = ped.view(labels2)
    ll2 = e.sweep(view2)
    before = float(np.sum(ll2))
    g = {int(x): j for j, x in enumerate(glob2)}
    if prev is not None:
        ls = e.eval(kid, [(PF_PROP_PRONG, -1, -1, prev)])[0]
        scored = before + ls - ll2[g[int(labels2[kid])]] - ll2[g[int(labels2[pa])]]
    elif was_selfing:
        ls = e.eval(kid, [(PF_PROP_SELFING, pa, -1, -1)])[0]
        scored = before + ls - ll2[g[int(labels2[kid])]] - ll2[g[int(labels2[pa])]]
    else:
        ls = e.eval(kid, [(PF_PROP_TRIO, pa, ma, -1)])[0]
        scored = before + ls - ll2[g[int(labels2[kid])]] - ll2[g[int(labels2[pa])]] - (ll2[g[int(labels2[ma])]] if labels2[ma] != labels2[pa] else 0.0)
    print("before", repr(before), "proposal data term", repr(float(ls)), "scored", repr(float(scored)), "after", repr(after), "scored - after", float(scored) - after)
    if len(sys.argv) > 4:
        import collections
        ck, cp = int(labels2[kid]), int(labels2[pa])
        for name, c in (("kid", ck), ("marriage", cp)):
            mem = [int(i) for i in np.flatnonzero(ped.active) if labels2[i] == c]
            ms = {m_: mm_ for m_, mm_ in ped.marriages.items() if labels2[mm_.pa] == c}
            print(name, "component", c, "members", len(mem), mem[:40])
            for m_, mm_ in sorted(ms.items()):
                print("   marriage", m_, "pa", mm_.pa, "ma", mm_.ma, "kids", mm_.kids, "founders", [int(ped.founder[x]) for x in (mm_.pa, mm_.ma)], [int(ped.founder[x]) for x in mm_.kids])
