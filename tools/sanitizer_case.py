"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): sweeps with several levels, large
sibships (KFOLD / FAM_EXCL / KID_CHUNK), spilled rows disabled/enabled, evals, swaps and the prefilter."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
from pedfac_b200 import Engine, PF_PROP_TRIO, PF_PROP_PRONG
from scenarios import random_forest

rng = np.random.default_rng(1)
for n, S, mk in ((60, 40, 6), (300, 33, 9)):
    sc = random_forest(rng, n_indiv=n, S=S, max_kids=mk, kinds=("hard", "dec16", "f32", "none"))
    e = Engine(sc.S, sc.ped.cap_indiv, sc.ped.cap_marr)
    sc.upload(e)
    labels, _ = sc.ped.components()
    view, glob = sc.ped.view(labels)
    ll = e.sweep(view)
    kid = int(np.flatnonzero(sc.ped.active)[0])
    others = [int(i) for i in np.flatnonzero(sc.ped.active) if labels[i] != labels[kid]][:40]
    props = [(PF_PROP_TRIO, c, -1, -1) for c in others] + [(PF_PROP_TRIO, -1, -1, -1)]
    ls = e.eval(kid, props)
    ll2, ls2 = e.sweep_eval(view, kid, props)
    assert np.array_equal(ll, ll2) and np.array_equal(ls, ls2)
    sites = sc.ped.swap_sites()
    if sites:
        x, z, mx, mz, m_of_x = sites[0]
        ped2 = sc.ped.copy(); pa, ma, prev = ped2.detach_kid(x); ped2.founder[x] = False
        l2, _ = ped2.components()
        if l2[pa] != l2[x] and l2[ma] != l2[x]:
            e.sweep(ped2.view(l2)[0])
            v3, _ = ped2.view(l2, only={int(l2[x]), int(l2[pa]), int(l2[ma])})
            e.eval_swaps(v3, [(1, 0, x, z, mz, mx, pa, ma, -1 if prev is None else prev), (3, 1, x, z, mz, mx, pa, ma, -1 if prev is None else prev)])
    kids = np.arange(0, min(n, 20), dtype=np.int32); cands = np.arange(0, min(n, 50), dtype=np.int32)
    e.prefilter(kids, (kids % 2).astype(np.uint8), cands, None, top_k=5)
    print("case", n, S, "ok", float(np.sum(ll)))
    e.close()
