"""Writers for the sampler's intermediate files — the file contract between the R front-end and
`pedigraph`: `geno.txt`, `prior.txt`, `rand` as R/renderGeno.R:219-239 and
R/runPedFac.R:85-87 write them. Used to feed synthetic data sets to the `pedigraph` executable."""
from __future__ import annotations

from pathlib import Path

import numpy as np

from .synth import SynthData


def write_run_dir(data: SynthData, run_dir, *, max_unobs: int = 1, min_age: float = 1.0, max_age: float = 1.0,
                  observe_frac: float = 0.8, max_gen: int = 0, seeds=(2590, 7579545), loci=None, seed_marriages: bool = False,
                  seed_one_kid: bool = False):
    """Writes run_dir/{geno.txt, prior.txt, rand}. Only observed individuals are listed (ids 1..n
    in slot order); `gen` is the generation index (0 = youngest), founders are the oldest observed
    generation when max_gen == 0 (R/renderGeno.R:84-86). With `seed_marriages` the optional
    marriage.txt (`kid pa ma` rows by user id; `nMar` = number of couples) seeds the chain with the
    true trios whose three members are all observed; `seed_one_kid` keeps only the first kid of every
    couple (the reference's marriage.txt reader never grows a marriage beyond its initial capacity
    of one kid — `_doubleKidCap` is called from `_PickAndUpdate` only — so it faults on more)."""
    run = Path(run_dir)
    run.mkdir(parents=True, exist_ok=True)
    sl = slice(None) if loci is None else slice(0, loci)
    S = data.n_loci if loci is None else loci
    obs = np.flatnonzero(data.observed)
    order = obs[np.argsort(data.slot[obs])]
    top = int(data.gen[obs].max())
    n_gen = top + max_gen + 1
    if data.soft_num is not None:      # every possible 4-decimal token, looked up instead of formatted
        tbl = np.array([f"{k / 10000.0:.4f}" for k in range(65536)])
        tblc = np.char.add(tbl, ",")
    else:
        tbl = np.array(["0", "1", "2", "3"])
    with open(run / "geno.txt", "w") as f:
        for rank, i in enumerate(order):
            founder = 1 if (max_gen == 0 and data.gen[i] == top) else 0
            if data.soft_num is not None:
                n = data.soft_num[i, sl]
                toks = np.char.add(np.char.add(tblc[n[:, 0]], tblc[n[:, 1]]), tbl[n[:, 2]])
            else:
                toks = tbl[data.codes[i, sl]]
            f.write(f"{rank + 1} 1 {int(data.sex[i])} {int(data.gen[i])} {founder} " + " ".join(toks.tolist()) + "\n")
    n_mar = 0
    if seed_marriages:
        uid = {int(i): k + 1 for k, i in enumerate(order)}
        rows = [(uid[int(i)], uid[int(data.pa[i])], uid[int(data.ma[i])]) for i in order
                if data.pa[i] >= 0 and int(data.pa[i]) in uid and int(data.ma[i]) in uid]
        if seed_one_kid:
            seen = set()
            rows = [r for r in rows if (r[1], r[2]) not in seen and not seen.add((r[1], r[2]))]
        n_mar = len({(p_, m_) for _, p_, m_ in rows})
        (run / "marriage.txt").write_text("".join(f"{k} {p_} {m_}\n" for k, p_, m_ in rows))
    obs_frac = np.zeros(n_gen)
    obs_frac[np.unique(data.gen[obs])] = observe_frac
    with open(run / "prior.txt", "w") as f:
        f.write(f"nIndiv {len(order)}\nnObsIndiv {len(order)}\nnGen {n_gen}\nnSNP {S}\nnMar {n_mar}\nmaxID {len(order) + 1}\n")
        f.write("aFreq " + " ".join(f"{a:g}" for a in data.afreq[sl]) + "\n")
        f.write("epsilon " + " ".join(f"{e:g}" for e in data.eps[sl]) + "\n")
        f.write("obsFrac " + " ".join(f"{x:g}" for x in np.round(obs_frac, 3)) + "\n")
        f.write(f"maxAge {max_age:g}\nminAge {min_age:g}\n")
        f.write(f"maxUnobsLayer {max_unobs}")
    (run / "rand").write_text(f"{seeds[0]} {seeds[1]} ")
    return order
