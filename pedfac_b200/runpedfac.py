"""Python mirror of the reference's R front-end (`runPedFac`, `checkGeno`, `writeIntermedGeno`,
`retrieveParent`, `retrieveGrandparents` — R/runPedFac.R, R/renderGeno.R, R/assignRelation.R):
same argument names and meaning, same validation messages, same intermediate and output files,
so that a user without R can drive the `pedigraph` executable the way the R package does. It
also implements what the reference documents but never reads: the optional `marker_info.txt`
(README.md:112-141, per-locus `gerror` / `afreq`).

Host-side text processing only; every likelihood is computed by `pedfac_b200/bin/pedigraph`
(libpedfac_cuda). The id map is written as `id.json` / `zero_rep.json` (the R package writes .rds).
"""
from __future__ import annotations

import json
import math
import subprocess
from collections import Counter, defaultdict
from pathlib import Path

import numpy as np

MISSING = {"N", "-1"}


def pedigraph_binary() -> Path:
    """R/runBinary.R:5-12 — where the sampler executable lives."""
    return Path(__file__).resolve().parent / "bin" / "pedigraph"


def _read_table(path):
    rows = [ln.split() for ln in Path(path).read_text().splitlines() if ln.strip()]
    if not rows:
        raise ValueError(f"{path} is empty")
    return rows


def check_geno(geno_path):
    """R/renderGeno.R:13-66 `checkGeno`."""
    geno_path = Path(geno_path)
    if not geno_path.exists():
        raise FileNotFoundError(f"the genotype file 'geno.path' provided - {geno_path} does not exist")
    rows = _read_table(geno_path)
    ids = [r[0] for r in rows]
    if len(set(ids)) != len(ids):
        raise ValueError(f"Duplicate ID entries are found in {geno_path}")
    for r in rows:
        if r[1] not in ("0", "1"):
            raise ValueError("The second column of the genotype files - whether indiv is observed - can accept value 0 or 1")
        if r[2] not in ("0", "1", "2"):
            raise ValueError("The third column of the genotype files - sex of indiv - can accept value 0, 1, or 2")
        try:
            float(r[3])
        except ValueError:
            raise ValueError("The fourth column of the genotype files - birth years - does not allow string character")
        if (len(r) - 4) % 2:
            raise ValueError("Each locus takes two columns")
        for a, b in zip(r[4::2], r[5::2]):
            if len(a.split(",")) == 2 or len(b.split(",")) == 2:
                raise ValueError("Detect missing entries in genotype likelihood format")
            if (a in MISSING) != (b in MISSING):
                raise ValueError("Genotype are reported inconsistently. Under missing genotype, both of allelic columns must either be set as N or -1.")
    return rows


def _fmt(x: float, nd: int) -> str:
    """R's print of round(x, nd): shortest decimal with at most nd digits."""
    s = f"{round(x, nd):.{nd}f}".rstrip("0").rstrip(".")
    return s if s not in ("", "-0") else "0"


def write_intermed_geno(param) -> dict:
    """R/renderGeno.R:68-240 `writeIntermedGeno`: recode alleles so that "0" is the major
    (ancestral) allele group, normalise genotype likelihoods to 4 decimals, and write geno.txt,
    prior.txt and the id map into param['output.path']."""
    rows = _read_table(param["geno.path"])
    out = Path(param["output.path"])
    n = len(rows)
    years = np.array([float(r[3]) for r in rows])
    gen = np.floor((years.max() - years) / param["min.age"]).astype(int)
    n_gen = int(gen.max()) + int(param["max.gen"]) + 1
    founder = (gen == gen.max()).astype(int) if param["max.gen"] == 0 else np.zeros(n, int)
    if param["observe.frac"] == -1:
        obs_frac = np.full(n_gen, -1.0)
    else:
        obs_frac = np.zeros(n_gen)
        obs_frac[gen] = param["observe.frac"]
    n_snp = (len(rows[0]) - 4) // 2

    # per-locus allele counts (hard alleles count 1 each, likelihood triples count their expected alleles)
    counts = [Counter() for _ in range(n_snp)]
    soft = [[None] * n_snp for _ in range(n)]
    for i, r in enumerate(rows):
        for s in range(n_snp):
            a, b = r[4 + 2 * s], r[5 + 2 * s]
            if "," in b:
                v = np.array([float(x) for x in b.split(",")])
                v = np.round(v / v.sum(), 4)
                soft[i][s] = v
                if a not in MISSING:
                    f0 = float(v[0] * 1.0 + v[1] * 0.5)
                    counts[s]["0"] += f0
                    counts[s]["1"] += 1.0 - f0
            else:
                for al in (a, b):
                    if al not in MISSING:
                        counts[s][al] += 1.0
    zero_rep, afreq = [], []
    for s in range(n_snp):
        tot = sum(counts[s].values())
        if tot == 0:
            zero_rep.append(set()); afreq.append(0.0); continue
        freq = sorted(((c / tot, al) for al, c in counts[s].items()), key=lambda t: (-t[0], t[1]))
        base = freq[0][0]
        zero = {freq[0][1]}
        if base < 0.5 and len(freq) > 1:
            # "base, jam, crust": add the rarest alleles while the cumulative frequency stays
            # below one half, then the one allele that crosses it with the largest cumulative sum
            rest = sorted(freq[1:], key=lambda t: (t[0], t[1]))
            cum, over = base, []
            for f, al in rest:
                cum += f
                if cum < 0.5:
                    zero.add(al)
                else:
                    over.append((cum, al))
            if over:
                zero.add(max(over)[1])
        zero_rep.append(zero)
        afreq.append(1.0 - sum(f for f, al in freq if al in zero))

    lines = []
    for i, r in enumerate(rows):
        toks = []
        for s in range(n_snp):
            a, b = r[4 + 2 * s], r[5 + 2 * s]
            if soft[i][s] is not None:
                v = soft[i][s] if "0" in zero_rep[s] or not zero_rep[s] else soft[i][s][::-1]
                toks.append(",".join(_fmt(float(x), 4) for x in v))
            else:
                def code(al):
                    return -1 if al in MISSING else (0 if al in zero_rep[s] else 1)
                g = code(a) + code(b)
                toks.append("3" if g == -2 else str(g))
        lines.append(f"{i + 1} {r[1]} {r[2]} {gen[i]} {founder[i]} " + " ".join(toks))
    (out / "geno.txt").write_text("\n".join(lines) + "\n")
    (out / "id.json").write_text(json.dumps({"id": [r[0] for r in rows]}))
    (out / "zero_rep.json").write_text(json.dumps([sorted(z) for z in zero_rep]))

    eps = [param["geno.err"]] * n_snp
    af = [round(a, 3) for a in afreq]
    if param.get("marker.path"):
        eps, af = apply_marker_info(param["marker.path"], eps, af)
    prior = [f"nIndiv {n}", f"nObsIndiv {sum(r[1] == '1' for r in rows)}", f"nGen {n_gen}", f"nSNP {n_snp}",
             f"nMar {param.get('n.marr', 0)}", f"maxID {n + 1}",
             "aFreq " + " ".join(_fmt(a, 3) for a in af),
             "epsilon " + " ".join(_fmt(e, 6) for e in eps),
             "obsFrac " + " ".join(_fmt(x, 3) for x in obs_frac),
             # no maxAge / minAge lines: R/renderGeno.R:224-239 never writes them, so the sampler runs
             # with its defaults 1.0 / 1.0 and min.age / max.age act only through `gen` above
             # (= floor((max year - year) / min.age), already in generation units)
             f"maxUnobsLayer {param['max.unobs']}"]
    (out / "prior.txt").write_text("\n".join(prior))
    return {"n.indiv": n, "n.snp": n_snp, "n.gen": n_gen, "gen": gen, "afreq": af, "epsilon": eps}


def apply_marker_info(marker_path, eps, afreq):
    """README.md:112-141: first line `name ...`, then optional `gerror ...` (genotyping error per
    locus -> prior.txt `epsilon`) and `afreq ...` (frequency of the "1" allele -> `aFreq`)."""
    rows = _read_table(marker_path)
    if rows[0][0] != "name":
        raise ValueError("marker_info.txt must begin with the tag 'name'")
    n = len(rows[0]) - 1
    if n != len(eps):
        raise ValueError(f"marker_info.txt names {n} loci, the genotype file has {len(eps)}")
    eps, afreq = list(eps), list(afreq)
    for r in rows[1:]:
        vals = [float(x) for x in r[1:]]
        if len(vals) != n or any(not (0.0 <= v <= 1.0) for v in vals):
            raise ValueError(f"marker_info.txt: '{r[0]}' needs {n} values between 0 and 1")
        if r[0] == "gerror":
            eps = vals
        elif r[0] == "afreq":
            afreq = vals
        else:
            raise ValueError(f"marker_info.txt: unknown tag '{r[0]}'")
    return eps, afreq


def _last_per_iter_kid(ped):
    last = {}
    for it, kid, pa, ma in ped:
        last[(it, kid)] = (pa, ma)
    return last


def retrieve_parent(last, max_id, ids):
    """R/assignRelation.R `retrieveParent`: posterior frequency of each (pa, ma) per observed kid."""
    n_iter = max(k[0] for k in last) + 1
    cnt = Counter()
    sub = lambda x: "-1" if x > max_id else ids[x - 1]
    for (it, kid), (pa, ma) in last.items():
        if kid <= max_id:
            cnt[(kid, sub(pa), sub(ma))] += 1
    rows = [(ids[k - 1], pa, ma, c / n_iter) for (k, pa, ma), c in cnt.items()]
    return sorted(rows, key=lambda r: -r[3])


def retrieve_grandparents(last, max_id, ids):
    """R/assignRelation.R `retrieveGrandparents`."""
    n_iter = max(k[0] for k in last) + 1
    cnt = Counter()
    sub = lambda x: "-1" if x > max_id else ids[x - 1]
    for (it, kid), (pa, ma) in last.items():
        if kid > max_id or (it, pa) not in last or (it, ma) not in last:
            continue
        gpp, gmp = last[(it, pa)]
        gpm, gmm = last[(it, ma)]
        cnt[(kid, sub(gpp), sub(gmp), sub(gpm), sub(gmm))] += 1
    rows = [(ids[k[0] - 1],) + k[1:] + (c / n_iter,) for k, c in cnt.items()]
    return sorted(rows, key=lambda r: -r[-1])


def run_pedfac(geno_path, marker_path="", output_path="", random_seed=26, n_iter=1, cyclic_choice=0,
               observe_frac=0.8, max_unobs=1, max_gen=0, min_age=1, max_age=1, haplo_method=0, geno_err=0.02,
               binary=None):
    """R/runPedFac.R:29-177 `runPedFac` — same arguments, checks, files and outputs."""
    geno_path = Path(geno_path)
    if not geno_path.exists():
        raise FileNotFoundError(f"the genotype file 'geno.path' provided - {geno_path} does not exist")
    if marker_path != "" and not Path(marker_path).exists():
        raise FileNotFoundError(f"the marker file 'marker.path' provided - {marker_path} does not exist")
    if output_path != "" and not Path(output_path).exists():
        raise FileNotFoundError(f"the output path 'output.path' provided - {output_path} does not exist")
    out = Path(output_path if output_path != "" else ".") / "run"
    out.mkdir(exist_ok=True)
    if isinstance(random_seed, str):
        raise ValueError("Random seed cannot be a string")
    if n_iter < 1:
        raise ValueError("Number of iteration, n.iter, must be equal or greater than 1")
    if cyclic_choice not in (0, 1, 2):
        raise ValueError("Out of acceptable range for cyclic.choice")
    if observe_frac > 1 or (observe_frac < 0 and observe_frac != -1):
        raise ValueError("Out of acceptable range for observe.frac")
    if max_unobs < 0:
        raise ValueError("max.unobs must be equal or greater than 0")
    if max_gen < 0:
        raise ValueError("max.gen must be equal or greater than 0")
    if min_age <= 0:
        raise ValueError("min.age must be greater than 0")
    if max_age <= 0:
        raise ValueError("max.age must be greater than 0")
    if max_age < min_age:
        raise ValueError("min.age is much greater than max.age")
    if haplo_method not in (0, 1, 2):
        raise ValueError("Out of acceptable range for haplo.method")
    if geno_err > 1 or geno_err < 0:
        raise ValueError("Out of acceptable range for geno.err")
    rng = np.random.default_rng(random_seed)
    seeds = np.ceil(rng.uniform(0, 1e6, size=2)).astype(int)      # R: ceiling(runif(2, 0, 1e6))
    (out / "rand").write_text(f"{seeds[0]} {seeds[1]} ")
    check_geno(geno_path)
    param = {"geno.path": geno_path, "marker.path": marker_path, "output.path": out, "observe.frac": observe_frac,
             "max.unobs": max_unobs, "max.gen": max_gen, "min.age": min_age, "max.age": max_age, "n.marr": 0,
             "geno.err": geno_err}
    write_intermed_geno(param)
    binary = Path(binary) if binary else pedigraph_binary()
    # like the reference, cyclic.choice / haplo.method are validated but not forwarded (R/runPedFac.R:99-101)
    r = subprocess.run([str(binary), "-d", str(out), "-n", str(n_iter), "-r", str(out / "rand")], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"pedigraph failed ({r.returncode}): {r.stderr[-500:]}")
    ped = [tuple(int(x) for x in ln.split()) for ln in (out / "out" / "ped.txt").read_text().splitlines() if ln.strip()]
    ids = json.loads((out / "id.json").read_text())["id"]
    max_id = len(ids)
    last = _last_per_iter_kid(ped)
    name = lambda x: f"unobs.{x}" if x > max_id else ids[x - 1]
    with open(out / "ped_sample.txt", "w") as f:
        for (it, kid), (pa, ma) in sorted(last.items()):
            f.write(f"{it} {name(kid)} {name(pa)} {name(ma)}\n")
    parents = retrieve_parent(last, max_id, ids)
    with open(out / "parent_assignment.txt", "w") as f:
        for k, pa, ma, p in parents:
            f.write(f"{k} {pa} {ma} {_fmt(p, 7)}\n")
    # the reference's full-/half-sib tables are always empty (R/assignRelation.R:62,102: SURVEY.md App. D.5)
    (out / "fullsib_assignment.txt").write_text("")
    (out / "halfsib_assignment.txt").write_text("")
    grand = retrieve_grandparents(last, max_id, ids)
    with open(out / "grandparent_assignment.txt", "w") as f:
        for row in grand:
            f.write(" ".join(str(x) for x in row[:-1]) + f" {_fmt(row[-1], 7)}\n")
    return {"run": out, "parents": parents, "grandparents": grand, "stdout": r.stdout}


if __name__ == "__main__":
    import argparse
    ap = argparse.ArgumentParser(description="runPedFac without R: genotype.txt -> pedigraph -> summaries")
    ap.add_argument("geno_path")
    ap.add_argument("--marker-path", default="")
    ap.add_argument("--output-path", default="")
    ap.add_argument("--n-iter", type=int, default=1)
    ap.add_argument("--random-seed", type=int, default=26)
    ap.add_argument("--observe-frac", type=float, default=0.8)
    ap.add_argument("--max-unobs", type=int, default=1)
    ap.add_argument("--max-gen", type=int, default=0)
    ap.add_argument("--geno-err", type=float, default=0.02)
    a = ap.parse_args()
    res = run_pedfac(a.geno_path, a.marker_path, a.output_path, a.random_seed, a.n_iter, observe_frac=a.observe_frac,
                     max_unobs=a.max_unobs, max_gen=a.max_gen, geno_err=a.geno_err)
    print(f"wrote {res['run']}")
