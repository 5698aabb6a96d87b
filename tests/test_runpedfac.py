"""The R-free front-end (pedfac_b200/runpedfac.py ~ R/runPedFac.R, R/renderGeno.R): input
checks, allele recoding, marker_info.txt, and the whole pipeline on the CPU sampler build."""
from pathlib import Path

import numpy as np
import pytest

from pedfac_b200 import runpedfac as rp
from pedfac_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
REF = ROOT / "oracle" / "pedigraph_ref"

# the four genotype encodings the reference's example/case_0 mixes (our own values):
# integer alleles, missing (N / -1), "0,1,2 a,b,c" likelihoods, haplotype strings
FIXTURE = """\
7 1 1 1986.5 1 0 0,1,2 0.2,0.2,0.2 ACGT ACGT 12 15
21 1 1 1987 -1 -1 0,1,2 0.1,0.5,0.9 TTAA TCTA 12 15
22 1 2 1987 N -1 0,1,2 0.1,0.3,0.4 TATA TCTA 12 12
23 1 1 1989 0 0 0,1,2 0,0.9,0.1 TATA TCTA 15 12
24 1 2 1990 1 1 0,1,2 0,0.8,0.2 TATA TATA 12 12
"""


def test_check_geno_messages(tmp_path):
    p = tmp_path / "genotype.txt"
    p.write_text(FIXTURE)
    assert len(rp.check_geno(p)) == 5
    p.write_text(FIXTURE.replace("21 1 1 1987", "7 1 1 1987"))
    with pytest.raises(ValueError, match="Duplicate ID"):
        rp.check_geno(p)
    p.write_text(FIXTURE.replace("22 1 2 1987 N -1", "22 1 2 1987 N 1"))
    with pytest.raises(ValueError, match="reported inconsistently"):
        rp.check_geno(p)
    p.write_text(FIXTURE.replace("23 1 1", "23 1 5"))
    with pytest.raises(ValueError, match="sex of indiv"):
        rp.check_geno(p)


def test_write_intermed_geno_recodes_like_the_r_front_end(tmp_path):
    p = tmp_path / "genotype.txt"
    p.write_text(FIXTURE)
    param = {"geno.path": p, "output.path": tmp_path, "observe.frac": 0.8, "max.unobs": 1, "max.gen": 0,
             "min.age": 1, "max.age": 1, "geno.err": 0.02, "marker.path": ""}
    info = rp.write_intermed_geno(param)
    rows = [ln.split() for ln in (tmp_path / "geno.txt").read_text().splitlines()]
    assert [r[0] for r in rows] == ["1", "2", "3", "4", "5"]                    # ids renumbered 1..n
    assert [r[3] for r in rows] == ["3", "3", "3", "1", "0"]                    # floor((1990 - year)/1)
    assert [r[4] for r in rows] == ["1", "1", "1", "0", "0"]                    # oldest generation = founders (max.gen 0)
    # locus 1: alleles 0 x3, 1 x3 -> tie broken by name: "0" is the zero allele; missing pairs -> 3
    assert [r[5] for r in rows] == ["1", "3", "3", "0", "2"]
    # locus 2: likelihood triples normalised to 4 decimals; the expected "1" allele is the more
    # frequent one here, so the triples are reversed to keep "0" the major allele
    assert rows[0][6] == "0.3333,0.3333,0.3333" and rows[3][6] == "0.1,0.9,0"
    # locus 3: haplotype strings are alleles too: TATA (4) is the major one
    assert rows[4][7] == "0" and rows[0][7] == "2"
    prior = dict(ln.split(" ", 1) for ln in (tmp_path / "prior.txt").read_text().splitlines())
    assert prior["nIndiv"] == "5" and prior["nSNP"] == "4" and prior["nGen"] == "4" and prior["maxID"] == "6"
    assert prior["epsilon"].split() == ["0.02"] * 4
    assert len(prior["aFreq"].split()) == 4 and prior["obsFrac"].split() == ["0.8", "0.8", "0", "0.8"]
    assert info["n.gen"] == 4


def test_marker_info_overrides_error_and_frequency(tmp_path):
    p = tmp_path / "genotype.txt"
    p.write_text(FIXTURE)
    m = tmp_path / "marker_info.txt"
    m.write_text("name L1 L2 L3 L4\ngerror 0.2 0.1 0.05 0.01\nafreq 0.01 0.04 0.1 0.23\n")
    param = {"geno.path": p, "output.path": tmp_path, "observe.frac": 0.8, "max.unobs": 1, "max.gen": 0,
             "min.age": 1, "max.age": 1, "geno.err": 0.02, "marker.path": m}
    rp.write_intermed_geno(param)
    prior = dict(ln.split(" ", 1) for ln in (tmp_path / "prior.txt").read_text().splitlines())
    assert prior["epsilon"].split() == ["0.2", "0.1", "0.05", "0.01"]
    assert prior["aFreq"].split() == ["0.01", "0.04", "0.1", "0.23"]
    m.write_text("name L1 L2\ngerror 0.2 0.1\n")
    with pytest.raises(ValueError, match="names 2 loci"):
        rp.write_intermed_geno(param)


def test_run_pedfac_end_to_end_on_the_cpu_build(tmp_path):
    d = synth.simulate(150, 120, soft=False, seed=3)
    obs = np.flatnonzero(d.observed)
    lines = []
    for i in obs:
        al = []
        for c in d.codes[i]:
            al += {0: ["0", "0"], 1: ["0", "1"], 2: ["1", "1"], 3: ["-1", "-1"]}[int(c)]
        lines.append(f"id{i} 1 {d.sex[i]} {2000 - d.gen[i]} " + " ".join(al))
    g = tmp_path / "genotype.txt"
    g.write_text("\n".join(lines) + "\n")
    res = rp.run_pedfac(g, output_path=str(tmp_path), n_iter=3, binary=REF)
    run = res["run"]
    for f in ["geno.txt", "prior.txt", "rand", "id.json", "out/ped.txt", "ped_sample.txt", "parent_assignment.txt",
              "grandparent_assignment.txt", "fullsib_assignment.txt", "halfsib_assignment.txt"]:
        assert (run / f).exists(), f
    best = {}
    for kid, pa, ma, prob in res["parents"]:
        best.setdefault(kid, (pa, ma, prob))
    ok = tot = 0
    for i in obs:
        if d.pa[i] >= 0 and d.observed[d.pa[i]] and d.observed[d.ma[i]]:
            tot += 1
            ok += set(best[f"id{i}"][:2]) == {f"id{d.pa[i]}", f"id{d.ma[i]}"}
    assert tot > 10 and ok / tot > 0.85
    with pytest.raises(ValueError, match="observe.frac"):
        rp.run_pedfac(g, output_path=str(tmp_path), observe_frac=1.5, binary=REF)


def test_min_age_only_changes_the_generation_column(tmp_path):
    """R/renderGeno.R:226-238 writes no maxAge / minAge line: min.age and max.age act on the R side
    only, through gen = floor((max year - year) / min.age). The sampler then sees generation units and
    its own age window stays at the binary's defaults (1, 1)."""
    p = tmp_path / "genotype.txt"
    p.write_text(FIXTURE)
    param = {"geno.path": p, "output.path": tmp_path, "observe.frac": 0.8, "max.unobs": 1, "max.gen": 0,
             "min.age": 2, "max.age": 2, "geno.err": 0.02, "marker.path": ""}
    rp.write_intermed_geno(param)
    rows = [ln.split() for ln in (tmp_path / "geno.txt").read_text().splitlines()]
    assert [r[3] for r in rows] == ["1", "1", "1", "0", "0"]                    # floor((1990 - year) / 2)
    keys = [ln.split(" ", 1)[0] for ln in (tmp_path / "prior.txt").read_text().splitlines()]
    assert "maxAge" not in keys and "minAge" not in keys
    prior = dict(ln.split(" ", 1) for ln in (tmp_path / "prior.txt").read_text().splitlines())
    assert prior["nGen"] == "2"
