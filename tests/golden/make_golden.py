#!/usr/bin/env python
"""Regenerates everything under tests/golden/ — run HERE (container with /root/reference), never on the GPU box.

1. copies the reference's own likelihood-test fixtures (test/sample_networks/*.nw + *_alignment.txt: DATA, not
   source) into tests/golden/fixtures/;
2. parses the golden stdout of libpll's regression suite (LIBPLL/../test/out/derivatives.out, pinv = 0 blocks;
   inline data of test/src/derivatives.c:60-150) into libpll_derivatives_golden.json;
3. runs the REAL forked libpll (oracle/_ref, kind "reference") under the restated NetRAX layer on every
   fixture pairing of test/src/LikelihoodTest.cpp:282-360 and stores network lnL, per-displayed-tree lnL /
   log-prob, scaler sums and CLV checksums into netrax_fixtures_golden.json.
"""
import json, os, re, shutil, sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
REF = "/root/reference"
LIBPLL_TEST = REF + "/libs/raxml-ng/libs/pll-modules/libs/libpll/test"

import numpy as np
from helpers import FIXTURE_PAIRS, load_fixture, fixture_summary  # noqa: E402
from oracle import oracle  # noqa: E402


def copy_fixtures():
    src = REF + "/test/sample_networks"
    for f in sorted(os.listdir(src)):
        shutil.copy(os.path.join(src, f), os.path.join(HERE, "fixtures", f))


def _parse_derivative_blocks(path, pinv=False):
    blocks, cur = [], None
    for line in open(path):
        m = re.match(r"\s*TEST alpha\(ncats\) =\s*([\d.]+)\(\s*(\d+)\) ; pinv = ([\d.]+)", line)
        if m:
            cur = {"alpha": float(m.group(1)), "ncats": int(m.group(2)), "pinv": float(m.group(3)), "inner": [], "tip": []}
            blocks.append(cur)
            continue
        m = re.match(r"Branch(\(Tip\))?\s+([\d.]+) :\s+(\S+)\s+(\S+)\s+(\S+)", line)
        if m and cur is not None:
            cur["tip" if m.group(1) else "inner"].append([float(m.group(2)), float(m.group(3)), float(m.group(4)), float(m.group(5))])
    return [b for b in blocks if (b["pinv"] > 0.0) == pinv]


def parse_libpll_pinv():
    """The pinv = 0.3 / 0.6 / 0.9 blocks of derivatives.out and derivatives-oddstates.out (+I: proportion of invariant sites)."""
    for src, ref, dst in (("derivatives.out", "libpll_derivatives_golden.json", "libpll_derivatives_pinv_golden.json"),
                          ("derivatives-oddstates.out", "libpll_derivatives_oddstates_golden.json", "libpll_derivatives_oddstates_pinv_golden.json")):
        base = json.load(open(os.path.join(HERE, ref)))
        base["blocks"] = _parse_derivative_blocks(LIBPLL_TEST + "/out/" + src, pinv=True)
        base["source"] = base["source"].replace("pinv=0 blocks", "pinv > 0 blocks")
        json.dump(base, open(os.path.join(HERE, dst), "w"), indent=1)
    return len(base["blocks"])


def parse_libpll_golden():
    blocks = _parse_derivative_blocks(LIBPLL_TEST + "/out/derivatives.out")
    json.dump({"source": "libpll test/out/derivatives.out (pinv=0 blocks); columns: branch, edge lnL, d(-lnL)/dt, d2(-lnL)/dt2",
               "tips": ["WAACTCGCTA--ATTCTAAT", "CACCATGCTA--ATTGTCTT", "AG-C-TGCAG--CTTCTACT", "CGTCTTGCAA--AT-C-AAG", "CGACTTGCCA--AT-T-AAG"],
               "freqs": [0.3, 0.4, 0.1, 0.2], "subst": [1, 2.5, 1, 1, 2.5, 1], "branch_lengths": [0.1, 0.2],
               "blocks": blocks}, open(os.path.join(HERE, "libpll_derivatives_golden.json"), "w"), indent=1)
    return len(blocks)


def parse_libpll_oddstates():
    """libpll test/out/derivatives-oddstates.out: the same experiment with FIVE states (states_padded = 8), inline data of
    test/src/derivatives-oddstates.c:100-145; characters through odd5_map (test/src/common.c:8-19: A..D = states 0..3,
    E = C|D, gap = all five)."""
    blocks = _parse_derivative_blocks(LIBPLL_TEST + "/out/derivatives-oddstates.out")
    json.dump({"source": "libpll test/out/derivatives-oddstates.out (pinv=0 blocks); 5 states; columns: branch, edge lnL, d(-lnL)/dt, d2(-lnL)/dt2",
               "states": 5, "char_masks": {"A": 1, "B": 2, "C": 4, "D": 8, "E": 12, "-": 31},
               "tips": ["DAACBCECBA--ABBCBAAB", "CACCABECBA--ABBEBCBB", "AE-C-BECAE--CBBCBACB", "CEBCBBECAA--AB-C-AAE", "CEACBBECCA--AB-B-AAE"],
               "freqs": [0.3, 0.25, 0.1, 0.2, 0.15],
               "subst": [1.452176, 0.937951, 0.462880, 0.617729, 1.745312, 0.937951, 0.462880, 0.617729, 1.745312, 1.000000],
               "branch_lengths": [0.1, 0.2], "blocks": blocks}, open(os.path.join(HERE, "libpll_derivatives_oddstates_golden.json"), "w"), indent=1)
    return len(blocks)


def parse_libpll_alpha_cats():
    """libpll test/out/alpha-cats.out: discrete-Gamma rates (MEAN and MEDIAN) and an edge lnL for 9 alphas x 5 category counts."""
    lines = open(LIBPLL_TEST + "/out/alpha-cats.out").read().splitlines()
    blocks, i = [], 0
    while i < len(lines):
        m = re.match(r"\s*TEST alpha\(ncats\) =\s*([\d.]+)\(\s*(\d+)\), mode = (\w+)", lines[i])
        if m:
            j = i + 1
            while not lines[j].strip():
                j += 1
            # the block header prints MODENAME(m) with m the LOOP INDEX over modes[] = {MEDIAN, MEAN} (test/src/alpha-cats.c:38-39,135-136),
            # so its label is swapped; the summary lines at the end use MODENAME(modes[m]) and are right
            mode = "MEDIAN" if m.group(3) == "MEAN" else "MEAN"
            blocks.append({"alpha": float(m.group(1)), "ncats": int(m.group(2)), "mode": mode, "rates": [float(x) for x in lines[j].split()]})
            i = j
        i += 1
    logl = {}
    for l in lines:
        m = re.match(r"ti/tv:alpha\(ncats\) =\s*([\d.]+)\(\s*(\d+)\), mode =\s*(\w+)\((\d)\)\s+logL:\s+(\S+)", l)
        if m:
            logl[(float(m.group(1)), int(m.group(2)), m.group(3))] = float(m.group(5))
    for b in blocks:
        b["logl"] = logl[(b["alpha"], b["ncats"], b["mode"])]
    json.dump({"source": "libpll test/out/alpha-cats.out (test/src/alpha-cats.c: 5 taxa x 20 sites, HKY titv 2.5, pi=(.3,.4,.1,.2), branch lengths m0=0.1 m1=0.2; per (alpha, categories, mode): discrete Gamma rates (6 decimals) and the edge lnL between clv6 and clv7 (6 decimals))",
               "tips": ["WAACTCGCTA--ATTCTAAT", "CACCATGCTA--ATTGTCTT", "AG-C-TGCAG--CTTCTACT", "CGTCTTGCAA--AT-C-AAG", "CGACTTGCCA--AT-T-AAG"],
               "freqs": [0.3, 0.4, 0.1, 0.2], "subst": [1, 2.5, 1, 1, 2.5, 1], "branch_lengths": [0.1, 0.2], "blocks": blocks},
              open(os.path.join(HERE, "libpll_alpha_cats_golden.json"), "w"), indent=0)
    return len(blocks)


def parse_libpll_protein_models():
    """libpll test/out/protein-models.out: edge lnL under 20 empirical amino-acid models; the model DATA tables come from the
    reference's compiled libpll (oracle/_ref/libpll_ref.so symbols pll_aa_rates_* / pll_aa_freqs_*, LIBPLL/pll.h:553-600)."""
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libpll_ref.so"))
    names = ["Dayhoff", "LG", "DCMut", "JTT", "MtREV", "WAG", "RtREV", "CpREV", "VT", "Blosum62", "MtMam", "MtArt", "MtZoa", "PMB", "HIVb",
             "HIVw", "JTT-DCMut", "FLU", "StmtREV", "DEN"]
    logl = {}
    for l in open(LIBPLL_TEST + "/out/protein-models.out"):
        m = re.match(r"Log-L \((\S+)\): (\S+)", l)
        if m:
            logl[m.group(1)] = float(m.group(2))
    src = open(LIBPLL_TEST + "/src/protein-models.c").read()
    seqs = re.findall(r'pll_set_tip_states \(partition, \d, pll_map_aa,\s*"([^"]+)"\);', src)
    models = {}
    for n in names:
        s = {"JTT-DCMut": "jttdcmut"}.get(n, n.lower())
        models[n] = {"rates": list((C.c_double * 190).in_dll(lib, "pll_aa_rates_" + s)), "freqs": list((C.c_double * 20).in_dll(lib, "pll_aa_freqs_" + s)),
                     "logl": logl[n]}
    json.dump({"source": "libpll test/out/protein-models.out (test/src/protein-models.c: 5 taxa x 113 amino-acid sites, Gamma alpha = 1 with 4 MEAN categories, branch lengths m0=0.1 m1=0.2, edge lnL between clv6=((t0,t1),t2) and clv7=(t3,t4) over matrix 0, 6 decimals); exchangeabilities / frequencies = the model DATA tables pll_aa_rates_* / pll_aa_freqs_* of the reference's compiled libpll",
               "alpha": 1.0, "ncats": 4, "branch_lengths": [0.1, 0.2], "tips": seqs, "aa_order": "ARNDCQEGHILKMFPSTWYV", "models": models},
              open(os.path.join(HERE, "libpll_protein_models_golden.json"), "w"))
    return len(models)


def _parse_matrix_block(lines, i, states, cats):
    """pll_show_pmatrix (LIBPLL/output.c): per category `states` rows of `states` numbers, then a blank line."""
    out = np.zeros((cats, states, states))
    for c in range(cats):
        for r in range(states):
            out[c, r] = [float(x) for x in lines[i].split()]
            i += 1
        i += 1
    return out, i


def parse_libpll_pmatrix():
    """libpll test/out/pmatrix.out (test/src/pmatrix.c): P-matrices printed with 9 decimals for DNA (4), PROT (20) and ODD (5)
    states x 3 frequency vectors (equal / skewed / extreme) x 3 exchangeability vectors (equal / skewed / extreme: 1e-3 .. 1e3)
    x 5 branch lengths (1e-6 .. 100) x 4 category rates (1e-31, 1e-6, 1, 100).  The frequency and rate vectors are printed
    with 6 decimals only, so the tests rebuild them from the formulas of init_freqs / init_rates (:112-170) and use the
    printed ones as a check."""
    lines = open(LIBPLL_TEST + "/out/pmatrix.out").read().splitlines()
    states_of = {"DNA": 4, "PROT": 20, "ODD": 5}
    out, i, n = {}, 0, 0
    while i < len(lines):
        m = re.match(r"datatype = (\w+)", lines[i])
        if not m:
            i += 1
            continue
        dt = m.group(1)
        S = states_of[dt]
        freqs = [float(x) for x in lines[i + 1].split("[")[1].split("]")[0].split()]
        subst = [float(x) for x in lines[i + 2].split("[")[1].split("]")[0].split()]
        i += 3
        mats, brlens = [], []
        for b in range(5):
            mm = re.match(r"P-matrix: (\d+), brlen = ([\d.]+)", lines[i])
            assert mm and int(mm.group(1)) == b, lines[i]
            brlens.append(float(mm.group(2)))
            mat, i = _parse_matrix_block(lines, i + 1, S, 4)
            mats.append(mat)
        k = sum(1 for key in out if key.startswith(dt + "_P_"))
        out[f"{dt}_P_{k}"] = np.stack(mats)                 # [5 branches][4 cats][S][S]
        out[f"{dt}_freqs_{k}"] = np.array(freqs)
        out[f"{dt}_subst_{k}"] = np.array(subst)
        n += 1
    out["branch_lengths"] = np.array([1e-6, 1e-2, 0.2, 1.0, 100.0])
    out["cat_rates"] = np.array([1e-31, 1e-6, 1.0, 100.0])
    np.savez_compressed(os.path.join(HERE, "libpll_pmatrix_golden.npz"), **out)
    return n


def parse_libpll_hky():
    """libpll test/out/hky.out (test/src/hky.c): the 5-taxon / 20-site data set of derivatives.c under HKY with 10 ti/tv
    ratios, pi = (.3,.4,.1,.2), Gamma alpha = 1 (4 MEAN categories), branch lengths m0 = 0.1, m1 = 0.2: per ratio the
    P-matrices (4 decimals), the inner CLVs 5 = (t0,t1), 6 = (clv5,t2), 7 = (t3,t4) (5 decimals) and the edge lnL between
    clv6 and clv7 over matrix 0 (4 decimals)."""
    lines = open(LIBPLL_TEST + "/out/hky.out").read().splitlines()
    titv, P, clvs, logl = [], [], [], []
    i = 0
    while i < len(lines):
        m = re.match(r"\s*TEST ti/tv = ([\d.]+)", lines[i])
        if m:
            titv.append(float(m.group(1)))
            P.append([])
            clvs.append({})
        m = re.match(r"\[(\d+)\] P-matrix for branch length ([\d.]+)", lines[i])
        if m:
            mat, i = _parse_matrix_block(lines, i + 1, 4, 4)
            P[-1].append(mat)
            continue
        m = re.match(r"\[(\d+)\] CLV (\d): \[(.*)\]", lines[i])
        if m:
            nums = [float(x) for x in re.findall(r"[-+]?\d+\.\d+", m.group(3))]
            clvs[-1][int(m.group(2))] = np.array(nums).reshape(20, 4, 4)
        m = re.match(r"ti/tv:\s+([\d.]+)\s+logL:\s+(\S+)", lines[i])
        if m:
            logl.append(float(m.group(2)))
        i += 1
    assert len(titv) == len(logl) == 10
    np.savez_compressed(os.path.join(HERE, "libpll_hky_golden.npz"),
                        titv=np.array([0.175, 1, 1.5, 2.25, 2.725, 4, 7.125, 8.19283745, 9.73647382, 10]),   # hky.c:28-30 (printed with 4 decimals)
                        titv_printed=np.array(titv), P=np.array(P), logl=np.array(logl),
                        clv5=np.stack([c[5] for c in clvs]), clv6=np.stack([c[6] for c in clvs]), clv7=np.stack([c[7] for c in clvs]))
    return len(titv)


def netrax_golden():
    out = {}
    for name, (nw, aln) in FIXTURE_PAIRS.items():
        for variant in (0, 1):
            net, part = load_fixture(nw, aln)
            eng = oracle.make_engine("ref", net, [part], variant=variant)
            out[f"{name}/{'AVERAGE' if variant == 0 else 'BEST'}"] = fixture_summary(eng)
    json.dump({"generator": "tests/golden/make_golden.py with oracle kind=reference (real forked libpll AVX2+PATTERN_TIP)",
               "model": "GTR(1,2.5,0.8,1.2,3,1) pi=(.3,.2,.2,.3) G4 alpha=0.5", "cases": out},
              open(os.path.join(HERE, "netrax_fixtures_golden.json"), "w"), indent=1)
    return len(out)


if __name__ == "__main__":
    copy_fixtures()
    print("libpll golden blocks:", parse_libpll_golden())
    print("libpll odd-states blocks:", parse_libpll_oddstates())
    print("libpll +I blocks:", parse_libpll_pinv())
    print("libpll alpha-cats blocks:", parse_libpll_alpha_cats())
    print("libpll protein models:", parse_libpll_protein_models())
    print("libpll pmatrix evaluations:", parse_libpll_pmatrix())
    print("libpll hky ratios:", parse_libpll_hky())
    print("netrax golden cases:", netrax_golden())
