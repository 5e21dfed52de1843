"""The input side of the likelihood API (netrax_b200/msa_io.py, SURVEY §8f f4 "data formats"): raxml-ng model strings and
partition files, FASTA / PHYLIP alignments, pattern compression and raxml-ng's starting values — checked against the rules
in RAXML/Model.cpp / PLLMOD/util/models_dna.c / PLLMOD/msa/pll_msa.c (restated independently below where a brute-force form
exists) and, end to end, against the golden lnLs of the reference's own fixtures."""
import os

import numpy as np
import pytest

from helpers import FIX, FIXTURE_PAIRS, load_fixture, load_golden
from netrax_b200.msa_io import (DNA_MODELS, apply_model_state, build_partitions, msa_stats, parse_model, parse_partition_file,
                                read_msa, read_phylip)
from netrax_b200.network_io import parse_extended_newick, read_fasta
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES

STD_MODEL = "GTR{1/2.5/0.8/1.2/3.0/1}+FU{0.3/0.2/0.2/0.3}+G4{0.5}"   # the model of the golden fixtures (tests/helpers.py)


def test_model_names_symmetries_and_free_parameter_counts():
    """PLLMOD/util/models_dna.c:37-100 + Model::num_free_params (RAXML/Model.cpp:1014-1055)."""
    want = {"JC": 0, "K80": 1, "F81": 3, "HKY": 4, "TN93": 5, "K81uf": 5, "TIM2": 3, "TVM": 7, "SYM": 5, "GTR": 8}
    for name, k in want.items():
        assert parse_model(name).free_params() == k, name
    assert len(DNA_MODELS) == 22
    assert parse_model("TrN").name == "TN93" and parse_model("dna").name == "GTR" and parse_model("gtr").name == "GTR"
    m = parse_model("GTR+G")
    assert (m.rate_cats, m.alpha, m.alpha_mode, m.gamma_mode, m.free_params()) == (4, 1.0, "ML", 0, 9)
    m = parse_model("GTR+G8a{0.3}+I+FC")
    assert (m.rate_cats, m.alpha, m.alpha_mode, m.gamma_mode, m.pinv_mode, m.freq_mode) == (8, 0.3, "user", 1, "ML", "empirical")
    assert m.free_params() == 3 + 5 + 1   # alpha is user-fixed
    m = parse_model("HKY{1/2.5}+FE+I{0.25}+B{1.5}")
    np.testing.assert_allclose(m.subst_rates, [1 / 1.0, 2.5, 1, 1, 2.5, 1])   # expanded through the symmetry, normalised by the last rate
    assert (m.freq_mode, m.pinv, m.pinv_mode, m.brlen_scaler, m.free_params()) == ("equal", 0.25, "user", 1.5, 0)
    m = parse_model("GTR{2/4/6/8/10/2}")
    np.testing.assert_allclose(m.subst_rates, [1, 2, 3, 4, 5, 1])   # set_user_srates normalises by the last rate
    m = parse_model("LG+G+F")
    assert (m.states, m.rate_cats, m.freq_mode, m.rate_mode, m.free_params()) == (20, 4, "empirical", "model", 19 + 1)
    assert parse_model("LG").free_params() == 0
    for bad, msg in (("FOO", "Invalid model name"), ("GTR+R4", r"\+R"), ("GTR+ASC_LEWIS", "ascertainment"), ("GTR{1/2}", "expected 6"),
                     ("GTR+FU{0.5/0.5}", "user frequencies"), ("GTR+X", "Invalid model options"), ("LG4X", "FreeRate"), ("PROTGTR", "BFGS")):
        with pytest.raises(ValueError, match=msg):
            parse_model(bad)


def test_phylip_sequential_and_interleaved_equal_fasta():
    fa = read_fasta(open(os.path.join(FIX, "small_fake_alignment.txt")).read())
    n, L = len(fa), len(next(iter(fa.values())))
    seq = f"{n} {L}\n" + "".join(f"{k}  {v}\n" for k, v in fa.items())
    cut = 7
    inter = f" {n}   {L}\n" + "".join(f"{k}\t{v[:cut]}\n" for k, v in fa.items()) + "\n" + "".join(f"{v[cut:cut + 5]} {v[cut + 5:]}\n" for v in fa.values())
    wrapped = f"{n} {L}\n" + "".join(f"{k} {v[:cut]}\n{v[cut:]}\n" for k, v in fa.items())
    for text in (seq, inter, wrapped):
        assert read_phylip(text) == fa
        assert read_msa(text) == fa
    assert read_msa(open(os.path.join(FIX, "small_fake_alignment.txt")).read()) == fa
    with pytest.raises(ValueError, match="shorter|expected"):
        read_phylip(f"{n} {L + 3}\n" + "".join(f"{k} {v}\n" for k, v in fa.items()))


def _stats_bruteforce(masks, weights, states):
    """pllmod_msa_compute_features written cell by cell (PLLMOD/msa/pll_msa.c:742-830)."""
    freqs, gaps, inv_w = np.zeros(states), 0.0, 0.0
    for j in range(masks.shape[1]):
        union = 0
        for i in range(masks.shape[0]):
            st = int(masks[i, j])
            pop = bin(st).count("1")
            if pop == states:
                gaps += weights[j]
                continue
            union |= st
            for k in range(states):
                if st >> k & 1:
                    freqs[k] += weights[j] / pop
        if bin(union).count("1") == 1:
            inv_w += weights[j]
    return freqs / (weights.sum() * masks.shape[0] - gaps), inv_w / weights.sum()


def test_msa_stats_match_the_cellwise_restatement():
    rng = np.random.default_rng(3)
    masks = rng.choice(np.array([1, 2, 4, 8, 5, 10, 15, 15], np.uint32), size=(6, 200))
    masks[:, :40] = masks[0:1, :40]                 # some invariant columns (incl. gap-only ones, which are NOT invariant)
    masks[2, :40] = 15
    w = rng.integers(1, 5, size=200).astype(np.uint32)
    f, inv = msa_stats(masks, w, 4)
    f0, inv0 = _stats_bruteforce(masks, w.astype(float), 4)
    np.testing.assert_allclose(f, f0, rtol=1e-13)
    assert inv == pytest.approx(inv0, rel=1e-13) and 0.1 < inv < 0.4
    assert f.sum() == pytest.approx(1.0, rel=1e-12)
    f, inv = msa_stats(np.array([[1, 1, 2], [1, 15, 8]], np.uint32), None, 4)   # A A C / A - T
    np.testing.assert_allclose(f, [3 / 5, 1 / 5, 0, 1 / 5])
    assert inv == pytest.approx(2 / 3)


@pytest.mark.parametrize("name", ["small", "celine", "three_reticulations"])
def test_files_to_partitions_reproduce_the_golden_fixture_likelihoods(name):
    """Network file + alignment file + model string -> the same partition the fixture loader builds by hand, and through the
    oracle the lnL pinned in tests/golden/netrax_fixtures_golden.json (values of the reference's own fixtures)."""
    from oracle import oracle
    nw, aln = FIXTURE_PAIRS[name]
    net = parse_extended_newick(open(os.path.join(FIX, nw)).read())
    msa = read_msa(open(os.path.join(FIX, aln)).read())
    parts, specs = build_partitions(msa, net.tip_labels, STD_MODEL)
    _, want = load_fixture(nw, aln)
    p = parts[0]
    assert np.array_equal(p.tip_masks, want.tip_masks) and np.array_equal(p.pattern_weights, want.pattern_weights)
    np.testing.assert_allclose(p.freqs, DNA_FREQS); np.testing.assert_allclose(p.subst, GTR_RATES)
    np.testing.assert_allclose(p.rates, GAMMA4_ALPHA05, rtol=1e-7)   # synth's constant is the 10-digit rounding of these rates
    assert specs[0].free_params() == 0
    e = oracle.make_engine("port", net, parts)
    ref_lnl = load_golden("netrax_fixtures_golden.json")["cases"][f"{name}/AVERAGE"]["lnl"]
    lnl = e.computeLoglikelihood(0, 1)
    w = oracle.make_engine("port", net, [want])
    assert lnl == pytest.approx(w.computeLoglikelihood(0, 1), rel=1e-7)
    assert lnl == pytest.approx(ref_lnl, rel=1e-7)
    e.close(); w.close()


def test_partition_file_ranges_models_and_starting_values():
    from oracle import oracle
    nw, aln = FIXTURE_PAIRS["celine"]
    net = parse_extended_newick(open(os.path.join(FIX, nw)).read())
    msa = read_msa(open(os.path.join(FIX, aln)).read())
    L = len(next(iter(msa.values())))
    half = L // 2
    text = f"""# two genes, the second one by codon position
    GTR+G+FC, gene1 = 1-{half}
    HKY+I, gene2_12 = {half + 1}-{L}\\3, {half + 2}-{L}\\3
    JC, gene2_3 = {half + 3}-{L}/3
    """
    prs = parse_partition_file(text)
    assert [p.name for p in prs] == ["gene1", "gene2_12", "gene2_3"]
    cols = [p.columns(L) for p in prs]
    assert sorted(np.concatenate(cols).tolist()) == list(range(L))        # every column exactly once
    parts, specs = build_partitions(msa, net.tip_labels, text)
    assert [int(p.pattern_weights.sum()) for p in parts] == [len(c) for c in cols]
    assert [p.rate_cats for p in parts] == [4, 1, 1]
    full = np.stack([np.frombuffer(msa[t].encode(), dtype="S1") for t in net.tip_labels])
    from netrax_b200.msa_io import encode
    f0, _ = msa_stats(np.stack([encode(b"".join(r[cols[0]]).decode(), "DNA") for r in full]), None, 4)
    np.testing.assert_allclose(parts[0].freqs, f0, rtol=1e-12)              # +FC: empirical frequencies of THAT partition
    np.testing.assert_allclose(parts[1].freqs, 0.25)                        # ML frequencies start equal
    _, inv = msa_stats(parts[1].tip_masks, parts[1].pattern_weights, 4)
    assert specs[1].pinv == pytest.approx(inv / 2) and specs[1].pinv_mode == "ML"   # half the empirical proportion
    np.testing.assert_allclose(parts[0].rates, oracle.make_engine("port", net, parts[:1]).api.gamma_rates(1.0, 4), rtol=1e-12)
    k = sum(s.free_params() for s in specs)
    assert k == (3 + 5 + 1) + (3 + 1 + 1) + 0
    for kind in ["port"] + (["ref"] if oracle.have_ref() else []):
        e = oracle.make_engine(kind, net, parts)
        assert apply_model_state(e, specs) == k
        assert e.get_alpha(0) == 1.0 and e.get_pinv(1) == pytest.approx(inv / 2) and e.get_pinv(0) == 0.0
        lnl = e.computeLoglikelihood(0, 1)
        n = float(L) * net.num_tips
        assert e.scoreNetwork() == pytest.approx(-2 * lnl + (k + net.num_reticulations + net.num_edges) * np.log(n), rel=1e-12)
        e.close()
    with pytest.raises(ValueError, match="outside the alignment"):
        build_partitions(msa, net.tip_labels, f"GTR, p = 1-{L + 1}")
    with pytest.raises(ValueError, match="missing from the alignment"):
        build_partitions(msa, list(net.tip_labels) + ["nobody"], "GTR")


def test_score_only_flow_over_the_oracle_engine():
    """The reference's --score_only flow (src/main.cpp:287-326) from files, driven here over the CPU oracle as the engine
    (the CLI, scripts/score_network.py, passes the CUDA engine; tests/test_gpu_parity.py runs the same flow on the device
    and compares).  Checks the bookkeeping: the BIC never gets worse, AIC / AICc / BIC agree with ComplexityScoring.cpp's
    formulae, the written network parses back with the optimised lengths."""
    from netrax_b200.score import score_only
    from oracle import oracle
    nw, aln = FIXTURE_PAIRS["small"]
    lines = []
    res = score_only(lambda net, parts, **kw: oracle.make_engine("port", net, parts, **kw), open(os.path.join(FIX, nw)).read(),
                     open(os.path.join(FIX, aln)).read(), "GTR{1/2.5/0.8/1.2/3.0/1}+FC+G", log=lines.append)
    assert res["bic"] <= res["start_bic"] + 1e-3 and res["logl"] >= res["start_logl"] - 1e-6
    k = res["param_count"]
    assert res["model_params"] == 3 + 1 and k == 4 + res["reticulations"] + parse_extended_newick(res["network"]).num_edges
    assert res["aic"] == pytest.approx(-2 * res["logl"] + 2 * k)
    assert 0.02 < res["alphas"][0] <= 100.0
    back = parse_extended_newick(res["network"])
    assert back.num_reticulations == res["reticulations"] and np.all(back.edge_length >= 1e-6)
    assert any(l.startswith("BIC Score: ") for l in lines) and any(l.startswith("Number of reticulations: 1") for l in lines)


GP = load_golden("libpll_protein_models_golden.json")


@pytest.mark.parametrize("model", list(GP["models"]))
def test_protein_model_names_reproduce_libpll_golden_lnl(model):
    """FASTA text + ``<NAME>+G4{1.0}`` for each of libpll's 20 empirical protein matrices -> partitions -> lnL equal to libpll's
    own regression output (test/out/protein-models.out, 6 decimals): pins aa_models.json, the name lookup (case-insensitive,
    ``JTT-DCMut`` with its dash) and the amino-acid encoding of the input layer."""
    from oracle import oracle
    b0, b1 = GP["branch_lengths"]
    net = parse_extended_newick(f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{b0 / 2},(T3:{b1},T4:{b1})X7:{b0 / 2});")
    fasta = "".join(f">T{i}\n{s}\n" for i, s in enumerate(GP["tips"]))
    parts, specs = build_partitions(read_msa(fasta), net.tip_labels, f"{model}+G{GP['ncats']}{{{GP['alpha']}}}", compress=False)
    assert specs[0].name == model.upper() and specs[0].free_params() == 0
    np.testing.assert_array_equal(parts[0].subst, np.asarray(GP["models"][model]["rates"]))
    eng = oracle.make_engine("port", net, parts)
    assert abs(eng.computeLoglikelihood(0, 1) - GP["models"][model]["logl"]) < 2e-6
    eng.close()
    parts_c, _ = build_partitions(read_msa(fasta), net.tip_labels, f"{model}+G{GP['ncats']}{{{GP['alpha']}}}")   # compressed: same lnL
    eng = oracle.make_engine("port", net, parts_c)
    assert abs(eng.computeLoglikelihood(0, 1) - GP["models"][model]["logl"]) < 2e-6
    eng.close()


def test_lg4m_is_a_per_category_mixture_through_set_submodels():
    """``LG4M`` (PLLMOD/util/models_aa.c:103-105): four matrices, category c uses matrix c; apply_model_state hands them to
    set_submodels (raxml-ng's ratecat_submodels -> libpll params_indices).  Checked against the sum over categories of
    single-matrix, single-rate evaluations."""
    from oracle import oracle
    from netrax_b200._capi import Partition
    kind = "ref" if oracle.have_ref() else "port"
    b0, b1 = GP["branch_lengths"]
    net = parse_extended_newick(f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{b0 / 2},(T3:{b1},T4:{b1})X7:{b0 / 2});")
    fasta = "".join(f">T{i}\n{s}\n" for i, s in enumerate(GP["tips"]))
    parts, specs = build_partitions(read_msa(fasta), net.tip_labels, "LG4M+G4{0.8}", compress=False)
    ms = specs[0]
    assert ms.ratecat_submodels == [0, 1, 2, 3] and len(ms.submodels) == 4 and ms.free_params() == 0
    eng = oracle.make_engine(kind, net, parts)
    apply_model_state(eng, specs)
    lnl = eng.computeLoglikelihood(0, 1)
    from helpers import mixture_lnl_by_categories
    short = Partition(20, 4, parts[0].tip_masks[:, :12], parts[0].freqs, parts[0].subst, parts[0].rates)   # 12 sites: 48 tiny engines
    e = oracle.make_engine(kind, net, [short])
    apply_model_state(e, specs)
    want = mixture_lnl_by_categories(lambda n, p: oracle.make_engine(kind, n, [p]), net, short, ms.ratecat_submodels,
                                     [f for _, f in ms.submodels], [r for r, _ in ms.submodels])
    assert e.computeLoglikelihood(0, 1) == pytest.approx(want, rel=1e-10)
    e.close()
    single = oracle.make_engine(kind, net, parts)
    assert abs(single.computeLoglikelihood(0, 1) - lnl) > 1e-3   # without the mixture: matrix 0 for every category
    eng.close(); single.close()


def test_score_network_cli_refuses_to_run_without_a_gpu():
    """scripts/score_network.py drives the CUDA engine only: on a box without a device it stops with the engine's message
    (no CPU fallback, nothing under oracle/ is imported)."""
    import subprocess
    import sys
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "scripts", "score_network.py"), "--msa", os.path.join(FIX, "small_fake_alignment.txt"),
                        "--start_network", os.path.join(FIX, "small.nw"), "--model", "GTR{1/2.5/0.8/1.2/3.0/1}+FC+G"], capture_output=True, text=True, timeout=300)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    for rel in (("scripts", "score_network.py"), ("netrax_b200", "score.py"), ("netrax_b200", "msa_io.py"), ("netrax_b200", "network_io.py")):
        src = open(os.path.join(root, *rel)).read()
        assert "from oracle" not in src and "import oracle" not in src, rel


def test_score_only_refuses_ml_estimated_rates_and_frequencies():
    """The reference's optimizeModel fits ML rates / frequencies with L-BFGS-B (out of scope): plain GTR+G must not be scored
    silently at its start values."""
    from netrax_b200.score import UnoptimisedModelError, score_only
    from oracle import oracle
    nw, aln = FIXTURE_PAIRS["small"]
    net_text, msa_text = open(os.path.join(FIX, nw)).read(), open(os.path.join(FIX, aln)).read()
    factory = lambda net, parts, **kw: oracle.make_engine("port", net, parts, **kw)   # noqa: E731
    with pytest.raises(UnoptimisedModelError, match="substitution rates"):
        score_only(factory, net_text, msa_text, "GTR+G", log=None)
    lines = []
    res = score_only(factory, net_text, msa_text, "GTR+G", log=lines.append, optimize=False, allow_unoptimised_ml_params=True)
    assert res["unoptimised_ml_params"] and any(l.startswith("WARNING: model parameters") for l in lines)
    assert score_only(factory, net_text, msa_text, "GTR{1/2.5/0.8/1.2/3.0/1}+FC+G", log=None, optimize=False)["unoptimised_ml_params"] == []
