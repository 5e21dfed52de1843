"""Oracle self-checks at the NetRAX layer (CPU only).

* port restatement vs golden files generated with the REAL forked libpll underneath (kind "reference");
* the reference's own invariants: improved == naive per-displayed-tree evaluation
  (test/src/LikelihoodTest.cpp:204-255), full == incremental (:257-279), virtual re-rooting preserves lnL
  on every edge (test/src/BrlenOptTest.cpp:297-367), changing + restoring a branch restores lnL;
* derivatives vs central finite differences of the edge-rooted lnL (sign convention Q6).
"""
import hashlib

import numpy as np
import pytest

from helpers import FIXTURE_PAIRS, fixture_summary, load_fixture, load_golden
from netrax_b200._capi import AVERAGE, BEST, LINKED, UNLINKED, Partition
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, caterpillar_network, random_network, simulate_alignment
from oracle import oracle

GOLD = load_golden("netrax_fixtures_golden.json")["cases"]
KINDS = ["port"] + (["ref"] if oracle.have_ref() else [])
SMALL = [k for k in FIXTURE_PAIRS if not k.startswith("celine")]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", list(FIXTURE_PAIRS))
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_fixture_matches_reference_golden(kind, name, variant):
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    eng = oracle.make_engine(kind, net, [part], variant=variant)
    got = fixture_summary(eng)
    exp = GOLD[f"{name}/{'AVERAGE' if variant == AVERAGE else 'BEST'}"]
    assert got["lnl"] == pytest.approx(exp["lnl"], rel=1e-12)
    assert [t["config"] for t in got["root_trees"]] == [t["config"] for t in exp["root_trees"]]
    for a, b in zip(got["root_trees"], exp["root_trees"]):
        assert a["logprob"] == pytest.approx(b["logprob"], rel=1e-14, abs=1e-300)
        assert a["partition_logl"] == pytest.approx(b["partition_logl"], rel=1e-12)
    assert got["nodes"].keys() == exp["nodes"].keys()
    for k in got["nodes"]:
        assert got["nodes"][k]["scaler_sum"] == exp["nodes"][k]["scaler_sum"]          # bit-exact integers
        assert got["nodes"][k]["clv_sha"] == exp["nodes"][k]["clv_sha"], k            # DNA CLVs: bit-exact
    eng.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", SMALL + ["celine_smaller_1"])
def test_improved_equals_naive(kind, name):
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    for variant in (AVERAGE, BEST):
        eng = oracle.make_engine(kind, net, [part], variant=variant)
        l = eng.computeLoglikelihood(0, 1)
        ln, tl, lp = oracle.naive_loglikelihood(eng)
        assert l == pytest.approx(ln, rel=1e-13)
        eng.close()


@pytest.mark.parametrize("kind", KINDS)
def test_incremental_equals_full_and_branch_restore(kind):
    net, part = load_fixture(*FIXTURE_PAIRS["three_reticulations"])
    eng = oracle.make_engine(kind, net, [part])
    l0 = eng.computeLoglikelihood(0, 1)
    assert eng.computeLoglikelihood(1, 1) == l0
    for e in range(net.num_edges):
        old = float(net.edge_length[e])
        eng.set_branch_length(e, old * 3 + 0.01)
        l1 = eng.computeLoglikelihood(1, 1)
        assert l1 == eng.computeLoglikelihood(0, 1)
        eng.set_branch_length(e, old)
        assert eng.computeLoglikelihood(1, 1) == pytest.approx(l0, rel=1e-14)
    eng.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("name", SMALL)
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_rerooting_preserves_lnl_on_every_edge(kind, name, variant):
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    eng = oracle.make_engine(kind, net, [part], variant=variant)
    l0 = eng.computeLoglikelihood(0, 1)
    for e in range(net.num_edges):
        eng.brlen_prepare(e)
        assert eng.computeLoglikelihoodBrlenOpt(e) == pytest.approx(l0, rel=1e-12), e
        assert eng.brlen_finish(e) == pytest.approx(l0, rel=1e-13)
    eng.close()


@pytest.mark.parametrize("kind", KINDS)
def test_derivatives_match_finite_differences_single_tree(kind):
    net, part = load_fixture(*FIXTURE_PAIRS["tree"])
    eng = oracle.make_engine(kind, net, [part])
    eng.computeLoglikelihood(0, 1)
    for e in range(net.num_edges):
        t = float(net.edge_length[e])
        eng.brlen_prepare(e)
        eng.computePartitionSumtables(e)
        d1, d2, *_ = eng.computeLoglikelihoodDerivatives(e)
        if t < 0.01:  # clamped 1e-6 branches: a central second difference is pure rounding noise
            eng.brlen_finish(e)
            continue
        h = 1e-4 * t
        vals = []
        for tt in (t - h, t, t + h):
            eng.brlen_set_length(e, tt)
            vals.append(eng.computeLoglikelihoodBrlenOpt(e))
        eng.brlen_set_length(e, t)
        fd1 = -(vals[2] - vals[0]) / (2 * h)          # libpll returns derivatives of MINUS lnL (Q6)
        fd2 = -(vals[2] - 2 * vals[1] + vals[0]) / (h * h)
        assert d1 == pytest.approx(fd1, rel=1e-4, abs=1e-5), e
        assert d2 == pytest.approx(fd2, rel=1e-3, abs=1e-3), e
        eng.brlen_finish(e)
    eng.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_derivative_mixing_matches_mpfr_semantics(kind, variant):
    """K7 for derivatives (LH/LikelihoodDerivatives.cpp:13-23,147-180): recompute the AVERAGE quotient rule /
    BEST pick (quirks Q2, Q6) from the raw per-sumtable (f, d1, d2) with mpmath at 53 bits, i.e. what
    mpfr::mpreal does in the reference."""
    mp = pytest.importorskip("mpmath")
    mp.mp.prec = 53
    net = random_network(9, 2, seed=4)
    m, w = simulate_alignment(net, 120, seed=2)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    eng = oracle.make_engine(kind, net, [part], variant=variant)
    eng.computeLoglikelihood(0, 1)
    checked = 0
    for e in range(net.num_edges):
        eng.brlen_prepare(e)
        n = eng.computePartitionSumtables(e)
        if n == 0:
            eng.brlen_finish(e)
            continue
        d1, d2, pd1, pd2, raw = eng.computeLoglikelihoodDerivatives(e)
        probs = [eng.read_sumtable(0, i)[1] for i in range(n)]
        if n == 1:
            assert (d1, d2) == (raw[0, 0, 1], raw[0, 0, 2])
        elif variant == AVERAGE:
            S = S1 = S2 = mp.mpf(0)
            for (f, a, b), pr in zip(raw[0], probs):
                lh = mp.exp(mp.mpf(float(f)))
                lhp = lh * float(a)
                lhpp = lhp * float(a) + lh * float(b)
                S += lh * pr; S1 += lhp * pr; S2 += lhpp * pr
            assert d1 == pytest.approx(float(S1 / S), rel=1e-13)
            assert d2 == pytest.approx(float((S2 * S - S1 * S1) / (S * S)), rel=1e-11)
            checked += 1
        else:
            best = max(range(n), key=lambda i: (raw[0, i, 0] * probs[i], -i))
            assert (d1, d2) == (raw[0, best, 1], raw[0, best, 2])
            checked += 1
        eng.brlen_finish(e)
    assert checked > 0
    eng.close()


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("cfg", [(12, 2, 300, 5), (25, 3, 200, 6), (40, 5, 120, 7)])
def test_port_equals_reference_on_synthetic(cfg):
    n, r, pat, seed = cfg
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    a, b = oracle.make_engine("port", net, [part]), oracle.make_engine("ref", net, [part])
    la, lb = a.computeLoglikelihood(0, 1), b.computeLoglikelihood(0, 1)
    assert la == pytest.approx(lb, rel=1e-13)
    for e in range(net.num_edges):
        assert np.array_equal(a.get_pmatrix(e), b.get_pmatrix(e))
    for v in range(net.num_tips, net.num_nodes):
        assert a.num_trees(v) == b.num_trees(v)
        for t in range(a.num_trees(v)):
            assert np.array_equal(a.read_scaler(v, t), b.read_scaler(v, t))
            assert np.array_equal(a.read_clv(v, t), b.read_clv(v, t))   # DNA: op order mirrored -> bit-exact
    e = int(net.ret_first_edge[0])
    for eng in (a, b):
        eng.brlen_prepare(e)
        eng.computePartitionSumtables(e)
    da, db = a.computeLoglikelihoodDerivatives(e), b.computeLoglikelihoodDerivatives(e)
    assert da[0] == pytest.approx(db[0], rel=1e-10) and da[1] == pytest.approx(db[1], rel=1e-10)
    assert a.computeLoglikelihoodBrlenOpt(e) == pytest.approx(b.computeLoglikelihoodBrlenOpt(e), rel=1e-13)


@pytest.mark.parametrize("kind", KINDS)
def test_scaler_stress_caterpillar(kind):
    net = caterpillar_network(400)
    m, w = simulate_alignment(net, 200, seed=11, random_cells=True)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    eng = oracle.make_engine(kind, net, [part])
    l = eng.computeLoglikelihood(0, 1)
    assert np.isfinite(l) and l < 0
    mx = max(int(eng.read_scaler(net.root, t).max()) for t in range(eng.num_trees(net.root)))
    assert mx >= 2
    ln, _, _ = oracle.naive_loglikelihood(eng)
    assert l == pytest.approx(ln, rel=1e-13)


def test_multi_partition_unlinked_best():
    net = random_network(10, 2, seed=3)
    parts, brl = [], []
    rng = np.random.default_rng(0)
    for p in range(3):
        m, w = simulate_alignment(net, 150, seed=20 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2, net.num_edges))
    eng = oracle.make_engine("port", net, parts, variant=BEST, linkage=UNLINKED, partition_brlens=brl)
    l = eng.computeLoglikelihood(0, 1)
    ln, _, _ = oracle.naive_loglikelihood(eng)
    assert l == pytest.approx(ln, rel=1e-13)
    assert l == pytest.approx(eng.partition_loglh().sum(), rel=1e-14)


def test_protein_port_equals_reference():
    """20 states (BASELINE config 4 shape, LG+G4): scalar restatement vs the reference's AVX2 kernels."""
    from netrax_b200.synth import lg_model
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    rates, freqs = lg_model()
    net = random_network(12, 2, seed=1)
    m, w = simulate_alignment(net, 300, seed=1, states=20, rates=rates, freqs=freqs)
    part = Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w)
    a, b = oracle.make_engine("port", net, [part]), oracle.make_engine("ref", net, [part])
    la, lb = a.computeLoglikelihood(0, 1), b.computeLoglikelihood(0, 1)
    assert la == pytest.approx(lb, rel=1e-12)
    for v in range(net.num_tips, net.num_nodes):
        for t in range(a.num_trees(v)):
            assert np.array_equal(a.read_scaler(v, t), b.read_scaler(v, t))
            np.testing.assert_allclose(a.read_clv(v, t), b.read_clv(v, t), rtol=1e-11, atol=1e-300)
    e = int(net.ret_first_edge[0])
    for eng in (a, b):
        eng.brlen_prepare(e); eng.computePartitionSumtables(e)
    da, db = a.computeLoglikelihoodDerivatives(e), b.computeLoglikelihoodDerivatives(e)
    assert da[0] == pytest.approx(db[0], rel=1e-9) and da[1] == pytest.approx(db[1], rel=1e-9)


# ---------------------------------------------------------------------------------------------- pseudo-likelihood
@pytest.mark.parametrize("kind", KINDS)
def test_pseudo_loglikelihood_equals_exact_on_a_tree(kind):
    """Without reticulations every weight is (1, 0, 0, 0): computePseudoLoglikelihood (LH/PseudoLoglikelihood.cpp:57-226)
    is the ordinary tree likelihood."""
    net = random_network(12, 0, seed=5)
    m, w = simulate_alignment(net, 300, seed=5)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    e = oracle.make_engine(kind, net, [part])
    assert e.computePseudoLoglikelihood(0, 1) == pytest.approx(e.computeLoglikelihood(0, 1), rel=1e-13)


def test_pseudo_loglikelihood_port_equals_reference_and_dispatch():
    """Scalar port vs real libpll under the same restated driver: CLVs and scalers of every node bit-identical (DNA);
    incremental == full after a branch-length and a reticulation-probability change; variant SARAH_PSEUDO routes
    computeLoglikelihood to it (LH/LikelihoodComputation.cpp:23-27)."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from netrax_b200._capi import SARAH_PSEUDO
    for n, r, seed in ((10, 2, 2), (14, 3, 3), (9, 4, 4)):
        net = random_network(n, r, seed=seed)
        m, w = simulate_alignment(net, 257, seed=seed)
        part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
        a, b = oracle.make_engine("port", net, [part], variant=SARAH_PSEUDO), oracle.make_engine("ref", net, [part], variant=SARAH_PSEUDO)
        la, lb = a.computePseudoLoglikelihood(0, 1), b.computePseudoLoglikelihood(0, 1)
        assert la == pytest.approx(lb, rel=1e-12)
        for v in range(net.num_tips, net.num_nodes):
            assert np.array_equal(a.read_pseudo_scaler(v), b.read_pseudo_scaler(v))
            np.testing.assert_allclose(a.read_pseudo_clv(v), b.read_pseudo_clv(v), rtol=1e-13, atol=0)
        for eng in (a, b):
            eng.set_branch_length(1, 0.37)
            eng.set_reticulation_prob(0, 0.8)
        inc = a.computePseudoLoglikelihood(1, 1)
        assert inc == pytest.approx(b.computePseudoLoglikelihood(1, 1), rel=1e-12)
        assert inc == pytest.approx(a.computePseudoLoglikelihood(0, 1), rel=1e-13) and inc != la
        c = oracle.make_engine("ref", net, [part], variant=SARAH_PSEUDO)
        assert c.computeLoglikelihood(0, 1) == pytest.approx(lb, rel=1e-13)


# ---- one rate matrix per rate category (LG4M / LG4X: raxml-ng ratecat_submodels -> libpll params_indices) ------------------
@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("states", [4, 20])
def test_submodels_oracle_semantics(kind, states):
    """Both oracle flavours (the scalar port's per-category matrices, pll_port.c port_set_submodels; real libpll with
    params_indices): (i) a mixture of identical matrices is the single-matrix model, bit for bit; (ii) a real mixture equals
    the lnL assembled from single-matrix, single-category, single-site evaluations; (iii) back to one matrix."""
    from helpers import encode_aa, mixture_lnl_by_categories, mixture_models
    from netrax_b200.network_io import parse_extended_newick
    net = parse_extended_newick("(((T0:0.1,T1:0.2):0.05,T2:0.3):0.1,(T3:0.15,T4:0.25):0.2);")   # a tree: one displayed tree, AVERAGE = plain lnL
    rng = np.random.default_rng(3)
    masks = (1 << rng.integers(0, states, size=(5, 9))).astype(np.uint32)
    masks[2, 4] = (1 << states) - 1   # a gap
    freqs, subst = mixture_models(states, 4, seed=11)
    rates, weights = np.array([0.2, 0.7, 1.3, 2.4]), np.array([0.4, 0.3, 0.2, 0.1])   # LG4X: free rates and weights
    part = Partition(states, 4, masks, freqs[0], subst[0], rates, rate_weights=weights)
    make = lambda net, part: oracle.make_engine(kind, net, [part])
    e = make(net, part)
    l_single = e.computeLoglikelihood(0, 1)
    e.set_submodels(0, [0, 1, 2, 3], np.stack([freqs[0]] * 4), np.stack([subst[0]] * 4))
    assert e.computeLoglikelihood(0, 1) == l_single
    cat_model = [2, 0, 3, 1]
    e.set_submodels(0, cat_model, freqs, subst)
    l_mix = e.computeLoglikelihood(0, 1)
    assert abs(l_mix - l_single) > 1e-3
    want = mixture_lnl_by_categories(make, net, part, cat_model, freqs, subst)
    assert l_mix == pytest.approx(want, rel=1e-12)
    e.set_submodels(0, [0, 0, 0, 0], freqs[:1], subst[:1])   # back to one matrix
    assert e.computeLoglikelihood(0, 1) == l_single
    if kind == "port":
        with pytest.raises(Exception, match="matri"):
            e.set_submodels(0, [0, 1, 2, 4], freqs, subst)   # category 3 names a matrix that does not exist
    e.close()


@pytest.mark.parametrize("states", [4, 20])
def test_submodels_port_equals_real_libpll_on_networks(states):
    """Per-category matrices in the scalar port against real libpll with the same params_indices on a network with two
    reticulations, with and without +I: lnL, the re-rooted edge lnL, every sumtable entry and the derivatives."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from helpers import mixture_models
    net = random_network(9, 2, seed=71)
    freqs, subst = mixture_models(states, 4, seed=12)
    m, w = simulate_alignment(net, 120, seed=71, states=states, rates=subst[0], freqs=freqs[0] / freqs[0].sum())
    part = Partition(states, 4, m, freqs[0], subst[0], GAMMA4_ALPHA05, pattern_weights=w)
    for pinv in (0.0, 0.2):
        a, b = oracle.make_engine("port", net, [part]), oracle.make_engine("ref", net, [part])
        for eng in (a, b):
            eng.set_submodels(0, [3, 1, 0, 2], freqs, subst)
            eng.set_pinv(0, pinv)
        assert a.computeLoglikelihood(0, 1) == pytest.approx(b.computeLoglikelihood(0, 1), rel=1e-11)
        for e in (0, net.num_edges - 1, int(net.ret_first_edge[0])):
            assert a.brlen_prepare(e) == pytest.approx(b.brlen_prepare(e), rel=1e-11)
            assert a.computeLoglikelihoodBrlenOpt(e) == pytest.approx(b.computeLoglikelihoodBrlenOpt(e), rel=1e-11)
            n_tables = a.computePartitionSumtables(e)
            assert n_tables == b.computePartitionSumtables(e)
            for k in range(n_tables):
                sb = b.read_sumtable(0, k)[0]   # entries are lefterm x righterm, each a cancelling sum: tolerance against the table's scale
                np.testing.assert_allclose(a.read_sumtable(0, k)[0], sb, rtol=1e-9, atol=1e-12 * np.abs(sb).max())
            da, db = a.computeLoglikelihoodDerivatives(e), b.computeLoglikelihoodDerivatives(e)
            np.testing.assert_allclose(da[4], db[4], rtol=1e-9, atol=1e-10)
            assert a.brlen_finish(e) == pytest.approx(b.brlen_finish(e), rel=1e-11)
        a.close(); b.close()


def scaled_linkage_case(seed=5):
    net = random_network(9, 2, seed=seed)
    parts = []
    for i in range(3):
        m, w = simulate_alignment(net, 150 + 40 * i, seed=seed + i)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    return net, parts, [0.5, 1.0, 1.7]


@pytest.mark.parametrize("kind", KINDS)
def test_scaled_branch_length_linkage(kind):
    """PLLMOD_COMMON_BRLEN_SCALED: the P-matrices of partition p use brlen_scalers[p] x the linked length
    (PLLMOD/tree/treeinfo.c:862-864) — identical to an unlinked analysis whose per-partition lengths are the scaled ones;
    scalers of 1 give the linked result; derivatives refuse, as the reference does (LH/LikelihoodDerivatives.cpp:42-46)."""
    from netrax_b200._capi import SCALED
    net, parts, scalers = scaled_linkage_case()
    lin = oracle.make_engine(kind, net, parts, linkage=LINKED)
    sc = oracle.make_engine(kind, net, parts, linkage=SCALED)
    assert sc.computeLoglikelihood(0, 1) == lin.computeLoglikelihood(0, 1)
    for p, s in enumerate(scalers):
        sc.set_brlen_scaler(p, s)
    un = oracle.make_engine(kind, net, parts, linkage=UNLINKED, partition_brlens=[net.edge_length * s for s in scalers])
    l_sc = sc.computeLoglikelihood(1, 1)
    assert l_sc == un.computeLoglikelihood(0, 1)
    np.testing.assert_array_equal(sc.partition_loglh(), un.partition_loglh())
    assert abs(l_sc - lin.computeLoglikelihood(0, 1)) > 1e-3
    e = int(net.ret_first_edge[0])
    sc.set_branch_length(e, 0.33); un_len = 0.33
    for p, s in enumerate(scalers):
        un.set_branch_length(e, un_len * s, partition=p)
    assert sc.computeLoglikelihood(1, 1) == un.computeLoglikelihood(1, 1)
    sc.brlen_prepare(e)
    sc.computePartitionSumtables(e)
    with pytest.raises(Exception, match="scaled branch lengths"):
        sc.computeLoglikelihoodDerivatives(e)
    with pytest.raises(Exception, match="scaled branch length mode"):
        lin.set_brlen_scaler(0, 2.0)
    for x in (lin, sc, un):
        x.close()


def test_pinv_port_equals_real_libpll_on_networks():
    """+I in the scalar port against the reference's real libpll under the same restated driver: a DNA network, a protein
    network and a 300-taxon caterpillar with invariant columns whose scaled sites take libpll's "undo the scaling on the
    non-invariant term only" branch of the edge lnL (core_likelihood_avx.c:493-501) — lnL, re-rooted edge lnL, sumtables and
    derivatives on a tip edge, an inner edge and a reticulation edge."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from netrax_b200.synth import caterpillar_network, lg_model
    cases = []
    net = random_network(14, 3, seed=61)
    m, w = simulate_alignment(net, 500, seed=61)
    cases.append((net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.25))
    netp = random_network(10, 2, seed=62)
    lr, lf = lg_model()
    mp, wp = simulate_alignment(netp, 150, seed=62, states=20, rates=np.asarray(lr), freqs=np.asarray(lf))
    cases.append((netp, Partition(20, 4, mp, lf, lr, GAMMA4_ALPHA05, pattern_weights=wp), 0.4))
    cat = caterpillar_network(300)
    m, w = simulate_alignment(cat, 200, seed=63, gap_frac=0.0)
    m[:, :60] = m[0, :60]
    cases.append((cat, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.3))
    # with Gamma categories the slowest category keeps invariant columns above the scaling threshold; ONE category on saturated
    # branches scales them too (scaler 2 on every column): the case that really takes the "non-invariant term only" branch
    sat = caterpillar_network(300, brlen=4.0)
    m1, w1 = simulate_alignment(sat, 200, seed=63, gap_frac=0.0)
    m1[:, :60] = m1[0, :60]
    cases.append((sat, Partition(4, 1, m1, DNA_FREQS, GTR_RATES, np.ones(1), pattern_weights=w1), 0.3))
    for net, part, pinv in cases:
        a, b = oracle.make_engine("port", net, [part]), oracle.make_engine("ref", net, [part])
        a.set_eigen(0, *b.get_eigen(0)) if hasattr(a, "set_eigen") else None
        a.set_pinv(0, pinv); b.set_pinv(0, pinv)
        la, lb = a.computeLoglikelihood(0, 1), b.computeLoglikelihood(0, 1)
        assert la == pytest.approx(lb, rel=1e-11), (part.states, pinv)
        if net is sat:   # the invariant columns (the first 60) are scaled here, and only here
            assert np.all(a.read_scaler(net.root, 0)[:60] > 0) and np.array_equal(a.read_scaler(net.root, 0), b.read_scaler(net.root, 0))
        elif net is cat:
            assert np.all(a.read_scaler(net.root, 0)[:60] == 0) and np.any(a.read_scaler(net.root, 0) > 0)
        for e in (0, net.num_edges - 1) + ((int(net.ret_first_edge[0]),) if net.num_reticulations else ()):
            assert a.brlen_prepare(e) == pytest.approx(b.brlen_prepare(e), rel=1e-11)
            assert a.computeLoglikelihoodBrlenOpt(e) == pytest.approx(b.computeLoglikelihoodBrlenOpt(e), rel=1e-11), (part.states, e)
            assert a.computePartitionSumtables(e) == b.computePartitionSumtables(e)
            da, db = a.computeLoglikelihoodDerivatives(e), b.computeLoglikelihoodDerivatives(e)
            np.testing.assert_allclose(da[4], db[4], rtol=1e-9, atol=1e-10)
            assert a.brlen_finish(e) == pytest.approx(b.brlen_finish(e), rel=1e-11)
        a.set_pinv(0, 0.0); b.set_pinv(0, 0.0)
        assert a.computeLoglikelihood(0, 1) == pytest.approx(b.computeLoglikelihood(0, 1), rel=1e-11)
        a.close(); b.close()


@pytest.mark.parametrize("kind", KINDS)
def test_empty_shard_of_a_partition_contributes_nothing(kind):
    """Site sharding may leave a rank without any pattern of a small partition (the reference's partitions[p] == NULL,
    "skip remote partitions", LH/ImprovedLoglikelihood.cpp:128-131): a zero-pattern shard evaluates to lnL 0 and leaves the
    other partitions alone, so the all-reduced sums are those of the ranks that own patterns."""
    net = random_network(8, 1, seed=4)
    m, w = simulate_alignment(net, 50, seed=4)
    full = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    alone = oracle.make_engine(kind, net, [full])
    both = oracle.make_engine(kind, net, [full, full.slice(0, 0)])
    assert both.computeLoglikelihood(0, 1) == alone.computeLoglikelihood(0, 1)
    np.testing.assert_array_equal(both.partition_loglh(), [alone.partition_loglh()[0], 0.0])
    alone.close(); both.close()


@pytest.mark.parametrize("kind", KINDS)
def test_persite_lnl_sums_to_the_tree_lnl(kind):
    """orc_persite_lnl: the per-site array of pll_compute_root_loglikelihood (LH/ImprovedLoglikelihood.cpp:448-453) sums to the
    per-tree partition lnL, pattern weights applied, scaled sites included."""
    for net, seed in ((random_network(12, 2, seed=5), 5), (caterpillar_network(300), 63)):
        m, w = simulate_alignment(net, 200, seed=seed)
        part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
        eng = oracle.make_engine(kind, net, [part])
        eng.computeLoglikelihood(0, 1)
        for t in range(eng.num_trees(net.root)):
            ps = oracle.persite_lnl(eng, t)[0]
            assert ps.shape == (200,) and np.all(ps < 0)
            assert ps.sum() == pytest.approx(eng.tree_info(net.root, t)[1][0], rel=1e-12)
        eng.close()


@pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built")
def test_persite_lnl_port_equals_real_libpll():
    """element-wise: the restated root-lnL kernel against libpll's own persite_lnl output (DNA, scaled sites, +I, protein)"""
    from netrax_b200.synth import lg_model
    rates, freqs = lg_model()
    cases = []
    net = random_network(12, 2, seed=5)
    m, w = simulate_alignment(net, 200, seed=5)
    cases.append((net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.0))
    cases.append((net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.25))
    cat = caterpillar_network(300)
    m, w = simulate_alignment(cat, 150, seed=63, gap_frac=0.0)
    cases.append((cat, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.0))
    m, w = simulate_alignment(net, 120, seed=6, states=20, rates=rates, freqs=freqs)
    cases.append((net, Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w), 0.0))
    for net, part, pinv in cases:
        a, b = oracle.make_engine("port", net, [part]), oracle.make_engine("ref", net, [part])
        if pinv:
            a.set_pinv(0, pinv); b.set_pinv(0, pinv)
        a.computeLoglikelihood(0, 1); b.computeLoglikelihood(0, 1)
        for t in range(a.num_trees(net.root)):
            np.testing.assert_allclose(oracle.persite_lnl(a, t)[0], oracle.persite_lnl(b, t)[0], rtol=1e-12, atol=0)
        a.close(); b.close()
