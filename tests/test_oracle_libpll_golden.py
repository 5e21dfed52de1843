"""Pins the oracle (both kinds) to the golden numbers of libpll's own regression suite
(LIBPLL/../test/out/derivatives.out, inline data of test/src/derivatives.c).  The golden test computes an
edge lnL and d/dt, d2/dt2 of -lnL on the inner edge (clv6,clv7) and on the tip edge (tip4,clv7) of a
5-taxon unrooted tree; here the same tree is rooted ON that edge (t split in two halves), so the network
lnL equals the golden edge lnL and the NetRAX branch-length derivative of the half edge equals the golden
derivative.  Golden precision: 6 decimals (lnL), 5 significant digits (derivatives)."""
import numpy as np
import pytest

from helpers import load_golden
from netrax_b200._capi import Partition
from netrax_b200.network_io import encode_dna, parse_extended_newick
from oracle import oracle

G = load_golden("libpll_derivatives_golden.json")
KINDS = ["port"] + (["ref"] if oracle.have_ref() else [])


def _engine(kind, block, t, tip_edge):
    b0, b1 = G["branch_lengths"]
    h = t / 2
    if not tip_edge:  # clv6=((t0,t1):b0,t2) | clv7=(t3,t4), joined by matrix 0 at length t
        nw = f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{h},(T3:{b1},T4:{b1})X7:{h});"
    else:             # after operation 3: clv7=(clv6:b0, t3:b0), joined with tip4 by matrix 1 at length t
        nw = f"(T4:{h},(((T0:{b1},T1:{b1}):{b0},T2:{b1}):{b0},T3:{b0})X7:{h});"
    net = parse_extended_newick(nw)
    order = [int(l[1:]) for l in net.tip_labels]
    masks = np.stack([encode_dna(G["tips"][i]) for i in order])
    rates = oracle.api("port").gamma_rates(block["alpha"], block["ncats"]) if block["ncats"] > 1 else np.ones(1)
    part = Partition(4, block["ncats"], masks, G["freqs"], G["subst"], rates)
    return net, oracle.make_engine(kind, net, [part])


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("bi", range(len(G["blocks"])))
@pytest.mark.parametrize("tip_edge", [False, True])
def test_libpll_golden_edge_lnl_and_derivatives(kind, bi, tip_edge):
    block = G["blocks"][bi]
    for t, f, d1, d2 in block["tip" if tip_edge else "inner"]:
        if t > 10:  # golden derivatives there are 1e-15 noise; P-matrix saturates
            continue
        net, eng = _engine(kind, block, t, tip_edge)
        lnl = eng.computeLoglikelihood(0, 1)
        assert abs(lnl - f) < 2e-6, (t, lnl, f)
        # the edge whose child is X7 (inner case: X6's sibling) — any of the two root edges works
        edge = [e for e in range(net.num_edges) if net.edge_source[e] == net.root][0]
        eng.brlen_prepare(edge)
        assert abs(eng.computeLoglikelihoodBrlenOpt(edge) - lnl) < 1e-9
        assert eng.computePartitionSumtables(edge) == 1
        g1, g2, *_ = eng.computeLoglikelihoodDerivatives(edge)
        assert g1 == pytest.approx(d1, rel=2e-4, abs=1e-9), (t, g1, d1)
        assert g2 == pytest.approx(d2, rel=2e-4, abs=1e-9), (t, g2, d2)
        assert abs(eng.brlen_finish(edge) - lnl) < 1e-9
        eng.close()


# ---- second golden set: libpll test/out/alpha-cats.out (Gamma rates, both discretisation modes, and an edge lnL) ----------
GA = load_golden("libpll_alpha_cats_golden.json")
MODE = {"MEAN": 0, "MEDIAN": 1}   # PLL_GAMMA_RATES_MEAN / _MEDIAN (LIBPLL/pll.h)


@pytest.mark.parametrize("kind", KINDS)
def test_libpll_golden_gamma_rates_oracle(kind):
    """pll_compute_gamma_cats for alpha in 0.1..100, 1..16 categories, MEAN and MEDIAN: the golden prints 6 decimals."""
    api = oracle.api(kind)
    for b in GA["blocks"]:
        want = np.array(b["rates"])
        got = api.gamma_rates(b["alpha"], b["ncats"], MODE[b["mode"]]) if b["ncats"] > 1 else np.ones(1)
        np.testing.assert_allclose(got, want, rtol=0, atol=5.1e-7, err_msg=str((b["alpha"], b["ncats"], b["mode"])))


def test_libpll_golden_gamma_rates_product_host():
    """The PRODUCT's own Gamma discretisation (host/model.cpp compute_gamma_cats, what setAlpha / optimize_alpha use)
    against the same golden vectors — host code, no device needed."""
    from netrax_b200 import engine
    api = engine.load()
    for b in GA["blocks"]:
        if b["ncats"] == 1:
            continue
        got = api.gamma_rates(b["alpha"], b["ncats"], MODE[b["mode"]])
        np.testing.assert_allclose(got, np.array(b["rates"]), rtol=0, atol=5.1e-7, err_msg=str((b["alpha"], b["ncats"], b["mode"])))


@pytest.mark.parametrize("kind", KINDS)
def test_libpll_golden_alpha_cats_edge_lnl(kind):
    """The golden edge lnL between clv6 = ((t0,t1),t2) and clv7 = (t3,t4) for every (alpha, categories, mode): same tree
    rooted on that edge (0.1 split in two halves), rates = the oracle's own Gamma rates for that mode."""
    b0, b1 = GA["branch_lengths"]
    h = b0 / 2
    nw = f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{h},(T3:{b1},T4:{b1})X7:{h});"
    net = parse_extended_newick(nw)
    order = [int(l[1:]) for l in net.tip_labels]
    masks = np.stack([encode_dna(GA["tips"][i]) for i in order])
    api = oracle.api(kind)
    for b in GA["blocks"]:
        rates = api.gamma_rates(b["alpha"], b["ncats"], MODE[b["mode"]]) if b["ncats"] > 1 else np.ones(1)
        part = Partition(4, b["ncats"], masks, GA["freqs"], GA["subst"], rates)
        eng = oracle.make_engine(kind, net, [part])
        assert abs(eng.computeLoglikelihood(0, 1) - b["logl"]) < 2e-6, (b["alpha"], b["ncats"], b["mode"])
        eng.close()


# ---- third golden set: libpll test/out/protein-models.out (20 empirical amino-acid models, 20-state kernels) -------------
GP = load_golden("libpll_protein_models_golden.json")


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("model", list(GP["models"]))
def test_libpll_golden_protein_models(kind, model):
    """Edge lnL of a 5-taxon, 113-site amino-acid alignment under each of libpll's 20 empirical models (Gamma alpha = 1,
    4 categories): the 20-state arithmetic of both oracle flavours against the reference's regression output (6 decimals)."""
    from helpers import protein_golden_case
    net, part = protein_golden_case(GP, model)
    eng = oracle.make_engine(kind, net, [part])
    assert abs(eng.computeLoglikelihood(0, 1) - GP["models"][model]["logl"]) < 2e-6
    eng.close()


def test_shipped_lg_model_is_the_reference_table():
    """netrax_b200/lg_model.json (the model data bench config 4 uses) == the LG table the golden file was produced with."""
    from netrax_b200.synth import lg_model
    rates, freqs = lg_model()
    np.testing.assert_array_equal(np.asarray(rates), np.asarray(GP["models"]["LG"]["rates"]))
    np.testing.assert_array_equal(np.asarray(freqs), np.asarray(GP["models"]["LG"]["freqs"]))


# ---- fourth golden set: libpll test/out/derivatives-oddstates.out (5 states, states_padded = 8) --------------------------
GO = load_golden("libpll_derivatives_oddstates_golden.json")


def oddstates_case(t, tip_edge, ncats, rates):
    b0, b1 = GO["branch_lengths"]
    h = t / 2
    if not tip_edge:
        nw = f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{h},(T3:{b1},T4:{b1})X7:{h});"
    else:
        nw = f"(T4:{h},(((T0:{b1},T1:{b1}):{b0},T2:{b1}):{b0},T3:{b0})X7:{h});"
    net = parse_extended_newick(nw)
    order = [int(l[1:]) for l in net.tip_labels]
    masks = np.stack([np.array([GO["char_masks"][c] for c in GO["tips"][i]], dtype=np.uint32) for i in order])
    return net, Partition(GO["states"], ncats, masks, GO["freqs"], GO["subst"], rates)


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("tip_edge", [False, True])
def test_libpll_golden_oddstates(kind, tip_edge):
    """Five states (padded to eight: the only golden set where states != states_padded): edge lnL and derivatives."""
    for block in GO["blocks"]:
        rates = oracle.api("port").gamma_rates(block["alpha"], block["ncats"]) if block["ncats"] > 1 else np.ones(1)
        for t, f, d1, d2 in block["tip" if tip_edge else "inner"]:
            if t > 10:
                continue
            net, part = oddstates_case(t, tip_edge, block["ncats"], rates)
            eng = oracle.make_engine(kind, net, [part])
            lnl = eng.computeLoglikelihood(0, 1)
            assert abs(lnl - f) < 2e-6, (block["alpha"], block["ncats"], t, lnl, f)
            edge = [e for e in range(net.num_edges) if net.edge_source[e] == net.root][0]
            eng.brlen_prepare(edge)
            assert abs(eng.computeLoglikelihoodBrlenOpt(edge) - lnl) < 1e-9
            assert eng.computePartitionSumtables(edge) == 1
            g1, g2, *_ = eng.computeLoglikelihoodDerivatives(edge)
            assert g1 == pytest.approx(d1, rel=2e-4, abs=1e-9), (t, g1, d1)
            assert g2 == pytest.approx(d2, rel=2e-4, abs=1e-9), (t, g2, d2)
            eng.close()


# ---- +I: the pinv = 0.3 / 0.6 / 0.9 blocks of derivatives.out and derivatives-oddstates.out ------------------------------
GI = load_golden("libpll_derivatives_pinv_golden.json")
GOI = load_golden("libpll_derivatives_oddstates_pinv_golden.json")


def pinv_case(G, t, tip_edge, ncats, rates):
    """Same 5-taxon tree; works for the 4-state and the 5-state data set."""
    if "char_masks" in G:
        return oddstates_case(t, tip_edge, ncats, rates)
    b0, b1 = G["branch_lengths"]
    h = t / 2
    if not tip_edge:
        nw = f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{h},(T3:{b1},T4:{b1})X7:{h});"
    else:
        nw = f"(T4:{h},(((T0:{b1},T1:{b1}):{b0},T2:{b1}):{b0},T3:{b0})X7:{h});"
    net = parse_extended_newick(nw)
    order = [int(l[1:]) for l in net.tip_labels]
    masks = np.stack([encode_dna(G["tips"][i]) for i in order])
    return net, Partition(4, ncats, masks, G["freqs"], G["subst"], rates)


def check_pinv_golden(make_engine, G, tip_edge, gamma):
    for block in G["blocks"]:
        rates = gamma(block["alpha"], block["ncats"]) if block["ncats"] > 1 else np.ones(1)
        for t, f, d1, d2 in block["tip" if tip_edge else "inner"]:
            if t > 10:
                continue
            net, part = pinv_case(G, t, tip_edge, block["ncats"], rates)
            eng = make_engine(net, part)
            eng.set_pinv(0, block["pinv"])
            lnl = eng.computeLoglikelihood(0, 1)
            assert abs(lnl - f) < 2e-6, (block["alpha"], block["ncats"], block["pinv"], t, lnl, f)
            edge = [e for e in range(net.num_edges) if net.edge_source[e] == net.root][0]
            eng.brlen_prepare(edge)
            assert abs(eng.computeLoglikelihoodBrlenOpt(edge) - lnl) < 1e-8
            assert eng.computePartitionSumtables(edge) == 1
            g1, g2, *_ = eng.computeLoglikelihoodDerivatives(edge)
            assert g1 == pytest.approx(d1, rel=2e-4, abs=1e-9), (block["pinv"], t, g1, d1)
            assert g2 == pytest.approx(d2, rel=2e-4, abs=1e-9), (block["pinv"], t, g2, d2)
            eng.close()


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("which", ["dna", "oddstates"])
@pytest.mark.parametrize("tip_edge", [False, True])
def test_libpll_golden_pinv(kind, which, tip_edge):
    """+I through both oracle flavours — the scalar port's invariant-site terms (pll_port.c: root / edge lnL, derivatives,
    invariant-pattern detection) and real libpll: the golden edge lnL / derivatives with pinv in {0.3, 0.6, 0.9}."""
    G = GI if which == "dna" else GOI
    check_pinv_golden(lambda net, part: oracle.make_engine(kind, net, [part]), G, tip_edge, oracle.api(kind).gamma_rates)
    net, part = pinv_case(G, 0.1, False, 1, np.ones(1))
    e = oracle.make_engine(kind, net, [part])
    with pytest.raises(Exception, match="[Ii]nvalid proportion|invariant"):
        e.set_pinv(0, 1.0)
    e.close()


# ---- fifth / sixth golden set: libpll test/out/pmatrix.out (K1 + eigendecomposition, 9 decimals) and hky.out ---------------
def _npz(name):
    import os
    from helpers import GOLDEN
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("datatype", ["DNA", "PROT", "ODD"])
def test_libpll_golden_pmatrix(kind, datatype):
    """P(t) for 4 / 20 / 5 states x equal / skewed / extreme frequencies and exchangeabilities x branch lengths 1e-6 .. 100 x
    category rates 1e-31 .. 100: eigendecomposition + pll_core_update_pmatrix against libpll's regression output."""
    from helpers import check_pmatrix_golden
    check_pmatrix_golden(lambda net, part: oracle.make_engine(kind, net, [part]), _npz("libpll_pmatrix_golden.npz"), datatype)


@pytest.mark.parametrize("kind", KINDS)
def test_libpll_golden_hky(kind):
    """HKY with 10 ti/tv ratios: P-matrices, the inner CLVs and the edge lnL of libpll's test/out/hky.out."""
    from helpers import check_hky_golden
    check_hky_golden(lambda net, part: oracle.make_engine(kind, net, [part]), _npz("libpll_hky_golden.npz"))
