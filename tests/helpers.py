"""Shared test helpers: fixture loading, the standard test model, summaries used for golden files."""
import hashlib
import json
import os

import numpy as np

from netrax_b200._capi import Partition
from netrax_b200.network_io import compress_patterns, encode_dna, parse_extended_newick, read_fasta
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FIX = os.path.join(GOLDEN, "fixtures")

# pairings of test/src/LikelihoodTest.cpp:282-360 (the reference's own likelihood tests)
FIXTURE_PAIRS = {
    "small": ("small.nw", "small_fake_alignment.txt"),
    "tree": ("tree.nw", "small_fake_alignment.txt"),
    "tiny": ("tiny.nw", "tiny_fake_alignment.txt"),
    "clv_averaging": ("clv_averaging.nw", "5_taxa_fake_alignment.txt"),
    "two_reticulations": ("two_reticulations.nw", "5_taxa_fake_alignment.txt"),
    "three_reticulations": ("three_reticulations.nw", "7_taxa_fake_alignment.txt"),
    "interleaved_reticulations": ("interleaved_reticulations.nw", "5_taxa_fake_alignment.txt"),
    "reticulation_in_reticulation": ("reticulation_in_reticulation.nw", "small_fake_alignment.txt"),
    "celine": ("celine.nw", "celine_fake_alignment.txt"),
    "celine_smaller_1": ("celine_smaller_1.nw", "celine_fake_alignment_smaller.txt"),
    "celine_nonzero_branches": ("celine_nonzero_branches.nw", "celine_fake_alignment.txt"),
}


def load_fixture(nw, aln, compress=True):
    net = parse_extended_newick(open(os.path.join(FIX, nw)).read())
    fa = read_fasta(open(os.path.join(FIX, aln)).read())
    masks = np.stack([encode_dna(fa[l]) for l in net.tip_labels])
    w = None
    if compress:
        masks, w = compress_patterns(masks)
    return net, Partition(4, 4, masks, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)


def fixture_summary(eng):
    """Numbers a golden file pins for one engine state after a full evaluation."""
    lnl = eng.computeLoglikelihood(0, 1)
    root = eng.net.root
    trees = []
    for t in range(eng.num_trees(root)):
        lp, pl, _ = eng.tree_info(root, t)
        trees.append({"config": eng.tree_config(root, t), "logprob": lp, "partition_logl": pl.tolist()})
    nodes = {}
    for v in range(eng.net.num_tips, eng.net.num_nodes):
        for t in range(eng.num_trees(v)):
            clv = eng.read_clv(v, t, 0)
            sc = eng.read_scaler(v, t, 0)
            nodes[f"{v}:{eng.tree_config(v, t)}"] = {"clv_sha": hashlib.sha256(clv.tobytes()).hexdigest()[:16],
                                                      "clv_sum": float(clv.sum()), "scaler_sum": int(sc.sum())}
    return {"lnl": lnl, "partition_loglh": eng.partition_loglh().tolist(), "root_trees": trees, "nodes": nodes}


def load_golden(name):
    return json.load(open(os.path.join(GOLDEN, name)))


AA_ORDER = "ARNDCQEGHILKMFPSTWYV"   # pll_map_aa (LIBPLL/maps.c:66-100)
_AA_AMBIG = {"B": "ND", "Z": "QE", "J": "IL"}


def encode_aa(seq: str):
    """Amino-acid characters -> 20-bit state masks exactly as pll_map_aa: gaps / X / ? / * are fully ambiguous."""
    import numpy as np
    out = np.zeros(len(seq), dtype=np.uint32)
    for i, ch in enumerate(seq.upper()):
        if ch in AA_ORDER:
            out[i] = 1 << AA_ORDER.index(ch)
        elif ch in _AA_AMBIG:
            out[i] = sum(1 << AA_ORDER.index(c) for c in _AA_AMBIG[ch])
        elif ch in "-X?*":
            out[i] = (1 << 20) - 1
        else:
            raise ValueError(f"illegal amino-acid character {ch!r}")
    return out


def protein_golden_case(GP, model):
    """5-taxon tree of libpll's test/src/protein-models.c rooted on the edge the golden lnL is computed on."""
    import numpy as np
    from netrax_b200._capi import Partition
    from netrax_b200.network_io import parse_extended_newick
    from oracle import oracle
    b0, b1 = GP["branch_lengths"]
    h = b0 / 2
    net = parse_extended_newick(f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{h},(T3:{b1},T4:{b1})X7:{h});")
    order = [int(l[1:]) for l in net.tip_labels]
    masks = np.stack([encode_aa(GP["tips"][i]) for i in order])
    rates = oracle.api("port").gamma_rates(GP["alpha"], GP["ncats"])
    m = GP["models"][model]
    return net, Partition(20, GP["ncats"], masks, m["freqs"], m["rates"], rates)


# ---- libpll test/src/pmatrix.c and hky.c, restated for the fifth / sixth golden set ---------------------------------------
def pmatrix_golden_inputs(n_states):
    """The 3 frequency and 3 exchangeability vectors of libpll's pmatrix test (test/src/pmatrix.c:112-170: equal / skewed /
    extreme), as the doubles the C code builds (the golden file prints them with 6 decimals only).  The skewed and the odd-state
    extreme frequency vectors do not sum to 1: pll_set_frequencies renormalises them, and so do both oracles and the product."""
    n_rates = n_states * (n_states - 1) // 2
    f_eq = np.full(n_states, 1.0 / n_states)
    f_sk = np.full(n_states, 1.0 / n_states)
    skew = 1.0 / (3.0 * n_states)
    for k in range(n_states):
        if k % 2 == 0:
            f_sk[k] += skew
        elif k != n_states - 1:
            f_sk[k] -= skew
    minfreq = 1e-3
    maxfreq = (1.0 - 0.5 * n_states * minfreq) / (0.5 * n_states)
    f_ex = np.array([minfreq if k % 2 == 0 else maxfreq for k in range(n_states)])
    r_eq = np.ones(n_rates)
    r_sk = np.ones(n_rates)
    for k in range(n_rates):
        if k % 2 == 0:
            r_sk[k] *= 5.0
        elif k != n_rates - 1:
            r_sk[k] /= 5.0
    r_ex = np.array([1e-3 if k % 2 == 0 else 1e3 for k in range(n_rates)])
    r_ex[n_rates - 1] = 1.0
    return [f_eq, f_sk, f_ex], [r_eq, r_sk, r_ex]


def pmatrix_golden_case(G, n_states):
    """A 4-taxon tree whose first five branches carry the golden test's five branch lengths; returns (net, masks, edge ids in
    golden order)."""
    bl = [float(x) for x in G["branch_lengths"]]
    net = parse_extended_newick(f"((T0:{bl[0]!r},T1:{bl[1]!r}):{bl[2]!r},(T2:{bl[3]!r},T3:{bl[4]!r}):0.05);")
    edges = [int(np.flatnonzero(net.edge_length == t)[0]) for t in bl]
    rng = np.random.default_rng(5)
    masks = (1 << rng.integers(0, n_states, size=(4, 5))).astype(np.uint32)   # N_SITES = 5; the data do not enter a P-matrix
    return net, masks, edges


def check_pmatrix_golden(make_engine, G, datatype, atol=6e-10):
    """All 9 (frequencies, exchangeabilities) settings of one datatype: every P-matrix entry against the 9 printed decimals."""
    S = {"DNA": 4, "PROT": 20, "ODD": 5}[datatype]
    SP = (S + 3) & ~3
    freqs, substs = pmatrix_golden_inputs(S)
    net, masks, edges = pmatrix_golden_case(G, S)
    worst = 0.0
    for j in range(3):
        for k in range(3):
            idx = j * 3 + k
            np.testing.assert_allclose(freqs[j] / freqs[j].sum(), G[f"{datatype}_freqs_{idx}"], atol=5.1e-7, rtol=0)
            np.testing.assert_allclose(substs[k], G[f"{datatype}_subst_{idx}"], atol=5.1e-7, rtol=0)
            part = Partition(S, 4, masks, freqs[j], substs[k], G["cat_rates"])
            eng = make_engine(net, part)
            eng.computeLoglikelihood(0, 1)
            for b, e in enumerate(edges):
                got = eng.get_pmatrix(e, 0).reshape(4, S, SP)[:, :, :S]
                want = G[f"{datatype}_P_{idx}"][b]
                assert np.all(np.isfinite(got)) and got.min() >= -1e-12, (datatype, idx, b)   # check_matrix of pmatrix.c:47-56
                worst = max(worst, float(np.abs(got - want).max()))
                np.testing.assert_allclose(got, want, atol=atol, rtol=0, err_msg=str((datatype, idx, b)))
            eng.close()
    return worst


def hky_golden_case(H, i):
    """Tree of libpll's test/src/hky.c rooted on the edge the golden lnL is computed on; ti/tv ratio number i."""
    from oracle import oracle
    tips = ["WAACTCGCTA--ATTCTAAT", "CACCATGCTA--ATTGTCTT", "AG-C-TGCAG--CTTCTACT", "CGTCTTGCAA--AT-C-AAG", "CGACTTGCCA--AT-T-AAG"]
    b0, b1 = 0.1, 0.2
    net = parse_extended_newick(f"(((T0:{b1},T1:{b1}):{b0},T2:{b1})X6:{b0 / 2},(T3:{b1},T4:{b1})X7:{b0 / 2});")
    order = [int(l[1:]) for l in net.tip_labels]
    masks = np.stack([encode_dna(tips[k]) for k in order])
    k = float(H["titv"][i])
    part = Partition(4, 4, masks, [0.3, 0.4, 0.1, 0.2], [1, k, 1, 1, k, 1], oracle.api("port").gamma_rates(1.0, 4))
    return net, part


def check_hky_golden(make_engine, H):
    """P-matrices (4 decimals), the three inner CLVs (5 decimals) and the edge lnL (4 decimals) for the 10 ti/tv ratios."""
    for i in range(len(H["titv"])):
        net, part = hky_golden_case(H, i)
        eng = make_engine(net, part)
        lnl = eng.computeLoglikelihood(0, 1)
        assert abs(lnl - H["logl"][i]) < 5.1e-5, (i, lnl, H["logl"][i])
        kids = {v: [int(net.edge_target[e]) for e in range(net.num_edges) if net.edge_source[e] == v] for v in range(net.num_nodes)}
        inner = [c for c in kids[net.root] if c >= net.num_tips]
        x6 = [v for v in inner if any(c >= net.num_tips for c in kids[v])][0]   # ((t0,t1),t2)
        x7 = [v for v in inner if v != x6][0]                                    # (t3,t4)
        x5 = [c for c in kids[x6] if c >= net.num_tips][0]                       # (t0,t1)
        for node, key in ((x5, "clv5"), (x6, "clv6"), (x7, "clv7")):
            assert eng.num_trees(node) == 1 and int(eng.read_scaler(node, 0).sum()) == 0
            got = eng.read_clv(node, 0).reshape(20, 4, 4)
            np.testing.assert_allclose(got, H[key][i], atol=5.1e-6, rtol=0, err_msg=f"{key} ratio {i}")
        # matrix 0 (t = 0.1): the branch below clv6's first child; matrix 1 (t = 0.2): any tip branch
        e01 = [e for e in range(net.num_edges) if net.edge_target[e] == x5][0]
        e02 = [e for e in range(net.num_edges) if net.edge_target[e] < net.num_tips][0]
        for e, b in ((e01, 0), (e02, 1)):
            np.testing.assert_allclose(eng.get_pmatrix(e, 0).reshape(4, 4, 4), H["P"][i][b], atol=5.1e-5, rtol=0)
        eng.close()


# ---- mixtures with one rate matrix per rate category (LG4M / LG4X shape) ---------------------------------------------------
def mixture_models(states, n, seed):
    """n reversible models of `states` states: seeded log-normal perturbations of a base model (GTR test model / LG)."""
    from netrax_b200.synth import lg_model
    rng = np.random.default_rng(seed)
    if states == 20:
        r0, f0 = lg_model()
    else:
        nr = states * (states - 1) // 2
        r0 = np.asarray(GTR_RATES, float) if states == 4 else np.linspace(0.5, 2.0, nr)
        f0 = np.asarray(DNA_FREQS, float) if states == 4 else np.full(states, 1.0 / states)
    subst = np.stack([r0 * np.exp(rng.normal(0, 0.5, r0.shape)) for _ in range(n)])
    subst /= subst[:, -1:]
    freqs = np.stack([f0 * np.exp(rng.normal(0, 0.3, f0.shape)) for _ in range(n)])
    freqs /= freqs.sum(axis=1, keepdims=True)
    return freqs, subst


def mixture_lnl_by_categories(make_engine, net, part, cat_model, freqs, subst):
    """The mixture lnL assembled from SINGLE-matrix, single-category, single-pattern evaluations only:
    lnL = sum_sites w_site log sum_c weight_c L_c(site), L_c = site likelihood under matrix cat_model[c] at rate rates[c]."""
    total = 0.0
    pw = part.pattern_weights if part.pattern_weights is not None else np.ones(part.sites, dtype=np.uint32)
    for s in range(part.sites):
        lk = 0.0
        for c in range(part.rate_cats):
            m = int(cat_model[c])
            one = Partition(part.states, 1, part.tip_masks[:, s:s + 1], freqs[m], subst[m], [part.rates[c]], rate_weights=[1.0])
            eng = make_engine(net, one)
            lk += part.rate_weights[c] * np.exp(eng.computeLoglikelihood(0, 1))
            eng.close()
        total += float(pw[s]) * np.log(lk)
    return total
