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
