"""The extended-Newick reader (netrax_b200/network_io.py, SURVEY §8f f4 "on-disk formats") on the format variants of the
reference's test/src/NetworkIOTest.cpp:189-290 (reticulation labels with / without name, length, support, probability in
every combination) plus its sanity checks (test/src/NetworkIOTest.cpp sanity_checks: tips first, every reticulation has
two distinct parents and one child, probabilities of the two arcs sum to 1), and on the reference's fixture networks."""
import os

import numpy as np
import pytest

from helpers import FIXTURE_PAIRS, FIX
from netrax_b200.network_io import parse_extended_newick

CASES = [  # (name in NetworkIOTest.cpp, input, tips, reticulations, first-parent prob or None)
    ("reticulationHasLeafChild", "((A:2,(B:1)X#H1)Q:2,(D:2,X#H1)R:2);", 3, 1, None),
    ("readSimpleNetworkReticulationNoExtra", "((A:2,((B:1,C:1)P:1)X#H1)Q:2,(D:2,X#H1)R:2);", 4, 1, None),
    ("readSimpleNetworkReticulationNoExtraNoLabel", "((A:2,((B:1,C:1)P:1)#H1)Q:2,(D:2,#H1)R:2);", 4, 1, None),
    ("readSimpleNetworkReticulationOnlyLength", "((A:2,((B:1,C:1)P:1)X#H1:0)Q:2,(D:2,X#H1:0)R:2);", 4, 1, None),
    ("readSimpleNetworkReticulationOnlyProb", "((A:2,((B:1,C:1)P:1)X#H1:::0.3)Q:2,(D:2,X#H1:::0.7)R:2);", 4, 1, 0.3),
    ("readSimpleNetworkReticulationOnlySupport", "((A:2,((B:1,C:1)P:1)X#H1::0)Q:2,(D:2,X#H1::0)R:2);", 4, 1, None),
    ("readSimpleNetworkReticulationLengthAndSupport", "((A:2,((B:1,C:1)P:1)X#H1:0:0)Q:2,(D:2,X#H1:0:0)R:2);", 4, 1, None),
    ("readSimpleNetworkReticulationLengthAndProb", "((A:2,((B:1,C:1)P:1)X#H1:0::1)Q:2,(D:2,X#H1:0::0)R:2);", 4, 1, 1.0),
    ("readSimpleNetworkReticulationSupportAndProb", "((A:2,((B:1,C:1)P:1)X#H1::0:1)Q:2,(D:2,X#H1::0:0)R:2);", 4, 1, 1.0),
    ("readSimpleNetworkLowercaseTaxa", "((a:2,((b:1,c:1)P:1)X#H1:0::0.3)Q:2,(d:2,X#H1:0::0.7)R:2);", 4, 1, 0.3),
    ("a tree", "((A:1,B:1):1,(C:1,D:1):1);", 4, 0, None),
]


def sanity_checks(net):
    assert net.num_nodes == net.num_edges + 1 - net.num_reticulations   # every node but the root has one incoming edge, reticulations two
    indeg = np.zeros(net.num_nodes, dtype=int)
    outdeg = np.zeros(net.num_nodes, dtype=int)
    for e in range(net.num_edges):
        indeg[net.edge_target[e]] += 1
        outdeg[net.edge_source[e]] += 1
    assert indeg[net.root] == 0
    for v in range(net.num_tips):
        assert outdeg[v] == 0 and indeg[v] == 1          # tips come first (clv_index < num_tips)
    for v in range(net.num_tips, net.num_nodes):
        assert outdeg[v] in (1, 2)
    for r in range(net.num_reticulations):
        v, e1, e2 = int(net.ret_node[r]), int(net.ret_first_edge[r]), int(net.ret_second_edge[r])
        assert indeg[v] == 2 and outdeg[v] == 1
        assert net.edge_target[e1] == v and net.edge_target[e2] == v and net.edge_source[e1] != net.edge_source[e2]
        assert net.edge_prob[e1] + net.edge_prob[e2] == pytest.approx(1.0)
    assert len(set(net.tip_labels)) == net.num_tips


@pytest.mark.parametrize("name,text,tips,rets,prob", CASES, ids=[c[0] for c in CASES])
def test_network_io_format_variants(name, text, tips, rets, prob):
    net = parse_extended_newick(text)
    assert net.num_tips == tips and net.num_reticulations == rets
    sanity_checks(net)
    if prob is not None:
        assert float(net.edge_prob[int(net.ret_first_edge[0])]) == pytest.approx(min(max(prob, 1e-6), 1.0 - 1e-6), rel=1e-12)   # clamped to [brprob_min, brprob_max]
    elif rets:
        assert float(net.edge_prob[int(net.ret_first_edge[0])]) == pytest.approx(0.5)   # unspecified: 0.5 / 0.5


@pytest.mark.parametrize("name", list(FIXTURE_PAIRS))
def test_reference_fixture_networks_parse(name):
    nw, _ = FIXTURE_PAIRS[name]
    net = parse_extended_newick(open(os.path.join(FIX, nw)).read())
    sanity_checks(net)


def _canonical(net):
    """Numbering-free description: per edge (tips below the child, tips below the parent, length, prob), sorted."""
    below = {}

    def tips_below(v):
        if v not in below:
            ch = [int(net.edge_target[e]) for e in range(net.num_edges) if int(net.edge_source[e]) == v]
            below[v] = frozenset([net.tip_labels[v]]) if not ch else frozenset().union(*(tips_below(c) for c in ch))
        return below[v]

    rows = []
    for e in range(net.num_edges):
        rows.append((sorted(tips_below(int(net.edge_target[e]))), sorted(tips_below(int(net.edge_source[e]))),
                     float(net.edge_length[e]), float(net.edge_prob[e])))
    return sorted(rows)


@pytest.mark.parametrize("name", sorted(n for n in os.listdir(FIX) if n.endswith(".nw")))
def test_extended_newick_writer_round_trips_reference_fixtures(name):
    """toExtendedNewick (src/io/NetworkIO.cpp:384-452,510-523): what we write parses back to the same network — same
    clusters under every edge, same lengths and inheritance probabilities, same displayed-tree count."""
    from netrax_b200.network_io import to_extended_newick
    net = parse_extended_newick(open(os.path.join(FIX, name)).read())
    text = to_extended_newick(net)
    back = parse_extended_newick(text)
    assert (back.num_tips, back.num_nodes, back.num_edges, back.num_reticulations) == (net.num_tips, net.num_nodes, net.num_edges, net.num_reticulations)
    assert sorted(back.tip_labels) == sorted(net.tip_labels)
    assert _canonical(back) == _canonical(net)
    assert text.count("#H") == 2 * net.num_reticulations and text.endswith(";")


def test_extended_newick_writer_takes_the_optimised_state_and_the_reference_precision():
    from netrax_b200.network_io import to_extended_newick
    net = parse_extended_newick("((A:0.1,(B:0.2)X#H1:0.3::0.4)P:0.5,(X#H1:0.6::0.6,C:0.7)Q:0.8)R;")
    brl = net.edge_length * 2.0
    text = to_extended_newick(net, branch_lengths=brl, reticulation_probs=[0.25], precision=6)
    back = parse_extended_newick(text)
    np.testing.assert_allclose(sorted(back.edge_length), sorted(brl), rtol=1e-6)
    e = int(back.ret_first_edge[0])
    assert back.edge_prob[e] == pytest.approx(0.25) and back.edge_prob[int(back.ret_second_edge[0])] == pytest.approx(0.75)
    assert "#H0:0.6::0.25" in text and "#H0:1.2::0.75" in text   # newickNodeName: "#H" + reticulation index, empty support field


def test_engine_newick_writes_the_current_state_and_averages_unlinked_lengths():
    """toExtendedNewick(ann_network) = updateNetwork + write (src/io/NetworkIO.cpp:493-523); unlinked analyses write the
    partition-weighted average (collect_average_branches, :455-491).  Driven over the oracle engine (the wrapper is shared)."""
    from netrax_b200._capi import UNLINKED, Partition
    from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment
    from oracle import oracle
    net = random_network(6, 1, seed=2)
    parts = []
    for k, n in enumerate((100, 300)):
        m, w = simulate_alignment(net, n, seed=20 + k)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    brl = [net.edge_length * 1.0, net.edge_length * 3.0]
    e = oracle.make_engine("port", net, parts, linkage=UNLINKED, partition_brlens=brl)
    e.set_reticulation_prob(0, 0.3)
    back = parse_extended_newick(e.toExtendedNewick())
    w = np.array([float(p.pattern_weights.sum()) for p in parts])
    want = (brl[0] * w[0] + brl[1] * w[1]) / w.sum()
    np.testing.assert_allclose(sorted(back.edge_length), sorted(want), rtol=1e-12)
    assert back.edge_prob[int(back.ret_first_edge[0])] in (pytest.approx(0.3), pytest.approx(0.7))
    e.close()


@pytest.mark.parametrize("taxa,rets,seed", [(8, 1, 1), (20, 4, 2), (50, 8, 3), (100, 8, 47)])
def test_extended_newick_writer_round_trips_synthetic_networks(taxa, rets, seed):
    """The bench's synthetic networks (up to the headline 100 taxa / 8 reticulations) survive write -> read: same clusters,
    lengths and inheritance probabilities on every edge, and the same number of displayed trees at the root."""
    from netrax_b200.network_io import to_extended_newick
    from netrax_b200.synth import random_network
    net = random_network(taxa, rets, seed=seed)
    back = parse_extended_newick(to_extended_newick(net))
    assert (back.num_tips, back.num_edges, back.num_reticulations) == (net.num_tips, net.num_edges, net.num_reticulations)
    got, want = _canonical(back), _canonical(net)
    assert [(r[0], r[1]) for r in got] == [(r[0], r[1]) for r in want]
    np.testing.assert_allclose([r[2] for r in got], [r[2] for r in want], rtol=0, atol=0)
    np.testing.assert_allclose([r[3] for r in got], [r[3] for r in want], rtol=1e-15)


def test_reticulation_probabilities_are_clamped_and_checked_like_the_reference():
    """RootedNetworkParser.cpp:317-345: probabilities not given -> 0.5 / 0.5; otherwise each is clamped to [1e-6, 1 - 1e-6] and
    the pair must sum to 1 within 1e-3 (the reference throws)."""
    from netrax_b200.network_io import parse_extended_newick
    nw = "((A:1,(B:1)X#H1:1::{p0}):1,(X#H1:1::{p1},C:1):1);"
    net = parse_extended_newick(nw.format(p0="0.4", p1="0.6"))
    assert net.edge_prob[net.ret_first_edge[0]] == pytest.approx(0.4) and net.edge_prob[net.ret_second_edge[0]] == pytest.approx(0.6)
    net = parse_extended_newick(nw.format(p0="1e-10", p1="1.0"))        # clamped, and 1e-6 + (1 - 1e-6) still sums to 1
    assert net.edge_prob[net.ret_first_edge[0]] == 1e-6 and np.log(net.edge_prob).min() > -14
    for bad in (("0.4", "0.4"), ("0.3", "0"), ("0.7", "0.7")):      # inconsistent, or one probability missing
        with pytest.raises(ValueError, match="do not sum up to 1"):
            parse_extended_newick(nw.format(p0=bad[0], p1=bad[1]))
    net = parse_extended_newick("((A:1,(B:1)X#H1:1):1,(X#H1:1,C:1):1);")   # not given at all
    assert net.edge_prob[net.ret_first_edge[0]] == 0.5
