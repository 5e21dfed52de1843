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
        assert float(net.edge_prob[int(net.ret_first_edge[0])]) == pytest.approx(prob)
    elif rets:
        assert float(net.edge_prob[int(net.ret_first_edge[0])]) == pytest.approx(0.5)   # unspecified: 0.5 / 0.5


@pytest.mark.parametrize("name", list(FIXTURE_PAIRS))
def test_reference_fixture_networks_parse(name):
    nw, _ = FIXTURE_PAIRS[name]
    net = parse_extended_newick(open(os.path.join(FIX, nw)).read())
    sanity_checks(net)
