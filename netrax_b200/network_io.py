"""Flat network description + extended-Newick reader / writer + FASTA reader (data formats).

File formats are not part of the hot path; this module feeds the reference's own fixtures
(test/sample_networks/*.nw, *_alignment.txt), user files (scripts/score_network.py) and synthetic networks through
the likelihood API and writes the optimised network back (to_extended_newick).  Numbering follows the reference's
``convertNetworkToplevel`` (src/io/NetworkIO.cpp:59-330): tips 0..n-1 in order of appearance, then
inner tree nodes with the root last, then reticulation nodes; the pmatrix index of a non-reticulation
node's parent edge equals its clv index; reticulation i owns edges base+2i (first parent) and
base+2i+1 (second parent).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

BRPROB_MIN, BRPROB_MAX = 1e-6, 1.0 - 1e-6   # NetraxOptions::brprob_min / brprob_max (src/NetraxOptions.hpp)
BRLEN_MIN = 1e-6   # raxml-ng RAXML_BRLEN_MIN (libs/raxml-ng/src/constants.hpp:14)
BRLEN_MAX = 100.0  # RAXML_BRLEN_MAX (:15)


@dataclass
class NetworkDesc:
    num_tips: int
    num_nodes: int
    root: int
    edge_source: np.ndarray   # uint32[E]  (parent end)
    edge_target: np.ndarray   # uint32[E]  (child end)
    edge_length: np.ndarray   # float64[E]
    edge_prob: np.ndarray     # float64[E] (first-parent edge carries the inheritance probability)
    ret_node: np.ndarray      # uint32[R]
    ret_first_edge: np.ndarray
    ret_second_edge: np.ndarray
    tip_labels: List[str] = field(default_factory=list)

    @property
    def num_edges(self) -> int:
        return int(self.edge_source.shape[0])

    @property
    def num_reticulations(self) -> int:
        return int(self.ret_node.shape[0])


class _PNode:
    __slots__ = ("label", "children", "is_ret", "ret_name", "parents", "lengths", "probs", "index")

    def __init__(self):
        self.label = ""
        self.children: List["_PNode"] = []
        self.is_ret = False
        self.ret_name = ""
        self.parents: List["_PNode"] = []
        self.lengths: List[float] = []
        self.probs: List[float] = []
        self.index = -1


def _split_top(s: str) -> List[str]:
    out, depth, start = [], 0, 0
    for i, c in enumerate(s):
        if c == "(":
            depth += 1
        elif c == ")":
            depth -= 1
        elif c == "," and depth == 0:
            out.append(s[start:i])
            start = i + 1
    out.append(s[start:])
    return out


def _parse_tail(tail: str) -> Tuple[str, str, float, float]:
    """'label#Hname:len:support:prob' -> (label, ret_name, length, prob); missing numbers -> 0."""
    fields = tail.split(":")
    name = fields[0].strip()
    vals = [0.0, 0.0, 0.0]
    for i, f in enumerate(fields[1:4]):
        f = f.strip()
        if f:
            vals[i] = float(f)
    ret_name = ""
    if "#" in name:
        name, ret_name = name.split("#", 1)
    return name, ret_name, vals[0], vals[2]


def parse_extended_newick(newick: str) -> NetworkDesc:
    s = newick.strip()
    if ";" not in s:
        raise ValueError("No semicolon found")
    s = s[: s.index(";")]
    rets: Dict[str, _PNode] = {}
    all_nodes: List[_PNode] = []

    def read(sub: str, parent: Optional[_PNode]) -> _PNode:
        sub = sub.strip()
        lp, rp = sub.find("("), sub.rfind(")")
        tail = sub[rp + 1:] if rp >= 0 else sub
        label, ret_name, length, prob = _parse_tail(tail)
        if ret_name:
            node = rets.get(ret_name)
            if node is None:
                node = _PNode()
                node.is_ret, node.ret_name, node.label = True, ret_name, label
                rets[ret_name] = node
                all_nodes.append(node)
        else:
            node = _PNode()
            node.label = label
            all_nodes.append(node)
        if parent is not None:
            node.parents.append(parent)
            node.lengths.append(length)
            node.probs.append(prob)
        if lp >= 0 and rp >= 0:
            for c in _split_top(sub[lp + 1: rp]):
                node.children.append(read(c, node))
        return node

    root = read(s, None)
    while len(root.children) == 1 and not root.is_ret:  # makeToplevel, NetworkIO.cpp:332-352
        all_nodes.remove(root)
        root = root.children[0]
        root.parents, root.lengths, root.probs = [], [], []
    if len(root.children) == 3:  # enforceToplevelBifurcation, RootedNetworkParser.cpp:273-286
        nn = _PNode()
        nn.children = root.children[1:]
        for c in nn.children:
            c.parents[c.parents.index(root)] = nn
        nn.parents, nn.lengths, nn.probs = [root], [0.0], [0.0]
        root.children = [root.children[0], nn]
        all_nodes.append(nn)
    if len(root.children) > 3:
        raise ValueError("The network is not bifurcating")

    tips = [n for n in all_nodes if not n.children]
    inner = [n for n in all_nodes if n.children and not n.is_ret and n is not root] + [root]
    retn = [n for n in all_nodes if n.is_ret]
    for n in retn:
        if len(n.parents) != 2 or len(n.children) != 1:
            raise ValueError(f"reticulation {n.ret_name} must have two parents and one child")
    for i, n in enumerate(tips + inner + retn):
        n.index = i
    nt, ni = len(tips), len(inner)
    E = nt + ni - 1 + 2 * len(retn)
    src = np.zeros(E, np.uint32); tgt = np.zeros(E, np.uint32)
    length = np.zeros(E); prob = np.ones(E)
    for n in tips + inner[:-1]:
        if len(n.parents) != 1:
            raise ValueError("tree node with != 1 parent")
        src[n.index], tgt[n.index], length[n.index] = n.parents[0].index, n.index, n.lengths[0]
    base = nt + ni - 1
    rf, rs = [], []
    for i, n in enumerate(retn):
        p0, p1 = n.probs
        if p0 == 0 and p1 == 0:       # RootedNetworkParser.cpp:317-324: probabilities not given -> 0.5 / 0.5
            p0 = p1 = 0.5
        else:                         # :325-345: clamp each to [brprob_min, brprob_max], then the two must sum to 1
            p0, p1 = min(max(p0, BRPROB_MIN), BRPROB_MAX), min(max(p1, BRPROB_MIN), BRPROB_MAX)
            if abs(1.0 - (p0 + p1)) >= 1e-3:
                raise ValueError(f"Reticulation probs do not sum up to 1 (reticulation {n.ret_name}: {p0} + {p1})")
        for k in (0, 1):
            e = base + 2 * i + k
            src[e], tgt[e], length[e] = n.parents[k].index, n.index, n.lengths[k]
        prob[base + 2 * i], prob[base + 2 * i + 1] = p0, 1.0 - p0
        rf.append(base + 2 * i); rs.append(base + 2 * i + 1)
    length = np.clip(length, BRLEN_MIN, BRLEN_MAX)  # src/graph/AnnotatedNetwork.cpp:420-437
    return NetworkDesc(nt, nt + ni + len(retn), root.index, src, tgt, length, prob,
                       np.array([n.index for n in retn], np.uint32), np.array(rf, np.uint32),
                       np.array(rs, np.uint32), [t.label for t in tips])


def read_fasta(text: str) -> Dict[str, str]:
    seqs: Dict[str, str] = {}
    name = None
    for line in text.splitlines():
        line = line.strip()
        if not line:
            continue
        if line.startswith(">"):
            name = line[1:].strip()
            seqs[name] = ""
        elif name is not None:
            seqs[name] += line
    return seqs


# IUPAC nucleotide codes -> 4-bit state masks (A=1, C=2, G=4, T=8), the values of libpll's pll_map_nt.
_DNA = {"A": 1, "C": 2, "G": 4, "T": 8, "U": 8, "M": 3, "R": 5, "W": 9, "S": 6, "Y": 10, "K": 12,
        "V": 7, "H": 11, "D": 13, "B": 14, "N": 15, "O": 15, "X": 15, "-": 15, "?": 15}
_AA_ORDER = "ARNDCQEGHILKMFPSTWYV"


def encode_dna(seq: str) -> np.ndarray:
    return np.array([_DNA[c.upper()] for c in seq], dtype=np.uint32)


def encode_aa(seq: str) -> np.ndarray:
    out = np.zeros(len(seq), np.uint32)
    for i, c in enumerate(seq.upper()):
        if c in _AA_ORDER:
            out[i] = 1 << _AA_ORDER.index(c)
        elif c == "B":
            out[i] = (1 << 2) | (1 << 3)
        elif c == "Z":
            out[i] = (1 << 5) | (1 << 6)
        elif c == "J":
            out[i] = (1 << 9) | (1 << 10)
        else:
            out[i] = (1 << 20) - 1
    return out


def compress_patterns(masks: np.ndarray) -> Tuple[np.ndarray, np.ndarray]:
    """[tips, sites] masks -> (unique columns [tips, patterns], weights uint32[patterns]), stable order."""
    cols, inverse = np.unique(masks.T, axis=0, return_inverse=True)
    first = np.full(cols.shape[0], masks.shape[1], np.int64)
    np.minimum.at(first, inverse.ravel(), np.arange(masks.shape[1]))
    order = np.argsort(first, kind="stable")
    weights = np.bincount(inverse.ravel(), minlength=cols.shape[0]).astype(np.uint32)
    return np.ascontiguousarray(cols[order].T.astype(np.uint32)), weights[order]


def to_extended_newick(net: NetworkDesc, branch_lengths: Optional[Sequence[float]] = None,
                       reticulation_probs: Optional[Sequence[float]] = None, precision: Optional[int] = None) -> str:
    """The writer side of the reference's network format (toExtendedNewick / printNodeNewick / newickNodeName,
    src/io/NetworkIO.cpp:384-452,510-523): children in parentheses, tips by label, a reticulation node printed under
    BOTH parents as ``#H<reticulation index>:length::prob`` (support left empty; first parent carries prob, second
    1 - prob) with its subtree expanded only at the first visit.  ``branch_lengths`` [E] / ``reticulation_probs`` [R]
    override the description's values, which is what updateNetwork (:493-508) does with the optimised state before
    writing.  ``precision`` = significant digits (the reference streams doubles with the default 6); None writes the
    shortest representation that reads back to the same double."""
    length = np.asarray(net.edge_length if branch_lengths is None else branch_lengths, float)
    if length.shape[0] < net.num_edges:
        raise ValueError("branch_lengths must cover every edge")
    first_prob = {int(net.ret_node[i]): (float(net.edge_prob[int(net.ret_first_edge[i])]) if reticulation_probs is None
                                         else float(reticulation_probs[i])) for i in range(net.num_reticulations)}
    ret_index = {int(v): i for i, v in enumerate(net.ret_node)}
    second_edges = set(int(e) for e in net.ret_second_edge)
    kids: Dict[int, List[int]] = {}
    for e in range(net.num_edges):
        kids.setdefault(int(net.edge_source[e]), []).append(e)

    def num(x: float) -> str:
        return repr(float(x)) if precision is None else f"{float(x):.{precision}g}"

    visited = set()

    def emit(node: int, via_edge: Optional[int]) -> str:
        out = ""
        is_ret = node in ret_index
        ch = kids.get(node, [])
        if is_ret and not ch:
            raise ValueError("Encountered a reticulation node that has no children")
        if ch and node not in visited:
            out += "(" + ",".join(emit(int(net.edge_target[e]), e) for e in ch) + ")"
            if is_ret:
                visited.add(node)
        if node < net.num_tips and net.tip_labels:
            out += net.tip_labels[node]
        if is_ret:
            p = first_prob[node]
            out += f"#H{ret_index[node]}:{num(length[via_edge])}::{num(1.0 - p if via_edge in second_edges else p)}"
        elif via_edge is not None:
            out += ":" + num(length[via_edge])
        return out

    return emit(int(net.root), None) + ";"
