"""Seeded synthetic inputs for tests and the bench (SURVEY.md §8d "Synthetic inputs").

topology : random rooted binary tree (Yule) + r random acyclic arc insertions — the analogue of the
           reference's build_random_annotated_network + add_extra_reticulations
           (src/graph/AnnotatedNetwork.cpp:263-306); branch lengths ~ Exp(mean 0.1) clamped to raxml-ng's
           [1e-6, 100]; first-parent probabilities ~ U(0.2, 0.8).
alignment: simulated down displayed tree 0 (every reticulation takes its first parent) under the
           model, 5 % of cells -> gap (fully ambiguous), de-duplicated to exactly `patterns` columns,
           weights = multiplicities; or uniform random cells (worst case for scaling).
model    : GTR rates (1, 2.5, 0.8, 1.2, 3.0, 1), pi = (0.3, 0.2, 0.2, 0.3), Gamma(alpha=0.5), 4 cats.
"""
from __future__ import annotations

from typing import List, Optional, Tuple

import numpy as np

from .network_io import BRLEN_MAX, BRLEN_MIN, NetworkDesc

GTR_RATES = np.array([1.0, 2.5, 0.8, 1.2, 3.0, 1.0])
DNA_FREQS = np.array([0.3, 0.2, 0.2, 0.3])
# discrete Gamma(0.5), 4 categories, mean mode — what pll_compute_gamma_cats(0.5, 4, PLL_GAMMA_RATES_MEAN)
# returns (libpll/src/gamma.c:267-330); tests/test_host_model.py re-derives it through the host library.
GAMMA4_ALPHA05 = np.array([0.03338775337571123, 0.25191592470299234, 0.8202684786530262, 2.894427849944841])


def lg_model():
    """(rates[190], freqs[20]) of the LG amino-acid model (netrax_b200/lg_model.json, extracted from the reference's
    libpll constants by tests/golden/make_lg_model.py).  The table sums to 1 + 1e-6 exactly as libpll ships it; the engine
    renormalises it the way pll_set_frequencies does."""
    import json
    import os
    d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "lg_model.json")))
    return np.asarray(d["rates"], float), np.asarray(d["freqs"], float)


def random_network(n_taxa: int, n_ret: int, seed: int = 42, mean_brlen: float = 0.1) -> NetworkDesc:
    rng = np.random.default_rng(seed)
    # --- Yule tree as parent->children lists over provisional ids -------------------------------
    children = {0: []}
    leaves = [0]
    nxt = 1
    while len(leaves) < n_taxa:
        v = leaves.pop(int(rng.integers(len(leaves))))
        a, b = nxt, nxt + 1
        nxt += 2
        children[v] = [a, b]
        children[a], children[b] = [], []
        leaves += [a, b]
    edges: List[List] = []  # [parent, child, length]
    for v, ch in children.items():
        for c in ch:
            edges.append([v, c, float(rng.exponential(mean_brlen))])
    is_ret = {}
    ret_list: List[Tuple[int, int, int, float]] = []  # (node, first_parent, second_parent, prob)

    def descendants(v: int) -> set:
        out, stack = set(), [v]
        while stack:
            x = stack.pop()
            if x in out:
                continue
            out.add(x)
            stack += [c for (p, c, _) in edges if p == x]
        return out

    tries = 0
    while len(ret_list) < n_ret:
        tries += 1
        if tries > 10000:
            raise RuntimeError("could not place reticulations")
        i1, i2 = (int(x) for x in rng.integers(len(edges), size=2))
        if i1 == i2:
            continue
        u1, v1, l1 = edges[i1]
        u2, v2, l2 = edges[i2]
        if u1 in descendants(v2):  # x would become a descendant of h -> cycle
            continue
        x, h = nxt, nxt + 1
        nxt += 2
        f1, f2 = float(rng.uniform(0.2, 0.8)), float(rng.uniform(0.2, 0.8))
        e1a, e1b = [u1, x, l1 * f1], [x, v1, l1 * (1 - f1)]
        e2a, e2b = [u2, h, l2 * f2], [h, v2, l2 * (1 - f2)]
        # if a split edge ends in a reticulation node, that node's first/second parent record follows the split
        def fix_parent(old_parent, node, new_parent):
            for k, (rn, fp, sp, pr) in enumerate(ret_list):
                if rn == node:
                    if fp == old_parent:
                        ret_list[k] = (rn, new_parent, sp, pr)
                    elif sp == old_parent:
                        ret_list[k] = (rn, fp, new_parent, pr)
        fix_parent(u1, v1, x)
        fix_parent(u2, v2, h)
        for idx, repl in sorted(((i1, [e1a, e1b]), (i2, [e2a, e2b])), reverse=True):
            edges[idx:idx + 1] = repl
        edges.append([x, h, float(rng.exponential(mean_brlen))])
        is_ret[h] = True
        ret_list.append((h, u2, x, float(rng.uniform(0.2, 0.8))))

    # --- final numbering: tips, inner tree nodes (root last), reticulations --------------------------
    nodes = set([0])
    for p, c, _ in edges:
        nodes.add(p); nodes.add(c)
    has_child = {p for p, _, _ in edges}
    tips = sorted(n for n in nodes if n not in has_child)
    retn = [r[0] for r in ret_list]
    inner = sorted(n for n in nodes if n in has_child and n not in is_ret and n != 0) + [0]
    idx = {n: i for i, n in enumerate(tips + inner + retn)}
    nt, ni, nr = len(tips), len(inner), len(retn)
    assert nt == n_taxa
    E = nt + ni - 1 + 2 * nr
    src = np.zeros(E, np.uint32); tgt = np.zeros(E, np.uint32); length = np.zeros(E); prob = np.ones(E)
    base = nt + ni - 1
    rinfo = {r[0]: (k, r[1], r[2], r[3]) for k, r in enumerate(ret_list)}
    seen = set()
    for p, c, l in edges:
        if c in rinfo:
            k, fp, sp, pr = rinfo[c]
            e = base + 2 * k + (0 if p == fp else 1)
            assert p in (fp, sp)
        else:
            e = idx[c]
        assert e not in seen
        seen.add(e)
        src[e], tgt[e], length[e] = idx[p], idx[c], l
    for k, r in enumerate(ret_list):
        prob[base + 2 * k], prob[base + 2 * k + 1] = r[3], 1.0 - r[3]
    assert len(seen) == E
    length = np.clip(length, BRLEN_MIN, BRLEN_MAX)
    return NetworkDesc(nt, nt + ni + nr, idx[0], src, tgt, length, prob, np.array([idx[r] for r in retn], np.uint32),
                       np.array([base + 2 * k for k in range(nr)], np.uint32),
                       np.array([base + 2 * k + 1 for k in range(nr)], np.uint32), [f"T{i}" for i in range(nt)])


def caterpillar_network(n_taxa: int, seed: int = 7, brlen: float = 0.6) -> NetworkDesc:
    """Scaler-stress fixture: caterpillar with one reticulation near the top; long branches force >= 1
    scaling event per site path (SURVEY §8d "scaler stress fixture")."""
    from .network_io import parse_extended_newick
    s = f"T0:{brlen}"
    for i in range(1, n_taxa - 2):
        s = f"({s},T{i}:{brlen}):{brlen}"
    a, b = n_taxa - 2, n_taxa - 1
    nw = f"(({s},(T{a}:{brlen})X#H1:{brlen}::0.4):{brlen},(X#H1:{brlen}::0.6,T{b}:{brlen}):{brlen});"
    return parse_extended_newick(nw)


def _q_matrix(rates: np.ndarray, freqs: np.ndarray) -> np.ndarray:
    n = len(freqs)
    q = np.zeros((n, n))
    k = 0
    for i in range(n):
        for j in range(i + 1, n):
            q[i, j] = rates[k] * freqs[j]
            q[j, i] = rates[k] * freqs[i]
            k += 1
    q -= np.diag(q.sum(1))
    q /= -(freqs * np.diag(q)).sum()
    return q


def _pmat(q: np.ndarray, t: float) -> np.ndarray:
    w, v = np.linalg.eig(q)
    p = (v * np.exp(w * t)) @ np.linalg.inv(v)
    p = np.clip(p.real, 0, None)
    return p / p.sum(1, keepdims=True)


def simulate_alignment(net: NetworkDesc, patterns: int, seed: int = 1, states: int = 4, rates=GTR_RATES, freqs=DNA_FREQS,
                       cat_rates=GAMMA4_ALPHA05, gap_frac: float = 0.05, random_cells: bool = False,
                       dedup: bool = True) -> Tuple[np.ndarray, np.ndarray]:
    """-> (tip_masks uint32[tips, patterns], weights uint32[patterns])."""
    rng = np.random.default_rng(seed)
    nt = net.num_tips
    full = (1 << states) - 1
    out_cols: Optional[np.ndarray] = None
    weights = None
    need = patterns
    chunks = []
    total_unique = 0
    q = None if random_cells else _q_matrix(np.asarray(rates, float), np.asarray(freqs, float))
    # displayed tree 0: drop second-parent edges
    second = set(int(e) for e in net.ret_second_edge)
    kids = {}
    for e in range(net.num_edges):
        if e in second:
            continue
        kids.setdefault(int(net.edge_source[e]), []).append((int(net.edge_target[e]), float(net.edge_length[e])))
    while need > 0:
        n = int(need * 1.08) + 64
        if random_cells:
            st = rng.integers(states, size=(nt, n))
        else:
            cat = rng.integers(len(cat_rates), size=n)
            st_all = {net.root: rng.choice(states, size=n, p=np.asarray(freqs) / np.sum(freqs))}
            stack = [net.root]
            st = np.zeros((nt, n), np.int64)
            while stack:
                v = stack.pop()
                sv = st_all.pop(v)
                if v < nt:
                    st[v] = sv
                    continue
                for c, t in kids.get(v, []):
                    # one pass per edge: cumulative rows of P(t * r_cat) gathered per column by (category, parent state)
                    cum = np.stack([np.cumsum(_pmat(q, t * r), axis=1) for r in cat_rates]).astype(np.float32)
                    u = rng.random(n, dtype=np.float32)
                    sc = np.minimum((u[:, None] > cum[cat, sv]).sum(1), states - 1)
                    st_all[c] = sc
                    stack.append(c)
        masks = (1 << st).astype(np.uint32)
        if gap_frac > 0:
            masks[rng.random(masks.shape) < gap_frac] = full
        chunks.append(masks)
        allm = np.concatenate(chunks, axis=1)
        if not dedup:
            return np.ascontiguousarray(allm[:, :patterns]), np.ones(patterns, np.uint32)
        from .network_io import compress_patterns
        cols, w = compress_patterns(allm)
        total_unique = cols.shape[1]
        if total_unique >= patterns:
            return np.ascontiguousarray(cols[:, :patterns]), np.ascontiguousarray(w[:patterns])
        need = patterns - total_unique
    raise AssertionError


def sum_trees_per_node(eng) -> int:
    """Σ_nodes trees(node) over inner nodes — the CLV-slot count the metric is defined on."""
    return sum(eng.num_trees(v) for v in range(eng.net.num_tips, eng.net.num_nodes))
