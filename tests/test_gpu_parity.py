"""GPU parity tests proper (-m gpu): the CUDA path, called through the C-ABI, against the oracle on the same
seeded inputs.  Bars (BASELINE.json north_star): lnL within 1e-10 relative, derivatives within 1e-8 relative,
scaler counts and BEST-tree selection bit-exact.  DNA CLVs are additionally required to be BIT-IDENTICAL when the
engine is fed the oracle's eigen-decomposition (the kernels mirror the reference's operation order)."""
import numpy as np
import pytest

from helpers import FIXTURE_PAIRS, fixture_summary, load_fixture, load_golden
from netrax_b200._capi import AVERAGE, BEST, LINKED, UNLINKED, Partition
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, caterpillar_network, random_network, simulate_alignment

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-10     # north_star: per-site and total lnL within 1e-10 relative
DERIV_RTOL = 1e-8    # north_star: derivatives within 1e-8 relative
# A replayed plan computes bit-identical CLVs and scalers; the per-tree lnL is a sum over patterns whose ORDER depends on the kernel
# that runs the replay (one-launch tile walk: per 32-pattern tile, then over tiles; level-by-level graph: k_term_lnl_sum's grid-stride
# order) — each deterministic, equal to rounding.
REPLAY_RTOL = 1e-13


def _gpu(net, parts, **kw):
    from netrax_b200.engine import NetraxB200
    return NetraxB200(net, parts, **kw)


def _oracle(net, parts, **kw):
    from oracle import oracle
    return oracle.make_engine("ref" if oracle.have_ref() else "port", net, parts, **kw)


def _inject_eigen(g, o):
    for p in range(g.P):
        g.set_eigen(p, *o.get_eigen(p))


def _compare_all_clvs(g, o, exact):
    net = g.net
    for v in range(net.num_tips, net.num_nodes):
        assert g.num_trees(v) == o.num_trees(v), v
        for t in range(g.num_trees(v)):
            assert g.tree_config(v, t) == o.tree_config(v, t)
            for p in range(g.P):
                assert np.array_equal(g.read_scaler(v, t, p), o.read_scaler(v, t, p)), (v, t, p)   # bit-exact integers
                a, b = g.read_clv(v, t, p), o.read_clv(v, t, p)
                if exact:
                    assert np.array_equal(a, b), (v, t, p, np.abs(a - b).max())
                else:
                    np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)


GOLD = load_golden("netrax_fixtures_golden.json")["cases"]


@pytest.mark.parametrize("name", list(FIXTURE_PAIRS))
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_reference_fixtures_match_golden(name, variant):
    """The committed golden files were produced by the reference's real libpll (oracle kind 'reference')."""
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    g = _gpu(net, [part], variant=variant)
    got = fixture_summary(g)
    exp = GOLD[f"{name}/{'AVERAGE' if variant == AVERAGE else 'BEST'}"]
    assert got["lnl"] == pytest.approx(exp["lnl"], rel=LNL_RTOL)
    assert [t["config"] for t in got["root_trees"]] == [t["config"] for t in exp["root_trees"]]
    for a, b in zip(got["root_trees"], exp["root_trees"]):
        assert a["logprob"] == pytest.approx(b["logprob"], rel=1e-13, abs=1e-300)
        assert a["partition_logl"] == pytest.approx(b["partition_logl"], rel=LNL_RTOL)
    assert got["nodes"].keys() == exp["nodes"].keys()
    for k in got["nodes"]:
        assert got["nodes"][k]["scaler_sum"] == exp["nodes"][k]["scaler_sum"], k
        assert got["nodes"][k]["clv_sum"] == pytest.approx(exp["nodes"][k]["clv_sum"], rel=1e-11), k
    g.close()


@pytest.mark.parametrize("cfg", [(8, 1, 257, 1), (20, 1, 1000, 2), (25, 3, 333, 3), (40, 5, 150, 4), (12, 0, 64, 5)])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_full_evaluation_bitexact_clvs(cfg, variant):
    n, r, pat, seed = cfg
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    o = _oracle(net, [part], variant=variant)
    g = _gpu(net, [part], variant=variant)
    # (1) own eigen-decomposition + device P-matrices: tolerance-level agreement
    lo, lg = o.computeLoglikelihood(0, 1), g.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    for e in range(net.num_edges + 1):
        np.testing.assert_allclose(g.get_pmatrix(e), o.get_pmatrix(e), rtol=1e-12, atol=1e-15)
    _compare_all_clvs(g, o, exact=False)
    # (2) the oracle's eigen-decomposition injected: P-matrices within 1 ulp-ish, CLVs bit-identical whenever P is
    _inject_eigen(g, o)
    lg = g.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=1e-13)
    same_p = all(np.array_equal(g.get_pmatrix(e), o.get_pmatrix(e)) for e in range(net.num_edges + 1))
    _compare_all_clvs(g, o, exact=same_p)
    root = net.root
    for t in range(g.num_trees(root)):
        assert g.tree_info(root, t)[1] == pytest.approx(o.tree_info(root, t)[1], rel=LNL_RTOL)
    if variant == BEST:  # BEST-tree selection bit-exact
        best_g = max(range(g.num_trees(root)), key=lambda t: (g.tree_info(root, t)[0] + g.tree_info(root, t)[1][0], -t))
        best_o = max(range(o.num_trees(root)), key=lambda t: (o.tree_info(root, t)[0] + o.tree_info(root, t)[1][0], -t))
        assert best_g == best_o
    # second full evaluation goes through the cached plan: identical result
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)
    g.close()


def _persite_cases():
    """(name, net, partition, pinv): one case per K3 kernel variant and per branch inside it."""
    from oracle import oracle
    from netrax_b200.synth import lg_model
    cases = []
    net = random_network(10, 1, seed=9)
    m, w = simulate_alignment(net, 300, seed=9)
    cases.append(("dna4 k_tree_lnl_dna4", net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.0))
    cases.append(("dna4 +I", net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.3))
    cat = caterpillar_network(300)
    m, w = simulate_alignment(cat, 257, seed=63, gap_frac=0.0)
    cases.append(("dna4 scaled sites", cat, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.0))
    cases.append(("protein k_tree_lnl_pc<20>",) + _protein_case(12, 2, 301, 11) + (0.0,))
    cases.append(("protein +I",) + _protein_case(9, 1, 130, 12) + (0.4,))
    pcat = caterpillar_network(120)
    rates, freqs = lg_model()
    m, w = simulate_alignment(pcat, 90, seed=13, states=20, rates=rates, freqs=freqs, gap_frac=0.0)
    cases.append(("protein scaled sites", pcat, Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w), 0.0))
    net = random_network(11, 2, seed=43)
    m, w = simulate_alignment(net, 333, seed=43)
    for cats in (1, 3, 8):   # 3 categories: thread-per-pattern generic kernel; 1 / 8: k_tree_lnl_pc<0>
        r = oracle.api("port").gamma_rates(0.7, cats) if cats > 1 else np.ones(1)
        cases.append((f"dna {cats} categories", net, Partition(4, cats, m, DNA_FREQS, GTR_RATES, r, pattern_weights=w), 0.0))
    return cases


def test_persite_lnl_matches_oracle():
    """north_star: PER-SITE lnL within 1e-10 relative.  Every site of every root displayed tree against the `persite_lnl` array
    the reference's pll_compute_root_loglikelihood fills (LH/ImprovedLoglikelihood.cpp:448-453, LIBPLL/likelihood.c:122-184,
    core_likelihood.c:190-200), for every K3 kernel variant, with and without site scaling and +I."""
    from oracle import oracle
    for name, net, part, pinv in _persite_cases():
        g, o = _gpu(net, [part]), _oracle(net, [part])
        _inject_eigen(g, o)
        if pinv:
            g.set_pinv(0, pinv); o.set_pinv(0, pinv)
        assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL), name
        assert g.num_trees(net.root) == o.num_trees(net.root)
        scaled = 0
        for t in range(g.num_trees(net.root)):
            ps, po = g.persite_lnl(t)[0], oracle.persite_lnl(o, t)[0]
            assert np.all(po < 0) and np.all(np.isfinite(po)), name
            np.testing.assert_allclose(ps, po, rtol=LNL_RTOL, atol=0, err_msg=f"{name}, root tree {t}")
            assert ps.sum() == pytest.approx(o.tree_info(net.root, t)[1][0], rel=LNL_RTOL)
            scaled += int(o.read_scaler(net.root, t).sum()) if "scaled" in name else 0
        if "scaled" in name:
            assert scaled > 0, name   # the case really exercises the scaler term
        g.close()


def test_persite_lnl_headline_topology():
    """The same element-wise comparison on BASELINE config 5's network (192 root displayed trees), 700 patterns."""
    import bench
    from oracle import oracle
    net, parts, _ = bench.make_inputs(dict(bench.CONFIGS[5]), 700)
    g, o = _gpu(net, parts), _oracle(net, parts)
    _inject_eigen(g, o)
    g.computeLoglikelihood(0, 1); o.computeLoglikelihood(0, 1)
    for t in range(g.num_trees(net.root)):
        np.testing.assert_allclose(g.persite_lnl(t)[0], oracle.persite_lnl(o, t)[0], rtol=LNL_RTOL, atol=0, err_msg=f"root tree {t}")
    g.close()


def test_incremental_and_cached_semantics():
    net, part = load_fixture(*FIXTURE_PAIRS["three_reticulations"])
    g, o = _gpu(net, [part]), _oracle(net, [part])
    l0 = g.computeLoglikelihood(0, 1)
    assert l0 == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    n0 = g.launch_count()
    assert g.computeLoglikelihood(1, 1) == l0 and g.launch_count() == n0      # cached: nothing launched
    for e in range(net.num_edges):
        old = float(net.edge_length[e])
        for eng in (g, o):
            eng.set_branch_length(e, old * 3 + 0.01)
        l1 = g.computeLoglikelihood(1, 1)
        assert l1 == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
        assert l1 == pytest.approx(g.computeLoglikelihood(0, 1), rel=1e-14)    # incremental == full
        for eng in (g, o):
            eng.set_branch_length(e, old)
        assert g.computeLoglikelihood(1, 1) == pytest.approx(l0, rel=1e-14)
        o.computeLoglikelihood(1, 1)
    # reticulation probability change re-mixes cached per-tree lnLs without touching CLVs
    n1 = g.launch_count()
    for eng in (g, o):
        eng.set_reticulation_prob(0, 0.31)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
    assert g.launch_count() == n1
    g.close()


@pytest.mark.parametrize("name", ["small", "clv_averaging", "two_reticulations", "three_reticulations", "interleaved_reticulations",
                                  "reticulation_in_reticulation", "tree"])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_brlen_flow_every_edge(name, variant):
    """optimize_branch's likelihood calls on every edge: re-rooting preserves lnL (BrlenOptTest.cpp:297-367),
    edge-rooted lnL, sumtables and derivatives match the oracle at several proposal lengths."""
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    g, o = _gpu(net, [part], variant=variant), _oracle(net, [part], variant=variant)
    _inject_eigen(g, o)
    l0 = g.computeLoglikelihood(0, 1)
    assert l0 == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    for e in range(net.num_edges):
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        lb = g.computeLoglikelihoodBrlenOpt(e)
        assert lb == pytest.approx(l0, rel=1e-11), e
        assert lb == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        ng, no = g.computePartitionSumtables(e), o.computePartitionSumtables(e)
        assert ng == no
        for i in range(ng):
            sg, pg, lg_, rg = g.read_sumtable(0, i)
            so, po, lo_, ro = o.read_sumtable(0, i)
            assert (lg_, rg) == (lo_, ro) and pg == pytest.approx(po, rel=1e-14)
            np.testing.assert_allclose(sg, so, rtol=1e-10, atol=1e-14 * np.abs(so).max())  # entries may cancel to ~0
        if ng:
            t0 = float(net.edge_length[e])
            for t in (t0, 0.05, 0.7):
                for eng in (g, o):
                    eng.brlen_set_length(e, t)
                dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
                np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)      # raw (f, d1, d2) per tree pair
                assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
                assert dg[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-7)
                assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
            for eng in (g, o):
                eng.brlen_set_length(e, t0)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(l0, rel=1e-13)
    g.close()


def test_scaler_stress_caterpillar_bitexact_scalers():
    net = caterpillar_network(400)
    m, w = simulate_alignment(net, 1000, seed=11, random_cells=True)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    root = net.root
    mx = 0
    for t in range(g.num_trees(root)):
        sg, so = g.read_scaler(root, t), o.read_scaler(root, t)
        assert np.array_equal(sg, so)
        mx = max(mx, int(sg.max()))
    assert mx >= 2
    g.close()


def test_multi_partition_unlinked_best():
    net = random_network(14, 3, seed=21)
    rng = np.random.default_rng(0)
    parts, brl = [], []
    for p, pat in enumerate((500, 333, 1, 777)):   # ragged partition sizes incl. a single-pattern one
        m, w = simulate_alignment(net, pat, seed=30 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS * 0 + [0.25 + 0.02 * p, 0.25 - 0.02 * p, 0.25, 0.25], GTR_RATES * (1 + 0.1 * p), GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2, net.num_edges))
    for variant in (BEST, AVERAGE):
        g = _gpu(net, parts, variant=variant, linkage=UNLINKED, partition_brlens=brl)
        o = _oracle(net, parts, variant=variant, linkage=UNLINKED, partition_brlens=brl)
        assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
        np.testing.assert_allclose(g.partition_loglh(), o.partition_loglh(), rtol=LNL_RTOL)
        e = int(net.ret_first_edge[1])
        for eng in (g, o):
            eng.brlen_prepare(e)
            eng.computePartitionSumtables(e)
        dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
        np.testing.assert_allclose(dg[2], do[2], rtol=DERIV_RTOL, atol=1e-7)
        np.testing.assert_allclose(dg[3], do[3], rtol=DERIV_RTOL, atol=1e-7)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
        g.close()


def test_model_change_full_reevaluation():
    from oracle import oracle
    net = random_network(9, 2, seed=5)
    m, w = simulate_alignment(net, 400, seed=5)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    for alpha in (0.3, 1.2):
        rates = oracle.api("port").gamma_rates(alpha, 4)
        for eng in (g, o):
            eng.set_model(0, [0.2, 0.3, 0.1, 0.4], [0.5, 2.0, 1.5, 0.7, 4.0, 1.0], rates, [0.25] * 4)
        assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    g.close()


def test_error_behaviour_mirrors_reference():
    from netrax_b200._capi import LikelihoodError
    net = random_network(6, 1, seed=2)
    m, w = simulate_alignment(net, 40, seed=2)
    bad = m.copy(); bad[0, 0] = 0   # illegal state code (pll_set_tip_states fails with "Illegal state code in tip")
    with pytest.raises(LikelihoodError, match="Illegal state code"):
        _gpu(net, [Partition(4, 4, bad, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05)])
    g = _gpu(net, [Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)])
    g.computeLoglikelihood(0, 1)
    tip_edge = 0   # edge into tip 0: source inner, target tip -> fine; a tip-tip pair cannot occur in a network
    g.brlen_prepare(tip_edge); g.computePartitionSumtables(tip_edge); g.brlen_finish(tip_edge)
    g.close()


# ---------------------------------------------------------------------------------------------- protein (20 states)
def _protein_case(n, r, pat, seed, random_cells=False, net=None):
    from netrax_b200.synth import lg_model
    rates, freqs = lg_model()
    net = net or random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed, states=20, rates=rates, freqs=freqs, random_cells=random_cells)
    return net, Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w)


@pytest.mark.parametrize("cfg", [(12, 2, 300, 1), (30, 2, 1001, 2), (9, 0, 77, 3)])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_protein_lg_g4_matches_oracle(cfg, variant):
    """BASELINE config 4 shape (LG+G4): the DMMA (FP64 tensor core) CLV kernel against the reference's AVX2 kernels.
    20-state sums are ordered differently (tensor-core k-steps vs AVX2 FMA lanes), so CLVs agree to rounding, not
    bit-for-bit; scaler counts must still be identical and lnL within 1e-10 (north_star)."""
    net, part = _protein_case(*cfg)
    g, o = _gpu(net, [part], variant=variant), _oracle(net, [part], variant=variant)
    _inject_eigen(g, o)
    lo, lg = o.computeLoglikelihood(0, 1), g.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    for v in range(net.num_tips, net.num_nodes):
        assert g.num_trees(v) == o.num_trees(v)
        for t in range(g.num_trees(v)):
            assert np.array_equal(g.read_scaler(v, t), o.read_scaler(v, t)), (v, t)
            np.testing.assert_allclose(g.read_clv(v, t), o.read_clv(v, t), rtol=1e-11, atol=1e-300)
    root = net.root
    for t in range(g.num_trees(root)):
        assert g.tree_info(root, t)[1] == pytest.approx(o.tree_info(root, t)[1], rel=LNL_RTOL)
    # branch-length flow on a reticulation edge and a tip edge
    edges = [0] + ([int(net.ret_first_edge[0])] if net.num_reticulations else [])
    for e in edges:
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        assert g.computePartitionSumtables(e) == o.computePartitionSumtables(e)
        dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
        assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
        assert dg[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-7)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


def test_protein_scaler_stress_bitexact_scalers():
    net, part = _protein_case(0, 0, 500, 13, random_cells=True, net=caterpillar_network(150))
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    mx = 0
    for t in range(g.num_trees(net.root)):
        sg, so = g.read_scaler(net.root, t), o.read_scaler(net.root, t)
        assert np.array_equal(sg, so)
        mx = max(mx, int(sg.max()))
    assert mx >= 2
    g.close()


def test_protein_dmma_equals_scalar_kernel(monkeypatch):
    """The tensor-core kernel against this repo's own scalar 20-state kernel (NRX_AA=generic): same scalers, CLVs to rounding."""
    net, part = _protein_case(15, 2, 513, 4)
    g = _gpu(net, [part])
    l1 = g.computeLoglikelihood(0, 1)
    monkeypatch.setenv("NRX_AA", "generic")
    s = _gpu(net, [part])
    l2 = s.computeLoglikelihood(0, 1)
    assert l1 == pytest.approx(l2, rel=1e-12)
    for v in range(net.num_tips, net.num_nodes):
        for t in range(g.num_trees(v)):
            assert np.array_equal(g.read_scaler(v, t), s.read_scaler(v, t))
            np.testing.assert_allclose(g.read_clv(v, t), s.read_clv(v, t), rtol=1e-12, atol=1e-300)
    g.close(); s.close()


def test_protein_brlen_flow_every_edge_tensor_core_kernels():
    """K4 (edge lnL) and K5 (sumtables) of 20-state partitions run on the FP64 tensor cores (k_aa20_dmma<AA_EDGE/AA_SUM>):
    on EVERY edge (inner-inner and tip-inner pairs, reticulation edges) the edge-rooted lnL, every sumtable entry and the
    derivatives at several proposal lengths must match the reference's AVX2 kernels under the oracle driver."""
    net, part = _protein_case(14, 2, 301, 21)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    l0 = g.computeLoglikelihood(0, 1)
    assert l0 == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    for e in range(net.num_edges):
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        lb = g.computeLoglikelihoodBrlenOpt(e)
        assert lb == pytest.approx(l0, rel=1e-11), e
        assert lb == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        ng, no = g.computePartitionSumtables(e), o.computePartitionSumtables(e)
        assert ng == no
        for i in range(ng):
            sg, so = g.read_sumtable(0, i)[0], o.read_sumtable(0, i)[0]
            np.testing.assert_allclose(sg, so, rtol=1e-9, atol=1e-13 * np.abs(so).max())  # eigenvector sums cancel
        if ng:
            t0 = float(net.edge_length[e])
            for t in (t0, 0.03, 0.9):
                for eng in (g, o):
                    eng.brlen_set_length(e, t)
                dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
                np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)
                assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
                assert dg[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-7)
                assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
            for eng in (g, o):
                eng.brlen_set_length(e, t0)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


def test_protein_dmma_k4_k5_equal_scalar_kernels(monkeypatch):
    """Tensor-core K4/K5 against this repo's own scalar 20-state kernels (NRX_AA=generic) incl. a ragged last tile."""
    net, part = _protein_case(12, 1, 203, 9)
    g = _gpu(net, [part])
    monkeypatch.setenv("NRX_AA", "generic")
    s = _gpu(net, [part])
    g.computeLoglikelihood(0, 1); s.computeLoglikelihood(0, 1)
    for e in (0, int(net.ret_first_edge[0]), net.num_edges - 1):
        assert g.brlen_prepare(e) == pytest.approx(s.brlen_prepare(e), rel=1e-12)
        assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(s.computeLoglikelihoodBrlenOpt(e), rel=1e-12)
        n = g.computePartitionSumtables(e)
        assert n == s.computePartitionSumtables(e)
        for i in range(n):
            a, b = g.read_sumtable(0, i)[0], s.read_sumtable(0, i)[0]
            np.testing.assert_allclose(a, b, rtol=1e-9, atol=1e-13 * np.abs(b).max())
        g.brlen_finish(e); s.brlen_finish(e)
    g.close(); s.close()


def test_mixed_dna_and_protein_partitions():
    """Two shape classes in one engine (4-state pipelined kernels + 20-state tensor-core kernels share the partial-sum
    layout of every reduction): full lnL and one branch's flow against the oracle."""
    net = random_network(12, 2, seed=31)
    md, wd = simulate_alignment(net, 410, seed=31)
    dna = Partition(4, 4, md, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=wd)
    _, aa = _protein_case(12, 2, 150, 31, net=net)
    g, o = _gpu(net, [dna, aa]), _oracle(net, [dna, aa])
    _inject_eigen(g, o)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    np.testing.assert_allclose(g.partition_loglh(), o.partition_loglh(), rtol=LNL_RTOL)
    for e in (1, int(net.ret_first_edge[1])):
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        assert g.computePartitionSumtables(e) == o.computePartitionSumtables(e)
        dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
        assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
        assert dg[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-7)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


@pytest.mark.parametrize("cats", [1, 2, 3, 8, 16])
def test_other_category_counts(cats):
    """GTR + Gamma with 1, 2, 8 or 16 rate categories runs the pipelined K2 kernel templated on the category count (round 2; the
    reference's AVX kernels loop over rate_cats at full speed, LIBPLL/core_partials_avx.c:402-565): CLVs and scalers bit-identical
    to libpll as for 4 categories, incl. plan replay with the fused K3.  3 categories: the generic thread-per-pattern kernels.
    K3-K6: thread per (pattern, category) for power-of-two counts."""
    from oracle import oracle
    net = random_network(11, 2, seed=40 + cats)
    m, w = simulate_alignment(net, 1333, seed=40 + cats)
    rates = oracle.api("port").gamma_rates(0.7, cats) if cats > 1 else np.ones(1)
    part = Partition(4, cats, m, DNA_FREQS, GTR_RATES, rates, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    lo = o.computeLoglikelihood(0, 1)
    lg = g.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    same_p = all(np.array_equal(g.get_pmatrix(e), o.get_pmatrix(e)) for e in range(net.num_edges + 1))
    _compare_all_clvs(g, o, exact=same_p and cats != 3)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)     # plan replay (graph + fused K3 for 1 / 2 / 8 / 16)
    _compare_all_clvs(g, o, exact=same_p and cats != 3)
    for e in (0, int(net.ret_first_edge[0]), net.num_edges - 1):
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        assert g.computePartitionSumtables(e) == o.computePartitionSumtables(e)
        dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
        np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


def test_category_counts_scaler_stress():
    """the per-pattern scaling vote over a 1-, 2-, 8- or 16-lane group: deep caterpillar, scaler counts bit-exact"""
    from oracle import oracle
    net = caterpillar_network(300)
    m, w = simulate_alignment(net, 300, seed=77, random_cells=True)
    for cats in (1, 2, 8, 16):
        rates = oracle.api("port").gamma_rates(0.5, cats) if cats > 1 else np.ones(1)
        part = Partition(4, cats, m, DNA_FREQS, GTR_RATES, rates, pattern_weights=w)
        g, o = _gpu(net, [part]), _oracle(net, [part])
        _inject_eigen(g, o)
        assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
        mx = 0
        for t in range(g.num_trees(net.root)):
            sg, so = g.read_scaler(net.root, t), o.read_scaler(net.root, t)
            assert np.array_equal(sg, so)
            mx = max(mx, int(so.max()))
        assert mx >= 1, cats
        g.close()


def test_batched_scoring_equals_sequential():
    """computeLoglikelihoodBatch over candidate networks (different topologies / branch lengths over one alignment, one
    engine and stream each): every result equals the single evaluation and the oracle's; full (plan replay, fused K3),
    incremental after branch-length changes, and cached re-evaluation."""
    from netrax_b200.engine import compute_loglikelihood_batch
    base = random_network(14, 2, seed=77)
    m, w = simulate_alignment(base, 700, seed=77)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    nets = [base] + [random_network(14, r, seed=80 + r) for r in (0, 1, 3)]
    gs = [_gpu(n, [part]) for n in nets]
    os_ = [_oracle(n, [part]) for n in nets]
    for g, o in zip(gs, os_):
        _inject_eigen(g, o)
    want = np.array([o.computeLoglikelihood(0, 1) for o in os_])
    got = compute_loglikelihood_batch(gs, 0, 1)          # first evaluation: records the plans
    np.testing.assert_allclose(got, want, rtol=LNL_RTOL)
    got2 = compute_loglikelihood_batch(gs, 0, 1)         # plan replay + fused K3, all four in flight
    # batched = throughput geometry (level-by-level plan + k_term_lnl_sum), single = one-launch tile walk: identical CLVs, the per-tree
    # lnL summed over patterns in a different order -> equal to rounding (REPLAY_RTOL), not bit for bit
    np.testing.assert_allclose(got2, np.array([g.computeLoglikelihood(0, 1) for g in gs]), rtol=REPLAY_RTOL)
    np.testing.assert_allclose(got2, want, rtol=LNL_RTOL)
    for k, (g, o) in enumerate(zip(gs, os_)):            # candidates diverge: one branch each, incremental re-evaluation
        for eng in (g, o):
            eng.set_branch_length(k + 1, 0.05 * (k + 1))
    want3 = np.array([o.computeLoglikelihood(1, 1) for o in os_])
    np.testing.assert_allclose(compute_loglikelihood_batch(gs, 1, 1), want3, rtol=LNL_RTOL)
    np.testing.assert_allclose(compute_loglikelihood_batch(gs, 1, 1), want3, rtol=LNL_RTOL)   # cached
    for g in gs:
        g.close()


def _fuzz_cases():
    rng = np.random.default_rng(2026)
    sizes = [1, 2, 63, 64, 65, 127, 128, 129, 191, 255, 256, 257, 383, 500]   # around the 64 / 128-pattern tile edges
    cases = []
    for k, pat in enumerate(sizes):
        cases.append((int(rng.integers(4, 28)), int(rng.integers(0, 5)), pat, 100 + k, bool(k % 3 == 0), AVERAGE if k % 2 else BEST))
    return cases


@pytest.mark.parametrize("n,r,pat,seed,random_cells,variant", _fuzz_cases())
def test_randomised_networks_bitexact(n, r, pat, seed, random_cells, variant):
    """Seeded random networks / alignment sizes around the kernel tile edges (ragged last tiles, a single pattern),
    uniform-random cells every third case (the worst case for scaling), random reticulation probabilities: CLVs and
    scalers bit-identical to the reference's libpll, lnL / derivatives within tolerance, re-rooting on two random edges."""
    rng = np.random.default_rng(seed)
    r = min(r, max(0, n - 3))
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed, random_cells=random_cells)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part], variant=variant), _oracle(net, [part], variant=variant)
    _inject_eigen(g, o)
    for i in range(net.num_reticulations):
        pr = float(rng.uniform(0.05, 0.95))
        g.set_reticulation_prob(i, pr); o.set_reticulation_prob(i, pr)
    lo, lg = o.computeLoglikelihood(0, 1), g.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    same_p = all(np.array_equal(g.get_pmatrix(e), o.get_pmatrix(e)) for e in range(net.num_edges + 1))
    _compare_all_clvs(g, o, exact=same_p)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)     # plan replay (tile walk / graph + fused K3)
    for e in sorted({int(x) for x in rng.integers(0, net.num_edges, 2)}):
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        ng = g.computePartitionSumtables(e)
        assert ng == o.computePartitionSumtables(e)
        if ng:
            dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
            np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


# ---------------------------------------------------------------------------------------------- pseudo-likelihood
@pytest.mark.parametrize("cfg", [(10, 2, 257, 2), (14, 3, 128, 3), (9, 4, 65, 4), (12, 0, 300, 5), (25, 5, 1, 6)])
def test_pseudo_loglikelihood_bitexact(cfg):
    """computePseudoLoglikelihood (src/likelihood/PseudoLoglikelihood.cpp:57-226) as ONE fused kernel per node batch
    (three libpll updates + merge_clvs in registers) against the reference's libpll under the restated driver: every node's
    merged CLV and scaler bit-identical (DNA), incremental == full, variant SARAH_PSEUDO dispatches to it."""
    from netrax_b200._capi import SARAH_PSEUDO
    n, r, pat, seed = cfg
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part], variant=SARAH_PSEUDO), _oracle(net, [part], variant=SARAH_PSEUDO)
    _inject_eigen(g, o)
    lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)     # dispatch (LikelihoodComputation.cpp:23-27)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    same_p = all(np.array_equal(g.get_pmatrix(e), o.get_pmatrix(e)) for e in range(net.num_edges + 1))
    for v in range(net.num_tips, net.num_nodes):
        assert np.array_equal(g.read_pseudo_scaler(v), o.read_pseudo_scaler(v)), v
        a, b = g.read_pseudo_clv(v), o.read_pseudo_clv(v)
        if same_p:
            assert np.array_equal(a, b), (v, np.abs(a - b).max())
        else:
            np.testing.assert_allclose(a, b, rtol=1e-12, atol=0)
    if r == 0:
        e2 = _gpu(net, [part])
        assert lg == pytest.approx(e2.computeLoglikelihood(0, 1), rel=1e-12)   # a tree: pseudo == exact
        e2.close()
    for eng in (g, o):
        eng.set_branch_length(1, 0.37)
        if r:
            eng.set_reticulation_prob(0, 0.8)
    inc_g, inc_o = g.computeLoglikelihood(1, 1), o.computeLoglikelihood(1, 1)
    assert inc_g == pytest.approx(inc_o, rel=LNL_RTOL)
    assert inc_g == pytest.approx(g.computePseudoLoglikelihood(0, 1), rel=1e-13)
    g.close()


def test_pseudo_loglikelihood_scaler_stress_and_protein():
    """Deep caterpillar with one reticulation (scaler counts >= 2 in the blended CLVs: the three updates scale separately and
    the last one's scaler wins, as in the reference) and a 20-state partition through the generic kernel."""
    from netrax_b200._capi import SARAH_PSEUDO
    net = caterpillar_network(400)
    m, w = simulate_alignment(net, 300, seed=11, random_cells=True)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part], variant=SARAH_PSEUDO), _oracle(net, [part], variant=SARAH_PSEUDO)
    _inject_eigen(g, o)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    mx = 0
    for v in range(net.num_tips, net.num_nodes):
        sg = g.read_pseudo_scaler(v)
        assert np.array_equal(sg, o.read_pseudo_scaler(v)), v
        mx = max(mx, int(sg.max()))
    assert mx >= 2
    g.close()
    net2, aa = _protein_case(10, 2, 97, 8)
    g, o = _gpu(net2, [aa], variant=SARAH_PSEUDO), _oracle(net2, [aa], variant=SARAH_PSEUDO)
    _inject_eigen(g, o)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    for v in range(net2.num_tips, net2.num_nodes):
        assert np.array_equal(g.read_pseudo_scaler(v), o.read_pseudo_scaler(v))
        np.testing.assert_allclose(g.read_pseudo_clv(v), o.read_pseudo_clv(v), rtol=1e-11, atol=1e-300)
    g.close()


def test_empty_partition_slice_is_skipped():
    """A site shard may own NO pattern of a partition (reference: partitions[p] == NULL, 'skip remote partitions',
    LH/ImprovedLoglikelihood.cpp:128-131): the engine must accept patterns = 0 and contribute exactly 0 to that partition."""
    net = random_network(10, 2, seed=8)
    m, w = simulate_alignment(net, 200, seed=8)
    full = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    empty = full.slice(0, 0)
    g = _gpu(net, [full, empty], variant=BEST, linkage=UNLINKED, partition_brlens=[net.edge_length, net.edge_length])
    o = _oracle(net, [full])
    lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
    pl = g.partition_loglh()
    assert pl[0] == pytest.approx(o.partition_loglh()[0], rel=LNL_RTOL)
    # the empty slice adds only the tree log-prior of the best tree (0 site terms)
    best_prior = max(g.tree_info(net.root, t)[0] for t in range(g.num_trees(net.root)))
    assert pl[1] == pytest.approx(best_prior, abs=1e-12)
    # (a partition that is empty on EVERY rank is rejected by the edge-rooted evaluation exactly as in the reference:
    #  "bad partition logl", LH/VirtualRerooting.cpp:339-341; locally-empty slices under sharding are covered by
    #  scripts/check_multi_gpu.py, where the all-reduced sum is non-zero)
    from netrax_b200._capi import LikelihoodError
    e = int(net.ret_first_edge[0])
    g.brlen_prepare(e)
    with pytest.raises(LikelihoodError, match="bad partition logl"):
        g.computeLoglikelihoodBrlenOpt(e)
    g.close()


# ---------------------------------------------------------------------------------------------- BASELINE full sizes
def test_headline_topology_matches_oracle():
    """BASELINE config 5's network (100 taxa, 8 reticulations: 776 displayed-tree CLVs, 192 root trees) at a pattern count
    the oracle finishes in seconds: lnL, every per-tree lnL, scalers and CLVs of the root against libpll."""
    import bench
    cfg = dict(bench.CONFIGS[5])
    net, parts, _ = bench.make_inputs(cfg, 700)
    g, o = _gpu(net, parts), _oracle(net, parts)
    _inject_eigen(g, o)
    lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    assert sum(g.num_trees(v) for v in range(net.num_tips, net.num_nodes)) == 776
    root = net.root
    assert g.num_trees(root) == o.num_trees(root) == 192
    for t in range(g.num_trees(root)):
        assert g.tree_config(root, t) == o.tree_config(root, t)
        assert g.tree_info(root, t)[1] == pytest.approx(o.tree_info(root, t)[1], rel=LNL_RTOL)
        assert np.array_equal(g.read_scaler(root, t), o.read_scaler(root, t))
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)   # plan replay (tile walk, or CUDA graph + PDL + fused K3)
    e = int(net.ret_first_edge[3])
    assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
    assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
    assert g.computePartitionSumtables(e) == o.computePartitionSumtables(e)
    dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
    assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
    assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


def test_full_size_config5_size_independent_properties():
    """BASELINE config 5 at its FULL size (1 M patterns, 102 GB of CLVs on one B200) through properties that need no
    oracle run: replay determinism, per-site terms sum to the tree lnL, additivity over pattern slices (what site sharding
    relies on), re-rooting preserves the network lnL, incremental == full.  NRX_TEST_FULL_PATTERNS overrides the size."""
    import os

    import bench
    import torch
    patterns = int(os.environ.get("NRX_TEST_FULL_PATTERNS", "1000000"))
    free, _total = torch.cuda.mem_get_info(0)
    need = 776 * patterns * 132 * 1.08
    if free < need:
        pytest.skip(f"needs {need / 1e9:.0f} GB of free device memory, have {free / 1e9:.0f} GB")
    cfg = dict(bench.CONFIGS[5])
    net, parts, _ = bench.make_inputs(cfg, patterns)
    g = _gpu(net, parts)
    l0 = g.computeLoglikelihood(0, 1)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(l0, rel=REPLAY_RTOL)   # replay: same CLVs, per-tree sums in the replay kernel's order
    root = net.root
    trees = [g.tree_info(root, t)[1][0] for t in range(g.num_trees(root))]
    for t in (0, g.num_trees(root) - 1):
        ps = g.persite_lnl(t)[0]
        assert float(ps[:patterns].sum()) == pytest.approx(trees[t], rel=1e-11)
    e = int(net.ret_first_edge[0])
    g.brlen_prepare(e)
    assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(l0, rel=1e-11)   # re-rooting preserves lnL (BrlenOptTest.cpp:297-367)
    assert g.brlen_finish(e) == pytest.approx(l0, rel=1e-12)
    g.set_branch_length(5, 0.123)
    inc = g.computeLoglikelihood(1, 1)
    assert inc == pytest.approx(g.computeLoglikelihood(0, 1), rel=1e-13) and inc != l0
    g.close()
    del g
    # additivity over pattern slices: per-tree lnLs of the two halves add up to the whole (the cross-rank SUM of §8e)
    half = patterns // 2
    acc = np.zeros(len(trees))
    for lo, hi in ((0, half), (half, patterns)):
        s = _gpu(net, [parts[0].slice(lo, hi)])
        s.computeLoglikelihood(0, 1)
        acc += np.array([s.tree_info(root, t)[1][0] for t in range(s.num_trees(root))])
        s.close()
        del s
    np.testing.assert_allclose(acc, np.array(trees), rtol=1e-11)


@pytest.mark.parametrize("config", [1, 2, 3, 4])
def test_baseline_configs_full_size_match_oracle(config):
    """BASELINE configs 1-4 at their FULL sizes, directly against the reference's libpll under the restated driver (the
    oracle needs seconds for these): network lnL, per-partition lnL, every root tree's lnL, root scalers; config 2
    additionally the branch-length derivative on one reticulation edge (the sweep of bench_configs.py edge by edge)."""
    import bench
    cfg = dict(bench.CONFIGS[config])
    net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    kw = dict(variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    g, o = _gpu(net, parts, **kw), _oracle(net, parts, **kw)
    _inject_eigen(g, o)
    lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    np.testing.assert_allclose(g.partition_loglh(), o.partition_loglh(), rtol=LNL_RTOL)
    root = net.root
    for t in range(g.num_trees(root)):
        np.testing.assert_allclose(g.tree_info(root, t)[1], o.tree_info(root, t)[1], rtol=LNL_RTOL)
        assert np.array_equal(g.read_scaler(root, t), o.read_scaler(root, t))
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)
    if config == 2:
        e = int(net.ret_first_edge[1])
        assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
        assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
        assert g.computePartitionSumtables(e) == o.computePartitionSumtables(e)
        dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
        np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)
        assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
        assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


# ---------------------------------------------------------------------------------------------- libpll's own golden vectors
def _golden_five_taxon(G, t, tip_edge, ncats, rates):
    from netrax_b200.network_io import encode_dna, parse_extended_newick
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


def test_libpll_golden_alpha_cats_on_gpu():
    """The PRODUCT against libpll's regression output test/out/alpha-cats.out (tests/golden/libpll_alpha_cats_golden.json):
    the edge lnL of the 5-taxon data set for 9 alphas x {1, 2, 4, 8, 16} categories x {MEAN, MEDIAN} with the product's own
    Gamma rates, eigendecomposition, P-matrices and kernels (4 categories: the pipelined 4x4 kernels; others: generic)."""
    from helpers import load_golden
    from netrax_b200.engine import load
    GA = load_golden("libpll_alpha_cats_golden.json")
    api = load()
    for b in GA["blocks"]:
        rates = api.gamma_rates(b["alpha"], b["ncats"], {"MEAN": 0, "MEDIAN": 1}[b["mode"]]) if b["ncats"] > 1 else np.ones(1)
        net, part = _golden_five_taxon(GA, GA["branch_lengths"][0], False, b["ncats"], rates)
        g = _gpu(net, [part])
        assert abs(g.computeLoglikelihood(0, 1) - b["logl"]) < 2e-6, (b["alpha"], b["ncats"], b["mode"])
        g.close()


@pytest.mark.parametrize("tip_edge", [False, True])
def test_libpll_golden_derivatives_on_gpu(tip_edge):
    """The PRODUCT against libpll's test/out/derivatives.out (tests/golden/libpll_derivatives_golden.json): edge lnL (6
    decimals) and first / second derivative (5 significant digits) on an inner and on a tip edge, for every (alpha,
    categories) block at three branch lengths — K1, K2, K4, K5, K6 and the host mixing pinned to the reference's own vectors."""
    from helpers import load_golden
    from netrax_b200.engine import load
    G = load_golden("libpll_derivatives_golden.json")
    api = load()
    for block in G["blocks"]:
        rates = api.gamma_rates(block["alpha"], block["ncats"]) if block["ncats"] > 1 else np.ones(1)
        for t, f, d1, d2 in block["tip" if tip_edge else "inner"][:6:2]:
            net, part = _golden_five_taxon(G, t, tip_edge, block["ncats"], rates)
            g = _gpu(net, [part])
            lnl = g.computeLoglikelihood(0, 1)
            assert abs(lnl - f) < 2e-6, (block["alpha"], block["ncats"], t, lnl, f)
            edge = [e for e in range(net.num_edges) if net.edge_source[e] == net.root][0]
            g.brlen_prepare(edge)
            assert abs(g.computeLoglikelihoodBrlenOpt(edge) - lnl) < 1e-9
            assert g.computePartitionSumtables(edge) == 1
            g1, g2, *_ = g.computeLoglikelihoodDerivatives(edge)
            assert g1 == pytest.approx(d1, rel=2e-4, abs=1e-9), (t, g1, d1)
            assert g2 == pytest.approx(d2, rel=2e-4, abs=1e-9), (t, g2, d2)
            g.close()


def test_libpll_golden_protein_models_on_gpu():
    """The PRODUCT's 20-state path (own Jacobi eigendecomposition, K1 + tip tables, tensor-core K2, K3) against libpll's
    test/out/protein-models.out for all 20 empirical amino-acid models; LG additionally through the tensor-core edge-lnL
    (K4) and sumtable / derivative kernels against the oracle."""
    from helpers import load_golden, protein_golden_case
    GP = load_golden("libpll_protein_models_golden.json")
    for model, m in GP["models"].items():
        net, part = protein_golden_case(GP, model)
        g = _gpu(net, [part])
        lnl = g.computeLoglikelihood(0, 1)
        assert abs(lnl - m["logl"]) < 2e-6, (model, lnl, m["logl"])
        if model == "LG":
            o = _oracle(net, [part])
            o.computeLoglikelihood(0, 1)
            edge = [e for e in range(net.num_edges) if net.edge_source[e] == net.root][0]
            assert g.brlen_prepare(edge) == pytest.approx(o.brlen_prepare(edge), rel=1e-9)
            assert abs(g.computeLoglikelihoodBrlenOpt(edge) - m["logl"]) < 2e-6     # K4 against the golden value
            assert g.computePartitionSumtables(edge) == o.computePartitionSumtables(edge)
            dg, do = g.computeLoglikelihoodDerivatives(edge), o.computeLoglikelihoodDerivatives(edge)
            assert dg[0] == pytest.approx(do[0], rel=1e-7, abs=1e-7)
            assert dg[1] == pytest.approx(do[1], rel=1e-7, abs=1e-7)
        g.close()


@pytest.mark.parametrize("tip_edge", [False, True])
def test_libpll_golden_oddstates_on_gpu(tip_edge):
    """The PRODUCT's generic kernels (5 states in 8-double rows: states != states_padded nowhere else in the suite) against
    libpll's test/out/derivatives-oddstates.out: edge lnL, first and second derivative on an inner and a tip edge."""
    from test_oracle_libpll_golden import GO, oddstates_case
    from netrax_b200.engine import load
    api = load()
    for block in GO["blocks"]:
        rates = api.gamma_rates(block["alpha"], block["ncats"]) if block["ncats"] > 1 else np.ones(1)
        for t, f, d1, d2 in block["tip" if tip_edge else "inner"]:
            if t > 10:
                continue
            net, part = oddstates_case(t, tip_edge, block["ncats"], rates)
            g = _gpu(net, [part])
            lnl = g.computeLoglikelihood(0, 1)
            assert abs(lnl - f) < 2e-6, (block["alpha"], block["ncats"], t, lnl, f)
            edge = [e for e in range(net.num_edges) if net.edge_source[e] == net.root][0]
            g.brlen_prepare(edge)
            assert abs(g.computeLoglikelihoodBrlenOpt(edge) - lnl) < 1e-9
            assert g.computePartitionSumtables(edge) == 1
            g1, g2, *_ = g.computeLoglikelihoodDerivatives(edge)
            assert g1 == pytest.approx(d1, rel=2e-4, abs=1e-9), (t, g1, d1)
            assert g2 == pytest.approx(d2, rel=2e-4, abs=1e-9), (t, g2, d2)
            g.close()


def test_engines_driven_from_concurrent_host_threads():
    """One engine per host thread on the same GPU (the reference runs one AnnotatedNetwork per worker thread,
    RAXML/ParallelContext.cpp): handles are independent — own stream, own staging ring, thread-local error state — so
    concurrent evaluations, re-rootings and optimisations give exactly the sequential results."""
    import threading
    nets = [random_network(12 + k, 2, seed=300 + k) for k in range(4)]
    parts = []
    for k, n in enumerate(nets):
        m, w = simulate_alignment(n, 900 + 37 * k, seed=300 + k)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))

    def work(eng, net):
        out = [eng.computeLoglikelihood(0, 1)]
        for _ in range(20):
            out.append(eng.computeLoglikelihood(0, 1))
        e = int(net.ret_first_edge[0])
        eng.brlen_prepare(e)
        out.append(eng.computeLoglikelihoodBrlenOpt(e))
        eng.computePartitionSumtables(e)
        out.extend(eng.computeLoglikelihoodDerivatives(e)[:2])
        out.append(eng.brlen_finish(e))
        out.append(eng.optimize_branches())
        return out

    want = []
    for n, p in zip(nets, parts):
        g = _gpu(n, [p])
        want.append(work(g, n))
        g.close()
    engines = [_gpu(n, [p]) for n, p in zip(nets, parts)]
    got, errors = [None] * 4, []

    def run(i):
        try:
            got[i] = work(engines[i], nets[i])
        except Exception as ex:  # noqa: BLE001
            errors.append(ex)

    threads = [threading.Thread(target=run, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
    for a, b in zip(got, want):
        assert a == b
    for g in engines:
        g.close()


@pytest.mark.parametrize("which", ["dna", "oddstates"])
@pytest.mark.parametrize("tip_edge", [False, True])
def test_libpll_golden_pinv_on_gpu(which, tip_edge):
    """+I (proportion of invariant sites) in the PRODUCT — K1 rate scaling, the invariant-site terms of K3 / K4 / K6, the
    invariant-pattern detection from the tips — against libpll's golden blocks with pinv in {0.3, 0.6, 0.9}: 4 states with
    1 / 2 / 4 categories (the 4x4 kernels and the generic ones) and 5 states (generic kernels, padded rows)."""
    from test_oracle_libpll_golden import GI, GOI, check_pinv_golden
    from netrax_b200.engine import load
    check_pinv_golden(lambda net, part: _gpu(net, [part]), GI if which == "dna" else GOI, tip_edge, load().gamma_rates)


def test_pinv_on_networks_matches_reference_oracle():
    """+I on networks: DNA (4x4 kernels; fused K3 switched off), protein (tensor-core edge lnL) and a deep caterpillar whose
    scaled sites exercise libpll's 'undo the scaling on the non-invariant term only' branch, against real libpll."""
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    cases = []
    net = random_network(14, 3, seed=61)
    m, w = simulate_alignment(net, 500, seed=61)
    cases.append((net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.25))
    cases.append(_protein_case(10, 2, 150, 62) + (0.4,))
    cat = caterpillar_network(300)
    m, w = simulate_alignment(cat, 200, seed=63, gap_frac=0.0)
    m[:, :60] = m[0, :60]          # 60 invariant columns on a tree deep enough to scale
    cases.append((cat, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w), 0.3))
    # with Gamma categories the slowest category keeps invariant columns above the scaling threshold; ONE category on saturated
    # branches scales them too (scaler 2 on every column): the case that really takes the "non-invariant term only" branch
    sat = caterpillar_network(300, brlen=4.0)
    m1, w1 = simulate_alignment(sat, 200, seed=63, gap_frac=0.0)
    m1[:, :60] = m1[0, :60]
    cases.append((sat, Partition(4, 1, m1, DNA_FREQS, GTR_RATES, np.ones(1), pattern_weights=w1), 0.3))
    for net, part, pinv in cases:
        g, o = _gpu(net, [part]), oracle.make_engine("ref", net, [part])
        _inject_eigen(g, o)
        g.set_pinv(0, pinv); o.set_pinv(0, pinv)
        lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
        assert lg == pytest.approx(lo, rel=LNL_RTOL), (part.states, pinv)
        assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)
        for e in (0, net.num_edges - 1) + ((int(net.ret_first_edge[0]),) if net.num_reticulations else ()):
            assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
            assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL), (part.states, e)
            assert g.computePartitionSumtables(e) == o.computePartitionSumtables(e)
            dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
            np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)
            assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
        g.set_pinv(0, 0.0); o.set_pinv(0, 0.0)
        assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
        g.close()


# ---- libpll's pmatrix.out / hky.out directly against the product (K1 incl. the host eigendecomposition; K2 CLVs; K4 lnL) ----
@pytest.mark.gpu
@pytest.mark.parametrize("datatype", ["DNA", "PROT", "ODD"])
def test_libpll_golden_pmatrix_on_gpu(datatype):
    """K1 (k_pmatrix) + the host's Jacobi eigendecomposition: every entry of the 135 golden P-matrices per datatype (4 / 20 / 5
    states; equal / skewed / extreme frequencies and exchangeabilities; t = 1e-6 .. 100; rates 1e-31 .. 100) to 9 decimals."""
    import os
    from helpers import GOLDEN, check_pmatrix_golden
    G = np.load(os.path.join(GOLDEN, "libpll_pmatrix_golden.npz"))
    worst = check_pmatrix_golden(lambda net, part: _gpu(net, [part]), G, datatype)
    assert worst < 6e-10


@pytest.mark.gpu
def test_libpll_golden_hky_on_gpu():
    """hky.out: P-matrices, the three inner CLVs (K2 tip-tip, tip-inner) and the edge lnL for 10 ti/tv ratios."""
    import os
    from helpers import GOLDEN, check_hky_golden
    check_hky_golden(lambda net, part: _gpu(net, [part]), np.load(os.path.join(GOLDEN, "libpll_hky_golden.npz")))


# ---- one rate matrix per rate category (LG4M / LG4X: raxml-ng ratecat_submodels -> libpll params_indices) ------------------
def _mixture_case(states, seed):
    from helpers import mixture_models
    if states == 20:
        net, part = _protein_case(11, 2, 260, seed)
    else:
        net = random_network(13, 3, seed=seed)
        m, w = simulate_alignment(net, 400, seed=seed)
        part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    part.rates = np.array([0.15, 0.6, 1.2, 2.9])            # LG4X: free rates and weights
    part.rate_weights = np.array([0.35, 0.3, 0.25, 0.1])
    freqs, subst = mixture_models(states, 4, seed)
    return net, part, [2, 0, 3, 1], freqs, subst


@pytest.mark.gpu
@pytest.mark.parametrize("states", [4, 20])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_submodels_match_reference_oracle(states, variant):
    """Per-category rate matrices against the reference's libpll called with the same params_indices: network lnL, per-tree
    lnLs, scalers (bit-exact), then on every edge the edge-rooted lnL, the sumtables and the derivatives; with and without +I.
    K2 keeps its fast kernels (pipelined DNA / tensor-core protein), K1 and K3-K6 run the per-category generic kernels."""
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    net, part, cat_model, freqs, subst = _mixture_case(states, 71 + states)
    g, o = _gpu(net, [part], variant=variant), oracle.make_engine("ref", net, [part], variant=variant)
    l_single = g.computeLoglikelihood(0, 1)
    assert l_single == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    for eng in (g, o):
        eng.set_submodels(0, cat_model, freqs, subst)
    for pinv in (0.0, 0.2):
        for eng in (g, o):
            eng.set_pinv(0, pinv)
        lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
        assert lg == pytest.approx(lo, rel=LNL_RTOL), (states, pinv)
        assert abs(lg - l_single) > 1e-3
        root = net.root
        assert g.num_trees(root) == o.num_trees(root)
        for t in range(g.num_trees(root)):
            np.testing.assert_allclose(g.tree_info(root, t)[1], o.tree_info(root, t)[1], rtol=LNL_RTOL)
            assert np.array_equal(g.read_scaler(root, t), o.read_scaler(root, t))
        for e in range(0, net.num_edges, 3 if pinv else 1):
            assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
            lb = g.computeLoglikelihoodBrlenOpt(e)
            assert lb == pytest.approx(lg, rel=1e-11), e
            assert lb == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
            ng = g.computePartitionSumtables(e)
            assert ng == o.computePartitionSumtables(e)
            for i in range(ng):
                # sumtable rows live in the eigenbasis: the host's Jacobi solver and libpll's QL order the eigenvectors of
                # matrices 1.. differently (only matrix 0 can be injected), signs cancel in the product -> compare sorted rows
                sp = (states + 3) & ~3
                sg, so = (np.sort(x.read_sumtable(0, i)[0].reshape(-1, sp), axis=1) for x in (g, o))
                np.testing.assert_allclose(sg, so, rtol=1e-8, atol=1e-12 * np.abs(so).max())
            if ng:
                t0 = float(net.edge_length[e])
                for t in (t0, 0.04, 0.8):
                    for eng in (g, o):
                        eng.brlen_set_length(e, t)
                    dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
                    np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-9)
                    assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7)
                    assert dg[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-7)
                for eng in (g, o):
                    eng.brlen_set_length(e, t0)
            assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    # branch-length optimisation runs on top of it; and the mixture can be taken away again
    e = int(net.ret_first_edge[0])
    assert g.optimize_branch(e) == pytest.approx(o.optimize_branch(e), rel=1e-9)
    for eng in (g, o):
        eng.set_pinv(0, 0.0)
        eng.set_submodels(0, [0, 0, 0, 0], freqs[:1], subst[:1])
        eng.set_model(0, part.freqs, part.subst, part.rates, part.rate_weights)
        eng.set_branch_length(e, float(net.edge_length[e]))
    assert g.computeLoglikelihood(0, 1) == pytest.approx(l_single, rel=1e-13)
    g.close(); o.close()


@pytest.mark.gpu
@pytest.mark.parametrize("states", [4, 20])
def test_submodels_equal_the_sum_over_single_matrix_categories(states):
    """Independent of any oracle: the product's mixture lnL on a tree equals the value assembled from the product's own
    SINGLE-matrix, single-category, single-pattern evaluations (lnL = sum_sites log sum_c w_c L_c(site))."""
    from helpers import mixture_lnl_by_categories, mixture_models
    from netrax_b200.network_io import parse_extended_newick
    net = parse_extended_newick("(((T0:0.1,T1:0.2):0.05,T2:0.3):0.1,(T3:0.15,T4:0.25):0.2);")
    rng = np.random.default_rng(3)
    masks = (1 << rng.integers(0, states, size=(5, 7))).astype(np.uint32)
    masks[2, 4] = (1 << states) - 1
    freqs, subst = mixture_models(states, 4, seed=11)
    part = Partition(states, 4, masks, freqs[0], subst[0], [0.2, 0.7, 1.3, 2.4], rate_weights=[0.4, 0.3, 0.2, 0.1])
    g = _gpu(net, [part])
    cat_model = [2, 0, 3, 1]
    g.set_submodels(0, cat_model, freqs, subst)
    want = mixture_lnl_by_categories(lambda net, part: _gpu(net, [part]), net, part, cat_model, freqs, subst)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(want, rel=1e-12)
    g.close()


@pytest.mark.gpu
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_scaled_branch_length_linkage_on_gpu(variant):
    """Scaled linkage: P-matrices of partition p from brlen_scalers[p] x linked length — equal to the oracle, and bit-identical
    to the product's own unlinked evaluation at the scaled lengths (same kernels, same inputs)."""
    from netrax_b200._capi import SCALED
    from test_oracle_netrax import scaled_linkage_case
    net, parts, scalers = scaled_linkage_case()
    g, o = _gpu(net, parts, variant=variant, linkage=SCALED), _oracle(net, parts, variant=variant, linkage=SCALED)
    _inject_eigen(g, o)
    for p, s in enumerate(scalers):
        g.set_brlen_scaler(p, s); o.set_brlen_scaler(p, s)
    lg = g.computeLoglikelihood(1, 1)
    assert lg == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
    np.testing.assert_allclose(g.partition_loglh(), o.partition_loglh(), rtol=LNL_RTOL)
    un = _gpu(net, parts, variant=variant, linkage=UNLINKED, partition_brlens=[net.edge_length * s for s in scalers])
    _inject_eigen(un, o)
    assert un.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=REPLAY_RTOL)
    e = int(net.ret_first_edge[0])
    g.set_branch_length(e, 0.33); o.set_branch_length(e, 0.33)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
    g.brlen_prepare(e)
    g.computePartitionSumtables(e)
    with pytest.raises(Exception, match="scaled branch lengths"):
        g.computeLoglikelihoodDerivatives(e)
    g.close(); un.close(); o.close()


def test_tile_walk_evaluation_equals_level_by_level_launches(monkeypatch):
    """Round 2, SURVEY §8f f3 / VERDICT r1 item 6: a replayed full evaluation of a small alignment is ONE launch (k_walk_dna4: a
    block walks the whole plan for its tile of patterns, children from shared memory) after the P-matrix launch.  Same CLVs and
    scalers bit for bit as the level-by-level launches (NRX_WALK=0) and as libpll, per-tree lnLs equal to rounding; also with
    several partitions (unlinked branch lengths) and a partial last tile."""
    cases = []
    for n, r, pat, seed in ((20, 1, 1000, 2), (25, 3, 333, 3), (40, 5, 150, 4), (12, 0, 33, 5)):
        net = random_network(n, r, seed=seed)
        m, w = simulate_alignment(net, pat, seed=seed)
        cases.append((net, [Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)], None, LINKED))
    net = random_network(10, 2, seed=3)
    parts, brl = [], []
    rng = np.random.default_rng(0)
    for p in range(3):
        m, w = simulate_alignment(net, 100 + 37 * p, seed=30 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
    cases.append((net, parts, brl, UNLINKED))
    for net, parts, brl, linkage in cases:
        res = {}
        for mode in ("0", "1"):
            o = _oracle(net, parts, linkage=linkage, partition_brlens=brl)
            lo = o.computeLoglikelihood(0, 1)
            monkeypatch.setenv("NRX_WALK", mode)
            g = _gpu(net, parts, linkage=linkage, partition_brlens=brl)
            _inject_eigen(g, o)
            g.computeLoglikelihood(0, 1)            # records the plan
            n0 = g.launch_count()
            lg = g.computeLoglikelihood(0, 1)       # replay
            launches = g.launch_count() - n0
            assert lg == pytest.approx(lo, rel=LNL_RTOL)
            same_p = all(np.array_equal(g.get_pmatrix(e, p), o.get_pmatrix(e, p)) for e in range(net.num_edges + 1) for p in range(g.P))
            _compare_all_clvs(g, o, exact=same_p)
            trees = [g.tree_info(net.root, t)[1].copy() for t in range(g.num_trees(net.root))]
            res[mode] = (lg, trees, launches)
            if mode == "1":
                assert launches == 1, launches      # ONE launch: P-matrices (deferred K1), every CLV and the root lnLs
                # an incremental evaluation after one branch changed goes back to per-node launches on the same slots
                e = net.num_edges // 2
                g.set_branch_length(e, 0.123); o.set_branch_length(e, 0.123)
                assert g.computeLoglikelihood(1, 1) == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
                g.set_branch_length(e, float(net.edge_length[e])); o.set_branch_length(e, float(net.edge_length[e]))
                assert g.computeLoglikelihood(1, 1) == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
                assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)   # and a walk again
            g.close()
        assert res["1"][0] == pytest.approx(res["0"][0], rel=REPLAY_RTOL)
        for a, b in zip(res["0"][1], res["1"][1]):
            np.testing.assert_allclose(a, b, rtol=REPLAY_RTOL, atol=0)
        assert res["1"][2] < res["0"][2]


def test_async_tip_upload_is_validated_on_the_device_and_serves_any_alphabet():
    """Round 2 (ADVICE): nrx_set_tipchars_u8 / nrx_set_tipcodes_u8 enqueue the copy and return; a device kernel checks
    every code (an illegal one fails the next synchronising call with pll_set_tip_states' message) and rebuilds the invariant-site
    table, so +I may be switched on after such an upload; 20-state partitions take 1-byte codes + the code -> state-set map."""
    from netrax_b200._capi import LikelihoodError
    net = random_network(12, 2, seed=21)
    m, w = simulate_alignment(net, 700, seed=21)
    m[:, :90] = m[0, :90]                           # some invariant columns for +I
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    l0 = g.computeLoglikelihood(0, 1)
    assert l0 == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    codes = np.ascontiguousarray(m.astype(np.uint8))
    w32 = np.ascontiguousarray(w.astype(np.uint32))
    bad = codes.copy(); bad[3, 17] = 0              # no state at all
    g.upload_alignment_u8(0, bad.ctypes.data, w32.ctypes.data)
    with pytest.raises(LikelihoodError, match="Illegal state code in tip"):
        g.computeLoglikelihood(0, 1)
    bad[3, 17] = 200                                # not a DNA code
    g.upload_alignment_u8(0, bad.ctypes.data, w32.ctypes.data)
    with pytest.raises(LikelihoodError, match="Illegal state code in tip"):
        g.computeLoglikelihood(0, 1)
    g.upload_alignment_u8(0, codes.ctypes.data, w32.ctypes.data)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(l0, rel=REPLAY_RTOL)
    g.set_pinv(0, 0.3); o.set_pinv(0, 0.3)          # +I after an asynchronous upload: the invariant table is current
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    g.close()
    # protein: codes + map
    net, part = _protein_case(10, 1, 257, 22)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    lo = o.computeLoglikelihood(0, 1)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lo, rel=LNL_RTOL)
    tipmap, inv = np.unique(part.tip_masks, return_inverse=True)
    codes = np.ascontiguousarray(inv.reshape(part.tip_masks.shape).astype(np.uint8))
    perm = np.arange(len(tipmap))[::-1].copy()      # a different code assignment than the engine derived itself
    codes_p = np.ascontiguousarray(perm[codes].astype(np.uint8)); tipmap_p = np.zeros(len(tipmap), np.uint32); tipmap_p[perm] = tipmap
    w32 = np.ascontiguousarray(part.pattern_weights.astype(np.uint32))
    g.upload_alignment_codes(0, codes_p.ctypes.data, tipmap_p, w32.ctypes.data)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lo, rel=LNL_RTOL)
    e = net.num_edges - 1
    assert g.brlen_prepare(e) == pytest.approx(o.brlen_prepare(e), rel=LNL_RTOL)
    assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
    assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


@pytest.mark.parametrize("cfg", [(25, 3, 333, 3), (40, 5, 150, 4), (30, 6, 200, 8), (16, 4, 1, 9)])
def test_node_centric_k2_bitexact(cfg, monkeypatch):
    """Round 2, VERDICT item 7: in a replayed plan the ops of a node that share children run on k_clv_node_dna4 (each distinct child
    staged once, P . child once per child, pair products streamed).  Same CLVs and scalers bit for bit as the per-op kernel
    (NRX_NODE=0) and as libpll; the tile walk is switched off so that the level-by-level plan is what runs."""
    n, r, pat, seed = cfg
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    o = _oracle(net, [part])
    lo = o.computeLoglikelihood(0, 1)
    monkeypatch.setenv("NRX_WALK", "0")
    res = {}
    for mode in ("0", "1"):
        monkeypatch.setenv("NRX_NODE", mode)
        g = _gpu(net, [part])
        _inject_eigen(g, o)
        g.computeLoglikelihood(0, 1)
        n0 = g.launch_count()
        lg = g.computeLoglikelihood(0, 1)      # replay of the plan
        res[mode] = (lg, g.launch_count() - n0, [g.tree_info(net.root, t)[1].copy() for t in range(g.num_trees(net.root))])
        assert lg == pytest.approx(lo, rel=LNL_RTOL)
        same_p = all(np.array_equal(g.get_pmatrix(e), o.get_pmatrix(e)) for e in range(net.num_edges + 1))
        _compare_all_clvs(g, o, exact=same_p)
        g.close()
    assert res["0"][0] == res["1"][0]            # same CLVs, same K3 kernel and order: identical lnL
    for a, b in zip(res["0"][2], res["1"][2]):
        assert np.array_equal(a, b)
    if r >= 4:
        assert res["1"][1] > res["0"][1]         # the node-centric launches exist (one extra launch per batch that has groups)


def _sweep_records(eng, net, order, accept):
    """The derivative sweep of bench.py with everything it returns recorded per edge; `accept` keeps the last proposal length
    (what a real optimize_branch does) instead of restoring the old one."""
    out = []
    for e in order:
        e = int(e)
        t0 = float(eng.branch_lengths()[e])
        rec = [eng.brlen_prepare(e), eng.computeLoglikelihoodBrlenOpt(e)]
        n = eng.computePartitionSumtables(e)
        if n:
            for k in range(3):
                eng.brlen_set_length(e, t0 * (1.0 + 0.1 * (k + 1)))
                d = eng.computeLoglikelihoodDerivatives(e)
                rec += [d[0], d[1]] + list(d[4].ravel())
            if not accept:
                eng.brlen_set_length(e, t0)
        rec.append(eng.brlen_finish(e))
        out.append(np.array(rec))
    return out


@pytest.mark.parametrize("cfg", [(30, 3, 300, 5), (24, 5, 200, 6), (40, 2, 150, 7)])
@pytest.mark.parametrize("accept", [False, True])
def test_shadow_rerooting_memo_matches_inplace_reference(cfg, accept):
    """Round 2, VERDICT item 5: virtual re-rooting writes into shadow slots and memoises the re-rooted trees of the path nodes.  A
    sweep over all branches in pre-order — lengths restored, or the last proposal kept as a real optimisation does — returns, per
    edge, the same old lnL / edge-rooted lnL / derivatives / final lnL as (a) the same engine with the memo switched off (bit for
    bit) and (b) the checker, which re-roots in place and recomputes as the reference does (VirtualRerooting.cpp:192-252)."""
    n, r, pat, seed = cfg
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, g0, o = _gpu(net, [part]), _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    _inject_eigen(g0, o)
    g0.set_reroot_cache_slots(0)
    for eng in (g, g0, o):
        eng.computeLoglikelihood(0, 1)
    order = g.brlen_sweep_order()
    assert sorted(int(e) for e in order) == list(range(net.num_edges))
    n0 = g.launch_count()
    rg = _sweep_records(g, net, order, accept)
    launches = g.launch_count() - n0
    n0 = g0.launch_count()
    rg0 = _sweep_records(g0, net, order, accept)
    launches0 = g0.launch_count() - n0
    ro = _sweep_records(o, net, order, accept)
    for e, a, b, c in zip(order, rg, rg0, ro):
        assert np.array_equal(a, b), int(e)                       # a memo hit returns exactly what the miss computed
        np.testing.assert_allclose(a[:2], c[:2], rtol=LNL_RTOL)
        np.testing.assert_allclose(a[-1], c[-1], rtol=LNL_RTOL)
        np.testing.assert_allclose(a[2:-1], c[2:-1], rtol=DERIV_RTOL, atol=1e-7)
    st, st0 = g.reroot_stats(), g0.reroot_stats()
    # restored lengths: about one new node per branch; kept proposals: every recomputed ancestor (both parents of a reticulation below)
    # changes the data identity of the path nodes that hang it off as a side child, so fewer calls hit
    assert (st["hits"] > st["misses"] if not accept else st["hits"] > 0) and st0["hits"] == 0 and st0["entries"] == 0
    assert launches < launches0
    # the root-directed CLVs were never overwritten: a full re-evaluation agrees, and so does every node's tree set
    lg, lo = g.computeLoglikelihood(1, 1), o.computeLoglikelihood(1, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(lg, rel=1e-13)
    _compare_all_clvs(g, o, exact=False)
    g.close(); g0.close()


def test_shadow_rerooting_restored_length_leaves_nothing_invalid():
    """prepare -> edge lnL -> derivatives at other lengths -> length restored -> finish: no CLV is recomputed (the reference's in-place
    re-rooting recomputes the whole path); a changed length recomputes the nodes above the edge, as before."""
    net = random_network(30, 3, seed=12)
    m, w = simulate_alignment(net, 500, seed=12)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    l0 = g.computeLoglikelihood(0, 1)
    o.computeLoglikelihood(0, 1)
    e = int(net.num_edges // 2)
    t0 = float(net.edge_length[e])
    g.brlen_prepare(e)
    g.computeLoglikelihoodBrlenOpt(e)
    if g.computePartitionSumtables(e):
        g.brlen_set_length(e, 2 * t0)
        g.computeLoglikelihoodDerivatives(e)
        g.brlen_set_length(e, t0)
    u0 = g.clv_update_count()
    assert g.brlen_finish(e) == l0
    assert g.clv_update_count() == u0
    # mid-session evaluation closes the session first (the reference would mix re-rooted and root-directed CLVs)
    g.brlen_prepare(e)
    assert g.computeLoglikelihood(1, 1) == l0
    # a changed length
    for eng in (g, o):
        eng.brlen_prepare(e)
        eng.computeLoglikelihoodBrlenOpt(e)
        eng.brlen_set_length(e, 1.7 * t0)
    lg, lo = g.brlen_finish(e), o.brlen_finish(e)
    assert lg == pytest.approx(lo, rel=LNL_RTOL) and lg != l0
    assert g.clv_update_count() > u0
    _compare_all_clvs(g, o, exact=False)
    g.close()


@pytest.mark.parametrize("kind", ["unlinked_best", "protein"])
def test_shadow_rerooting_memo_other_shapes(kind):
    """The memo keys on the branch lengths of EVERY partition and is independent of the kernels underneath: unlinked branch lengths
    over ragged partitions with the BEST variant, and a protein partition (tensor-core K2 / K4 / K5), swept in pre-order with the
    last proposal kept — memo on == memo off bit for bit, both equal to the in-place checker."""
    from netrax_b200.synth import lg_model
    if kind == "protein":
        net = random_network(12, 2, seed=31)
        rates, freqs = lg_model()
        m, w = simulate_alignment(net, 160, seed=31, states=20, rates=rates, freqs=freqs)
        parts, kw = [Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w)], {}
    else:
        net = random_network(16, 3, seed=32)
        rng = np.random.default_rng(3)
        parts, brl = [], []
        for p, pat in enumerate((300, 97, 211)):
            m, w = simulate_alignment(net, pat, seed=40 + p)
            parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES * (1 + 0.1 * p), GAMMA4_ALPHA05, pattern_weights=w))
            brl.append(net.edge_length * rng.uniform(0.5, 2, net.num_edges))
        kw = dict(variant=BEST, linkage=UNLINKED, partition_brlens=brl)
    g, g0, o = _gpu(net, parts, **kw), _gpu(net, parts, **kw), _oracle(net, parts, **kw)
    _inject_eigen(g, o)
    _inject_eigen(g0, o)
    g0.set_reroot_cache_slots(0)
    for eng in (g, g0, o):
        eng.computeLoglikelihood(0, 1)
    order = g.brlen_sweep_order()
    rg, rg0, ro = (_sweep_records(eng, net, order, True) for eng in (g, g0, o))
    for e, a, b, c in zip(order, rg, rg0, ro):
        assert np.array_equal(a, b), int(e)
        np.testing.assert_allclose(a[:2], c[:2], rtol=LNL_RTOL)
        np.testing.assert_allclose(a[-1], c[-1], rtol=LNL_RTOL)
        np.testing.assert_allclose(a[2:-1], c[2:-1], rtol=DERIV_RTOL, atol=1e-7)
    assert g.reroot_stats()["hits"] > 0
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL)
    g.close(); g0.close()


def test_score_only_evaluation_skips_root_clv_stores_and_heals():
    """VERDICT r1 item 7 (second part): with score_only on, a full evaluation that replays the fused-K3 plan does not store the CLVs of
    the root displayed trees — the lnL is the same to the last bit (same kernels, same per-site terms), the root slots are stale, and
    the next incremental evaluation / CLV read-back re-evaluates with the stores on before anything reads them."""
    net = random_network(40, 4, seed=9)
    m, w = simulate_alignment(net, 6000, seed=9)     # large enough that the level-by-level plan (not the tile walk) runs
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    lo = o.computeLoglikelihood(0, 1)
    g.computeLoglikelihood(0, 1)
    l_full = g.computeLoglikelihood(0, 1)            # replayed plan, stores on
    root = net.root
    before = [g.read_clv(root, t).copy() for t in range(g.num_trees(root))]
    # poison one branch so that stale root CLVs would be visible, then score-only
    e = 0
    t0 = float(net.edge_length[e])
    g.set_branch_length(e, 3.0 * t0)
    o.set_branch_length(e, 3.0 * t0)
    g.set_score_only(True)
    l_so = g.computeLoglikelihood(0, 1)
    assert l_so == pytest.approx(o.computeLoglikelihood(0, 1), rel=LNL_RTOL) and l_so != l_full
    g.set_score_only(False)
    l_ref = g.computeLoglikelihood(0, 1)
    assert l_so == l_ref                             # bit-identical lnL with and without the stores
    g.set_score_only(True)
    assert g.computeLoglikelihood(0, 1) == l_ref
    # read-back heals: the CLVs handed out are the ones of the current branch lengths, equal to the checker's
    _compare_all_clvs(g, o, exact=False)
    after = [g.read_clv(root, t) for t in range(g.num_trees(root))]
    assert any(not np.array_equal(a, b) for a, b in zip(before, after))
    # incremental evaluation after a score-only one
    assert g.computeLoglikelihood(0, 1) == l_ref     # score-only again (stale)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(l_ref, rel=REPLAY_RTOL)
    for eng in (g, o):
        eng.brlen_prepare(e)
    assert g.computeLoglikelihoodBrlenOpt(e) == pytest.approx(o.computeLoglikelihoodBrlenOpt(e), rel=LNL_RTOL)
    assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
    g.close()


@pytest.mark.parametrize("cfg", [(30, 3, 300, 5), (24, 5, 200, 6), (40, 0, 150, 7)])
def test_lazy_rerooting_sweep_matches_reference_flow(cfg):
    """Lazy re-rooting: a pre-order sweep that KEEPS every proposal, without the evaluations from the root around each branch.  Per
    edge the edge-rooted lnL and every derivative equal the checker's (which evaluates from the root before and after every branch,
    as the reference does); the lnL a lazy finish returns is the edge-rooted one = the checker's root lnL to rounding; after the
    sweep a plain incremental evaluation settles everything and agrees with the checker, CLV for CLV."""
    n, r, pat, seed = cfg
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    for eng in (g, o):
        eng.computeLoglikelihood(0, 1)
    order = g.brlen_sweep_order()
    _sweep_records(g, net, order, False)          # first pass: every branch's re-rooting plan becomes known (topology-only)
    g.set_lazy_rerooting(True)
    u0, l0 = g.clv_update_count(), g.launch_count()
    rg = _sweep_records(g, net, order, True)
    lazy_updates, lazy_launches = g.clv_update_count() - u0, g.launch_count() - l0
    st = g.lazy_reroot_stats()
    assert st["sessions"] > net.num_edges // 3 and st["fallbacks"] < st["sessions"], st
    ro = _sweep_records(o, net, order, True)
    for e, a, c in zip(order, rg, ro):
        np.testing.assert_allclose(a[1], c[1], rtol=LNL_RTOL, err_msg=str(int(e)))           # edge-rooted lnL
        np.testing.assert_allclose(a[-1], c[-1], rtol=LNL_RTOL, err_msg=str(int(e)))         # final lnL
        np.testing.assert_allclose(a[2:-1], c[2:-1], rtol=DERIV_RTOL, atol=1e-7)
    g.set_lazy_rerooting(False)
    lg, lo = g.computeLoglikelihood(1, 1), o.computeLoglikelihood(1, 1)
    assert lg == pytest.approx(lo, rel=LNL_RTOL)
    _compare_all_clvs(g, o, exact=False)
    # the same sweep without the lazy mode recomputes the path above every branch
    g2 = _gpu(net, [part])
    _inject_eigen(g2, o)
    g2.computeLoglikelihood(0, 1)
    _sweep_records(g2, net, order, False)
    u0, l0 = g2.clv_update_count(), g2.launch_count()
    _sweep_records(g2, net, order, True)
    assert lazy_updates < g2.clv_update_count() - u0 and lazy_launches < g2.launch_count() - l0
    g.close(); g2.close()


def test_lazy_optimize_branch_matches_reference_flow():
    """optimize_branch (Newton-Raphson) over all branches in pre-order with lazy re-rooting: same final lengths and lnL as the checker's
    reference flow."""
    net = random_network(25, 2, seed=13)
    m, w = simulate_alignment(net, 400, seed=13)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _gpu(net, [part]), _oracle(net, [part])
    _inject_eigen(g, o)
    for eng in (g, o):
        eng.computeLoglikelihood(0, 1)
    order = [int(e) for e in g.brlen_sweep_order()]
    _sweep_records(g, net, order, False)
    g.set_lazy_rerooting(True)
    for e in order:
        lg, lo = g.optimize_branch(e), o.optimize_branch(e)
        assert lg == pytest.approx(lo, rel=1e-9), e
    np.testing.assert_allclose(g.branch_lengths(), o.branch_lengths(), rtol=1e-6, atol=1e-9)
    g.set_lazy_rerooting(False)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(o.computeLoglikelihood(1, 1), rel=1e-9)
    g.close()


@pytest.mark.parametrize("name", ["small", "two_reticulations", "three_reticulations", "reticulation_in_reticulation", "tree"])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_edge_lnl_and_sumtables_in_one_pass(name, variant):
    """computeLoglikelihoodBrlenOptAndSumtables (k_edge_sum_dna4q: the edge lnL and the sumtable of a displayed-tree pair from ONE read
    of its two CLVs) on every edge of the reference's fixtures: the lnL equals the separate call's and the checker's, every sumtable
    entry equals the separate K5's bit for bit and the checker's to rounding, the derivatives taken from them too."""
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    g, g2, o = _gpu(net, [part], variant=variant), _gpu(net, [part], variant=variant), _oracle(net, [part], variant=variant)
    _inject_eigen(g, o)
    _inject_eigen(g2, o)
    for eng in (g, g2, o):
        eng.computeLoglikelihood(0, 1)
    for e in range(net.num_edges):
        for eng in (g, g2, o):
            eng.brlen_prepare(e)
        lf, nf = g.computeLoglikelihoodBrlenOptAndSumtables(e)
        ls, ns = g2.computeLoglikelihoodBrlenOpt(e), g2.computePartitionSumtables(e)
        lo, no = o.computeLoglikelihoodBrlenOpt(e), o.computePartitionSumtables(e)
        assert nf == ns == no
        assert lf == pytest.approx(ls, rel=1e-13) and lf == pytest.approx(lo, rel=LNL_RTOL)
        for i in range(nf):
            sf, pf, lf_, rf = g.read_sumtable(0, i)
            ss, ps, ls_, rs = g2.read_sumtable(0, i)
            so, po, lo_, ro = o.read_sumtable(0, i)
            assert (lf_, rf) == (ls_, rs) == (lo_, ro) and pf == ps
            assert np.array_equal(sf, ss), (e, i)
            np.testing.assert_allclose(sf, so, rtol=1e-10, atol=1e-14 * np.abs(so).max())
        if nf:
            for eng in (g, g2, o):
                eng.brlen_set_length(e, 0.13)
            df, ds, do = g.computeLoglikelihoodDerivatives(e), g2.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
            assert df[0] == ds[0] and df[1] == ds[1]
            assert df[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-7) and df[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-7)
            for eng in (g, g2, o):
                eng.brlen_set_length(e, float(net.edge_length[e]))
        for eng in (g, g2, o):
            eng.brlen_finish(e)
    g.close(); g2.close()


def test_edge_lnl_and_sumtables_in_one_pass_pinv_and_other_shapes():
    """+I (terma / terminv through the quad reduction), unlinked ragged partitions, and a protein partition (not covered by the fused
    kernel: the call falls back to the two engine calls) — same numbers as the separate calls and as the checker."""
    from netrax_b200.synth import lg_model
    cases = []
    net = random_network(18, 3, seed=51)
    m, w = simulate_alignment(net, 257, seed=51)
    cases.append((net, [Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)], {}, 0.3))
    rng = np.random.default_rng(5)
    parts, brl = [], []
    for p, pat in enumerate((130, 61)):
        m, w = simulate_alignment(net, pat, seed=60 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES * (1 + 0.2 * p), GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2, net.num_edges))
    cases.append((net, parts, dict(linkage=UNLINKED, partition_brlens=brl), 0.0))
    pnet = random_network(9, 1, seed=52)
    rates, freqs = lg_model()
    pm, pw = simulate_alignment(pnet, 90, seed=52, states=20, rates=rates, freqs=freqs)
    cases.append((pnet, [Partition(20, 4, pm, freqs, rates, GAMMA4_ALPHA05, pattern_weights=pw)], {}, 0.0))
    for net, parts, kw, pinv in cases:
        g, o = _gpu(net, parts, **kw), _oracle(net, parts, **kw)
        _inject_eigen(g, o)
        if pinv:
            for eng in (g, o):
                eng.set_pinv(0, pinv)
        for eng in (g, o):
            eng.computeLoglikelihood(0, 1)
        for e in range(0, net.num_edges, 2):
            for eng in (g, o):
                eng.brlen_prepare(e)
            lf, nf = g.computeLoglikelihoodBrlenOptAndSumtables(e)
            lo, no = o.computeLoglikelihoodBrlenOpt(e), o.computePartitionSumtables(e)
            assert nf == no and lf == pytest.approx(lo, rel=LNL_RTOL)
            if nf:
                dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
                np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-7)
            assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL)
        g.close()


def test_staged_alignment_upload_double_buffer():
    """nrxh_stage_alignment_u8 / nrxh_commit_staged_alignment: the next alignment is copied on the engine's copy stream while the
    current one is evaluated; only the commit makes it live.  Two different alignments alternate: every evaluation sees exactly the
    alignment committed before it (lnL of the checker for that alignment), also with +I (the invariant-site table is rebuilt at
    commit), and an illegal staged code fails the evaluation after its commit."""
    from netrax_b200._capi import LikelihoodError
    net = random_network(14, 2, seed=23)
    ma, wa = simulate_alignment(net, 900, seed=23)
    mb, wb = simulate_alignment(net, 900, seed=24)
    ma[:, :70] = ma[0, :70]
    parts = {k: Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w) for k, (m, w) in (("a", (ma, wa)), ("b", (mb, wb)))}
    oa, ob = _oracle(net, [parts["a"]]), _oracle(net, [parts["b"]])
    g = _gpu(net, [parts["a"]])
    _inject_eigen(g, oa)   # both checkers hold the same model, hence the same decomposition
    la, lb = oa.computeLoglikelihood(0, 1), ob.computeLoglikelihood(0, 1)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(la, rel=LNL_RTOL)
    bufs = {k: (np.ascontiguousarray(m.astype(np.uint8)), np.ascontiguousarray(w.astype(np.uint32))) for k, (m, w) in (("a", (ma, wa)), ("b", (mb, wb)))}
    want = {"a": la, "b": lb}
    g.stage_alignment_u8(0, bufs["b"][0].ctypes.data, bufs["b"][1].ctypes.data)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(la, rel=LNL_RTOL)          # staged, not committed: still alignment a
    seq = ["b", "a", "b", "b", "a"]
    for k, name in enumerate(seq):
        g.commit_staged_alignment()
        if k + 1 < len(seq):
            nxt = seq[k + 1]
            g.stage_alignment_u8(0, bufs[nxt][0].ctypes.data, bufs[nxt][1].ctypes.data)   # overlaps with the evaluation below
        assert g.computeLoglikelihood(0, 1) == pytest.approx(want[name], rel=LNL_RTOL), (k, name)
    for eng in (g, oa):
        eng.set_pinv(0, 0.25)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(oa.computeLoglikelihood(0, 1), rel=LNL_RTOL)   # last committed: a
    bad = bufs["b"][0].copy(); bad[2, 5] = 0
    g.stage_alignment_u8(0, bad.ctypes.data, 0)
    g.commit_staged_alignment()
    with pytest.raises(LikelihoodError, match="Illegal state code in tip"):
        g.computeLoglikelihood(0, 1)
    g.close()


def test_config2_full_size_sweep_slice_matches_oracle():
    """BASELINE config 2 at its FULL size (100 k patterns): the bench's derivative sweep on a slice of the pre-order — eight consecutive
    branches (their re-rooting paths share memoised nodes), proposals kept, edge lnL and sumtables made in one pass, then again with lazy
    re-rooting — against the reference's libpll under the restated in-place driver, edge by edge."""
    import bench
    cfg = dict(bench.CONFIGS[2])
    net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    kw = dict(variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    g, o = _gpu(net, parts, **kw), _oracle(net, parts, **kw)
    _inject_eigen(g, o)
    for eng in (g, o):
        eng.computeLoglikelihood(0, 1)
    order = [int(e) for e in g.brlen_sweep_order()]
    for lazy, chunk in ((False, order[20:28]), (True, order[28:36])):
        if lazy:
            _sweep_records(g, net, chunk, False)      # the branches' re-rooting plans become known
            g.set_lazy_rerooting(True)
        for e in chunk:
            t0 = float(g.branch_lengths()[e])
            assert o.branch_lengths()[e] == t0
            for eng in (g, o):
                eng.brlen_prepare(e)
            lg, ng = g.computeLoglikelihoodBrlenOptAndSumtables(e)
            lo, no = o.computeLoglikelihoodBrlenOpt(e), o.computePartitionSumtables(e)
            assert ng == no and lg == pytest.approx(lo, rel=LNL_RTOL), e
            if ng:
                for k in range(2):
                    for eng in (g, o):
                        eng.brlen_set_length(e, t0 * (1.15 + 0.2 * k))
                    dg, do = g.computeLoglikelihoodDerivatives(e), o.computeLoglikelihoodDerivatives(e)
                    np.testing.assert_allclose(dg[4], do[4], rtol=DERIV_RTOL, atol=1e-7)
                    assert dg[0] == pytest.approx(do[0], rel=DERIV_RTOL, abs=1e-6) and dg[1] == pytest.approx(do[1], rel=DERIV_RTOL, abs=1e-6)
            assert g.brlen_finish(e) == pytest.approx(o.brlen_finish(e), rel=LNL_RTOL), e
    g.set_lazy_rerooting(False)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(o.computeLoglikelihood(1, 1), rel=LNL_RTOL)
    assert g.reroot_stats()["hits"] > 0 and g.lazy_reroot_stats()["sessions"] > 0
    g.close()


def test_batched_scoring_with_score_only():
    """Candidate scoring as the search would use it: several networks in flight (computeLoglikelihoodBatch), score-only replays (root
    CLVs not stored) — same lnLs as the plain sequential evaluations, and an incremental batch afterwards heals the stale roots."""
    from netrax_b200.engine import compute_loglikelihood_batch
    base = random_network(30, 3, seed=91)
    m, w = simulate_alignment(base, 5000, seed=91)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    nets = [base] + [random_network(30, r, seed=92 + r) for r in (1, 2, 4)]
    gs = [_gpu(n, [part]) for n in nets]
    want = np.array([g.computeLoglikelihood(0, 1) for g in gs])
    compute_loglikelihood_batch(gs, 0, 1)                                  # plans recorded
    for g in gs:
        g.set_score_only(True)
    got = compute_loglikelihood_batch(gs, 0, 1)
    np.testing.assert_allclose(got, want, rtol=REPLAY_RTOL)
    np.testing.assert_array_equal(got, compute_loglikelihood_batch(gs, 0, 1))
    for k, g in enumerate(gs):
        g.set_branch_length(k + 2, 0.07 * (k + 1))
    inc = compute_loglikelihood_batch(gs, 1, 1)                            # incremental after score-only: full re-evaluation with the stores on
    for g in gs:
        g.set_score_only(False)
    np.testing.assert_allclose(inc, np.array([g.computeLoglikelihood(0, 1) for g in gs]), rtol=REPLAY_RTOL)
    for g in gs:
        g.close()
