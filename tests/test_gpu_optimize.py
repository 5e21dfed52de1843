"""GPU parity (-m gpu) for the immediate callers of the path: branch-length optimisation (Newton-Raphson on the
device sumtable/derivative kernels, Brent on the full lnL) and reticulation-probability optimisation (Brent over the
re-mixed cached per-tree lnLs), product vs oracle on the same inputs.  The optimisers amplify nothing: every iterate
is a function of lnL / derivative values that already agree to 1e-10 / 1e-8, so final lnLs must agree to 1e-9 relative
and the optimised parameters to the solvers' own tolerance."""
import numpy as np
import pytest

from helpers import FIXTURE_PAIRS, load_fixture
from netrax_b200._capi import AVERAGE, BEST, BRENT_NORMAL, BRENT_REROOT, NEWTON_RAPHSON, UNLINKED, LikelihoodError, Partition
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment

pytestmark = pytest.mark.gpu


def _gpu(net, parts, **kw):
    from netrax_b200.engine import NetraxB200
    return NetraxB200(net, parts, **kw)


def _oracle(net, parts, **kw):
    from oracle import oracle
    return oracle.make_engine("ref" if oracle.have_ref() else "port", net, parts, **kw)


def _pair(net, parts, **kw):
    g, o = _gpu(net, parts, **kw), _oracle(net, parts, **kw)
    for p in range(g.P):
        g.set_eigen(p, *o.get_eigen(p))
    return g, o


@pytest.mark.parametrize("name", ["small", "two_reticulations", "three_reticulations", "interleaved_reticulations", "celine"])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimize_branch_every_edge_matches_oracle(name, variant):
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    g, o = _pair(net, [part], variant=variant)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=1e-10)
    for e in range(net.num_edges):
        lg, lo = g.optimize_branch(e), o.optimize_branch(e)
        assert lg == pytest.approx(lo, rel=1e-9), e
        assert g.branch_lengths()[e] == pytest.approx(o.branch_lengths()[e], rel=1e-5, abs=2e-7), e
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=1e-10)


@pytest.mark.parametrize("method", [NEWTON_RAPHSON, BRENT_NORMAL])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimize_branches_and_reticulations_match_oracle(method, variant):
    net = random_network(16, 3, seed=5)
    m, w = simulate_alignment(net, 800, seed=5)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _pair(net, [part], variant=variant)
    l0 = g.computeLoglikelihood(0, 1)
    lg, lo = g.optimize_branches(method=method), o.optimize_branches(method=method)
    assert lg >= l0 - 1e-3
    assert lg == pytest.approx(lo, rel=1e-9)
    np.testing.assert_allclose(g.branch_lengths(), o.branch_lengths(), rtol=1e-5, atol=2e-7)
    launches = g.launch_count()
    rg, ro = g.optimize_reticulations(), o.optimize_reticulations()
    assert rg == pytest.approx(ro, rel=1e-9) and rg >= lg - 1e-3
    np.testing.assert_allclose(g.reticulation_probs(), o.reticulation_probs(), rtol=1e-7)
    # row f2: reticulation-probability optimisation re-mixes cached per-tree lnLs on the host — not one kernel launch
    assert g.launch_count() == launches
    assert rg == pytest.approx(g.computeLoglikelihood(0, 1), rel=1e-12)


def test_optimize_branches_unlinked_partitions_match_oracle():
    net = random_network(10, 2, seed=9)
    parts, brl = [], []
    rng = np.random.default_rng(2)
    for p in range(3):
        m, w = simulate_alignment(net, 300, seed=30 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
    g, o = _pair(net, parts, linkage=UNLINKED, partition_brlens=brl)
    lg, lo = g.optimize_branches(), o.optimize_branches()
    assert lg == pytest.approx(lo, rel=1e-9)
    for p in range(3):
        np.testing.assert_allclose(g.branch_lengths(p), o.branch_lengths(p), rtol=1e-5, atol=2e-7)


def test_brent_reroot_raises_the_reference_error():
    net, part = load_fixture(*FIXTURE_PAIRS["small"])
    g = _gpu(net, [part])
    with pytest.raises(LikelihoodError, match="Cannot reuse old displayed trees"):
        g.optimize_branch(0, method=BRENT_REROOT)
    # the engine stays usable: a full re-evaluation recovers
    assert np.isfinite(g.computeLoglikelihood(0, 1))


@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimize_alpha_matches_oracle(variant):
    """Model-parameter loop: the ALPHA step of optimize_params on the device — every Brent iterate is a
    new set of Gamma rates + one full re-evaluation (plan replay) — against pll-modules' real minimiser over libpll."""
    net = random_network(12, 2, seed=3)
    parts = []
    for k in range(2):
        m, w = simulate_alignment(net, 600, seed=30 + k)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    g, o = _pair(net, parts, variant=variant)
    for eng in (g, o):
        eng.set_alpha(0, 2.0)
        eng.set_alpha(1, 0.1)
    l0g, l0o = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
    assert l0g == pytest.approx(l0o, rel=1e-10)
    lg, lo = g.optimize_alpha(), o.optimize_alpha()
    assert lg >= l0g - 1e-6
    assert lg == pytest.approx(lo, rel=1e-9)
    for p in range(2):
        assert g.get_alpha(p) == pytest.approx(o.get_alpha(p), rel=1e-5)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(lg, rel=1e-13)
    # the loop around it (optimizeAllNonTopology's order: model, reticulation probabilities, branch lengths)
    rg, ro = g.optimize_reticulations(), o.optimize_reticulations()
    assert rg == pytest.approx(ro, rel=1e-9) and rg >= lg - 1e-3
    g.close()


@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimize_all_non_topology_matches_oracle(variant):
    """The whole non-topology optimisation round (model = alpha, reticulation probabilities, branch lengths, scored by
    BIC; src/optimization/Optimization.cpp:118-214) on the device against the oracle."""
    net = random_network(10, 2, seed=9)
    m, w = simulate_alignment(net, 400, seed=9)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g, o = _pair(net, [part], variant=variant)
    for eng in (g, o):
        eng.set_alpha(0, 1.5)
        eng.set_scoring_sizes(9)
    b0g, b0o = g.scoreNetwork(), o.scoreNetwork()
    assert b0g == pytest.approx(b0o, rel=1e-10)
    bg, bo = g.optimizeAllNonTopology(1), o.optimizeAllNonTopology(1)
    assert bg <= b0g + 1e-3
    assert bg == pytest.approx(bo, rel=1e-8)
    assert g.get_alpha(0) == pytest.approx(o.get_alpha(0), rel=1e-4)
    np.testing.assert_allclose(g.branch_lengths(), o.branch_lengths(), rtol=1e-4, atol=2e-6)
    g.close()


def test_optimize_pinv_matches_oracle():
    """The PINV step of optimize_params (ModelOptimization.cpp:67-76) on the device — every Brent iterate rescales the
    rates by 1 / (1 - pinv) and re-evaluates with the invariant-site terms — against pll-modules' real minimiser over
    libpll's +I kernels."""
    from test_oracle_optimize import _pinv_case
    net, parts = _pinv_case()
    g, o = _pair(net, parts)
    for eng in (g, o):
        eng.set_pinv(0, 0.05)
    l0g, l0o = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
    assert l0g == pytest.approx(l0o, rel=1e-10)
    lg, lo = g.optimize_pinv(), o.optimize_pinv()
    assert lg >= l0g - 1e-6
    assert lg == pytest.approx(lo, rel=1e-9)
    assert g.get_pinv(0) == pytest.approx(o.get_pinv(0), rel=1e-4)
    assert g.get_pinv(1) == 0.0
    assert g.computeLoglikelihood(1, 1) == pytest.approx(lg, rel=1e-13)
    g.close()


@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimize_scalers_matches_oracle(variant):
    """optimize_scalers (BranchLengthOptimization.cpp:581-599 -> pllmod_algo_opt_brlen_scalers_treeinfo): scalers forced
    into range, one Brent search per partition (each iterate refreshes that partition's P-matrices from scaler x linked
    length and replays the evaluation plan), normalisation to a site-weighted mean of 1."""
    from netrax_b200._capi import SCALED
    from test_oracle_netrax import scaled_linkage_case
    net, parts, _ = scaled_linkage_case()
    g, o = _pair(net, parts, variant=variant, linkage=SCALED)
    for eng in (g, o):
        for p, s in enumerate([3.0, 0.3, 150.0]):
            eng.set_brlen_scaler(p, s)
        eng.set_scoring_sizes(9)
    b0g, b0o = g.scoreNetwork(), o.scoreNetwork()
    assert b0g == pytest.approx(b0o, rel=1e-10)
    bg, bo = g.optimize_scalers(), o.optimize_scalers()
    assert bg <= b0g
    assert bg == pytest.approx(bo, rel=1e-9)
    np.testing.assert_allclose(g.brlen_scalers(), o.brlen_scalers(), rtol=1e-4)
    np.testing.assert_allclose(g.branch_lengths(), o.branch_lengths(), rtol=1e-4)
    wsum = np.array([float(p.pattern_weights.sum()) for p in parts])
    assert float((g.brlen_scalers() * wsum).sum() / wsum.sum()) == pytest.approx(1.0, rel=1e-12)
    assert g.computeLoglikelihood(0, 1) == pytest.approx(o.computeLoglikelihood(0, 1), rel=1e-10)
    g.close()


def test_optimize_scalers_is_a_noop_without_scaled_linkage():
    net = random_network(8, 1, seed=4)
    m, w = simulate_alignment(net, 300, seed=4)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    g = _gpu(net, [part, part])
    b0 = g.scoreNetwork()
    assert g.optimize_scalers() == b0
    np.testing.assert_array_equal(g.brlen_scalers(), [1.0, 1.0])
    g.close()


def test_score_only_from_files_matches_oracle():
    """netrax --score_only (src/main.cpp:287-326) from the reference's fixture files: network file + FASTA + model string ->
    partitions -> optimizeModel -> BIC / lnL -> optimizeAllNonTopology(SLOW) -> BIC / lnL / AIC / AICc + the written network,
    the whole flow on the device against the same flow over the oracle."""
    import os
    from helpers import FIX
    from netrax_b200.score import score_only
    from oracle import oracle
    nw, aln = FIXTURE_PAIRS["small"]
    net_text, msa_text = open(os.path.join(FIX, nw)).read(), open(os.path.join(FIX, aln)).read()
    model = "GTR{1/2.5/0.8/1.2/3.0/1}+FC+G"
    rg = score_only(lambda net, parts, **kw: _gpu(net, parts, **kw), net_text, msa_text, model, log=None)
    ro = score_only(lambda net, parts, **kw: oracle.make_engine("ref" if oracle.have_ref() else "port", net, parts, **kw),
                    net_text, msa_text, model, log=None)
    for key in ("start_bic", "start_logl"):
        assert rg[key] == pytest.approx(ro[key], rel=1e-9)
    for key in ("bic", "logl", "aic", "aicc"):
        assert rg[key] == pytest.approx(ro[key], rel=1e-7)
    assert rg["alphas"][0] == pytest.approx(ro["alphas"][0], rel=1e-3)
    assert rg["bic"] <= rg["start_bic"] + 1e-3


def test_params_to_optimize_flags_and_capi_range_checks():
    """(1) which partitions optimize_pinv treats as free comes from pll-modules' params_to_optimize flags, not from
    "the value is > 0": a +I partition that starts at proportion 0 is optimised once flagged, and a flagged-off partition is left
    alone; (2) the flat C-ABI rejects out-of-range edge / reticulation / node indices and probabilities outside
    [brprob_min, brprob_max] instead of writing out of bounds."""
    from netrax_b200._capi import LikelihoodError
    from test_oracle_optimize import _pinv_case
    net, parts = _pinv_case()
    g, o = _pair(net, parts)
    for eng in (g, o):
        eng.set_params_to_optimize(0, alpha=False, pinv=True)    # +I, starting at 0
        eng.set_params_to_optimize(1, alpha=False, pinv=False)
    assert g.get_pinv(0) == 0.0
    lg, lo = g.optimize_pinv(), o.optimize_pinv()
    assert lg == pytest.approx(lo, rel=1e-9)
    assert g.get_pinv(0) == pytest.approx(o.get_pinv(0), rel=1e-4) and g.get_pinv(0) > 0.01
    assert g.get_pinv(1) == 0.0
    # range checks
    with pytest.raises(LikelihoodError, match="out of range"):
        g.set_branch_length(net.num_edges + 5, 0.1)
    with pytest.raises(LikelihoodError, match="unlinked"):
        g.set_branch_length(0, 0.1, partition=0)                 # linked linkage: no per-partition lengths
    with pytest.raises(LikelihoodError, match="negative"):
        g.set_branch_length(0, -1.0)
    if net.num_reticulations:
        with pytest.raises(LikelihoodError, match="out of range"):
            g.set_reticulation_prob(net.num_reticulations, 0.5)
        for bad in (0.0, 1.0, -0.1, float("nan")):
            with pytest.raises(LikelihoodError, match="brprob"):
                g.set_reticulation_prob(0, bad)
    assert g.num_trees(net.num_nodes + 3) == -1
    with pytest.raises(LikelihoodError, match="out of range"):
        g.tree_config(net.num_nodes + 3, 0)
    assert g.computeLoglikelihood(1, 1) == pytest.approx(lg, rel=1e-12)   # nothing was corrupted
    g.close()
