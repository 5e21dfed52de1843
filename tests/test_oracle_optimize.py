"""Oracle tests for the immediate callers of the path (optimize_branch(es), optimize_reticulation(s)).

Pins: (i) the C restatement of pll-modules' Newton-Raphson / Brent (oracle/opt_port.c) against the REFERENCE's
own opt_algorithms.c compiled into oracle/_ref, on analytic targets — same optimum AND same evaluation sequence;
(ii) the restated NetRAX optimiser loops over the scalar port against the same loops over the real libpll + real
pll-modules minimisers; (iii) the invariants of the reference's tests: a branch-length optimisation never makes the
lnL worse (test/src/BrlenOptTest.cpp:95-120 `ASSERT_GE(new_logl, old_logl)`), optimising reticulation probabilities
never makes it worse (src/optimization/Optimization.cpp:93-106)."""
import ctypes as C
import math

import numpy as np
import pytest

from helpers import FIXTURE_PAIRS, load_fixture
from netrax_b200._capi import AVERAGE, BEST, BRENT_NORMAL, BRENT_REROOT, NEWTON_RAPHSON, UNLINKED, LikelihoodError, Partition
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment
from oracle import oracle

needs_ref = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (needs /root/reference at build time)")
TARGET_T = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double)
DERIV_T = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))


def _minimisers(kind):
    lib = oracle.api(kind).lib
    lib.orc_test_brent.restype = C.c_int
    lib.orc_test_brent.argtypes = [C.c_int, C.c_double, C.c_double, C.c_double, C.c_double, TARGET_T, C.POINTER(C.c_double)]
    lib.orc_test_newton.restype = C.c_int
    lib.orc_test_newton.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_uint, DERIV_T, C.POINTER(C.c_int)]
    return lib


BRENT_CASES = [
    (lambda x: (x - 0.3) ** 2, 1e-6, 0.5, 1 - 1e-6, 0.1),
    (lambda x: (x - 0.3) ** 2, 1e-6, 0.5, 1 - 1e-6, 1e-4),
    (lambda x: -math.log(0.2 * x + 0.5 * (1 - x)), 1e-6, 0.7, 1 - 1e-6, 0.1),          # monotone: optimum at a bound
    (lambda x: math.cosh(3 * (x - 2.5)) + 0.1 * x, 1e-6, 0.1, 100.0, 0.1),
    (lambda x: abs(x - 1e-3) ** 1.5, 1e-6, 1e-6, 100.0, 1e-3),
    (lambda x: 1.0, 0.0, 0.0, 1.0, 0.1),                                                 # flat, xguess == 0 branch of the bracketing
]


@needs_ref
@pytest.mark.parametrize("case", range(len(BRENT_CASES)))
def test_brent_restatement_equals_reference(case):
    f, lo, guess, hi, tol = BRENT_CASES[case]
    lib = _minimisers("ref")
    seqs = []
    for use_ref in (1, 0):
        calls = []

        def target(_, x):
            calls.append(x)
            return f(x)
        out = C.c_double()
        assert lib.orc_test_brent(use_ref, lo, guess, hi, tol, TARGET_T(target), C.byref(out))
        seqs.append((out.value, calls))
    assert seqs[0][0] == seqs[1][0]
    assert seqs[0][1] == seqs[1][1]          # identical evaluation sequence, including the post-convergence repeats
    assert len(seqs[0][1]) == 5 + 101 + 1    # the reference always spends 107 target evaluations


NEWTON_CASES = [
    (lambda x: (2 * (x - 0.37), 2.0), 1e-6, 0.1, 100.0, 1e-7, 32),
    (lambda x: (math.exp(x) - 3.0, math.exp(x)), 1e-6, 5.0, 100.0, 1e-7, 32),
    (lambda x: (-1.0 / x + 0.5, 1.0 / (x * x)), 1e-6, 0.1, 100.0, 1e-7, 32),
    (lambda x: (1.0, -1.0), 1e-6, 0.5, 100.0, 1e-7, 8),     # negative curvature: marches to the lower bound
    (lambda x: (math.sin(5 * x), 5 * math.cos(5 * x)), 1e-6, 1.0, 100.0, 1e-7, 4),   # hits the iteration limit
    # f = df = 0 (a saturated branch): dx = -0/0 = NaN, which PLL_MAX(PLL_MIN(dx, dxmax), -dxmax) turns into +dxmax
    (lambda x: (0.0, 0.0) if x < 30 else (x - 40.0, 1.0), 1e-6, 5.0, 100.0, 1e-7, 32),
    (lambda x: (0.0, 0.0), 1e-6, 99.0, 100.0, 1e-7, 32),
    (lambda x: (1.0, 0.0), 1e-6, 50.0, 100.0, 1e-7, 8),     # df = 0, f != 0: dx = -inf, clamped to -dxmax
]


@needs_ref
@pytest.mark.parametrize("case", range(len(NEWTON_CASES)))
def test_newton_restatement_equals_reference(case):
    f, lo, guess, hi, tol, iters = NEWTON_CASES[case]
    lib = _minimisers("ref")
    res = []
    for use_ref in (1, 0):
        calls = []

        def deriv(_, x, d1, d2):
            calls.append(x[0])
            d1[0], d2[0] = f(x[0])
        x, st = C.c_double(guess), C.c_int()
        assert lib.orc_test_newton(use_ref, lo, C.byref(x), hi, tol, iters, DERIV_T(deriv), C.byref(st))
        res.append((x.value, st.value, calls))
    assert res[0] == res[1]


def _small_case(seed=3, taxa=12, ret=2, sites=500):
    net = random_network(taxa, ret, seed=seed)
    m, w = simulate_alignment(net, sites, seed=seed)
    return net, Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)


@needs_ref
@pytest.mark.parametrize("method", [NEWTON_RAPHSON, BRENT_NORMAL])
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimisers_port_equals_reference_libraries(method, variant):
    net, part = _small_case()
    out = []
    for kind in ("port", "ref"):
        e = oracle.make_engine(kind, net, [part], variant=variant)
        l0 = e.computeLoglikelihood(0, 1)
        l1 = e.optimize_branches(method=method)
        l2 = e.optimize_reticulations()
        assert l1 >= l0 - 1e-3 and l2 >= l1 - 1e-3
        assert l2 == pytest.approx(e.computeLoglikelihood(0, 1), rel=1e-12)   # incremental state == full re-evaluation
        out.append((l1, l2, e.branch_lengths(), e.reticulation_probs()))
    assert out[0][0] == pytest.approx(out[1][0], rel=1e-11)
    assert out[0][1] == pytest.approx(out[1][1], rel=1e-11)
    np.testing.assert_allclose(out[0][2], out[1][2], rtol=1e-9, atol=1e-12)
    np.testing.assert_allclose(out[0][3], out[1][3], rtol=1e-9)


@pytest.mark.parametrize("name", ["small", "two_reticulations", "three_reticulations", "celine"])
def test_optimize_branch_never_worse_on_reference_fixtures(name):
    """BrlenOptTest.cpp:95-120: after optimising the branches the lnL is not worse; every single optimize_branch call
    also returns a value equal to a full re-evaluation of the updated network."""
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    e = oracle.make_engine("ref" if oracle.have_ref() else "port", net, [part])
    prev = e.computeLoglikelihood(0, 1)
    for edge in range(net.num_edges):
        l = e.optimize_branch(edge)
        assert l >= prev - 1e-6, (edge, l, prev)
        prev = l
    assert prev == pytest.approx(e.computeLoglikelihood(0, 1), rel=1e-12)
    lens = e.branch_lengths()
    assert lens.min() >= 1e-6 and lens.max() <= 100.0


def test_brent_reroot_throws_like_the_reference():
    """optimize_branch_brent ends with invalidatePmatrixIndex (BranchLengthOptimization.cpp:153), which invalidates the
    root CLVs; the following computeLoglikelihoodBrlenOpt refuses them (VirtualRerooting.cpp:370-379).  BRENT_REROOT
    is therefore unusable in the reference as shipped, and the restatement reproduces the same exception."""
    net, part = _small_case()
    e = oracle.make_engine("port", net, [part])
    with pytest.raises(LikelihoodError, match="Cannot reuse old displayed trees"):
        e.optimize_branch(0, method=BRENT_REROOT)


def test_unlinked_newton_uses_partition_zero_derivative():
    """Quirk Q7 (BranchLengthOptimization.cpp:190-195): with unlinked branch lengths every partition's Newton-Raphson run
    is driven by partition 0's derivatives; the lnL guard keeps the result from getting worse."""
    net = random_network(8, 1, seed=11)
    parts, brl = [], []
    rng = np.random.default_rng(1)
    for p in range(3):
        m, w = simulate_alignment(net, 200, seed=20 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
    e = oracle.make_engine("port", net, parts, linkage=UNLINKED, partition_brlens=brl)
    l0 = e.computeLoglikelihood(0, 1)
    l1 = e.optimize_branch(2)
    assert l1 >= l0 - 1e-9
    assert l1 == pytest.approx(e.computeLoglikelihood(0, 1), rel=1e-12)


# ---------------------------------------------------------------------------------------------- model loop: alpha
@pytest.mark.parametrize("variant", [AVERAGE, BEST])
def test_optimize_alpha_port_equals_reference_minimiser(variant):
    """The ALPHA step of optimize_params (ModelOptimization.cpp:56-65): the restated Brent-multi driver (opt_port.c)
    against pll-modules' real pllmod_opt_minimize_brent_multi + libpll's Gamma rates (_ref), two partitions that start
    from different wrong shapes; the data were simulated with alpha = 0.5."""
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    net = random_network(12, 2, seed=3)
    parts = []
    for k in range(2):
        m, w = simulate_alignment(net, 600, seed=30 + k)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    res = {}
    for kind in ("port", "ref"):
        e = oracle.make_engine(kind, net, parts, variant=variant)
        e.set_alpha(0, 2.0)
        e.set_alpha(1, 0.1)
        l0 = e.computeLoglikelihood(0, 1)
        l1 = e.optimize_alpha()
        assert l1 >= l0 - 1e-6
        res[kind] = (l0, l1, e.get_alpha(0), e.get_alpha(1), e.computeLoglikelihood(0, 1))
        assert res[kind][4] == pytest.approx(l1, rel=1e-12)
    assert res["port"][0] == pytest.approx(res["ref"][0], rel=1e-10)
    assert res["port"][1] == pytest.approx(res["ref"][1], rel=1e-9)
    assert res["port"][2] == pytest.approx(res["ref"][2], rel=1e-5)
    assert res["port"][3] == pytest.approx(res["ref"][3], rel=1e-5)
    for a in res["ref"][2:4]:
        assert 0.3 < a < 0.8   # the generating shape is 0.5


def test_optimize_alpha_skips_partitions_without_shape():
    """Partitions whose rates were given directly (alpha = 0: not in params_to_optimize) are left alone."""
    net = random_network(8, 1, seed=4)
    m, w = simulate_alignment(net, 300, seed=4)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    e = oracle.make_engine("port", net, [part])
    l0 = e.computeLoglikelihood(0, 1)
    assert e.optimize_alpha() == pytest.approx(l0, rel=1e-13)
    assert e.get_alpha(0) == 0.0


def test_optimize_all_non_topology_improves_bic_and_flavours_agree():
    """scoreNetwork (BIC, ComplexityScoring.cpp:57-67) + optimizeAllNonTopology (Optimization.cpp:118-214) over the scalar
    port and over real libpll + real pll-modules minimisers: same trajectory end point; the BIC never gets worse."""
    net = random_network(10, 2, seed=9)
    m, w = simulate_alignment(net, 400, seed=9)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    res = {}
    for kind in (["port", "ref"] if oracle.have_ref() else ["port"]):
        e = oracle.make_engine(kind, net, [part])
        e.set_alpha(0, 1.5)
        e.set_scoring_sizes(9)   # GTR (5) + frequencies (3) + alpha (1)
        b0 = e.scoreNetwork()
        l0 = e.computeLoglikelihood(1, 1)
        k = 9 + net.num_reticulations + net.num_edges
        assert b0 == pytest.approx(-2 * l0 + k * math.log(float(w.sum()) * net.num_tips), rel=1e-13)
        b1 = e.optimizeAllNonTopology(1)
        assert b1 <= b0 + 1e-3
        res[kind] = (b1, e.get_alpha(0))
    if "ref" in res:
        assert res["port"][0] == pytest.approx(res["ref"][0], rel=1e-9)
        assert res["port"][1] == pytest.approx(res["ref"][1], rel=1e-5)


# ------------------------------------------------------------------------- model loop: +I and the branch-length scalers
def _pinv_case():
    net = random_network(10, 2, seed=6)
    parts = []
    for k in range(2):
        m, w = simulate_alignment(net, 500, seed=60 + k)
        w = w.copy()
        if k == 0:   # pile weight on the constant columns so that +I has something to find
            const = np.all(m == m[0:1], axis=0) & (np.bitwise_count(m[0]) == 1)
            w[const] *= 6
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    return net, parts


def test_optimize_pinv_port_equals_reference_minimiser():
    """The PINV step of optimize_params (ModelOptimization.cpp:67-76): pllmod_algo_opt_onedim_treeinfo(PINV) restated over the
    scalar port (its own +I terms and Brent-multi) against pll-modules' real Brent-multi over libpll's +I kernels (_ref).
    Partitions without +I (pinv = 0) are not in params_to_optimize and stay untouched; the end point is a maximum of the lnL
    in the proportion."""
    net, parts = _pinv_case()
    res = {}
    for kind in (["port", "ref"] if oracle.have_ref() else ["port"]):
        e = oracle.make_engine(kind, net, parts)
        e.set_pinv(0, 0.05)
        l0 = e.computeLoglikelihood(0, 1)
        l1 = e.optimize_pinv()
        assert l1 >= l0 - 1e-6
        assert e.get_pinv(1) == 0.0
        x = e.get_pinv(0)
        assert 0.1 < x < 0.99   # the inflated constant columns pull the proportion up from 0.05
        assert e.computeLoglikelihood(0, 1) == pytest.approx(l1, rel=1e-12)
        for d in (-0.02, 0.02):
            e.set_pinv(0, x + d)
            assert e.computeLoglikelihood(0, 1) < l1
        res[kind] = (l0, l1, x)
        e.close()
    if "ref" in res:
        assert res["port"][0] == pytest.approx(res["ref"][0], rel=1e-10)
        assert res["port"][1] == pytest.approx(res["ref"][1], rel=1e-9)
        assert res["port"][2] == pytest.approx(res["ref"][2], rel=1e-4)


def test_optimize_scalers_port_equals_reference_minimiser():
    """optimize_scalers (BranchLengthOptimization.cpp:581-599) = pllmod_algo_opt_brlen_scalers_treeinfo (pllmod_algorithm.c:
    869-960): out-of-range scalers forced into [0.01, 100] with the branches scaled by the inverse, one Brent search per
    partition, scalers normalised to a site-weighted mean of 1 with the branches scaled by the mean — the product
    scaler x length, and with it the lnL, survives the normalisation."""
    from netrax_b200._capi import SCALED
    from test_oracle_netrax import scaled_linkage_case
    net, parts, _ = scaled_linkage_case()
    res = {}
    for kind in (["port", "ref"] if oracle.have_ref() else ["port"]):
        e = oracle.make_engine(kind, net, parts, linkage=SCALED)
        for p, s in enumerate([3.0, 0.3, 150.0]):   # 150 > RAXML_BRLEN_SCALER_MAX: fix_brlen_scalers rescales all three
            e.set_brlen_scaler(p, s)
        e.set_scoring_sizes(9)
        b0, l0 = e.scoreNetwork(), e.computeLoglikelihood(1, 1)
        b1 = e.optimize_scalers()
        l1 = e.computeLoglikelihood(1, 1)
        sc, bl = e.brlen_scalers(), e.branch_lengths()
        assert b1 <= b0 and l1 >= l0
        wsum = np.array([float(p.pattern_weights.sum()) for p in parts])
        assert float((sc * wsum).sum() / wsum.sum()) == pytest.approx(1.0, rel=1e-12)
        assert e.computeLoglikelihood(0, 1) == pytest.approx(l1, rel=1e-10)   # cached value vs a full re-evaluation of the normalised state
        res[kind] = (b1, sc, bl)
        e.close()
    if "ref" in res:
        assert res["port"][0] == pytest.approx(res["ref"][0], rel=1e-9)
        np.testing.assert_allclose(res["port"][1], res["ref"][1], rtol=1e-4)
        np.testing.assert_allclose(res["port"][2], res["ref"][2], rtol=1e-4)


def test_optimize_scalers_is_a_noop_without_scaled_linkage():
    net = random_network(8, 1, seed=4)
    m, w = simulate_alignment(net, 300, seed=4)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    e = oracle.make_engine("port", net, [part, part])
    b0 = e.scoreNetwork()
    assert e.optimize_scalers() == b0
    np.testing.assert_array_equal(e.brlen_scalers(), [1.0, 1.0])
    e.close()


@pytest.mark.parametrize("kind", ["port"] + (["ref"] if oracle.have_ref() else []))
def test_params_to_optimize_flags_select_the_free_pinv_partitions(kind):
    """pllmod_algo_opt_onedim_treeinfo selects partitions by params_to_optimize[p] & param (PLLMOD/algorithm/pllmod_algorithm.c:765-772),
    not by the parameter's current value: a +I partition starting at proportion 0 is optimised, an unflagged one is not."""
    net, parts = _pinv_case()
    e = oracle.make_engine(kind, net, parts)
    l0 = e.computeLoglikelihood(0, 1)
    assert e.optimize_pinv() == pytest.approx(l0, rel=1e-13) and e.get_pinv(0) == 0.0     # by value: nothing to optimise
    e.set_params_to_optimize(0, alpha=False, pinv=True)
    e.set_params_to_optimize(1, alpha=False, pinv=False)
    l1 = e.optimize_pinv()
    assert l1 > l0 + 1e-3 and e.get_pinv(0) > 0.01 and e.get_pinv(1) == 0.0
    e.close()
