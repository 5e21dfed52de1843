"""CPU-only checks of the product's host side: the C-ABI libraries load and export every symbol the headers
declare (no compute calls without a GPU), the host model helpers (gamma rates) match the oracle, and the
engine refuses to run without a CUDA device instead of falling back."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import netrax_b200.engine as eng
from oracle import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared(header, prefix):
    txt = open(os.path.join(ROOT, "include", header)).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(" + prefix + r"[a-z_0-9]+)\s*\(", txt)))


@pytest.mark.parametrize("header,prefix,so", [("nrx_engine.h", "nrx_", eng.ENGINE_SO), ("netrax_b200.h", "nrxh_", eng.HOST_SO)])
def test_shared_libraries_export_every_declared_symbol(header, prefix, so):
    assert os.path.exists(so), f"{so} not built (python -c 'import __graft_entry__ as g; g.build()')"
    C.CDLL(eng.ENGINE_SO, mode=C.RTLD_GLOBAL)
    lib = C.CDLL(so)
    names = [n for n in _declared(header, prefix) if not n.endswith("_cb")]
    assert len(names) > 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/{header} but not exported by {os.path.basename(so)}"


def test_gamma_rates_match_oracle():
    api = eng.load()
    for alpha in (0.02, 0.1, 0.5, 0.75, 1.5, 10.0, 50.0):
        for cats in (1, 2, 4, 8):
            for mode in (0, 1):
                # same published algorithms, independently written: equal to rounding (exactly equal for alpha >= 0.1)
                np.testing.assert_allclose(api.gamma_rates(alpha, cats, mode), oracle.api("port").gamma_rates(alpha, cats, mode), rtol=1e-11)


@pytest.mark.skipif(eng.device_count() > 0, reason="a CUDA device is present")
def test_no_cpu_fallback():
    from netrax_b200._capi import LikelihoodError, Partition
    from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment
    net = random_network(6, 1, seed=1)
    m, w = simulate_alignment(net, 50, seed=1)
    with pytest.raises(LikelihoodError, match="no usable CUDA device"):
        eng.NetraxB200(net, [Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)])


@pytest.mark.parametrize("datatype", ["DNA", "PROT", "ODD"])
def test_product_eigendecomposition_reproduces_libpll_golden_pmatrices(datatype):
    """The PRODUCT's host eigendecomposition (host/model.cpp: cyclic Jacobi, not libpll's Householder + QL) against libpll's
    test/out/pmatrix.out: P(t) = I + V^-1 diag(expm1(lambda r t)) V assembled here in numpy from the host's eigen output —
    the formula K1 evaluates on the device (LIBPLL/core_pmatrix.c:24-244) — for 4 / 20 / 5 states, equal / skewed / extreme
    frequencies and exchangeabilities, branch lengths 1e-6 .. 100, category rates 1e-31 .. 100; 9 printed decimals."""
    import os
    from helpers import GOLDEN, pmatrix_golden_inputs
    from netrax_b200 import engine
    G = np.load(os.path.join(GOLDEN, "libpll_pmatrix_golden.npz"))
    S = {"DNA": 4, "PROT": 20, "ODD": 5}[datatype]
    freqs, substs = pmatrix_golden_inputs(S)
    for j in range(3):
        for k in range(3):
            ev, iev, evals = engine.eigen_decompose(S, freqs[j], substs[k])
            V, Vinv, lam = ev[:, :S], iev[:, :S], evals[:S]
            for b, t in enumerate(G["branch_lengths"]):
                for c, r in enumerate(G["cat_rates"]):
                    P = np.eye(S) + Vinv @ np.diag(np.expm1(lam * r * t)) @ V
                    np.testing.assert_allclose(P, G[f"{datatype}_P_{j * 3 + k}"][b][c], atol=6e-10, rtol=0,
                                               err_msg=str((datatype, j, k, t, r)))


def _product_minimisers():
    import ctypes as C
    from netrax_b200 import engine
    lib = engine.load().lib
    TARGET_T = C.CFUNCTYPE(C.c_double, C.c_void_p, C.c_double)
    DERIV_T = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double))
    lib.nrxh_minimize_brent.restype = C.c_int
    lib.nrxh_minimize_brent.argtypes = [C.c_double, C.c_double, C.c_double, C.c_double, TARGET_T, C.c_void_p, C.POINTER(C.c_double)]
    lib.nrxh_minimize_newton.restype = C.c_int
    lib.nrxh_minimize_newton.argtypes = [C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, C.c_uint, DERIV_T, C.c_void_p, C.POINTER(C.c_int)]
    return lib, TARGET_T, DERIV_T


def test_product_newton_equals_the_reference_minimiser_call_for_call():
    """The host library's Newton-Raphson (optimize.cpp newtonMulti) against pll-modules' real pllmod_opt_minimize_newton_multi
    (oracle/_ref) on the cases of test_oracle_optimize.py, incl. the f = df = 0 ones where the step is NaN and libpll's
    PLL_MIN / PLL_MAX operand order decides the next iterate: same iterates, same final x, same status."""
    import ctypes as C
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from test_oracle_optimize import NEWTON_CASES, _minimisers
    ref = _minimisers("ref")
    lib, _, DERIV_T = _product_minimisers()
    for f, lo, guess, hi, tol, iters in NEWTON_CASES:
        res = []
        for which in ("ref", "product"):
            calls = []

            def deriv(_, x, d1, d2):
                calls.append(x[0])
                d1[0], d2[0] = f(x[0])
            x, st = C.c_double(guess), C.c_int()
            if which == "ref":
                assert ref.orc_test_newton(1, lo, C.byref(x), hi, tol, iters, DERIV_T(deriv), C.byref(st))
            else:
                assert lib.nrxh_minimize_newton(lo, C.byref(x), hi, tol, iters, DERIV_T(deriv), None, C.byref(st))
            res.append((x.value, st.value, calls))
        assert res[0] == res[1]


def test_product_brent_equals_the_reference_minimiser_up_to_convergence():
    """brentSingle against pllmod_opt_minimize_brent: the same optimum and the same evaluation sequence — the product's is the
    reference's with the post-convergence repeats of the last proposal cut (deviation D1), then the final target(xopt)."""
    import ctypes as C
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from test_oracle_optimize import BRENT_CASES, TARGET_T as ORC_TARGET_T, _minimisers
    ref = _minimisers("ref")
    lib, TARGET_T, _ = _product_minimisers()
    for f, lo, guess, hi, tol in BRENT_CASES:
        seqs = []
        for which in ("ref", "product"):
            calls = []

            def target(_, x):
                calls.append(x)
                return f(x)
            out = C.c_double()
            if which == "ref":
                assert ref.orc_test_brent(1, lo, guess, hi, tol, ORC_TARGET_T(target), C.byref(out))
            else:
                assert lib.nrxh_minimize_brent(lo, guess, hi, tol, TARGET_T(target), None, C.byref(out))
            seqs.append((out.value, calls))
        (xr, cr), (xp, cp) = seqs
        assert xp == xr
        assert cp[-1] == xr and cr[-1] == xr                     # both finish with target(xopt)
        body = cp[:-1]
        assert body == cr[:len(body)]                            # identical bracketing + iterates
        assert all(c == cr[len(body) - 1] for c in cr[len(body):-1]) or len(body) == len(cr) - 1   # what was cut: repeats of the last proposal


def test_product_brent_multi_equals_the_reference_minimiser_call_for_call():
    """brentMulti (the driver under optimize_alpha / optimize_pinv / optimize_scalers) against pll-modules' real
    pllmod_opt_minimize_brent_multi and its restatement in the oracle port, three variables with different curvature that
    converge at different iterations: identical x vectors call after call (converged variables keep being passed, the callee
    skips them), identical optima."""
    import ctypes as C
    import math
    from oracle import oracle
    lib = _product_minimisers()[0]
    MT = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int))
    lib.nrxh_minimize_brent_multi.restype = C.c_int
    lib.nrxh_minimize_brent_multi.argtypes = [C.c_uint, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, MT, C.c_void_p]
    funcs = [lambda x: (x - 0.5) ** 2, lambda x: math.cosh(2 * (x - 3.0)) + 0.05 * x, lambda x: -math.log(0.2 * x + 0.1) + 0.3 * x, lambda x: 1.0]
    cases = [([2.0, 0.1, 1.0, 1.0], 0.0201, 100.0, 0.001), ([0.05, 0.3, 0.0, 0.5], 0.0, 0.99, 0.001), ([3.0, 0.3, 150.0, 1.0], 0.01, 100.0, 0.001)]
    for guess, lo, hi, tol in cases:
        n = len(guess)
        runs = {}
        kinds = ["product", "port"] + (["ref"] if oracle.have_ref() else [])
        for which in kinds:
            calls = []

            def target(_, x, fx, conv):
                xs = [x[j] for j in range(n)]
                flags = None if not conv else [conv[j] for j in range(n)]
                calls.append((xs, flags))
                unconverged = 0
                for j in range(n):
                    if conv and conv[j]:
                        continue
                    unconverged = 1
                if fx:
                    for j in range(n):
                        fx[j] = funcs[j](x[j])
                if conv:
                    conv[n] = 0 if unconverged else 1
                return 0.0
            x = (C.c_double * n)(*guess)
            if which == "product":
                assert lib.nrxh_minimize_brent_multi(n, lo, x, hi, tol, MT(target), None)
            else:
                olib = oracle.api("ref" if which == "ref" else "port").lib
                olib.orc_test_brent_multi.restype = C.c_int
                olib.orc_test_brent_multi.argtypes = [C.c_int, C.c_uint, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, MT]
                assert olib.orc_test_brent_multi(1 if which == "ref" else 0, n, lo, x, hi, tol, MT(target))
            runs[which] = (list(x), calls)
        for which in kinds[1:]:
            assert runs["product"][0] == runs[which][0], which
            assert runs["product"][1] == runs[which][1], which
        assert 5 + 2 < len(runs["product"][1]) <= 5 + 101 + 1


def test_product_minimisers_fuzzed_against_the_reference():
    """300 seeded random targets — quadratics, exponentials, flat and step functions, oscillations, zones returning NaN or inf —
    with random bounds, guesses (also out of range) and tolerances: the product's Newton and Brent produce the reference's
    iterates and results bit for bit (NaN-aware comparison through repr)."""
    import ctypes as C
    import math
    import random
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    from test_oracle_optimize import TARGET_T as ORC_TARGET_T, _minimisers
    ref = _minimisers("ref")
    lib, TARGET_T, DERIV_T = _product_minimisers()
    rnd = random.Random(1)

    def make():
        kind = rnd.choice(["quad", "exp", "flat", "step", "nanzone", "inf", "osc", "lin"])
        a, b, c = rnd.uniform(0, 5), rnd.uniform(0.1, 10), rnd.uniform(-3, 3)
        ex = lambda x: math.exp(min(50, b * (x - a)))
        f = {"quad": lambda x: b * (x - a) ** 2 + c, "exp": lambda x: ex(x) - c * x, "flat": lambda x: c,
             "step": lambda x: c if x < a else c + b, "nanzone": lambda x: float("nan") if a < x < a + b else (x - a) ** 2,
             "inf": lambda x: float("inf") if x > a + b else (x - a) ** 2, "osc": lambda x: math.sin(b * x) + 0.01 * x,
             "lin": lambda x: b * x + c}[kind]
        d = {"quad": lambda x: (2 * b * (x - a), 2 * b), "exp": lambda x: (b * ex(x) - c, b * b * ex(x)), "flat": lambda x: (0.0, 0.0),
             "step": lambda x: (0.0, 0.0) if x < a else (1.0, 0.0), "nanzone": lambda x: (float("nan"), 1.0) if a < x < a + b else (2 * (x - a), 2.0),
             "inf": lambda x: (float("inf"), 1.0) if x > a + b else (2 * (x - a), 2.0),
             "osc": lambda x: (b * math.cos(b * x) + 0.01, -b * b * math.sin(b * x)), "lin": lambda x: (b, 0.0)}[kind]
        return f, d

    for _ in range(300):
        f, d = make()
        lo, hi = rnd.choice([(1e-6, 100.0), (1e-6, 1 - 1e-6), (0.0, 0.99), (0.01, 100.0)])
        guess = rnd.uniform(lo, hi) if rnd.random() < 0.8 else rnd.choice([lo, hi, hi * 2, -1.0])
        tol, iters = rnd.choice([0.1, 1e-3, 1e-4, 1e-7]), rnd.choice([4, 8, 32])
        res = []
        for which in (0, 1):
            calls = []

            def deriv(_, x, d1, d2):
                calls.append(repr(x[0]))
                d1[0], d2[0] = d(x[0])
            x, st = C.c_double(guess), C.c_int()
            if which == 0:
                ref.orc_test_newton(1, lo, C.byref(x), hi, tol, iters, DERIV_T(deriv), C.byref(st))
            else:
                lib.nrxh_minimize_newton(lo, C.byref(x), hi, tol, iters, DERIV_T(deriv), None, C.byref(st))
            res.append((repr(x.value), st.value, calls))
        assert res[0] == res[1]
        seqs = []
        for which in (0, 1):
            calls = []

            def target(_, x):
                calls.append(repr(x))
                return f(x)
            out = C.c_double()
            if which == 0:
                ref.orc_test_brent(1, lo, guess, hi, tol, ORC_TARGET_T(target), C.byref(out))
            else:
                lib.nrxh_minimize_brent(lo, guess, hi, tol, TARGET_T(target), None, C.byref(out))
            seqs.append((repr(out.value), calls))
        (xr, cr), (xp, cp) = seqs
        assert xr == xp and cp[:-1] == cr[:len(cp) - 1]


def test_product_brent_multi_fuzzed_against_the_reference():
    """150 seeded cases, 1-5 variables with random targets (incl. NaN zones, steps, flat functions), random bounds / guesses /
    tolerances: the product's Brent-multi driver and pll-modules' pllmod_opt_minimize_brent_multi issue identical calls
    (x vectors and convergence flags) and return identical optima."""
    import ctypes as C
    import math
    import random
    from oracle import oracle
    if not oracle.have_ref():
        pytest.skip("oracle/_ref not built")
    lib = _product_minimisers()[0]
    MT = C.CFUNCTYPE(C.c_double, C.c_void_p, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_int))
    lib.nrxh_minimize_brent_multi.restype = C.c_int
    lib.nrxh_minimize_brent_multi.argtypes = [C.c_uint, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, MT, C.c_void_p]
    olib = oracle.api("ref").lib
    olib.orc_test_brent_multi.restype = C.c_int
    olib.orc_test_brent_multi.argtypes = [C.c_int, C.c_uint, C.c_double, C.POINTER(C.c_double), C.c_double, C.c_double, MT]
    rnd = random.Random(7)

    def make():
        kind = rnd.choice(["quad", "cosh", "flat", "step", "nanzone", "lin", "osc"])
        a, b, c = rnd.uniform(0, 5), rnd.uniform(0.1, 10), rnd.uniform(-3, 3)
        return {"quad": lambda x: b * (x - a) ** 2 + c, "cosh": lambda x: math.cosh(min(30, b * (x - a))) + c * x, "flat": lambda x: c,
                "step": lambda x: c if x < a else c + b, "nanzone": lambda x: float("nan") if a < x < a + 0.3 * b else (x - a) ** 2,
                "lin": lambda x: b * x + c, "osc": lambda x: math.sin(b * x) + 0.01 * x}[kind]

    for _ in range(150):
        n = rnd.choice([1, 2, 3, 5])
        funcs = [make() for _ in range(n)]
        lo, hi = rnd.choice([(0.0201, 100.0), (0.0, 0.99), (0.01, 100.0), (1e-6, 1 - 1e-6)])
        guess = [rnd.uniform(lo, hi) if rnd.random() < 0.85 else rnd.choice([lo, hi, 2 * hi, 0.0]) for _ in range(n)]
        tol = rnd.choice([0.1, 1e-3, 1e-4])
        runs = []
        for which in (0, 1):
            calls = []

            def target(_, x, fx, conv):
                calls.append(([repr(x[j]) for j in range(n)], None if not conv else [conv[j] for j in range(n)]))
                unconverged = 0 if conv and all(conv[j] for j in range(n)) else 1
                if fx:
                    for j in range(n):
                        fx[j] = funcs[j](x[j])
                if conv:
                    conv[n] = 0 if unconverged else 1
                return 0.0
            x = (C.c_double * n)(*guess)
            if which == 0:
                assert lib.nrxh_minimize_brent_multi(n, lo, x, hi, tol, MT(target), None)
            else:
                assert olib.orc_test_brent_multi(1, n, lo, x, hi, tol, MT(target))
            runs.append(([repr(v) for v in x], calls))
        assert runs[0] == runs[1]
