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
