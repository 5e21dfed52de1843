"""The C++ upper seam as a NetRAX caller sees it: tests/cpp/upper_seam_caller.cpp includes only
netrax_likelihood_api.hpp, links libnetrax_b200.so and calls computeLoglikelihood / updateCLVsVirtualRerootTrees /
computeLoglikelihoodBrlenOpt / computePartitionSumtables / computeLoglikelihoodDerivatives / optimize_branches /
optimize_reticulations / scoreNetwork / network_logl_wrapper with the reference's names and argument lists.
CPU: it compiles, links and — without a CUDA device — fails loudly (no CPU fallback).  GPU: its numbers match the oracle."""
import os
import subprocess

import numpy as np
import pytest

from helpers import FIXTURE_PAIRS, load_fixture

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "upper_seam_caller.cpp")
LIBDIR = os.path.join(ROOT, "netrax_b200")


def _build(tmp_path):
    exe = str(tmp_path / "upper_seam_caller")
    cmd = ["g++", "-O1", "-std=c++17", SRC, "-o", exe, f"-L{LIBDIR}", "-lnetrax_b200", "-lnrx_engine", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]
    return exe


def _write_input(path, net, part):
    with open(path, "w") as f:
        f.write(f"{net.num_tips} {net.num_nodes} {net.root} {net.num_edges} {net.num_reticulations} {part.sites}\n")
        for e in range(net.num_edges):
            f.write(f"{int(net.edge_source[e])} {int(net.edge_target[e])} {float(net.edge_length[e])!r} {float(net.edge_prob[e])!r}\n")
        for r in range(net.num_reticulations):
            f.write(f"{int(net.ret_node[r])} {int(net.ret_first_edge[r])} {int(net.ret_second_edge[r])}\n")
        f.write(" ".join(str(int(x)) for x in part.tip_masks.reshape(-1)) + "\n")
        w = part.pattern_weights if part.pattern_weights is not None else np.ones(part.sites, dtype=np.uint32)
        f.write(" ".join(str(int(x)) for x in w) + "\n")
        for arr in (part.freqs, part.subst, part.rates):
            f.write(" ".join(repr(float(x)) for x in arr) + "\n")


def test_cpp_caller_compiles_links_and_has_no_cpu_fallback(tmp_path):
    exe = _build(tmp_path)
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present: covered by the gpu test")
    net, part = load_fixture(*FIXTURE_PAIRS["small"])
    inp = str(tmp_path / "in.txt")
    _write_input(inp, net, part)
    r = subprocess.run([exe, inp], capture_output=True, text=True)
    assert r.returncode == 1 and "no CPU fallback" in r.stderr, (r.returncode, r.stderr)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["small", "two_reticulations", "celine"])
def test_cpp_caller_matches_oracle(tmp_path, name):
    from oracle import oracle
    exe = _build(tmp_path)
    net, part = load_fixture(*FIXTURE_PAIRS[name])
    inp = str(tmp_path / "in.txt")
    _write_input(inp, net, part)
    r = subprocess.run([exe, inp], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = {k: float(v) for k, v in (line.split() for line in r.stdout.strip().splitlines())}
    o = oracle.make_engine("ref" if oracle.have_ref() else "port", net, [part])
    lo = o.computeLoglikelihood(0, 1)
    assert out["logl_full"] == pytest.approx(lo, rel=1e-9)            # own Jacobi eigen-decomposition: 1e-10-level agreement
    assert out["logl_incremental"] == out["logl_full"]
    assert out["reroot_max_abs_diff"] <= 1e-8 * abs(lo)
    assert out["likelihood_target_function"] == pytest.approx(out["logl_after_probs"], rel=1e-12)
    o.brlen_prepare(0); o.computeLoglikelihoodBrlenOpt(0); o.computePartitionSumtables(0)
    do = o.computeLoglikelihoodDerivatives(0)
    o.brlen_finish(0)
    assert out["edge0_logl_prime"] == pytest.approx(do[0], rel=1e-7, abs=1e-6)
    assert out["edge0_logl_prime_prime"] == pytest.approx(do[1], rel=1e-7, abs=1e-6)
    o.set_scoring_sizes(0)
    assert out["bic_before"] == pytest.approx(o.scoreNetwork(), rel=1e-9)
    lb = o.optimize_branches()
    lr = o.optimize_reticulations()
    assert out["logl_after_brlen"] >= out["logl_full"] - 1e-3
    assert out["logl_after_brlen"] == pytest.approx(lb, rel=1e-7)
    assert out["logl_after_probs"] == pytest.approx(lr, rel=1e-7)
    assert out["bic_after"] <= out["bic_before"] + 1e-3
