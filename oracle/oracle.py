"""ORACLE loader (test infrastructure, NOT product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import
this.  Two flavours of the same restated NetRAX layer:
  kind "port"      — oracle/liboracle_port.so, arithmetic from our scalar restatement (pll_port.c)
  kind "reference" — oracle/_ref/liboracle_ref.so, arithmetic from the reference's real forked libpll
                     (compiled from /root/reference by oracle/Makefile; prebuilt file travels to the GPU box)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import Optional

import numpy as np

from netrax_b200._capi import FlatAPI, LikelihoodEngine  # generic ctypes glue only

HERE = os.path.dirname(os.path.abspath(__file__))
PORT_SO = os.path.join(HERE, "liboracle_port.so")
REF_SO = os.path.join(HERE, "_ref", "liboracle_ref.so")


def build(verbose: bool = False) -> None:
    """Compile the C/C++ restatement and, when /root/reference is present, oracle/_ref."""
    r = subprocess.run(["make", "-C", HERE, "all"], capture_output=True, text=True)
    if verbose or r.returncode:
        print(r.stdout[-4000:], r.stderr[-4000:])
    if r.returncode:
        raise RuntimeError("oracle build failed")


_apis = {}


def api(kind: str = "port") -> FlatAPI:
    if kind not in _apis:
        path = PORT_SO if kind == "port" else REF_SO
        if not os.path.exists(path):
            build()
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        lib = C.CDLL(path)
        a = FlatAPI(lib, "orc_")
        lib.orc_naive_loglikelihood.restype = C.c_int
        lib.orc_naive_loglikelihood.argtypes = [C.c_void_p, C.POINTER(C.c_double), C.c_void_p, C.c_void_p]
        _apis[kind] = a
    return _apis[kind]


def have_ref() -> bool:
    return os.path.exists(REF_SO)


def make_engine(kind, net, partitions, **kw) -> LikelihoodEngine:
    """kind 'port' -> scalar restatement; 'ref' -> real libpll underneath."""
    a = api("port" if kind == "port" else "ref")
    return LikelihoodEngine(a, net, partitions, backend=("port" if kind == "port" else "ref"), **kw)


def naive_loglikelihood(eng: LikelihoodEngine):
    """Per-displayed-tree evaluation (role of NaiveLoglikelihood.cpp): (lnL, tree_logl[T,P], tree_logprob[T])."""
    T = 1 << eng.net.num_reticulations
    tl = np.zeros(T * eng.P)
    lp = np.zeros(T)
    out = C.c_double()
    eng.api.check(eng.api.lib.orc_naive_loglikelihood(eng.h, C.byref(out), tl.ctypes.data_as(C.c_void_p), lp.ctypes.data_as(C.c_void_p)))
    return out.value, tl.reshape(T, eng.P), lp


def persite_lnl(eng: LikelihoodEngine, tree: int) -> np.ndarray:
    """Per-site lnL of root displayed tree `tree`, [P, max sites] — what pll_compute_root_loglikelihood writes into the
    `persite_lnl` array of LH/ImprovedLoglikelihood.cpp:448-453 (pattern weight applied)."""
    fn = eng.api.lib.orc_persite_lnl
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint, np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS"), C.c_uint]
    stride = max(p.sites for p in eng.partitions)
    out = np.zeros((eng.P, stride))
    eng.api.check(fn(eng.h, tree, out.reshape(-1), stride))
    return out
