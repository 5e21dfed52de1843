"""ctypes binder for a flat "network likelihood" C-ABI.

The product's host library (libnetrax_b200.so, prefix ``nrxh_``, declared in include/netrax_b200.h)
exports the NetRAX likelihood API (computeLoglikelihood, computeLoglikelihoodBrlenOpt,
computePartitionSumtables, computeLoglikelihoodDerivatives, ...) as plain-C entry points.  The test
oracle exports the same function set under the prefix ``orc_`` so parity tests can run identical call
sequences against both; this module knows nothing about either implementation.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence, Tuple

import numpy as np

from .network_io import NetworkDesc

AVERAGE, BEST, SARAH_PSEUDO = 0, 1, 2  # LikelihoodVariant (src/likelihood/LikelihoodVariant.hpp)
BRENT_NORMAL, BRENT_REROOT, NEWTON_RAPHSON = 0, 1, 2   # BrlenOptMethod (src/NetraxOptions.hpp:17-21)
LINKED, SCALED, UNLINKED = 0, 1, 2   # PLLMOD_COMMON_BRLEN_*

_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")


class LikelihoodError(RuntimeError):
    """Mirrors the std::runtime_error the reference throws from its likelihood layer."""


def _as(a, dt):
    return np.ascontiguousarray(np.asarray(a, dtype=dt))


class FlatAPI:
    def __init__(self, lib: C.CDLL, prefix: str):
        self.lib, self.prefix = lib, prefix
        g = self._fn
        g("last_error", C.c_char_p)
        g("new", C.c_void_p, C.c_char_p)
        g("free", None, C.c_void_p)
        g("set_network", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, C.c_uint, _u32p, _u32p, _f64p, _f64p,
          C.c_uint, _u32p, _u32p, _u32p)
        g("add_partition", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, _u32p, C.c_void_p, _f64p, _f64p, _f64p, _f64p)
        g("set_options", C.c_int, C.c_void_p, C.c_int, C.c_int)
        g("set_partition_brlens", C.c_int, C.c_void_p, C.c_uint, _f64p)
        g("init", C.c_int, C.c_void_p)
        g("compute_loglikelihood", C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double))
        g("num_partitions", C.c_uint, C.c_void_p)
        g("root", C.c_uint, C.c_void_p)
        g("num_nodes", C.c_uint, C.c_void_p)
        g("num_trees", C.c_int, C.c_void_p, C.c_uint)
        g("tree_config", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_char_p, C.c_uint)
        g("tree_info", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.POINTER(C.c_double), C.c_void_p, C.POINTER(C.c_int))
        g("read_clv", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, _f64p)
        g("read_scaler", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_uint, _u32p)
        g("partition_loglh", C.c_int, C.c_void_p, _f64p)
        g("set_branch_length", C.c_int, C.c_void_p, C.c_int, C.c_uint, C.c_double)
        g("set_reticulation_prob", C.c_int, C.c_void_p, C.c_uint, C.c_double)
        g("set_model", C.c_int, C.c_void_p, C.c_uint, _f64p, _f64p, _f64p, _f64p)
        g("get_eigen", C.c_int, C.c_void_p, C.c_uint, _f64p, _f64p, _f64p)
        g("get_pmatrix", C.c_int, C.c_void_p, C.c_uint, C.c_uint, _f64p)
        g("brlen_prepare", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double))
        g("brlen_logl", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double))
        g("brlen_sumtables", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_uint))
        g("brlen_read_sumtable", C.c_int, C.c_void_p, C.c_uint, C.c_uint, C.c_void_p, C.POINTER(C.c_double),
          C.POINTER(C.c_uint), C.POINTER(C.c_uint))
        g("brlen_set_length", C.c_int, C.c_void_p, C.c_int, C.c_uint, C.c_double)
        g("brlen_derivatives", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_double),
          C.c_void_p, C.c_void_p, C.c_void_p)
        g("brlen_finish", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double))
        if self.has("brlen_sweep_order"):   # product only: the checker re-roots in place, as the reference does
            g("brlen_sweep_order", C.c_int, C.c_void_p, C.c_void_p)
            g("reroot_stats", C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong), C.POINTER(C.c_uint), C.POINTER(C.c_uint))
            g("set_reroot_cache_slots", C.c_int, C.c_void_p, C.c_longlong)
            g("set_score_only", C.c_int, C.c_void_p, C.c_int)
            g("brlen_logl_sumtables", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double), C.POINTER(C.c_uint))
            g("set_lazy_rerooting", C.c_int, C.c_void_p, C.c_int)
            g("lazy_reroot_stats", C.c_int, C.c_void_p, C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong))
        g("optimize_branch", C.c_int, C.c_void_p, C.c_uint, C.c_int, C.c_uint, C.POINTER(C.c_double))
        g("optimize_branches", C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double))
        g("optimize_reticulation", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double))
        g("optimize_reticulations", C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double))
        g("compute_pseudo_loglikelihood", C.c_int, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_double))
        g("read_pseudo_clv", C.c_int, C.c_void_p, C.c_uint, C.c_uint, _f64p)
        g("read_pseudo_scaler", C.c_int, C.c_void_p, C.c_uint, C.c_uint, np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS"))
        g("score_network", C.c_int, C.c_void_p, C.POINTER(C.c_double))
        g("set_scoring_sizes", C.c_int, C.c_void_p, C.c_ulonglong, C.c_ulonglong)
        g("optimize_all_non_topology", C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_double))
        g("set_pinv", C.c_int, C.c_void_p, C.c_uint, C.c_double)
        g("set_brlen_scaler", C.c_int, C.c_void_p, C.c_uint, C.c_double)
        g("set_params_to_optimize", C.c_int, C.c_void_p, C.c_uint, C.c_int)
        g("set_submodels", C.c_int, C.c_void_p, C.c_uint, C.c_uint, _u32p, _f64p, _f64p)
        g("set_alpha", C.c_int, C.c_void_p, C.c_uint, C.c_double)
        g("get_alpha", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double))
        g("optimize_alpha", C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double))
        g("optimize_pinv", C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_double, C.POINTER(C.c_double))
        g("get_pinv", C.c_int, C.c_void_p, C.c_uint, C.POINTER(C.c_double))
        g("optimize_scalers", C.c_int, C.c_void_p, C.POINTER(C.c_double))
        g("get_brlen_scalers", C.c_int, C.c_void_p, _f64p)
        g("get_branch_lengths", C.c_int, C.c_void_p, C.c_int, _f64p)
        g("get_reticulation_probs", C.c_int, C.c_void_p, _f64p)
        g("clv_update_count", C.c_ulonglong, C.c_void_p)
        g("reset_counters", None, C.c_void_p)
        g("gamma_rates", C.c_int, C.c_double, C.c_uint, C.c_int, _f64p)

    def _fn(self, name, restype, *argtypes):
        f = getattr(self.lib, self.prefix + name)
        f.restype, f.argtypes = restype, list(argtypes)
        setattr(self, "_" + name, f)
        return f

    def has(self, name: str) -> bool:
        return hasattr(self.lib, self.prefix + name)

    def check(self, ok):
        if not ok:
            raise LikelihoodError(self._last_error().decode())

    def gamma_rates(self, alpha: float, cats: int, mode: int = 0) -> np.ndarray:
        out = np.zeros(cats)
        self.check(self._gamma_rates(alpha, cats, mode, out))
        return out


class Partition:
    """Inputs of one alignment partition: what create_pll_partition (src/RaxmlWrapper.cpp:686-760) sets."""

    def __init__(self, states: int, rate_cats: int, tip_masks: np.ndarray, freqs, subst, rates, rate_weights=None,
                 pattern_weights=None):
        self.states, self.rate_cats = int(states), int(rate_cats)
        self.tip_masks = _as(tip_masks, np.uint32)          # [tips, patterns]
        self.sites = int(self.tip_masks.shape[1])
        self.freqs, self.subst = _as(freqs, np.float64), _as(subst, np.float64)
        self.rates = _as(rates, np.float64)
        self.rate_weights = _as(rate_weights if rate_weights is not None else np.full(rate_cats, 1.0 / rate_cats), np.float64)
        self.pattern_weights = None if pattern_weights is None else _as(pattern_weights, np.uint32)

    def slice(self, lo: int, hi: int) -> "Partition":
        """Contiguous pattern range [lo, hi): one rank's share under site sharding."""
        pw = None if self.pattern_weights is None else self.pattern_weights[lo:hi]
        return Partition(self.states, self.rate_cats, self.tip_masks[:, lo:hi], self.freqs, self.subst, self.rates,
                         self.rate_weights, pw)


class LikelihoodEngine:
    """One AnnotatedNetwork + its likelihood state behind a FlatAPI (product or oracle)."""

    def __init__(self, api: FlatAPI, net: NetworkDesc, partitions: Sequence[Partition], variant: int = AVERAGE,
                 linkage: int = LINKED, backend: str = "", partition_brlens: Optional[Sequence[np.ndarray]] = None):
        self.api, self.net, self.partitions = api, net, list(partitions)
        self.variant, self.linkage = variant, linkage
        self.h = api._new(backend.encode())
        if not self.h:
            raise LikelihoodError(api._last_error().decode())
        a = api
        a.check(a._set_network(self.h, net.num_tips, net.num_nodes, net.root, net.num_edges, _as(net.edge_source, np.uint32),
                               _as(net.edge_target, np.uint32), _as(net.edge_length, np.float64), _as(net.edge_prob, np.float64),
                               net.num_reticulations, _as(net.ret_node, np.uint32), _as(net.ret_first_edge, np.uint32),
                               _as(net.ret_second_edge, np.uint32)))
        for p in self.partitions:
            pw = None if p.pattern_weights is None else p.pattern_weights.ctypes.data_as(C.c_void_p)
            a.check(a._add_partition(self.h, p.states, p.rate_cats, p.sites, p.tip_masks.reshape(-1), pw, p.freqs, p.subst,
                                     p.rates, p.rate_weights))
        a.check(a._set_options(self.h, variant, linkage))
        if partition_brlens is not None:
            for i, b in enumerate(partition_brlens):
                a.check(a._set_partition_brlens(self.h, i, _as(b, np.float64)))
        a.check(a._init(self.h))
        self.P = len(self.partitions)

    def set_reduce(self, reduce):
        """Install the reference's parallel_reduce_cb (src/RaxmlWrapper.cpp:717-718): `reduce(arr)` must SUM the float64
        numpy array `arr` in place across all site shards (ranks).  All ranks then issue identical call sequences."""
        cb_t = C.CFUNCTYPE(None, C.c_void_p, C.POINTER(C.c_double), C.c_size_t, C.c_int)

        def _cb(ctx, data, count, op):
            reduce(np.ctypeslib.as_array(data, shape=(count,)))

        self._reduce_cb = cb_t(_cb)   # keep alive
        fn = getattr(self.api.lib, self.api.prefix + "set_reduce_callback")
        fn.restype = C.c_int
        fn.argtypes = [C.c_void_p, cb_t, C.c_void_p]
        self.api.check(fn(self.h, self._reduce_cb, None))

    def close(self):
        if getattr(self, "h", None):
            self.api._free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- src/likelihood/LikelihoodComputation.hpp:17-18 ----
    def computeLoglikelihood(self, incremental: int = 1, update_pmatrices: int = 1) -> float:
        out = C.c_double()
        self.api.check(self.api._compute_loglikelihood(self.h, incremental, update_pmatrices, C.byref(out)))
        return out.value

    def partition_loglh(self) -> np.ndarray:
        out = np.zeros(self.P)
        self.api.check(self.api._partition_loglh(self.h, out))
        return out

    def num_trees(self, node: int) -> int:
        return self.api._num_trees(self.h, node)

    def tree_config(self, node: int, tree: int) -> str:
        buf = C.create_string_buffer(1 << 16)
        self.api.check(self.api._tree_config(self.h, node, tree, buf, len(buf)))
        return buf.value.decode()

    def tree_info(self, node: int, tree: int) -> Tuple[float, np.ndarray, int]:
        lp, fl = C.c_double(), C.c_int()
        pl = np.zeros(self.P)
        self.api.check(self.api._tree_info(self.h, node, tree, C.byref(lp), pl.ctypes.data_as(C.c_void_p), C.byref(fl)))
        return lp.value, pl, fl.value

    def clv_entries(self, p: int) -> int:
        part = self.partitions[p]
        return part.sites * part.rate_cats * ((part.states + 3) & ~3)

    def read_clv(self, node: int, tree: int, p: int = 0) -> np.ndarray:
        out = np.zeros(self.clv_entries(p))
        self.api.check(self.api._read_clv(self.h, node, tree, p, out))
        return out

    def read_scaler(self, node: int, tree: int, p: int = 0) -> np.ndarray:
        out = np.zeros(self.partitions[p].sites, np.uint32)
        self.api.check(self.api._read_scaler(self.h, node, tree, p, out))
        return out

    def set_branch_length(self, edge: int, value: float, partition: int = -1):
        self.api.check(self.api._set_branch_length(self.h, partition, edge, value))

    def set_reticulation_prob(self, r: int, prob: float):
        self.api.check(self.api._set_reticulation_prob(self.h, r, prob))

    def set_model(self, p: int, freqs, subst, rates, rate_weights):
        self.api.check(self.api._set_model(self.h, p, _as(freqs, np.float64), _as(subst, np.float64), _as(rates, np.float64),
                                           _as(rate_weights, np.float64)))

    def get_eigen(self, p: int = 0):
        part = self.partitions[p]
        sp = (part.states + 3) & ~3
        ev, iev, evals = np.zeros(part.states * sp), np.zeros(part.states * sp), np.zeros(sp)
        self.api.check(self.api._get_eigen(self.h, p, ev, iev, evals))
        return ev, iev, evals

    def get_pmatrix(self, edge: int, p: int = 0) -> np.ndarray:
        part = self.partitions[p]
        out = np.zeros(part.rate_cats * part.states * ((part.states + 3) & ~3))
        self.api.check(self.api._get_pmatrix(self.h, p, edge, out))
        return out

    # ---- branch-length optimisation flow (src/optimization/BranchLengthOptimization.cpp:345-420) ----
    def brlen_prepare(self, edge: int) -> float:
        out = C.c_double()
        self.api.check(self.api._brlen_prepare(self.h, edge, C.byref(out)))
        return out.value

    def computeLoglikelihoodBrlenOpt(self, edge: int) -> float:
        out = C.c_double()
        self.api.check(self.api._brlen_logl(self.h, edge, C.byref(out)))
        return out.value

    def computePartitionSumtables(self, edge: int) -> int:
        n = C.c_uint()
        self.api.check(self.api._brlen_sumtables(self.h, edge, C.byref(n)))
        self._n_sumtables = n.value
        return n.value

    def computeLoglikelihoodBrlenOptAndSumtables(self, edge: int):
        """(edge-rooted lnL, number of sumtables): both from ONE pass over the displayed-tree pairs' CLVs."""
        out, n = C.c_double(), C.c_uint()
        self.api.check(self.api._brlen_logl_sumtables(self.h, edge, C.byref(out), C.byref(n)))
        self._n_sumtables = n.value
        return out.value, n.value

    def read_sumtable(self, p: int, idx: int):
        out = np.zeros(self.clv_entries(p))
        prob, lt, rt = C.c_double(), C.c_uint(), C.c_uint()
        self.api.check(self.api._brlen_read_sumtable(self.h, p, idx, out.ctypes.data_as(C.c_void_p), C.byref(prob), C.byref(lt), C.byref(rt)))
        return out, prob.value, lt.value, rt.value

    def brlen_set_length(self, edge: int, value: float, partition: int = -1):
        self.api.check(self.api._brlen_set_length(self.h, partition, edge, value))

    def computeLoglikelihoodDerivatives(self, edge: int):
        d1, d2 = C.c_double(), C.c_double()
        pd1, pd2 = np.zeros(self.P), np.zeros(self.P)
        raw = np.zeros(self.P * 3 * max(1, getattr(self, "_n_sumtables", 1)))
        self.api.check(self.api._brlen_derivatives(self.h, edge, C.byref(d1), C.byref(d2), pd1.ctypes.data_as(C.c_void_p),
                                                   pd2.ctypes.data_as(C.c_void_p), raw.ctypes.data_as(C.c_void_p)))
        return d1.value, d2.value, pd1, pd2, raw.reshape(self.P, -1, 3)

    def brlen_finish(self, edge: int) -> float:
        out = C.c_double()
        self.api.check(self.api._brlen_finish(self.h, edge, C.byref(out)))
        return out.value

    def brlen_sweep_order(self) -> np.ndarray:
        """All branches in depth-first pre-order: consecutive branches share their re-rooting paths (memoised on the device)."""
        out = np.zeros(self.net.num_edges, dtype=np.uint32)
        self.api.check(self.api._brlen_sweep_order(self.h, out.ctypes.data_as(C.c_void_p)))
        return out

    def reroot_stats(self) -> dict:
        h, m, n, s = C.c_ulonglong(), C.c_ulonglong(), C.c_uint(), C.c_uint()
        self.api.check(self.api._reroot_stats(self.h, C.byref(h), C.byref(m), C.byref(n), C.byref(s)))
        return {"hits": h.value, "misses": m.value, "entries": n.value, "cached_slots": s.value}

    def set_score_only(self, on: bool = True):
        """Full evaluations replaying the cached plan skip the CLV stores of the root displayed trees (candidate scoring)."""
        self.api.check(self.api._set_score_only(self.h, 1 if on else 0))

    def set_lazy_rerooting(self, on: bool = True):
        """brlen_prepare / brlen_finish / optimize_branch without the evaluations from the network root around every branch."""
        self.api.check(self.api._set_lazy_rerooting(self.h, 1 if on else 0))

    def lazy_reroot_stats(self) -> dict:
        a, b = C.c_ulonglong(), C.c_ulonglong()
        self.api.check(self.api._lazy_reroot_stats(self.h, C.byref(a), C.byref(b)))
        return {"sessions": a.value, "fallbacks": b.value}

    def set_reroot_cache_slots(self, max_slots: int = -1):
        self.api.check(self.api._set_reroot_cache_slots(self.h, max_slots))

    # ---- the immediate callers (src/optimization/BranchLengthOptimization.cpp, ReticulationOptimization.cpp) ----
    def optimize_branch(self, edge: int, method: int = NEWTON_RAPHSON, max_iters: int = 32) -> float:
        out = C.c_double()
        self.api.check(self.api._optimize_branch(self.h, edge, method, max_iters, C.byref(out)))
        return out.value

    def optimize_branches(self, max_iters: int = 32, max_iters_outside: int = 32, radius: int = -1, method: int = NEWTON_RAPHSON) -> float:
        """optimizeBranches (src/optimization/Optimization.cpp:17-38): max_iters = brlen_smooth_factor * RAXML_BRLEN_SMOOTHINGS (32),
        radius = PLLMOD_OPT_BRLEN_OPTIMIZE_ALL (-1)."""
        out = C.c_double()
        self.api.check(self.api._optimize_branches(self.h, max_iters, max_iters_outside, radius, method, C.byref(out)))
        return out.value

    def optimize_reticulation(self, r: int) -> float:
        out = C.c_double()
        self.api.check(self.api._optimize_reticulation(self.h, r, C.byref(out)))
        return out.value

    def computePseudoLoglikelihood(self, incremental: int = 1, update_pmatrices: int = 1) -> float:
        """src/likelihood/PseudoLoglikelihood.cpp:57-226: one merged CLV per node, weights = reticulation probabilities."""
        out = C.c_double()
        self.api.check(self.api._compute_pseudo_loglikelihood(self.h, incremental, update_pmatrices, C.byref(out)))
        return out.value

    def read_pseudo_clv(self, node: int, p: int = 0) -> np.ndarray:
        part = self.partitions[p]
        out = np.zeros(part.sites * part.rate_cats * ((part.states + 3) & ~3))
        self.api.check(self.api._read_pseudo_clv(self.h, node, p, out))
        return out

    def read_pseudo_scaler(self, node: int, p: int = 0) -> np.ndarray:
        out = np.zeros(self.partitions[p].sites, dtype=np.uint32)
        self.api.check(self.api._read_pseudo_scaler(self.h, node, p, out))
        return out

    def scoreNetwork(self) -> float:
        """BIC of the network (src/likelihood/ComplexityScoring.cpp:57-67); smaller is better."""
        out = C.c_double()
        self.api.check(self.api._score_network(self.h, C.byref(out)))
        return out.value

    def set_scoring_sizes(self, total_num_model_parameters: int, total_num_sites: int = 0):
        self.api.check(self.api._set_scoring_sizes(self.h, total_num_model_parameters, total_num_sites))

    def optimizeAllNonTopology(self, type: int = 1) -> float:
        """src/optimization/Optimization.cpp:118-214 (0 QUICK, 1 NORMAL, 2 SLOW); returns the final BIC."""
        out = C.c_double()
        self.api.check(self.api._optimize_all_non_topology(self.h, type, C.byref(out)))
        return out.value

    def set_pinv(self, p: int, prop_invar: float):
        """+I: proportion of invariant sites (pll_update_invariant_sites_proportion)."""
        self.api.check(self.api._set_pinv(self.h, p, prop_invar))

    def set_params_to_optimize(self, p: int, alpha: bool, pinv: bool):
        """pllmod_treeinfo_t::params_to_optimize of partition p for optimize_alpha / optimize_pinv (instead of "the value is > 0")."""
        self.api.check(self.api._set_params_to_optimize(self.h, p, (1 if alpha else 0) | (2 if pinv else 0)))

    def set_brlen_scaler(self, p: int, scaler: float):
        """pllmod_treeinfo_t::brlen_scalers[p] under scaled branch-length linkage (linkage = SCALED)."""
        self.api.check(self.api._set_brlen_scaler(self.h, p, scaler))

    def set_submodels(self, p: int, ratecat_submodels, freqs, subst):
        """One rate matrix per rate category (LG4M / LG4X; raxml-ng Model::ratecat_submodels -> libpll params_indices):
        freqs [n][states], subst [n][states (states - 1) / 2], category c uses matrix ratecat_submodels[c]."""
        freqs, subst = _as(np.atleast_2d(freqs), np.float64), _as(np.atleast_2d(subst), np.float64)
        self.api.check(self.api._set_submodels(self.h, p, freqs.shape[0], _as(ratecat_submodels, np.uint32), freqs, subst))

    def set_alpha(self, p: int, alpha: float):
        """treeinfo_set_alpha: Gamma shape -> discrete rates of partition p (mean mode)."""
        self.api.check(self.api._set_alpha(self.h, p, alpha))

    def get_alpha(self, p: int) -> float:
        out = C.c_double()
        self.api.check(self.api._get_alpha(self.h, p, C.byref(out)))
        return out.value

    def optimize_alpha(self, min_alpha: float = 0.0201, max_alpha: float = 100.0, tolerance: float = 0.001) -> float:
        """The ALPHA step of optimize_params (ModelOptimization.cpp:56-65): Brent over all partitions' Gamma shapes."""
        out = C.c_double()
        self.api.check(self.api._optimize_alpha(self.h, min_alpha, max_alpha, tolerance, C.byref(out)))
        return out.value

    def optimize_pinv(self, min_pinv: float = 0.0, max_pinv: float = 0.99, tolerance: float = 0.001) -> float:
        """The PINV step of optimize_params (ModelOptimization.cpp:67-76): Brent over the +I partitions' proportions."""
        out = C.c_double()
        self.api.check(self.api._optimize_pinv(self.h, min_pinv, max_pinv, tolerance, C.byref(out)))
        return out.value

    def get_pinv(self, p: int) -> float:
        out = C.c_double()
        self.api.check(self.api._get_pinv(self.h, p, C.byref(out)))
        return out.value

    def optimize_scalers(self) -> float:
        """optimize_scalers (BranchLengthOptimization.cpp:581-599): pllmod_algo_opt_brlen_scalers_treeinfo under scaled
        linkage with several partitions, otherwise a no-op; returns the BIC."""
        out = C.c_double()
        self.api.check(self.api._optimize_scalers(self.h, C.byref(out)))
        return out.value

    def brlen_scalers(self) -> np.ndarray:
        out = np.zeros(self.P)
        self.api.check(self.api._get_brlen_scalers(self.h, out))
        return out

    def optimize_reticulations(self, max_iters: int = 10) -> float:
        out = C.c_double()
        self.api.check(self.api._optimize_reticulations(self.h, max_iters, C.byref(out)))
        return out.value

    def branch_lengths(self, partition: int = -1) -> np.ndarray:
        out = np.zeros(self.net.num_edges)
        self.api.check(self.api._get_branch_lengths(self.h, partition, out))
        return out

    def reticulation_probs(self) -> np.ndarray:
        out = np.zeros(max(1, self.net.num_reticulations))
        self.api.check(self.api._get_reticulation_probs(self.h, out))
        return out[: self.net.num_reticulations]

    def toExtendedNewick(self, precision: Optional[int] = None, average_unlinked: bool = True) -> str:
        """toExtendedNewick(AnnotatedNetwork&) (src/io/NetworkIO.cpp:517-523): updateNetwork (:493-508) copies the linked
        branch lengths and the reticulation probabilities of the current (optimised) state into the network, then writes it.
        With unlinked branch lengths the search first replaces the linked lengths by the partitions' average weighted with
        their share of the alignment (collect_average_branches, :455-491, called from src/search/ScoreImprovement.cpp:38);
        ``average_unlinked`` does the same here."""
        from .network_io import to_extended_newick
        lengths = self.branch_lengths()
        if self.linkage == UNLINKED and average_unlinked:
            w = np.array([float(p.pattern_weights.sum()) if p.pattern_weights is not None else float(p.sites) for p in self.partitions])
            lengths = sum(self.branch_lengths(p) * (w[p] / w.sum()) for p in range(self.P))
        return to_extended_newick(self.net, lengths, self.reticulation_probs(), precision)

    def clv_update_count(self) -> int:
        return int(self.api._clv_update_count(self.h))

    def reset_counters(self):
        self.api._reset_counters(self.h)
