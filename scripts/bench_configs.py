#!/usr/bin/env python
"""Measure every BASELINE.json config on one B200 next to the CPU reference arm (real libpll AVX2 under the restated
NetRAX driver, all host cores, bounded sample).  bench.py stays the driver's contract (config 5); this script fills
profiles/ with the other configs' numbers.  Usage: python scripts/bench_configs.py [--configs 1,2,3] [--out file.json]

Per config it reports: full network lnL evaluations/s + CLV site-updates/s, and for config 2 the branch-length
derivative sweep (for EVERY edge: virtual re-rooting, edge-rooted lnL, sumtables, 3 Newton-iterate derivative
evaluations, restore) in sweeps/s and edges/s."""
import argparse
import json
import multiprocessing as mp
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from netrax_b200._capi import UNLINKED  # noqa: E402


from bench import derivative_sweep  # noqa: E402,F401  (the sweep definition lives with the driver-run bench)


def optimisation_round(eng):
    """One optimizeBranches + optimizeReticulationProbs round as optimizeAllNonTopology issues them
    (src/optimization/Optimization.cpp:17-38,88-106): Newton-Raphson over every branch (max_iters = 32), then Brent over
    every reticulation probability (<= 10 rounds).  Returns wall seconds and the lnL trajectory."""
    l0 = eng.computeLoglikelihood(0, 1)
    t = time.perf_counter(); l1 = eng.optimize_branches(); tb = time.perf_counter() - t
    t = time.perf_counter(); l2 = eng.optimize_reticulations(); tr = time.perf_counter() - t
    out = {"brlen_round_s": tb, "retprob_round_s": tr, "lnl_start": l0, "lnl_after_brlen": l1, "lnl_after_retprob": l2}
    if hasattr(eng, "optimize_alpha"):  # the ALPHA step of optimize_params: start from a wrong shape (2.0)
        for p in range(len(eng.partitions)):
            eng.set_alpha(p, 2.0)
        t = time.perf_counter(); l3 = eng.optimize_alpha(); ta = time.perf_counter() - t
        out.update({"alpha_opt_s": ta, "lnl_after_alpha": l3, "alpha": eng.get_alpha(0)})
    return out


def _cpu_worker(args):
    kind, cfg, patterns, widx, reps, sweep = args
    from oracle import oracle
    net, parts, brl = bench.make_inputs(cfg, widx * patterns, (widx + 1) * patterns)
    eng = oracle.make_engine(kind, net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    eng.computeLoglikelihood(0, 1)
    out = {"lnl": [], "sweep": []}
    eng.reset_counters()
    for _ in range(reps):
        t = time.perf_counter(); eng.computeLoglikelihood(0, 1); out["lnl"].append(time.perf_counter() - t)
    out["updates"] = eng.clv_update_count() // reps
    if sweep:
        t = time.perf_counter(); derivative_sweep(eng, net); out["sweep"].append(time.perf_counter() - t)
        out["opt"] = optimisation_round(eng)
    return out


def cpu_arm(cfg, cores, patterns_per_core, reps, sweep):
    from oracle import oracle
    kind = "ref" if oracle.have_ref() else "port"
    with mp.get_context("fork").Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, cfg, patterns_per_core, i, reps, sweep) for i in range(cores)])
    lnl_t = float(np.median([max(r["lnl"][k] for r in res) for k in range(reps)]))
    out = {"kind": "reference" if kind == "ref" else "port", "cores": cores, "patterns_per_core": patterns_per_core,
           "lnl_eval_s_on_sample": lnl_t, "site_updates_per_s": sum(r["updates"] for r in res) / lnl_t}
    if sweep:
        out["sweep_s_on_sample"] = max(r["sweep"][0] for r in res)
        out["brlen_round_s_on_sample"] = max(r["opt"]["brlen_round_s"] for r in res)
        out["retprob_round_s_on_sample"] = max(r["opt"]["retprob_round_s"] for r in res)
        out["alpha_opt_s_on_sample"] = max(r["opt"].get("alpha_opt_s", 0.0) for r in res)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="1,2,3")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--out", default="")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--patterns", type=int, default=0, help="override the config's pattern count")
    args = ap.parse_args()
    from netrax_b200.engine import NetraxB200
    results = {}
    for c in [int(x) for x in args.configs.split(",")]:
        cfg = dict(bench.CONFIGS[c])
        if args.patterns:
            cfg["patterns"] = args.patterns
            cfg["name"] += f" [patterns overridden: {args.patterns}]"
        net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
        eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
        for _ in range(3):
            eng.computeLoglikelihood(0, 1)
        slots = sum(eng.num_trees(v) for v in range(net.num_tips, net.num_nodes))
        updates = slots * sum(p.sites for p in parts)
        l0 = eng.launch_count()
        eng.profile_enable(True)
        eng.timer_start()
        for _ in range(args.reps):
            eng.computeLoglikelihood(0, 1)
        ms = eng.timer_stop() / args.reps
        prof = eng.profile_read()
        eng.profile_enable(False)
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
        k2_gbs = prof["clv_bytes"] / (prof["clv_ms"] / 1e3) / 1e9
        r = {"workload": cfg["name"], "sum_trees_per_node": slots, "root_trees": eng.num_trees(net.root),
             "gpu": {"ms_per_lnl_eval": ms, "lnl_evals_per_s": 1e3 / ms, "site_updates_per_s": updates / (ms / 1e3),
                     "launches_per_eval": (eng.launch_count() - l0) / args.reps,
                     "k2_roofline": {"achieved_GBps": k2_gbs, "peak_GBps": peak, "frac": k2_gbs / peak, "share_of_eval": prof["clv_ms"] / (ms * args.reps),
                                     "site_updates_per_s_in_kernel": prof["clv_site_updates"] / (prof["clv_ms"] / 1e3)}}}
        sweep = (c == 2)
        if sweep:
            derivative_sweep(eng, net)  # warm-up (allocates re-rooting slots and sumtables)
            l0 = eng.launch_count()
            t = time.perf_counter()
            eng.timer_start()
            derivative_sweep(eng, net)
            ms_s = eng.timer_stop()
            wall = time.perf_counter() - t
            r["gpu"].update({"ms_per_derivative_sweep": ms_s, "wall_ms_per_derivative_sweep": 1e3 * wall, "edges": net.num_edges,
                             "edges_per_s": net.num_edges / (ms_s / 1e3), "launches_per_sweep": eng.launch_count() - l0})
            l0 = eng.launch_count()
            r["gpu"]["optimisation_round"] = optimisation_round(eng)
            r["gpu"]["optimisation_round"]["launches"] = eng.launch_count() - l0
        eng.close()
        if not args.no_cpu:
            cores = bench.host_cores()
            ppc = min(4000 if sweep else 16000, max(64, -(-cfg["patterns"] // cores)))
            cpu = cpu_arm(cfg, cores, ppc, 3, sweep)
            scale = cfg["patterns"] / (ppc * cores)   # the sample covers ppc*cores patterns of the config's total
            cpu["lnl_evals_per_s_full_config_est"] = 1.0 / (cpu["lnl_eval_s_on_sample"] * scale)
            r["cpu"] = cpu
            r["speedup_site_updates"] = r["gpu"]["site_updates_per_s"] / cpu["site_updates_per_s"]
            if sweep:
                cpu["sweep_s_full_config_est"] = cpu["sweep_s_on_sample"] * scale
                cpu["brlen_round_s_full_config_est"] = cpu["brlen_round_s_on_sample"] * scale
                cpu["alpha_opt_s_full_config_est"] = cpu["alpha_opt_s_on_sample"] * scale
                if r["gpu"]["optimisation_round"].get("alpha_opt_s"):
                    r["speedup_alpha_opt"] = cpu["alpha_opt_s_full_config_est"] / r["gpu"]["optimisation_round"]["alpha_opt_s"]
                r["speedup_brlen_round"] = cpu["brlen_round_s_full_config_est"] / r["gpu"]["optimisation_round"]["brlen_round_s"]
                r["speedup_derivative_sweep"] = cpu["sweep_s_full_config_est"] / (r["gpu"]["ms_per_derivative_sweep"] / 1e3)
        results[f"config{c}"] = r
        print(json.dumps({f"config{c}": r}), flush=True)
    if args.out:
        json.dump(results, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
