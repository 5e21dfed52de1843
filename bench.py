#!/usr/bin/env python
"""bench.py — network-likelihood hot path on N B200s vs the reference's CPU path on the box's host cores.

A "step" = ONE full network lnL evaluation, computeLoglikelihood(ann, incremental=0, update_pmatrices=1):
P-matrices for every edge + every CLV of every displayed tree at every node + per-tree root lnL + the
cross-rank reduction + AVERAGE/BEST mixing (BASELINE.md §4 "What is timed").

Workload (headline): BASELINE.json configs[4], the configuration the metric is quoted on — DNA GTR+G4, 100 taxa,
8 reticulations, 1M site patterns sharded across 1/2/4/8 B200 (STRONG scaling: ONE seeded global alignment, generated
in fixed chunks that do not depend on the world size; rank r owns the contiguous slice [r*G/N, (r+1)*G/N) of it, so the
printed lnL is the same number at every N up to summation order).  `--config 2` etc. select another headline.

metric  = CLV site-updates/s (sum over nodes of displayed trees(node) x patterns, per second, whole job);
          lnl_evals_per_sec is reported beside it.
value   = inputs resident in HBM, timed with CUDA events on the engine's stream, max over ranks.
e2e     = the same step through the host C-ABI with HOST buffers: every step re-uploads the rank's alignment
          slice (tipchars + pattern weights, pinned host memory) and reads the lnL back.
roofline= K2 (k_clv_dna4_pipe2) only: COMPULSORY bytes per launch (every distinct child CLV / tip row of the launch read
          once + every parent written once; the ops of a node share children through the L2) / CUDA-event time of the K2
          launches, against MEASURED_PEAKS.json:hbm_gbs.  The SURVEY §8d per-op ("algorithmic") figure is printed beside it.
configs = (N = 1 only) the other four BASELINE configs + config 2's branch-length derivative sweep, each with ms per
          evaluation, site-updates/s, launches, per-kernel-family compulsory-byte roofline fractions, a parity check
          against the oracle on a pattern prefix and the CPU arm on the same inputs.
parity  = lnL / per-tree lnL / scalers of the CUDA path against oracle/_ref (real libpll) on a 700-pattern prefix of
          the SAME global alignment — the checker, outside every timed region.
cpu_baseline / --impl reference = the restated NetRAX layer over the REAL forked libpll (oracle/_ref, kind
          "reference"; the scalar port if _ref is absent), site-sharded over all host cores, bounded sample
          (a prefix of the same global alignment, one contiguous slice per worker).
"""
from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from netrax_b200._capi import AVERAGE, BEST, LINKED, UNLINKED, Partition  # noqa: E402
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, lg_model, random_network, simulate_alignment  # noqa: E402

CONFIGS = {
    1: dict(name="config1: DNA GTR+G4, 20 taxa, 1 reticulation, 10k patterns, AVERAGE", taxa=20, ret=1, patterns=10_000, parts=1, variant=AVERAGE, linkage=LINKED),
    2: dict(name="config2: DNA GTR+G4, 50 taxa, 4 reticulations, 100k patterns, AVERAGE", taxa=50, ret=4, patterns=100_000, parts=1, variant=AVERAGE, linkage=LINKED),
    3: dict(name="config3: DNA 10 partitions x 50k patterns, unlinked brlens, 3 reticulations, BEST", taxa=50, ret=3, patterns=50_000, parts=10, variant=BEST, linkage=UNLINKED),
    4: dict(name="config4: Protein LG+G4, 30 taxa, 2 reticulations, 20k patterns, AVERAGE", taxa=30, ret=2, patterns=20_000, parts=1, variant=AVERAGE, linkage=LINKED, states=20),
    5: dict(name="config5: DNA GTR+G4, 100 taxa, 8 reticulations, 1M site patterns sharded across the GPUs", taxa=100, ret=8, patterns=1_000_000, parts=1, variant=AVERAGE, linkage=LINKED),
}
CHUNK = 31_250   # the global alignment is the concatenation of independently seeded chunks of this many patterns


def _global_columns(cfg, net, p, lo, hi):
    """Columns [lo, hi) of partition p of THE global alignment of this config: chunk c (patterns [c*CHUNK, (c+1)*CHUNK)) is
    simulated with seed 1000 * (c + 1) + p, whatever the world size / worker count, and trimmed to the requested range."""
    states = cfg.get("states", 4)
    cols = []
    for c in range(lo // CHUNK, max(lo // CHUNK + 1, -(-hi // CHUNK))):
        c_lo, c_hi = c * CHUNK, min((c + 1) * CHUNK, cfg["patterns"])
        if c_hi <= lo or c_lo >= hi:
            continue
        if states == 20:
            rates, freqs = lg_model()
            m, _ = simulate_alignment(net, c_hi - c_lo, seed=1000 * (c + 1) + p, dedup=False, states=20, rates=rates, freqs=freqs)
        else:
            m, _ = simulate_alignment(net, c_hi - c_lo, seed=1000 * (c + 1) + p, dedup=False)
        cols.append(m[:, max(lo, c_lo) - c_lo: min(hi, c_hi) - c_lo])
    if not cols:
        return np.zeros((net.num_tips, 0), np.uint32)
    return np.ascontiguousarray(np.concatenate(cols, axis=1))


def make_inputs(cfg, lo, hi=None):
    """Network, model and the pattern slice [lo, hi) of every partition of the config's global alignment.
    make_inputs(cfg, n) == the prefix [0, n) (tests use this form)."""
    if hi is None:
        lo, hi = 0, lo
    net = random_network(cfg["taxa"], cfg["ret"], seed=42 + cfg["taxa"])
    parts, brl = [], []
    rng = np.random.default_rng(5)
    for p in range(cfg["parts"]):
        m = _global_columns(cfg, net, p, lo, hi)
        w = np.ones(m.shape[1], np.uint32)
        if cfg.get("states", 4) == 20:
            rates, freqs = lg_model()
            parts.append(Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w))
        else:
            parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
    return net, parts, (brl if cfg["linkage"] == UNLINKED else None)


def derivative_sweep(eng, net, iters=3, accept=False):
    """For EVERY edge: virtual re-rooting, edge-rooted lnL, sumtables, `iters` Newton-iterate derivative evaluations, restore —
    the loop of optimize_branch (src/optimization/BranchLengthOptimization.cpp:345-420) with a fixed iterate count.  The reference
    visits the candidates in unordered_set order (:423-476), i.e. any order is the reference's; the product proposes the pre-order
    in which consecutive branches share their re-rooting paths (nrxh_brlen_sweep_order), the checker walks the same list.
    accept=False restores the old length (a converged optimisation round); accept=True keeps the last proposal, as an
    optimisation that moves every branch does — the nodes above the edge are then recomputed, as in the reference."""
    order = eng.brlen_sweep_order() if hasattr(eng.api, "_brlen_sweep_order") else range(net.num_edges)
    lengths = eng.branch_lengths()
    fused_call = hasattr(eng.api, "_brlen_logl_sumtables") and not os.environ.get("NRX_BENCH_SEPARATE_K4_K5")
    for e in order:
        e = int(e)
        t0 = float(lengths[e])
        eng.brlen_prepare(e)
        if fused_call:   # product: the edge-rooted lnL and the sumtables come out of one pass over the pairs' CLVs
            n_tables = eng.computeLoglikelihoodBrlenOptAndSumtables(e)[1]
        else:
            eng.computeLoglikelihoodBrlenOpt(e)
            n_tables = eng.computePartitionSumtables(e)
        if n_tables:
            for k in range(iters):
                eng.brlen_set_length(e, t0 * (1.0 + 0.1 * (k + 1)))
                eng.computeLoglikelihoodDerivatives(e)
            if not accept:
                eng.brlen_set_length(e, t0)
        eng.brlen_finish(e)
    if accept:   # put the lengths back so that the next sweep starts from the same state
        for e in range(net.num_edges):
            eng.set_branch_length(e, float(lengths[e]))


# ---------------------------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    kind, cfg, patterns, widx, reps, sweep = args
    from oracle import oracle
    net, parts, brl = make_inputs(cfg, widx * patterns, (widx + 1) * patterns)   # worker widx: its slice of the global prefix
    eng = oracle.make_engine(kind, net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    lnl = eng.computeLoglikelihood(0, 1)  # warm-up
    eng.reset_counters()
    times = []
    for _ in range(reps):
        t = time.perf_counter()
        eng.computeLoglikelihood(0, 1)
        times.append(time.perf_counter() - t)
    updates = eng.clv_update_count() // max(1, reps)
    ts = None
    if sweep:
        t = time.perf_counter()
        derivative_sweep(eng, net)
        ts = time.perf_counter() - t
    return times, updates, ts, lnl


def cpu_reference(cfg, cores, patterns_per_core, reps, sweep=False):
    """All host cores, one worker per core, each owning a slice of every partition (the reference's MPI site
    parallelism, RAXML/ParallelContext.cpp:354-487); per-step time = max over workers."""
    from oracle import oracle
    kind = "ref" if oracle.have_ref() else "port"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, cfg, patterns_per_core, i, reps, sweep) for i in range(cores)])
    step_times = [max(r[0][k] for r in res) for k in range(reps)]
    updates = sum(r[1] for r in res)
    out = {"kind": "reference" if kind == "ref" else "port", "step_times": step_times, "site_updates_per_step": updates,
           "sample": f"prefix of {cores * patterns_per_core} of the config's {cfg['patterns']} global patterns (per partition): {cores} workers x "
                     f"{patterns_per_core} patterns, {reps} full evaluations each (libpll AVX2 kernels under the restated NetRAX driver, "
                     f"not the netrax binary; the arm is memory-bound on the host: 16 -> 32 cores gave 1.3x in round 1)"}
    if sweep:
        out["sweep_s"] = max(r[2] for r in res)
    return out


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


# ---------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap,utilization.gpu"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        busy = [r for r in self.rows if len(r) >= 8 and r[7].strip().isdigit() and int(r[7]) >= 50]
        if busy:   # "under load": samples taken while the GPU was busy (the sampler also spans the gaps between the timed regions)
            self.rows = busy
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


# ---------------------------------------------------------------------------------------------- helpers of our arm
def family_table(prof, reps, total_ms, peak):
    """per kernel family: ms, launches, compulsory-byte GB/s and its fraction of the measured HBM peak"""
    rows = {}
    for k, v in prof.items():
        if v["launches"] == 0 or v["ms"] <= 0:
            continue
        gbs = v["compulsory_bytes"] / (v["ms"] / 1e3) / 1e9
        rows[k] = {"ms": v["ms"] / reps, "launches": v["launches"] / reps, "share": v["ms"] / reps / total_ms if total_ms else None,
                   "compulsory_GBps": gbs, "frac_of_hbm_peak": gbs / peak,
                   "algorithmic_over_compulsory": v["bytes"] / v["compulsory_bytes"] if v["compulsory_bytes"] else None}
    return rows


def parity_check(cfg, n_patterns, device):
    """The CUDA path against the oracle (real libpll when oracle/_ref is present) on the first n_patterns columns of the
    config's global alignment: network lnL, every per-tree partition lnL, every root scaler array.  Checker only."""
    from netrax_b200.engine import NetraxB200
    from oracle import oracle
    net, parts, brl = make_inputs(cfg, min(n_patterns, cfg["patterns"]))
    kind = "ref" if oracle.have_ref() else "port"
    g = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], device=device, partition_brlens=brl)
    o = oracle.make_engine(kind, net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    for p in range(g.P):
        g.set_eigen(p, *o.get_eigen(p))
    lg, lo = g.computeLoglikelihood(0, 1), o.computeLoglikelihood(0, 1)
    root = net.root
    worst, scalers_equal, best_equal = 0.0, True, None
    for t in range(g.num_trees(root)):
        a, b = g.tree_info(root, t)[1], o.tree_info(root, t)[1]
        worst = max(worst, float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-300))))
        scalers_equal = scalers_equal and all(np.array_equal(g.read_scaler(root, t, p), o.read_scaler(root, t, p)) for p in range(g.P))
    out = {"oracle": "reference libpll (oracle/_ref)" if kind == "ref" else "scalar port", "patterns": int(parts[0].sites), "lnl": lg, "oracle_lnl": lo,
           "lnl_rel_diff": abs(lg - lo) / abs(lo), "root_trees": g.num_trees(root), "per_tree_lnl_max_rel_diff": worst,
           "root_scalers_bit_equal": bool(scalers_equal), "trees_per_node_equal": all(g.num_trees(v) == o.num_trees(v) for v in range(net.num_tips, net.num_nodes))}
    out["pass"] = bool(out["lnl_rel_diff"] <= 1e-10 and worst <= 1e-10 and scalers_equal and out["trees_per_node_equal"])
    g.close(); o.close()
    return out


def measure_config(c, device, peak, reps, with_cpu, cores):
    """One BASELINE config on one GPU: full evaluations (+ config 2: the derivative sweep) with per-family rooflines."""
    from netrax_b200.engine import NetraxB200
    cfg = dict(CONFIGS[c])
    net, parts, brl = make_inputs(cfg, cfg["patterns"])
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], device=device, partition_brlens=brl)
    for _ in range(3):
        lnl = eng.computeLoglikelihood(0, 1)
    slots = sum(eng.num_trees(v) for v in range(net.num_tips, net.num_nodes))
    updates = slots * sum(p.sites for p in parts)
    eng.profile_enable(True)
    l0 = eng.launch_count()
    t = time.perf_counter()
    eng.timer_start()
    for _ in range(reps):
        eng.computeLoglikelihood(0, 1)
    ms = eng.timer_stop() / reps
    wall = 1e3 * (time.perf_counter() - t) / reps
    launches = (eng.launch_count() - l0) / reps
    fam = family_table(eng.profile_read_all(), reps, ms, peak)
    eng.profile_enable(False)
    r = {"workload": cfg["name"], "patterns": cfg["patterns"], "partitions": cfg["parts"], "sum_trees_per_node": slots, "root_trees": eng.num_trees(net.root),
         "lnl": lnl, "ms_per_eval": ms, "wall_ms_per_eval": wall, "lnl_evals_per_sec": 1e3 / ms, "site_updates_per_sec": updates / (ms / 1e3),
         "launches_per_eval": launches, "kernel_families": fam}
    if c in (2, 4):   # config 4: the old-length-restored flavour only (the protein K4 / K5 / K6 fractions under the driver's clock)
        for key, accept, lazy in ((("derivative_sweep", False, False), ("derivative_sweep_accept", True, False), ("derivative_sweep_accept_lazy", True, True)) if c == 2
                                  else (("derivative_sweep", False, False),)):
            eng.set_lazy_rerooting(lazy)
            derivative_sweep(eng, net, accept=accept)  # warm-up (allocates re-rooting slots and sumtables)
            eng.computeLoglikelihood(1, 1)
            st0 = eng.reroot_stats()
            eng.profile_enable(True)
            l0 = eng.launch_count()
            t = time.perf_counter()
            eng.timer_start()
            derivative_sweep(eng, net, accept=accept)
            ms_s = eng.timer_stop()
            wall_s = 1e3 * (time.perf_counter() - t)
            st1 = eng.reroot_stats()
            r[key] = {"what": "every edge in pre-order: re-rooting + edge lnL + sumtables + 3 Newton-iterate derivative evaluations + "
                              + ("last proposal kept (nodes above the edge recomputed)" if accept else "old length restored")
                              + (" — lazy re-rooting: no evaluation from the root around each branch, stale root-directed CLVs recomputed when a later re-rooting reads them" if lazy else ""),
                      "edges": int(net.num_edges), "ms": ms_s, "wall_ms": wall_s, "launches": eng.launch_count() - l0,
                      "edges_per_sec": net.num_edges / (wall_s / 1e3),
                      "reroot_memo": {"hits": st1["hits"] - st0["hits"], "misses": st1["misses"] - st0["misses"], "cached_slots": st1["cached_slots"]},
                      "kernel_families": family_table(eng.profile_read_all(), 1, ms_s, peak)}
            eng.profile_enable(False)
        eng.set_lazy_rerooting(False)
    eng.close()
    r["parity"] = parity_check(cfg, 500, device)
    if with_cpu:
        ppc = min(4000 if c == 2 else 16000, max(64, -(-cfg["patterns"] // cores)))
        cpu = cpu_reference(cfg, cores, ppc, 3, sweep=(c == 2))
        t_eval = float(np.median(cpu["step_times"]))
        scale = cfg["patterns"] / (ppc * cores)   # the sample covers ppc*cores of the config's patterns
        r["cpu"] = {"kind": cpu["kind"], "cores": cores, "sample": cpu["sample"], "site_updates_per_sec": cpu["site_updates_per_step"] / t_eval,
                    "ms_per_eval_on_sample": 1e3 * t_eval, "sample_fraction_of_config": 1.0 / scale}
        r["speedup_site_updates_vs_cpu"] = r["site_updates_per_sec"] / r["cpu"]["site_updates_per_sec"]
        if c == 2:
            r["cpu"]["sweep_ms_on_sample"] = 1e3 * cpu["sweep_s"]
            r["derivative_sweep"]["speedup_vs_cpu_scaled_to_config"] = (1e3 * cpu["sweep_s"] * scale) / r["derivative_sweep"]["wall_ms"]
    return r


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="netrax_b200", choices=["netrax_b200", "reference"])
    ap.add_argument("--config", type=int, default=5, choices=sorted(CONFIGS))
    ap.add_argument("--patterns", type=int, default=0, help="global pattern count per partition (default: the config's)")
    ap.add_argument("--cpu-patterns-per-core", type=int, default=0, help="0: the arm's global pattern count / host cores, capped at 16000")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-pipelined", action="store_true", help="e2e region: double-buffered uploads (step k + 1's copy overlaps step k's kernels) — measured SLOWER on config 5, see DESIGN.md §6")
    ap.add_argument("--no-score-only", action="store_true", help="skip the score-only region (root displayed trees' CLVs not stored)")
    ap.add_argument("--no-configs", action="store_true", help="skip the `configs` block (BASELINE configs 1-4 + sweep, N = 1 only)")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    cfg = dict(CONFIGS[args.config])
    if args.patterns:
        cfg["patterns"] = args.patterns
    if not args.cpu_patterns_per_core:
        args.cpu_patterns_per_core = min(16000, max(64, -(-cfg["patterns"] // host_cores())))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # strong scaling: rank r owns the contiguous slice [r*G/N, (r+1)*G/N) of every partition (reference: C1 site sharding)
    G = cfg["patterns"]
    lo, hi = rank * G // world, (rank + 1) * G // world
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    config = {"workload": cfg["name"], "global_patterns": cfg["patterns"], "patterns_per_gpu": -(-cfg["patterns"] // world), "partitions": cfg["parts"],
              "lh_model": "AVERAGE" if cfg["variant"] == AVERAGE else "BEST", "parallelism": f"site-sharding x{world}",
              "alignment": f"one seeded global alignment (chunks of {CHUNK} patterns, seed = 1000 * (chunk + 1) + partition), sliced contiguously by rank",
              "l2": "per-step working set (all CLV slots) is GBs >> 126 MB L2: inputs larger than L2, no flush needed"}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        cores = host_cores()
        r = cpu_reference(cfg, cores, args.cpu_patterns_per_core, args.warmup + args.steps)
        st = r["step_times"][args.warmup:]
        ms = 1e3 * float(np.mean(st))
        v = r["site_updates_per_step"] / (ms / 1e3)
        line = {"impl": "reference", "metric": "clv_site_updates_per_sec", "value": v, "unit": "site-updates/s", "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "lnl_evals_per_sec_on_sample": 1e3 / ms,
                "note": "ms_per_step is the time of one evaluation of the SAMPLE (see cpu_baseline.sample), not of the whole config; compare site-updates/s",
                "cpu_baseline": {"value": v, "unit": "site-updates/s", "cores": cores, "kind": r["kind"], "sample": r["sample"]},
                "e2e": {"value": v, "unit": "site-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from netrax_b200.engine import NetraxB200

    net, parts, brl = make_inputs(cfg, lo, hi)
    comm = None
    if world > 1:
        # the reference's parallel_reduce_cb (MPI_Allreduce SUM) -> ONE ncclAllReduce inside the engine per evaluation;
        # torch.distributed only carries the 128-byte NCCL unique id to the ranks (and the timing max-reduce below)
        from netrax_b200.engine import comm_unique_id
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local_rank}")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        comm = (bytes(uid.cpu().numpy().tobytes()), rank, world)

    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], device=local_rank, partition_brlens=brl, comm=comm)

    def barrier():
        eng.api.check(1)
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    # pinned host copies of this rank's alignment slice for the e2e leg
    # (4 states: the mask is the code; other alphabets: 1-byte codes + the code -> state-set map)
    tipmaps = [None] * len(parts)
    tip_u8 = []
    for i, p in enumerate(parts):
        if p.states == 4:
            tip_u8.append(torch.from_numpy(p.tip_masks.astype(np.uint8)).pin_memory())
        else:
            tm, inv = np.unique(p.tip_masks, return_inverse=True)
            tipmaps[i] = tm.astype(np.uint32)
            tip_u8.append(torch.from_numpy(inv.reshape(p.tip_masks.shape).astype(np.uint8)).pin_memory())
    w_u32 = [torch.from_numpy((p.pattern_weights if p.pattern_weights is not None else np.ones(p.sites, np.uint32)).astype(np.int32)).pin_memory() for p in parts]

    # clocks are sampled from the warm-up to the end of the e2e region (every part of it is the same step under load)
    sampler = ClockSampler(local_rank)
    sampler.start()
    lnl = None
    for _ in range(args.warmup):
        lnl = eng.computeLoglikelihood(0, 1)
    slots_sum = sum(eng.num_trees(v) for v in range(net.num_tips, net.num_nodes))
    updates_per_step_local = slots_sum * sum(p.sites for p in parts)

    # ---- timed region 1: device-resident inputs ----
    eng.profile_enable(True)
    l0 = eng.launch_count()
    barrier()
    eng.timer_start()
    t_wall = time.perf_counter()
    for _ in range(args.steps):
        lnl = eng.computeLoglikelihood(0, 1)
    ms_dev = eng.timer_stop()
    barrier()
    ms_wall = 1e3 * (time.perf_counter() - t_wall)
    launches = eng.launch_count() - l0
    prof_all = eng.profile_read_all()
    prof = prof_all["K2_clv_update"]
    eng.profile_enable(False)

    # ---- timed region 2: end to end with host buffers ----
    ms_e2e, lnl_e2e = None, None
    # every step uploads its alignment slice from pinned host memory and reads its lnL back: the upload sits in front of its own
    # evaluation on the engine stream.  --e2e-pipelined (4-state partitions): double-buffered instead — the copy of step k + 1 runs on the
    # engine's copy stream while step k computes (stage / commit).  Measured on config 5: 42.9 instead of 34.8 ms per step — K2 keeps the
    # HBM at 98 % of its pin bandwidth and the DMA writes of the overlapped copy crawl, so the copy becomes the critical path.
    pipelined = all(t is None for t in tipmaps) and args.e2e_pipelined
    if tip_u8 is not None:
        def stage():
            for p in range(len(parts)):
                eng.stage_alignment_u8(p, tip_u8[p].data_ptr(), w_u32[p].data_ptr())
        barrier()
        eng.timer_start()
        if pipelined:
            stage()
        for k in range(args.steps):
            if pipelined:
                eng.commit_staged_alignment()
                if k + 1 < args.steps:
                    stage()
            else:
                for p in range(len(parts)):
                    if tipmaps[p] is None:
                        eng.upload_alignment_u8(p, tip_u8[p].data_ptr(), w_u32[p].data_ptr())
                    else:
                        eng.upload_alignment_codes(p, tip_u8[p].data_ptr(), tipmaps[p], w_u32[p].data_ptr())
            lnl_e2e = eng.computeLoglikelihood(0, 1)
        ms_e2e = eng.timer_stop()
        barrier()
        assert abs(lnl_e2e - lnl) <= 1e-9 * abs(lnl)
    # ---- region 3 (reported beside the headline, never instead of it): score-only evaluations — candidate scoring reads nothing but
    # the lnL, so the replayed plan does not store the CLVs of the root displayed trees (nrxh_set_score_only) ----
    ms_so, lnl_so = None, None
    if hasattr(eng.api, "_set_score_only") and not args.no_score_only:
        eng.set_score_only(True)
        for _ in range(2):
            eng.computeLoglikelihood(0, 1)
        barrier()
        eng.timer_start()
        for _ in range(args.steps):
            lnl_so = eng.computeLoglikelihood(0, 1)
        ms_so = eng.timer_stop()
        barrier()
        eng.set_score_only(False)
        assert lnl_so == lnl, (lnl_so, lnl)   # same kernels, same per-site terms: the same number
    clocks = sampler.stop()
    h2d = (sum(int(t.numel()) for t in tip_u8) if tip_u8 else 0) + sum(4 * int(t.numel()) for t in w_u32) + 8 * (net.num_edges + 1) * len(parts)
    d2h = 8 * eng.num_trees(net.root) * len(parts) + 8

    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e or 0.0, ms_wall, ms_so or 0.0], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_wall, ms_so = (float(x) for x in t.cpu())
        u = torch.tensor([float(updates_per_step_local)], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(u)
        updates_per_step = float(u.cpu()[0])
    else:
        updates_per_step = float(updates_per_step_local)

    if rank == 0:
        ms_step = ms_dev / args.steps
        value = updates_per_step / (ms_step / 1e3)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        nl = max(1, prof["launches"])
        avg_launch_s = prof["ms"] / nl / 1e3 if prof["ms"] > 0 else float("inf")
        achieved = prof["compulsory_bytes"] / nl / avg_launch_s / 1e9
        algorithmic = prof["bytes"] / nl / avg_launch_s / 1e9
        # DRAM traffic of K2: dram__bytes_read.sum + dram__bytes_write.sum per launch from an ncu pass over one evaluation
        # step of THIS workload at N = 1 (profiles/k2_traffic.json, which names the command); at N > 1 each GPU holds 1/N of
        # the patterns and the per-launch traffic scales with them (flagged "scaled")
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
            if args.config == tr.get("config", 5) and cfg["patterns"] == tr.get("global_patterns"):
                traffic = tr["dram_bytes_per_launch"] / world
                traffic_src = tr["source"] + ("" if world == 1 else f" (scaled by 1/{world}: patterns per GPU)")
        except Exception:
            pass
        line = {"metric": "clv_site_updates_per_sec", "value": value, "unit": "site-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "lnl_evals_per_sec": 1e3 / ms_step, "lnl": lnl, "sum_trees_per_node": slots_sum, "root_trees": eng.num_trees(net.root),
                "wall_ms_per_step": ms_wall / args.steps,
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"kernel": "k_clv_dna4_pipe2 (K2, CLV update)" if cfg.get("states", 4) == 4 else "k_aa20_dmma<AA_CLV> (K2, CLV update)",
                             "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                             "numerator": "compulsory bytes per launch: every distinct child CLV + scaler / tip row read once, every parent CLV + scaler written once",
                             "compulsory_bytes_per_launch": prof["compulsory_bytes"] / nl,
                             "traffic_over_compulsory": (traffic / (prof["compulsory_bytes"] / nl)) if traffic else None,
                             "dram_frac": (traffic / avg_launch_s / 1e9 / peak) if traffic and peak else None,
                             "algorithmic_per_op": {"what": "SURVEY §8d per-op bytes (every op charged both children); re-reads of shared children are L2 hits, so this is NOT a fraction of the HBM peak",
                                                    "bytes_per_launch": prof["bytes"] / nl, "GBps": algorithmic},
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                             "launches": int(prof["launches"]), "avg_launch_ms": prof["ms"] / nl,
                             "site_updates_per_sec_in_kernel": prof["units"] / (prof["ms"] / 1e3) if prof["ms"] > 0 else None,
                             "share_of_step": prof["ms"] / ms_dev if ms_dev > 0 else None},
                "kernel_families": family_table(prof_all, args.steps, ms_step, peak)}
        if ms_e2e:
            e2e_value = updates_per_step / (ms_e2e / args.steps / 1e3)
            line["e2e"] = {"value": e2e_value, "unit": "site-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                           "ms_per_step": ms_e2e / args.steps, "lnl_evals_per_sec": 1e3 / (ms_e2e / args.steps),
                           "uploads": ("double-buffered: step k + 1's host->device copy overlaps step k's kernels (nrxh_stage_alignment_u8 / nrxh_commit_staged_alignment); "
                                       "one upload per step, all inside the timed region") if pipelined else "in front of each evaluation on the engine stream"}
        if ms_so:
            line["score_only"] = {"what": "the same full evaluation with nrxh_set_score_only: the CLVs of the root displayed trees (root_trees of sum_trees_per_node "
                                          "slots) are computed and reduced to their per-site lnL but not stored; NOT the headline — batched candidate scoring (SURVEY §8f f3)",
                                  "ms_per_step": ms_so / args.steps, "value": updates_per_step / (ms_so / args.steps / 1e3), "unit": "site-updates/s", "lnl": lnl_so}
    eng.close()
    del eng
    if rank == 0:
        if not args.no_parity:
            line["parity"] = parity_check(cfg, 700, local_rank)
        cores = host_cores()
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference(cfg, cores, args.cpu_patterns_per_core, 1 + 5)
            st = r["step_times"][1:]
            v = r["site_updates_per_step"] / float(np.median(st))
            line["cpu_baseline"] = {"value": v, "unit": "site-updates/s", "cores": cores, "kind": r["kind"], "sample": r["sample"]}
        if world == 1 and not args.no_configs:
            line["configs"] = {}
            for c in (1, 2, 3, 4):
                if c == args.config:
                    continue
                try:
                    line["configs"][str(c)] = measure_config(c, local_rank, float(peaks.get("hbm_gbs", 6650.0)), 20, not args.no_cpu_baseline, cores)
                except Exception as ex:   # a failing side measurement must not take the headline line with it
                    line["configs"][str(c)] = {"error": f"{type(ex).__name__}: {ex}"}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
