#!/usr/bin/env python
"""bench.py — network-likelihood hot path on N B200s vs the reference's CPU path on the box's host cores.

A "step" = ONE full network lnL evaluation, computeLoglikelihood(ann, incremental=0, update_pmatrices=1):
P-matrices for every edge + every CLV of every displayed tree at every node + per-tree root lnL + the
cross-rank reduction + AVERAGE/BEST mixing (BASELINE.md §4 "What is timed").

Workload (default): BASELINE.json configs[4], the configuration the metric is quoted on — DNA GTR+G4, 100 taxa,
8 reticulations, 1M site patterns sharded across 1/2/4/8 B200 (STRONG scaling: the 1M patterns are split evenly
over the ranks; at N=1 all 776 CLV slots x 1M patterns = 102 GB live on one GPU).  `--config 2` (50 taxa /
4 reticulations / 100 k patterns) etc. select the others; `--patterns N` overrides the global pattern count.

metric  = CLV site-updates/s (sum over nodes of displayed trees(node) x patterns, per second, whole job);
          lnl_evals_per_sec is reported beside it.
value   = inputs resident in HBM, timed with CUDA events on the engine's stream, max over ranks.
e2e     = the same step through the host C-ABI with HOST buffers: every step re-uploads the rank's alignment
          slice (tipchars + pattern weights, pinned host memory) and reads the lnL back.
roofline= K2 (k_clv_dna4_pipe2) only: algorithmic bytes / CUDA-event time of the K2 launches.
cpu_baseline / --impl reference = the restated NetRAX layer over the REAL forked libpll (oracle/_ref, kind
          "reference"; the scalar port if _ref is absent), site-sharded over all host cores, bounded sample.
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


def make_inputs(cfg, patterns_local, rank):
    """Seeded synthetic inputs of the named shape; every rank simulates only its own slice (different seed per
    rank = different columns, same network and model)."""
    net = random_network(cfg["taxa"], cfg["ret"], seed=42 + cfg["taxa"])
    parts, brl = [], []
    rng = np.random.default_rng(5)
    for p in range(cfg["parts"]):
        if cfg.get("states", 4) == 20:
            rates, freqs = lg_model()
            m, w = simulate_alignment(net, patterns_local, seed=1000 * (rank + 1) + p, dedup=False, states=20, rates=rates, freqs=freqs)
            parts.append(Partition(20, 4, m, freqs, rates, GAMMA4_ALPHA05, pattern_weights=w))
            continue
        m, w = simulate_alignment(net, patterns_local, seed=1000 * (rank + 1) + p, dedup=False)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
    return net, parts, (brl if cfg["linkage"] == UNLINKED else None)


# ---------------------------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    kind, cfg, patterns, widx, reps = args
    from oracle import oracle
    net, parts, brl = make_inputs(cfg, patterns, 100 + widx)
    eng = oracle.make_engine(kind, net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    eng.computeLoglikelihood(0, 1)  # warm-up
    eng.reset_counters()
    times = []
    for _ in range(reps):
        t = time.perf_counter()
        eng.computeLoglikelihood(0, 1)
        times.append(time.perf_counter() - t)
    return times, eng.clv_update_count() // reps


def cpu_reference(cfg, cores, patterns_per_core, reps):
    """All host cores, one worker per core, each owning a slice of every partition (the reference's MPI site
    parallelism, RAXML/ParallelContext.cpp:354-487); per-step time = max over workers."""
    from oracle import oracle
    kind = "ref" if oracle.have_ref() else "port"
    ctx = mp.get_context("fork")
    with ctx.Pool(cores) as pool:
        res = pool.map(_cpu_worker, [(kind, cfg, patterns_per_core, i, reps) for i in range(cores)])
    step_times = [max(r[0][k] for r in res) for k in range(reps)]
    updates = sum(r[1] for r in res)
    return {"kind": "reference" if kind == "ref" else "port", "step_times": step_times, "site_updates_per_step": updates,
            "sample": f"{cores} workers x {patterns_per_core} patterns of the same network/model, {reps} full evaluations each "
                      f"(libpll AVX2 kernels under the restated NetRAX driver, not the netrax binary)"}


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
    local_patterns = (rank + 1) * G // world - rank * G // world
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    config = {"workload": cfg["name"], "global_patterns": cfg["patterns"], "patterns_per_gpu": -(-cfg["patterns"] // world), "partitions": cfg["parts"],
              "lh_model": "AVERAGE" if cfg["variant"] == AVERAGE else "BEST", "parallelism": f"site-sharding x{world}",
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

    net, parts, brl = make_inputs(cfg, local_patterns, rank)
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
    tip_u8 = [torch.from_numpy(p.tip_masks.astype(np.uint8)).pin_memory() for p in parts]
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
    prof = eng.profile_read()
    eng.profile_enable(False)

    # ---- timed region 2: end to end with host buffers ----
    barrier()
    eng.timer_start()
    for _ in range(args.steps):
        for p in range(len(parts)):
            eng.upload_alignment_u8(p, tip_u8[p].data_ptr(), w_u32[p].data_ptr())
        lnl_e2e = eng.computeLoglikelihood(0, 1)
    ms_e2e = eng.timer_stop()
    barrier()
    clocks = sampler.stop()
    assert abs(lnl_e2e - lnl) <= 1e-9 * abs(lnl)
    h2d = sum(int(t.numel()) for t in tip_u8) + sum(4 * int(t.numel()) for t in w_u32) + 8 * (net.num_edges + 1) * len(parts)
    d2h = 8 * eng.num_trees(net.root) * len(parts) + 8

    if dist is not None:
        t = torch.tensor([ms_dev, ms_e2e, ms_wall], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_dev, ms_e2e, ms_wall = (float(x) for x in t.cpu())
        u = torch.tensor([float(updates_per_step_local)], dtype=torch.float64, device=f"cuda:{local_rank}")
        dist.all_reduce(u)
        updates_per_step = float(u.cpu()[0])
    else:
        updates_per_step = float(updates_per_step_local)

    if rank == 0:
        ms_step = ms_dev / args.steps
        value = updates_per_step / (ms_step / 1e3)
        e2e_value = updates_per_step / (ms_e2e / args.steps / 1e3)
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = prof["clv_bytes"] / (prof["clv_ms"] / 1e3) / 1e9 if prof["clv_ms"] > 0 else 0.0
        # DRAM traffic of K2 from the committed ncu --set full capture (profiles/k2_traffic.json): the ratio measured
        # traffic / algorithmic bytes of one evaluation step is a property of the plan (which ops share children),
        # independent of the pattern count, so it scales the per-launch algorithmic bytes of THIS run.
        traffic, traffic_src = None, None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "k2_traffic.json")))
            if args.config == 5:
                traffic = tr["traffic_over_algorithmic"] * prof["clv_bytes"] / max(1, prof["clv_launches"])
                traffic_src = tr["source"]
        except Exception:
            pass
        line = {"metric": "clv_site_updates_per_sec", "value": value, "unit": "site-updates/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f64", "data": "synthetic", "config": config,
                "lnl_evals_per_sec": 1e3 / ms_step, "lnl": lnl, "sum_trees_per_node": slots_sum, "root_trees": eng.num_trees(net.root),
                "wall_ms_per_step": ms_wall / args.steps,
                "e2e": {"value": e2e_value, "unit": "site-updates/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": ms_e2e / args.steps, "lnl_evals_per_sec": 1e3 / (ms_e2e / args.steps)},
                "gpu_launches": int(launches),
                "clocks": clocks,
                "roofline": {"kernel": "k_clv_dna4_pipe2 (K2, CLV update)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                             "frac": achieved / peak if peak else None, "traffic": traffic, "traffic_source": traffic_src,
                             "dram_frac": (traffic / (prof["clv_ms"] / max(1, prof["clv_launches"]) / 1e3) / 1e9 / peak) if traffic and peak and prof["clv_ms"] > 0 else None,
                             "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)" if peaks else "fallback 6650 GB/s (B200_PROFILING.md)",
                             "launches": int(prof["clv_launches"]), "avg_launch_ms": prof["clv_ms"] / max(1, prof["clv_launches"]),
                             "algorithmic_bytes_per_launch": prof["clv_bytes"] / max(1, prof["clv_launches"]),
                             "site_updates_per_sec_in_kernel": prof["clv_site_updates"] / (prof["clv_ms"] / 1e3) if prof["clv_ms"] > 0 else None,
                             "share_of_step": prof["clv_ms"] / ms_dev if ms_dev > 0 else None}}
        if world == 1 and not args.no_cpu_baseline:
            cores = host_cores()
            r = cpu_reference(cfg, cores, args.cpu_patterns_per_core, 1 + 5)
            st = r["step_times"][1:]
            v = r["site_updates_per_step"] / float(np.median(st))
            line["cpu_baseline"] = {"value": v, "unit": "site-updates/s", "cores": cores, "kind": r["kind"], "sample": r["sample"]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
