#!/usr/bin/env python
"""Batched candidate scoring on a BASELINE config: K candidate networks (same taxa / alignment shape,
different topologies), full re-evaluation of every candidate per round — sequential computeLoglikelihood calls vs one
computeLoglikelihoodBatch call (one CUDA stream per candidate, all enqueued before the first result is collected).
  python scripts/batch_scoring.py [--config 1] [--candidates 8] [--rounds 50]"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from netrax_b200.synth import random_network  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=1)
    ap.add_argument("--candidates", type=int, default=8)
    ap.add_argument("--rounds", type=int, default=50)
    args = ap.parse_args()
    from netrax_b200.engine import NetraxB200, compute_loglikelihood_batch
    cfg = dict(bench.CONFIGS[args.config])
    _, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    engs = []
    for k in range(args.candidates):
        net = random_network(cfg["taxa"], cfg["ret"], seed=500 + k)
        engs.append(NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"]))
    for _ in range(3):
        seq = [e.computeLoglikelihood(0, 1) for e in engs]
        bat = compute_loglikelihood_batch(engs, 0, 1)
    assert list(bat) == seq, (bat, seq)
    t = time.perf_counter()
    for _ in range(args.rounds):
        for e in engs:
            e.computeLoglikelihood(0, 1)
    t_seq = (time.perf_counter() - t) / args.rounds
    t = time.perf_counter()
    for _ in range(args.rounds):
        compute_loglikelihood_batch(engs, 0, 1)
    t_bat = (time.perf_counter() - t) / args.rounds
    print(json.dumps({"workload": cfg["name"], "candidates": args.candidates, "sequential_ms_per_round": 1e3 * t_seq,
                      "batched_ms_per_round": 1e3 * t_bat, "sequential_evals_per_s": args.candidates / t_seq,
                      "batched_evals_per_s": args.candidates / t_bat, "speedup": t_seq / t_bat}))
    for e in engs:
        e.close()


if __name__ == "__main__":
    main()
