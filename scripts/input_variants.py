#!/usr/bin/env python
"""SURVEY §8d input variants on one BASELINE config: (a) the default synthetic inputs (alignment simulated down displayed
tree 0, reticulation probabilities ~ U(0.2, 0.8)), (b) every reticulation probability exactly 0.5 (the reference
experiments' setting), (c) uniform-random cells (no phylogenetic signal: the worst case for numerical scaling).
Reports ms per full evaluation and how many (root tree, pattern) entries carry a non-zero scaler.
  python scripts/input_variants.py [--config 2]"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from netrax_b200._capi import Partition  # noqa: E402
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, simulate_alignment  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--reps", type=int, default=30)
    args = ap.parse_args()
    from netrax_b200.engine import NetraxB200
    cfg = dict(bench.CONFIGS[args.config])
    net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    out = {"workload": cfg["name"], "variants": {}}
    for name in ("default", "probs_0.5", "random_cells"):
        ps = parts
        if name == "random_cells":
            ps = []
            for p in range(cfg["parts"]):
                m, w = simulate_alignment(net, cfg["patterns"], seed=7000 + p, dedup=False, random_cells=True)
                ps.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
        eng = NetraxB200(net, ps, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
        if name == "probs_0.5":
            for r in range(net.num_reticulations):
                eng.set_reticulation_prob(r, 0.5)
        for _ in range(3):
            lnl = eng.computeLoglikelihood(0, 1)
        eng.timer_start()
        for _ in range(args.reps):
            eng.computeLoglikelihood(0, 1)
        ms = eng.timer_stop() / args.reps
        scaled = sum(int(np.count_nonzero(eng.read_scaler(net.root, t))) for t in range(eng.num_trees(net.root)))
        max_scaler = max(int(eng.read_scaler(net.root, t).max()) for t in range(eng.num_trees(net.root)))
        out["variants"][name] = {"ms_per_eval": ms, "lnl": lnl, "root_entries_with_scaler": scaled, "max_scaler": max_scaler,
                                 "root_trees": eng.num_trees(net.root)}
        eng.close()
    print(json.dumps(out))


if __name__ == "__main__":
    main()
