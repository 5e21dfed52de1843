#!/usr/bin/env python
"""K2 roofline for DNA with other rate-category counts (config-2 topology, 100 k patterns): python scripts/cats_roofline.py [--md out.md]"""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import bench
from netrax_b200._capi import Partition
from netrax_b200.synth import DNA_FREQS, GTR_RATES
from netrax_b200.engine import NetraxB200, load


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--md", default="")
    ap.add_argument("--patterns", type=int, default=100_000)
    args = ap.parse_args()
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    cfg = dict(bench.CONFIGS[2]); cfg["patterns"] = args.patterns
    net, parts, _ = bench.make_inputs(cfg, args.patterns)
    rows = []
    for cats in (1, 2, 4, 8, 16, 3):
        rates = load().gamma_rates(0.5, cats) if cats > 1 else np.ones(1)
        part = Partition(4, cats, parts[0].tip_masks, DNA_FREQS, GTR_RATES, rates, pattern_weights=parts[0].pattern_weights)
        eng = NetraxB200(net, [part])
        for _ in range(3):
            eng.computeLoglikelihood(0, 1)
        eng.profile_enable(True)
        eng.timer_start()
        for _ in range(10):
            eng.computeLoglikelihood(0, 1)
        ms = eng.timer_stop() / 10
        k2 = eng.profile_read_all()["K2_clv_update"]
        eng.profile_enable(False)
        gbs = k2["compulsory_bytes"] / (k2["ms"] / 1e3) / 1e9
        rows.append((cats, ms, k2["ms"] / 10, gbs, gbs / peak))
        print(json.dumps({"cats": cats, "ms_per_eval": ms, "k2_ms": k2["ms"] / 10, "k2_compulsory_GBps": gbs, "frac_of_hbm_peak": gbs / peak}), flush=True)
        eng.close()
    if args.md:
        with open(args.md, "w") as f:
            f.write(f"# K2 by rate-category count: DNA, config-2 network (50 taxa, 4 reticulations), {args.patterns} patterns, one B200 (peak {peak:.0f} GB/s measured copy)\n\n")
            f.write("1 / 2 / 4 / 8 / 16 categories: k_clv_dna4_pipe2<2, CATS>; 3 categories: k_clv_generic (thread per pattern).\n\n| categories | ms / evaluation | K2 ms | K2 compulsory GB/s | frac of peak |\n|---|---|---|---|---|\n")
            for r in rows:
                f.write(f"| {r[0]} | {r[1]:.3f} | {r[2]:.3f} | {r[3]:.0f} | {r[4]:.2f} |\n")


if __name__ == "__main__":
    main()
