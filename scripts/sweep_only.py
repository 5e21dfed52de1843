#!/usr/bin/env python
"""One branch-length derivative sweep (scripts/bench_configs.py:derivative_sweep) on a BASELINE config, for profiling:
  python scripts/sweep_only.py [--config 2] [--patterns N] [--edges K] [--mode sweep|eval|opt]
Prints device ms, wall ms and launches of the timed pass.  Run under `ncu --metrics gpu__time_duration.sum` for the
launch list (numbers printed under a profiler are not bench values)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bench import derivative_sweep  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    ap.add_argument("--patterns", type=int, default=0)
    ap.add_argument("--mode", default="sweep")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--no-warmup", action="store_true")
    args = ap.parse_args()
    from netrax_b200.engine import NetraxB200
    cfg = dict(bench.CONFIGS[args.config])
    if args.patterns:
        cfg["patterns"] = args.patterns
    net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    for _ in range(3):
        eng.computeLoglikelihood(0, 1)

    def run():
        if args.mode == "sweep":
            derivative_sweep(eng, net)
        elif args.mode == "eval":
            eng.computeLoglikelihood(0, 1)
        else:
            eng.optimize_branches()

    if not args.no_warmup:
        run()
    l0 = eng.launch_count()
    t = time.perf_counter()
    eng.timer_start()
    for _ in range(args.reps):
        run()
    ms = eng.timer_stop() / args.reps
    wall = 1e3 * (time.perf_counter() - t) / args.reps
    print(json.dumps({"config": cfg["name"], "mode": args.mode, "device_ms": ms, "wall_ms": wall,
                      "launches": (eng.launch_count() - l0) / args.reps, "edges": int(net.num_edges)}))
    eng.close()


if __name__ == "__main__":
    main()
