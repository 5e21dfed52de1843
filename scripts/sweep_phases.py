#!/usr/bin/env python
"""Wall-clock per phase of the branch-length derivative sweep (host + device, each phase ends in a host-visible result
except `sumtables`, which only enqueues): where the host leaves the GPU idle.  python scripts/sweep_phases.py [--config 2]"""
import argparse
import collections
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, default=2)
    args = ap.parse_args()
    from netrax_b200.engine import NetraxB200
    cfg = dict(bench.CONFIGS[args.config])
    net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    eng.computeLoglikelihood(0, 1)
    acc = collections.OrderedDict()

    def timed(name, fn):
        t = time.perf_counter()
        r = fn()
        acc[name] = acc.get(name, 0.0) + time.perf_counter() - t
        return r

    for rep in range(2):
        acc.clear()
        t_all = time.perf_counter()
        for e in range(net.num_edges):
            t0 = float(net.edge_length[e])
            timed("brlen_prepare (re-rooting)", lambda: eng.brlen_prepare(e))
            timed("computeLoglikelihoodBrlenOpt", lambda: eng.computeLoglikelihoodBrlenOpt(e))
            if timed("computePartitionSumtables", lambda: eng.computePartitionSumtables(e)):
                for k in range(3):
                    timed("brlen_set_length", lambda: eng.brlen_set_length(e, t0 * (1.0 + 0.1 * (k + 1))))
                    timed("computeLoglikelihoodDerivatives", lambda: eng.computeLoglikelihoodDerivatives(e))
                timed("brlen_set_length", lambda: eng.brlen_set_length(e, t0))
            timed("brlen_finish (restore)", lambda: eng.brlen_finish(e))
        total = time.perf_counter() - t_all
    print(json.dumps({"config": cfg["name"], "sweep_wall_ms": 1e3 * total, "phase_ms": {k: 1e3 * v for k, v in acc.items()}}))
    eng.close()


if __name__ == "__main__":
    main()
