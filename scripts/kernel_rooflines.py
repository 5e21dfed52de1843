#!/usr/bin/env python
"""Per-kernel roofline table: algorithmic GB/s (SURVEY §8d bytes / CUDA-event time of the kernel family's launches on
the engine stream) as a fraction of MEASURED_PEAKS.json:hbm_gbs, for one full evaluation and one branch-length
derivative sweep (scripts/bench_configs.py:derivative_sweep) of a BASELINE config.
  python scripts/kernel_rooflines.py --configs 2,4 [--out gpurun_out/x.json] [--md profiles/x.md]
Event brackets serialise nothing (same stream), so the sum of the family times ~ the device time of the pass."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from bench import derivative_sweep  # noqa: E402


def peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    return json.load(open(p)).get("hbm_gbs", 6650.0) if os.path.exists(p) else 6650.0


def measure(eng, fn, reps):
    fn()  # warm-up (allocations, plan capture)
    eng.profile_enable(True)
    l0 = eng.launch_count()
    t = time.perf_counter()
    eng.timer_start()
    for _ in range(reps):
        fn()
    ms = eng.timer_stop() / reps
    wall = 1e3 * (time.perf_counter() - t) / reps
    prof = eng.profile_read_all()
    eng.profile_enable(False)
    rows = {}
    for k, v in prof.items():
        if v["launches"] == 0:
            continue
        gbs = v["compulsory_bytes"] / (v["ms"] / 1e3) / 1e9 if v["ms"] > 0 else 0.0
        rows[k] = {"ms_per_pass": v["ms"] / reps, "launches_per_pass": v["launches"] / reps, "units_per_pass": v["units"] / reps,
                   "compulsory_GB_per_pass": v["compulsory_bytes"] / reps / 1e9, "algorithmic_GB_per_pass": v["bytes"] / reps / 1e9,
                   "GBps": gbs, "frac_of_peak": gbs / peak_gbs(), "share_of_pass": v["ms"] / reps / ms}
    return {"device_ms": ms, "wall_ms": wall, "launches": (eng.launch_count() - l0) / reps, "kernels": rows}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="2")
    ap.add_argument("--patterns", type=int, default=0)
    ap.add_argument("--out", default="")
    ap.add_argument("--md", default="")
    args = ap.parse_args()
    from netrax_b200.engine import NetraxB200
    res = {"peak_GBps": peak_gbs()}
    for c in [int(x) for x in args.configs.split(",")]:
        cfg = dict(bench.CONFIGS[c])
        if args.patterns:
            cfg["patterns"] = args.patterns
        net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
        eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
        eng.computeLoglikelihood(0, 1)
        r = {"workload": cfg["name"], "patterns": cfg["patterns"]}
        r["full_evaluation"] = measure(eng, lambda: eng.computeLoglikelihood(0, 1), 10)
        if cfg["patterns"] * cfg["parts"] <= 200_000:
            derivative_sweep(eng, net)   # warm-up: re-rooting slots, sumtables, the re-root memo
            r["derivative_sweep"] = measure(eng, lambda: derivative_sweep(eng, net), 1)
            derivative_sweep(eng, net, accept=True)
            r["derivative_sweep_accept"] = measure(eng, lambda: derivative_sweep(eng, net, accept=True), 1)
            eng.set_lazy_rerooting(True)
            derivative_sweep(eng, net, accept=True)
            r["derivative_sweep_accept_lazy"] = measure(eng, lambda: derivative_sweep(eng, net, accept=True), 1)
            eng.set_lazy_rerooting(False)
            r["reroot_memo"] = eng.reroot_stats()
        eng.close()
        res[f"config{c}"] = r
        print(json.dumps({f"config{c}": r}), flush=True)
    if args.out:
        json.dump(res, open(args.out, "w"), indent=1)
    if args.md:
        with open(args.md, "w") as f:
            f.write(f"# Per-kernel rooflines (scripts/kernel_rooflines.py; CUDA events on the engine stream; peak = {peak_gbs():.0f} GB/s measured copy bandwidth)\n\n")
            f.write("GB/s and the fraction use COMPULSORY bytes (per launch: every distinct operand CLV / tip row once + every output once); the SURVEY §8d per-op figure is the `algorithmic GB` column.\n")
            for c, r in res.items():
                if not c.startswith("config"):
                    continue
                for phase in ("full_evaluation", "derivative_sweep", "derivative_sweep_accept", "derivative_sweep_accept_lazy"):
                    if phase not in r:
                        continue
                    m = r[phase]
                    f.write(f"\n## {r['workload']} — {phase.replace('_', ' ')}: {m['device_ms']:.3f} ms device, {m['wall_ms']:.3f} ms wall, {m['launches']:.0f} launches\n\n")
                    f.write("| kernel family | launches | ms | share | compulsory GB | algorithmic GB | GB/s | frac of peak |\n|---|---|---|---|---|---|---|---|\n")
                    for k, v in sorted(m["kernels"].items(), key=lambda kv: -kv[1]["ms_per_pass"]):
                        f.write(f"| {k} | {v['launches_per_pass']:.0f} | {v['ms_per_pass']:.3f} | {100 * v['share_of_pass']:.1f} % | {v['compulsory_GB_per_pass']:.3f} | {v['algorithmic_GB_per_pass']:.3f} | {v['GBps']:.0f} | {v['frac_of_peak']:.2f} |\n")


if __name__ == "__main__":
    main()
