"""Where the host time of the derivative sweep goes: wall time per API call of the sweep (config 2), at full size and at 64 patterns
(device work ~ 0: what is left is host algebra, launches and round trips)."""
import collections, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from netrax_b200.engine import NetraxB200  # noqa: E402

for patterns in (100000, 64):
    cfg = dict(bench.CONFIGS[2]); cfg["patterns"] = patterns
    net, parts, brl = bench.make_inputs(cfg, patterns)
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    eng.computeLoglikelihood(0, 1)
    bench.derivative_sweep(eng, net)
    T = collections.defaultdict(float)
    order = eng.brlen_sweep_order(); lengths = eng.branch_lengths()
    def timed(name, f, *a):
        t = time.perf_counter(); r = f(*a); T[name] += time.perf_counter() - t; return r
    l0 = eng.launch_count(); t_all = time.perf_counter()
    for e in order:
        e = int(e); t0 = float(lengths[e])
        timed("prepare", eng.brlen_prepare, e)
        if os.environ.get("NRX_BENCH_SEPARATE_K4_K5"):
            timed("brlen_logl", eng.computeLoglikelihoodBrlenOpt, e)
            n_tables = timed("sumtables", eng.computePartitionSumtables, e)
        else:
            n_tables = timed("brlen_logl+sumtables", eng.computeLoglikelihoodBrlenOptAndSumtables, e)[1]
        if n_tables:
            for k in range(3):
                timed("set_length", eng.brlen_set_length, e, t0 * (1.0 + 0.1 * (k + 1)))
                timed("derivatives", eng.computeLoglikelihoodDerivatives, e)
            timed("set_length", eng.brlen_set_length, e, t0)
        timed("finish", eng.brlen_finish, e)
    total = time.perf_counter() - t_all
    print(json.dumps({"patterns": patterns, "edges": int(net.num_edges), "total_ms": 1e3 * total, "launches": eng.launch_count() - l0,
                      "per_call_ms": {k: round(1e3 * v, 3) for k, v in T.items()}}), flush=True)
    eng.close()
