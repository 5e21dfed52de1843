set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk or fixtures or randomised or bitexact" 2>&1 | tail -4
python - <<'PY'
import time, json, bench
from netrax_b200.engine import NetraxB200
for c in (1, 2):
    cfg = dict(bench.CONFIGS[c])
    if c == 2: cfg["patterns"] = 10000
    net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    for _ in range(20): eng.computeLoglikelihood(0, 1)
    reps = 2000
    l0 = eng.launch_count(); t = time.perf_counter(); eng.timer_start()
    for _ in range(reps): eng.computeLoglikelihood(0, 1)
    ms = eng.timer_stop() / reps; wall = 1e3 * (time.perf_counter() - t) / reps
    eng.profile_enable(True)
    for _ in range(200): eng.computeLoglikelihood(0, 1)
    fam = eng.profile_read_all(); eng.profile_enable(False)
    print(json.dumps({"config": c, "patterns": cfg["patterns"], "device_us": 1e3 * ms, "wall_us": 1e3 * wall, "launches": (eng.launch_count() - l0) / reps,
                      "k2_us": 1e3 * fam.get("K2_clv_update", {}).get("ms", 0) / 200 if isinstance(fam, dict) else None}))
    eng.close()
PY
