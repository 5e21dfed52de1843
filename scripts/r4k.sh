timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "fixtures or brlen_flow or reticulation or batched or headline or optimize_all or multi_partition" 2>&1 | tail -4
python - <<'PY'
import json, time, bench
from netrax_b200.engine import NetraxB200
cfg = dict(bench.CONFIGS[5])
for pat in (2048, 125000):
    net, parts, brl = bench.make_inputs(cfg, pat)
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    for _ in range(5): l = eng.computeLoglikelihood(0, 1)
    t = time.perf_counter(); eng.timer_start()
    for _ in range(50): l = eng.computeLoglikelihood(0, 1)
    ms = eng.timer_stop() / 50; wall = 1e3 * (time.perf_counter() - t) / 50
    print(json.dumps({"patterns": pat, "ms_per_eval": ms, "wall_ms": wall, "lnl": l}))
    eng.close()
PY
