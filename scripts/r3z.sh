timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "unlinked or multi_partition or config or scaled or golden_pmatrix or brlen_flow" 2>&1 | tail -5
python - <<'PY'
import json, time, bench
from netrax_b200.engine import NetraxB200
cfg = dict(bench.CONFIGS[3])
net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
for env in ("", "1"):
    import os
    if env: os.environ["NRX_NO_K1_MULTI"] = env
    else: os.environ.pop("NRX_NO_K1_MULTI", None)
    eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
    for _ in range(5): l = eng.computeLoglikelihood(0, 1)
    n0 = eng.launch_count(); eng.timer_start()
    for _ in range(50): l = eng.computeLoglikelihood(0, 1)
    ms = eng.timer_stop() / 50
    print(json.dumps({"no_multi": env, "ms_per_eval": ms, "launches": (eng.launch_count() - n0) / 50, "lnl": l}))
    eng.close()
PY
