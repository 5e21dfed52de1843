"""config 1, forty full evaluations (the ncu target of scripts/r3i.sh)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from netrax_b200.engine import NetraxB200  # noqa: E402
cfg = dict(bench.CONFIGS[1])
net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
for _ in range(40):
    eng.computeLoglikelihood(0, 1)
eng.close()
