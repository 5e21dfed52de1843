import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import bench
from netrax_b200.engine import NetraxB200
cfg = dict(bench.CONFIGS[4]); cfg["patterns"] = int(os.environ.get("PAT", "20000"))
net, parts, brl = bench.make_inputs(cfg, cfg["patterns"])
eng = NetraxB200(net, parts, variant=cfg["variant"], linkage=cfg["linkage"], partition_brlens=brl)
lib = eng.api.lib
lib.nrxh_engine.restype = C.c_void_p
e = lib.nrxh_engine(eng.h)
nrx = C.CDLL(os.path.join(os.path.dirname(bench.__file__), "netrax_b200", "libnrx_engine.so"))
nrx.nrx_supports_fused_lnl.argtypes = [C.c_void_p]
print("supports_fused_lnl", nrx.nrx_supports_fused_lnl(C.c_void_p(e)))
for i in range(3):
    print(eng.computeLoglikelihood(0, 1))
eng.profile_enable(True)
eng.computeLoglikelihood(0, 1)
print({k: v["launches"] for k, v in eng.profile_read_all().items() if v["launches"]})
