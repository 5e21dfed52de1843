"""Multi-GPU parity (-m gpu, needs >= 2 devices; skipped on a 1-GPU box): pattern shards + in-engine NCCL
all-reduce reproduce the single-GPU results.  The body runs under torchrun in scripts/check_multi_gpu.py."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpu_sharded_equals_single_gpu():
    import netrax_b200.engine as eng
    if eng.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29533", os.path.join(ROOT, "scripts", "check_multi_gpu.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("{")][-1]
    rep = json.loads(line)
    assert rep["ok"] and rep["world"] == 2
    for c in rep["cases"]:
        assert c["rel_diff"] <= (1e-9 if c["variant"] == "callers" else 1e-12)   # "callers": lnL after a Brent optimisation of alpha
    assert rep["cases"][-1]["variant"] == "callers"
