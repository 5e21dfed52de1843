set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "shadow or brlen or optimize or sweep or reroot" 2>&1 | tail -15
timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r3b_roof_cfg2.md > gpurun_out/r3b_roof.log 2>&1
tail -3 gpurun_out/r3b_roof.log | cut -c1-300
