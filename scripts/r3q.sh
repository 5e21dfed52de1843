timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "lazy or shadow" 2>&1 | tail -25
timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r3q_roof_cfg2.md > gpurun_out/r3q_roof.log 2>&1
grep -E "sweep|K2_clv|K3_tree|K4" gpurun_out/r3q_roof_cfg2.md | cut -c1-160
