timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "one_pass or optimize_branch or lazy or shadow" 2>&1 | tail -15
for sep in "" 1; do echo "separate=$sep"; NRX_BENCH_SEPARATE_K4_K5=$sep python scripts/sweep_host_profile.py; done
timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r3w_roof_cfg2.md > gpurun_out/r3w_roof.log 2>&1
grep -E "sweep|K45|K4_|K5_|K6_|K2_" gpurun_out/r3w_roof_cfg2.md | cut -c1-150
