timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "brlen or derivative or one_pass or golden or pinv or optimize_branch or lazy or shadow" 2>&1 | tail -4
for r in 1 0; do
NRX_K6_RING=$r timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r4c_roof_cfg2_ring$r.md > gpurun_out/r4c_roof_ring$r.log 2>&1
echo "ring=$r"; grep -E "sweep:|K6" gpurun_out/r4c_roof_cfg2_ring$r.md | head -4 | cut -c1-150
done
