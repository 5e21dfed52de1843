set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "protein or aa or lg or pinv or persite" 2>&1 | tail -4
for qd in 1 0; do for pat in 20000 200000; do
NRX_QUAD=$qd timeout -k 10 600 python scripts/kernel_rooflines.py --configs 4 --patterns $pat --md gpurun_out/r3l_roof_cfg4_${pat}_quad$qd.md > gpurun_out/r3l_roof_${pat}_quad$qd.log 2>&1
echo "quad $qd patterns $pat"; grep -E "evaluation:|sweep:|K3|K6|K4" gpurun_out/r3l_roof_cfg4_${pat}_quad$qd.md
done; done
