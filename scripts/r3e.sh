set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "shadow or brlen or optimize or persite or pinv or golden" 2>&1 | tail -5
for qb in 296 592 1184 2368; do
NRX_QUAD_BLOCKS=$qb timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r3e_roof_cfg2_qb$qb.md > gpurun_out/r3e_roof_qb$qb.log 2>&1
grep -A9 "derivative sweep:" gpurun_out/r3e_roof_cfg2_qb$qb.md | grep -E "sweep:|K6|K4"
done
