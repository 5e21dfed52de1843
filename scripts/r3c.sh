set -x
timeout -k 10 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8
for qd in 1 0; do
NRX_QUAD=$qd timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r3c_roof_cfg2_quad$qd.md > gpurun_out/r3c_roof_quad$qd.log 2>&1
done
