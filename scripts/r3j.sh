set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "protein or aa or lg or config" 2>&1 | tail -4
for pdl in 1 0; do
NRX_PDL=$pdl timeout -k 10 600 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r3j_roof_cfg4_pdl$pdl.md > gpurun_out/r3j_roof_pdl$pdl.log 2>&1
grep -A6 "full evaluation" gpurun_out/r3j_roof_cfg4_pdl$pdl.md | grep -E "evaluation|K2"
NRX_PDL=$pdl timeout -k 10 600 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r3j_roof_cfg4_200k_pdl$pdl.md > gpurun_out/r3j_roof_200k_pdl$pdl.log 2>&1
grep -A6 "full evaluation" gpurun_out/r3j_roof_cfg4_200k_pdl$pdl.md | grep -E "evaluation|K2"
done
