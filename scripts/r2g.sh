set -x
NRX_AA=v2 timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "protein or mixed" 2>&1 | tail -4
for b in 888 1776; do NRX_AA=v2 NRX_AA2_BLOCKS=$b timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2g_roof_aa200k_v2_b$b.md > gpurun_out/r2g_roof_v2.log 2>&1; grep -E "full evaluation|K2_clv" gpurun_out/r2g_roof_aa200k_v2_b$b.md; done
NRX_AA=v2 timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2g_roof_aa20k_v2.md >> gpurun_out/r2g_roof_v2.log 2>&1; grep -E "full evaluation|K2_clv" gpurun_out/r2g_roof_aa20k_v2.md
NRX_AA=v2 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_aa20_mma -s 30 -c 4 -o gpurun_out/r2g_aa_mma_200000 -f python scripts/sweep_only.py --config 4 --patterns 200000 --mode eval --no-warmup > gpurun_out/r2g_ncu.log 2>&1
