set -x
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for b in 1184; do NRX_AA2_BLOCKS=$b timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2i_roof_aa200k.md > gpurun_out/r2i_roof.log 2>&1; grep -E "full evaluation|derivative sweep|K2_clv|K5_sum" gpurun_out/r2i_roof_aa200k.md; done
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2i_roof_aa20k.md >> gpurun_out/r2i_roof.log 2>&1; grep -E "full evaluation|derivative sweep|K2_clv|K5_sum" gpurun_out/r2i_roof_aa20k.md
