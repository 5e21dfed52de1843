set -x
timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "protein" 2>&1 | tail -3
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2l_roof_aa200k.md > gpurun_out/r2l_roof.log 2>&1; grep -E "derivative sweep|K4_edge" gpurun_out/r2l_roof_aa200k.md
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2l_roof_aa20k.md >> gpurun_out/r2l_roof.log 2>&1; grep -E "derivative sweep|K4_edge" gpurun_out/r2l_roof_aa20k.md
NRX_AA=v1 timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2l_roof_aa200k_v1.md > gpurun_out/r2l_roof.log 2>&1; grep -E "derivative sweep|K4_edge" gpurun_out/r2l_roof_aa200k_v1.md
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:"k_aa20_mma<2" -s 2 -c 3 -o gpurun_out/r2l_aa_edge_200000 -f python scripts/sweep_only.py --config 4 --patterns 200000 --mode sweep --no-warmup > gpurun_out/r2l_ncu.log 2>&1
