set -x
timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "protein or pinv" 2>&1 | tail -3
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2m_roof_aa200k.md > gpurun_out/r2m_roof.log 2>&1; grep -E "derivative sweep|K4_edge" gpurun_out/r2m_roof_aa200k.md
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2m_roof_aa20k.md >> gpurun_out/r2m_roof.log 2>&1; grep -E "derivative sweep|K4_edge" gpurun_out/r2m_roof_aa20k.md
timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_aa20_mma -s 4 -c 6 -o gpurun_out/r2m_aa_edge_200000 -f python scripts/sweep_only.py --config 4 --patterns 200000 --mode sweep --no-warmup > gpurun_out/r2m_ncu.log 2>&1
