set -x
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2k_roof_aa200k.md > gpurun_out/r2k_roof.log 2>&1; grep -E "full evaluation|derivative sweep|K2_clv|K5_sum|K4_edge|K6_der|K3_tree" gpurun_out/r2k_roof_aa200k.md
timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2k_roof_aa20k.md >> gpurun_out/r2k_roof.log 2>&1; grep -E "full evaluation|derivative sweep|K2_clv|K5_sum|K4_edge|K6_der|K3_tree" gpurun_out/r2k_roof_aa20k.md
timeout -k 10 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_protein_brlen_flow_every_edge_tensor_core_kernels" 2>&1 | tail -3
