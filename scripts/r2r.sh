set -x
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_walk or bitexact_clvs or reference_fixtures or incremental" 2>&1 | tail -5
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for w in 1; do NRX_WALK=$w timeout -k 10 300 python scripts/kernel_rooflines.py --configs 1,2,3 --md gpurun_out/r2r_roof_walk$w.md > gpurun_out/r2r_roof_walk$w.log 2>&1; grep -E "full evaluation|K2_clv" gpurun_out/r2r_roof_walk$w.md; done
NRX_WALK=1 timeout -k 10 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_walk" 2>&1 | tail -3
