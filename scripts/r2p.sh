set -x
timeout -k 10 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tile_walk or bitexact_clvs" 2>&1 | tail -5
for w in 1; do NRX_WALK=$w timeout -k 10 300 python scripts/kernel_rooflines.py --configs 1,2,3 --md gpurun_out/r2p_roof_walk$w.md > gpurun_out/r2p_roof_walk$w.log 2>&1; grep -E "full evaluation|K2_clv" gpurun_out/r2p_roof_walk$w.md; done
NRX_WALK=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_walk -s 5 -c 1 -o gpurun_out/r2p_walk_cfg1 -f python scripts/sweep_only.py --config 1 --mode eval --no-warmup > gpurun_out/r2p_ncu.log 2>&1
