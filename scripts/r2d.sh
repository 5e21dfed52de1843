# round-2 call D: DMMA peak; ncu of k_aa20_mma; fused last-block reduction: full GPU suite + A/B on the config-2 sweep
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_bench scripts/micro/dmma_bench.cu && timeout 120 /tmp/dmma_bench > gpurun_out/r2d_dmma_bench.txt 2>&1
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2d_pytest.log
cat gpurun_out/r2d_pytest.log
for f in 1 0; do NRX_FUSE_REDUCE=$f timeout -k 10 300 python scripts/kernel_rooflines.py --configs 1,2 --md gpurun_out/r2d_roof_fuse$f.md > gpurun_out/r2d_roof_fuse$f.log 2>&1; done
NRX_AA=v2 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_aa20_mma -s 30 -c 4 -o gpurun_out/r2d_aa_mma_200000 -f python scripts/sweep_only.py --config 4 --patterns 200000 --mode eval --no-warmup > gpurun_out/r2d_ncu.log 2>&1
