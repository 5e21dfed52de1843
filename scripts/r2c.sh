# round-2 call C: DMMA peak microbenchmark + the warp-specialised protein kernel k_aa20_mma (NRX_AA=v2): parity, then A/B against v1
set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_bench scripts/micro/dmma_bench.cu && timeout 120 /tmp/dmma_bench > gpurun_out/r2c_dmma_bench.txt 2>&1
NRX_AA=v2 timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "protein or mixed or pinv or submodels or pseudo or baseline_configs or golden_protein" 2>&1 | tail -8 > gpurun_out/r2c_pytest_v2.log
cat gpurun_out/r2c_pytest_v2.log
for v in v1 v2; do
  NRX_AA=$v timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2c_roof_aa20k_$v.md > gpurun_out/r2c_roof_$v.log 2>&1
  NRX_AA=$v timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2c_roof_aa200k_$v.md >> gpurun_out/r2c_roof_$v.log 2>&1
done
for b in 296 592 1184; do NRX_AA=v2 NRX_AA2_BLOCKS=$b timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2c_roof_aa200k_v2_b$b.md >> gpurun_out/r2c_roof_v2.log 2>&1; done
NRX_AA=v2 timeout -k 10 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_protein_lg_g4_matches_oracle" 2>&1 | tail -15 > gpurun_out/r2c_memcheck.log
