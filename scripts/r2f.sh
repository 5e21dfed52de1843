set -x
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/dmma_bench scripts/micro/dmma_bench.cu && timeout 120 /tmp/dmma_bench > gpurun_out/r2f_dmma_bench.txt 2>&1
