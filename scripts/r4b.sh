ncu --set full --clock-control none --import-source on -k regex:"k_derivatives_dna4q|k_edge_sum_dna4q" -s 40 -c 4 -o gpurun_out/r4b_sweep_kernels -f python scripts/kernel_rooflines.py --configs 2 > gpurun_out/r4b_ncu.log 2>&1
ncu --set full --clock-control none -k regex:"k_term_lnl_sum" -s 4 -c 1 -o gpurun_out/r4b_k3f -f python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity --no-score-only > gpurun_out/r4b_ncu2.log 2>&1
ls -la gpurun_out/r4b_*.ncu-rep
