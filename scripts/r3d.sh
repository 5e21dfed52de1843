set -x
ncu --set full --clock-control none --import-source on -k regex:"k_derivatives_dna4q|k_edge_lnl_dna4q|k_sumtable_dna4" -s 60 -c 6 -o gpurun_out/r3d_k456 -f python scripts/kernel_rooflines.py --configs 2 > gpurun_out/r3d_ncu.log 2>&1
ls -la gpurun_out/r3d_k456.ncu-rep
