# round-2 call A: sanity (GPU tests on HEAD) + the ncu evidence VERDICT r1 asked for on the protein kernels
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2a_pytest.log
NCU="ncu --set full --clock-control none --import-source on"
for pat in 20000 200000; do
  $NCU -k regex:k_aa20_dmma -s 30 -c 4 -o gpurun_out/r2a_aa_clv_$pat -f python scripts/sweep_only.py --config 4 --patterns $pat --mode eval --no-warmup > gpurun_out/r2a_ncu_clv_$pat.log 2>&1
  $NCU -k regex:"k_derivatives_pc|k_tree_lnl_pc" -s 12 -c 4 -o gpurun_out/r2a_aa_k6_$pat -f python scripts/sweep_only.py --config 4 --patterns $pat --mode sweep --no-warmup > gpurun_out/r2a_ncu_k6_$pat.log 2>&1
done
$NCU -k regex:k_aa20_dmma -s 120 -c 8 -o gpurun_out/r2a_aa_sweep_200000 -f python scripts/sweep_only.py --config 4 --patterns 200000 --mode sweep --no-warmup > gpurun_out/r2a_ncu_sweep.log 2>&1
python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2a_roof_aa20k.md > gpurun_out/r2a_roof.log 2>&1
python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2a_roof_aa200k.md >> gpurun_out/r2a_roof.log 2>&1
ls -la gpurun_out/
