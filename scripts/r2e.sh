# round-2 call E: k_aa20_mma with the LDGSTS loader + single bulk store: parity, A/B vs v1, ncu
set -x
NRX_AA=v2 timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "protein or mixed or pinv or submodels or pseudo or baseline_configs or golden_protein" 2>&1 | tail -8 > gpurun_out/r2e_pytest_v2.log
cat gpurun_out/r2e_pytest_v2.log
for v in v1 v2; do
  NRX_AA=$v timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2e_roof_aa20k_$v.md > gpurun_out/r2e_roof_$v.log 2>&1
  NRX_AA=$v timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2e_roof_aa200k_$v.md >> gpurun_out/r2e_roof_$v.log 2>&1
done
for b in 296 1184 2368; do NRX_AA=v2 NRX_AA2_BLOCKS=$b timeout -k 10 300 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2e_roof_aa200k_v2_b$b.md >> gpurun_out/r2e_roof_v2.log 2>&1; done
NRX_AA=v2 timeout -k 10 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "test_protein_lg_g4_matches_oracle" 2>&1 | tail -15 > gpurun_out/r2e_memcheck.log
NRX_AA=v2 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:k_aa20_mma -s 30 -c 4 -o gpurun_out/r2e_aa_mma_200000 -f python scripts/sweep_only.py --config 4 --patterns 200000 --mode eval --no-warmup > gpurun_out/r2e_ncu.log 2>&1
