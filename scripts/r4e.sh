for f in fused nofuse; do
if [ $f = nofuse ]; then export NRX_NO_AA_FUSED_LNL=1; else unset NRX_NO_AA_FUSED_LNL; fi
timeout -k 10 600 python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r4e_roof_cfg4_$f.md > gpurun_out/r4e_roof_$f.log 2>&1
echo "$f"; grep -A8 "full evaluation" gpurun_out/r4e_roof_cfg4_$f.md | grep -E "evaluation|K2|K3" | cut -c1-160
timeout -k 10 600 python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r4e_roof_cfg4_200k_$f.md > gpurun_out/r4e_roof_200k_$f.log 2>&1
grep -A8 "full evaluation" gpurun_out/r4e_roof_cfg4_200k_$f.md | grep -E "evaluation|K2|K3" | cut -c1-160
done
unset NRX_NO_AA_FUSED_LNL
timeout -k 10 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "protein or aa or lg or config or submodels or score_only_from or persite or batched" 2>&1 | tail -5
