for b in 296 444 592; do
for pat in 20000 200000; do
NRX_AA2_BLOCKS=$b timeout -k 10 600 python scripts/kernel_rooflines.py --configs 4 --patterns $pat --md gpurun_out/r3k_roof_cfg4_${pat}_b$b.md > gpurun_out/r3k_roof_${pat}_b$b.log 2>&1
echo "blocks $b patterns $pat"; grep -A6 "full evaluation" gpurun_out/r3k_roof_cfg4_${pat}_b$b.md | grep -E "evaluation|K2"
done; done
