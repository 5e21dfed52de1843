for qb in 148 296 444 592; do
NRX_QUAD_BLOCKS=$qb timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r4m_roof_cfg2_qb$qb.md > gpurun_out/r4m_roof_qb$qb.log 2>&1
echo "quad blocks $qb"; grep -A8 "derivative sweep:" gpurun_out/r4m_roof_cfg2_qb$qb.md | grep -E "sweep:|K6|K45" | cut -c1-140
done
