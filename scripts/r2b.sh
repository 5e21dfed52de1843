# round-2 call B: per-site parity tests + new bench line (configs / parity blocks) + DMMA peak + K2 DRAM traffic at 1 M patterns
set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2b_pytest.log
scripts/bin/dmma_bench > gpurun_out/r2b_dmma_bench.txt 2>&1
python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err
python bench.py --impl reference --steps 5 > gpurun_out/r2b_ref.json 2>> gpurun_out/r2b_bench.err
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:k_clv_dna4 -s 48 -c 16 --csv --log-file gpurun_out/r2b_k2_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r2b_ncu_bench.log 2>&1
python scripts/kernel_rooflines.py --configs 4 --md gpurun_out/r2b_roof_aa20k.md > gpurun_out/r2b_roof.log 2>&1
python scripts/kernel_rooflines.py --configs 4 --patterns 200000 --md gpurun_out/r2b_roof_aa200k.md >> gpurun_out/r2b_roof.log 2>&1
python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r2b_roof_cfg2.md >> gpurun_out/r2b_roof.log 2>&1
