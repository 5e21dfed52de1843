python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err
python bench.py --impl reference --steps 5 > gpurun_out/r2x_ref.json 2>> gpurun_out/r2x_bench.err
python scripts/kernel_rooflines.py --configs 2,4 --md gpurun_out/r2x_roof.md > gpurun_out/r2x_roof.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r2x_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r2x_ncu_bench.log 2>&1
