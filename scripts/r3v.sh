set -x
python bench.py > gpurun_out/r3v_bench.json 2> gpurun_out/r3v_bench.err
python bench.py --impl reference --steps 5 > gpurun_out/r3v_ref.json 2>> gpurun_out/r3v_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -s 100 -c 400 --csv --log-file gpurun_out/r3v_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-configs --no-parity --no-score-only > gpurun_out/r3v_ncu_bench.log 2>&1
python scripts/kernel_rooflines.py --configs 2,4 --md gpurun_out/r3v_roof.md > gpurun_out/r3v_roof.log 2>&1
tail -c 300 gpurun_out/r3v_bench.err
