set -x
( time timeout -k 10 1200 python bench.py > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err ) 2>&1 | tail -4
tail -c 600 gpurun_out/r3f_bench.err
