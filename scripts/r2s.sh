set -x
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout -k 10 600 python bench.py --config 4 --no-configs --steps 10 > gpurun_out/r2s_bench_cfg4.json 2> gpurun_out/r2s_bench_cfg4.err; tail -2 gpurun_out/r2s_bench_cfg4.err; cut -c1-600 gpurun_out/r2s_bench_cfg4.json
timeout -k 10 600 python bench.py --config 1 --no-configs --steps 50 > gpurun_out/r2s_bench_cfg1.json 2> gpurun_out/r2s_bench_cfg1.err; cut -c1-400 gpurun_out/r2s_bench_cfg1.json
