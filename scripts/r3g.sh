set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q -k "shadow or brlen or optimize or persite or fused or incremental" 2>&1 | tail -4
( time timeout -k 10 1200 python bench.py > gpurun_out/r3g_bench.json 2> gpurun_out/r3g_bench.err ) 2>&1 | tail -4
python scripts/sweep_host_profile.py
