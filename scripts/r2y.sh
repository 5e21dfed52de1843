set -x
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "node_centric or headline or bitexact_clvs or randomised" 2>&1 | tail -6
for nm in 0 1; do NRX_NODE=$nm timeout -k 10 600 python bench.py --steps 10 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r2y_bench_node$nm.json 2> gpurun_out/r2y_bench_node$nm.err; python - <<PY
import json
d=json.loads([l for l in open('gpurun_out/r2y_bench_node$nm.json') if l.startswith('{')][-1])
print('NODE=$nm', d['ms_per_step'], d['value'], d['lnl'], d['gpu_launches'], d['roofline']['frac'], d['clocks'])
PY
done
NRX_NODE=1 NRX_WALK=0 timeout -k 10 300 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "node_centric" 2>&1 | tail -3
