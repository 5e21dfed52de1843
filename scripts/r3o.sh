set -x
timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "score_only or incremental or fused or batched" 2>&1 | tail -12
timeout -k 10 600 python bench.py --steps 10 --no-cpu-baseline --no-configs --no-parity > gpurun_out/r3o_bench.json 2> gpurun_out/r3o_bench.err
tail -c 400 gpurun_out/r3o_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3o_bench.json') if l.startswith('{')][-1])
print(d['ms_per_step'], d['value'], d.get('score_only'))
PY
