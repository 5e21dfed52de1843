set -x
timeout -k 10 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r3m_bench_n2.json 2> gpurun_out/r3m_bench_n2.err
tail -c 300 gpurun_out/r3m_bench_n2.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r3m_bench_n2.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','lnl','gpu_launches')}, d['e2e']['value'], d.get('parity'))
PY
