timeout -k 10 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${NG:-8} --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus ${NG:-8} --steps 20 --warmup 3 > gpurun_out/r4g_bench_n8.json 2> gpurun_out/r4g_bench_n8.err
tail -c 300 gpurun_out/r4g_bench_n8.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4g_bench_n8.json') if l.startswith('{')][-1])
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','lnl','gpu_launches')}, d['e2e'], d.get('score_only',{}).get('ms_per_step'), d['parity']['pass'], d['clocks'])
PY
