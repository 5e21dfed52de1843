set -x
timeout -k 10 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -4
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r2u_bench_n2.json 2> gpurun_out/r2u_bench_n2.err
tail -3 gpurun_out/r2u_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2u_bench_n2.json'))
print({k:d[k] for k in ['value','ms_per_step','lnl','n_gpus','gpu_launches']}); print(d['e2e']); print(d['parity']); print(d['roofline']['frac'], d['roofline']['traffic'])
PY
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --impl reference --steps 3 --warmup 3 2>/dev/null | cut -c1-300
