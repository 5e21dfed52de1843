set -x
timeout -k 10 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 --patterns 200000 > gpurun_out/r2v_bench_n2.json 2> gpurun_out/r2v_bench_n2.err
echo "exit code $?"
wc -c gpurun_out/r2v_bench_n2.json gpurun_out/r2v_bench_n2.err
tail -20 gpurun_out/r2v_bench_n2.err
