NRX_WALK=1 timeout -k 10 600 ncu --set full --clock-control none --import-source on -k regex:walk -s 1 -c 1 -o gpurun_out/r2q_walk_cfg1 -f python scripts/sweep_only.py --config 1 --mode eval --no-warmup > gpurun_out/r2q_ncu.log 2>&1
tail -5 gpurun_out/r2q_ncu.log
