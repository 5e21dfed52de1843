ncu --set full --clock-control none --import-source on -k regex:k_walk_dna4 -s 30 -c 1 -o gpurun_out/r3i_walk -f python scripts/_walk_cfg1.py > gpurun_out/r3i_ncu.log 2>&1
