timeout -k 10 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
( time python bench.py > gpurun_out/r4f_bench.json 2> gpurun_out/r4f_bench.err ) 2>&1 | tail -4
tail -c 300 gpurun_out/r4f_bench.err
