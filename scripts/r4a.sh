( time timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "config2_full_size_sweep" 2>&1 | tail -12 ) 2>&1 | tail -16
