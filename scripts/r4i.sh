timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched" 2>&1 | tail -8
