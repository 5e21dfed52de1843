timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "staged or async_tip" 2>&1 | tail -12
