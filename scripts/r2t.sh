set -x
timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "category" 2>&1 | tail -5
timeout -k 10 600 python scripts/cats_roofline.py --md gpurun_out/r2t_cats_roofline.md 2>&1 | tail -8
timeout -k 10 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
