for z in 0 1; do echo "ZEROCOPY=$z"; NRX_ZEROCOPY=$z timeout -k 10 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batched_scoring" 2>&1 | grep -E "^E|passed|failed" | head -20; done
