timeout -k 10 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_optimize.py -m gpu -x -q 2>&1 | tail -4
for z in 1 0; do echo "zero-copy $z"; NRX_ZEROCOPY=$z python scripts/sweep_host_profile.py; NRX_ZEROCOPY=$z bash scripts/r3h.sh 2>&1 | grep "^{"; done
