timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "walk or fixtures or randomised or bitexact or fused or incremental or score_only" 2>&1 | tail -4
bash scripts/r3h.sh 2>&1 | grep "^{"
timeout -k 10 600 python bench.py --steps 10 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print(d['ms_per_step'], {k:(round(v['ms'],3), round(v['frac_of_hbm_peak'],3)) for k,v in d['kernel_families'].items()}, d['score_only']['ms_per_step'])"
