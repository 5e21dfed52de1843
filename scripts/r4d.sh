timeout -k 10 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "shadow or lazy or config2_full_size_sweep" 2>&1 | tail -4
timeout -k 10 600 python scripts/kernel_rooflines.py --configs 2 --md gpurun_out/r4d_roof_cfg2.md > gpurun_out/r4d_roof.log 2>&1
grep -E "sweep" gpurun_out/r4d_roof_cfg2.md | cut -c1-170
python - <<'PY'
import json
for l in open('gpurun_out/r4d_roof.log'):
    if l.startswith('{'):
        d=json.loads(l); print(d['config2'].get('reroot_memo'))
PY
