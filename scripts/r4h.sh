( time python bench.py --no-cpu-baseline > gpurun_out/r4h_bench.json 2> gpurun_out/r4h_bench.err ) 2>&1 | tail -4
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4h_bench.json') if l.startswith('{')][-1])
v=d['configs']['4']
print(v['ms_per_eval'], v['derivative_sweep']['ms'], v['derivative_sweep']['launches'], {f:round(t['frac_of_hbm_peak'],2) for f,t in v['derivative_sweep']['kernel_families'].items()})
PY
