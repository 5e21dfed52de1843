( time python bench.py > gpurun_out/r4l_bench.json 2> gpurun_out/r4l_bench.err ) 2>&1 | tail -4
tail -c 200 gpurun_out/r4l_bench.err
python - <<'PY'
import json
d=json.loads([l for l in open('gpurun_out/r4l_bench.json') if l.startswith('{')][-1])
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','lnl')}, d['roofline']['frac'], d['e2e']['value'], d['cpu_baseline']['value'], d['score_only']['ms_per_step'], d['parity']['pass'])
for k,v in d['configs'].items():
    print(k, v.get('error'), round(v.get('ms_per_eval',0),4), v.get('parity',{}).get('pass'), [ (s, round(v[s]['ms'],1), v[s]['launches']) for s in ('derivative_sweep','derivative_sweep_accept','derivative_sweep_accept_lazy') if s in v])
PY
