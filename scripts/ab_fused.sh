# A/B: fused K3 epilogue vs separate K3 (env NRX_NO_FUSED_LNL), alternating runs to average out box power-cap noise
for i in 1 2; do for v in "" 1; do echo "== NRX_NO_FUSED_LNL=$v"; env ${v:+NRX_NO_FUSED_LNL=1} python bench.py --no-cpu-baseline --steps 20 2>&1 | python -c "
import sys,json,os
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step',round(d['ms_per_step'],3),'k2 ms/step',round(r['avg_launch_ms']*r['launches']/d['steps'],3),'frac',round(r['frac'],3),'clk',d['clocks'])
"; done; done
