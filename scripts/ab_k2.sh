# A/B harness for K2 launch-geometry knobs (env NRX_K2, NRX_K2_BLOCKS); prints ms/step and K2 roofline per setting
for b in ${BLOCKS:-296 444 888 1184 2368 4736}; do echo "== pipe blocks=$b"; NRX_K2_BLOCKS=$b python bench.py --steps 20 --no-cpu-baseline $EXTRA 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step',round(d['ms_per_step'],3),'value %.3e'%d['value'],'roof GB/s',round(r['achieved']),'frac',round(r['frac'],3),'share',round(r['share_of_step'],3))
    else: print(l.rstrip()[-300:])
"; done
