python -m pytest tests -m gpu -x -q 2>&1 | tail -5
for v in 0 22 24 15 42; do echo "== NRX_K2=$v"; NRX_K2=$v python bench.py --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step',round(d['ms_per_step'],3),'value %.3e'%d['value'],'roof GB/s',round(r['achieved']),'frac',round(r['frac'],3),'share',round(r['share_of_step'],3),'e2e %.3e'%d['e2e']['value'])
    else: print(l.rstrip()[-300:])
"; done
for b in 148 444 592; do echo "== pipe blocks=$b"; NRX_K2_BLOCKS=$b python bench.py --steps 20 --no-cpu-baseline 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d['roofline']; print('ms/step',round(d['ms_per_step'],3),'value %.3e'%d['value'],'roof GB/s',round(r['achieved']),'frac',round(r['frac'],3))
    else: print(l.rstrip()[-300:])
"; done
