# A/B harness for the protein DMMA kernel's block-count target (env NRX_AA_BLOCKS) at config-4 size and at 200k patterns
python -m pytest tests -m gpu -x -q -k protein 2>&1 | tail -2
for b in ${BLOCKS:-444 888 1776 3552}; do for pat in 20000 200000; do echo "== aa blocks=$b patterns=$pat"; NRX_AA_BLOCKS=$b python scripts/bench_configs.py --configs 4 --no-cpu --patterns $pat 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l)['config4']['gpu']; print('ms/eval',round(d['ms_per_lnl_eval'],3),'site-updates/s %.3e'%d['site_updates_per_s'])
    else: print(l.rstrip()[-300:])
"; done; done
