run() { python bench.py --steps 6 --no-cpu-baseline --no-configs --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('$1', round(d['ms_per_step'],3), d['gpu_launches'], round(d['roofline']['frac'],3), d['clocks']['sm_mhz'], d['lnl'])"; }
NRX_NODE=0 run "node=0"
NRX_NODE=1 run "node=1 maxc16"
NRX_NODE=1 NRX_NODE_MAXC=8 run "node=1 maxc8"
NRX_NODE=1 NRX_NODE_MAXC=12 run "node=1 maxc12"
NRX_NODE=1 NRX_NODE_MAXC=8 NRX_NODE_BLOCKS=7104 run "node=1 maxc8 blocks 7104"
NRX_NODE=1 NRX_NODE_MAXC=16 NRX_NODE_BLOCKS=1776 run "node=1 maxc16 blocks 1776"
