python scripts/dbg_walk.py
NRX_WALK=1 timeout -k 10 300 compute-sanitizer --tool racecheck --print-limit 6 python scripts/dbg_walk.py 2>&1 | grep -v "^=========     \(Device\|Host\) Frame" | head -60
