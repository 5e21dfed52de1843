import os, sys, numpy as np
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from netrax_b200._capi import Partition
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment
from netrax_b200.engine import NetraxB200
from oracle import oracle
for (n, r, pat, seed) in ((20, 1, 1000, 2), (12, 0, 33, 5)):
    net = random_network(n, r, seed=seed)
    m, w = simulate_alignment(net, pat, seed=seed)
    part = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    o = oracle.make_engine("ref", net, [part]); lo = o.computeLoglikelihood(0, 1)
    for mode in ("0", "1"):
        os.environ["NRX_WALK"] = mode
        g = NetraxB200(net, [part])
        for p in range(g.P): g.set_eigen(p, *o.get_eigen(p))
        a = g.computeLoglikelihood(0, 1); b = g.computeLoglikelihood(0, 1)
        e = net.num_edges // 2
        g.set_branch_length(e, 0.123); o.set_branch_length(e, 0.123)
        c = g.computeLoglikelihood(1, 1); co = o.computeLoglikelihood(1, 1)
        g.set_branch_length(e, float(net.edge_length[e])); o.set_branch_length(e, float(net.edge_length[e]))
        d = g.computeLoglikelihood(1, 1); do = o.computeLoglikelihood(1, 1)
        f = g.computeLoglikelihood(0, 1)
        print(n, r, pat, 'mode', mode, 'full', a, b, 'oracle', lo, '| changed', c, co, '| restored', d, do, '| full again', f)
        g.close()
