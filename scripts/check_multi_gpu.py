#!/usr/bin/env python
"""Multi-GPU parity check (run under torchrun, one rank per GPU): the pattern-sharded engine with the NCCL
communicator attached must reproduce the single-GPU result on the same alignment — network lnL, every per-tree
partition lnL, and the branch-length derivatives — to the all-reduce's rounding (<= 1e-12 relative), and BEST-tree
selection must be identical.  Rank 0 prints one JSON line."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from netrax_b200._capi import AVERAGE, BEST, LINKED, UNLINKED, Partition  # noqa: E402
from netrax_b200.engine import NetraxB200, comm_unique_id  # noqa: E402
from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment  # noqa: E402


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))

    def fresh_uid():  # one NCCL unique id per communicator (an id must not be reused)
        uid = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{lr}")
        if rank == 0:
            uid.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(uid, 0)
        return uid.cpu().numpy().tobytes()

    report = {"world": world, "cases": []}
    for variant, linkage, nparts in ((AVERAGE, LINKED, 1), (BEST, UNLINKED, 3)):
        net = random_network(24, 3, seed=77)
        rng = np.random.default_rng(3)
        full, brl = [], []
        for p in range(nparts):
            # odd sizes: ragged shards; the last partition of the 3-partition case has ONE pattern, so rank 0 owns an
            # empty slice of it (the reference's "skip remote partitions" case, LH/ImprovedLoglikelihood.cpp:128-131)
            m, w = simulate_alignment(net, 1 if (nparts == 3 and p == 2) else 3001 + 517 * p, seed=50 + p)
            full.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES * (1 + 0.1 * p), GAMMA4_ALPHA05, pattern_weights=w))
            brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
        brl = brl if linkage == UNLINKED else None
        shard = [q.slice(rank * q.sites // world, (rank + 1) * q.sites // world) for q in full]
        g = NetraxB200(net, shard, variant=variant, linkage=linkage, device=lr, partition_brlens=brl, comm=(fresh_uid(), rank, world))
        lnl = g.computeLoglikelihood(0, 1)
        trees = np.array([g.tree_info(net.root, t)[1] for t in range(g.num_trees(net.root))])
        e = int(net.ret_first_edge[0])
        g.brlen_prepare(e)
        lb = g.computeLoglikelihoodBrlenOpt(e)
        g.computePartitionSumtables(e)
        d = g.computeLoglikelihoodDerivatives(e)
        lf = g.brlen_finish(e)
        case = {"variant": int(variant), "lnl": lnl}
        if rank == 0:
            s = NetraxB200(net, full, variant=variant, linkage=linkage, device=lr, partition_brlens=brl)
            lnl1 = s.computeLoglikelihood(0, 1)
            trees1 = np.array([s.tree_info(net.root, t)[1] for t in range(s.num_trees(net.root))])
            s.brlen_prepare(e)
            lb1 = s.computeLoglikelihoodBrlenOpt(e)
            s.computePartitionSumtables(e)
            d1 = s.computeLoglikelihoodDerivatives(e)
            lf1 = s.brlen_finish(e)
            assert abs(lnl - lnl1) <= 1e-12 * abs(lnl1), (lnl, lnl1)
            np.testing.assert_allclose(trees, trees1, rtol=1e-12)
            assert abs(lb - lb1) <= 1e-12 * abs(lb1) and abs(lf - lf1) <= 1e-12 * abs(lf1)
            np.testing.assert_allclose(d[2], d1[2], rtol=1e-9, atol=1e-9)
            np.testing.assert_allclose(d[3], d1[3], rtol=1e-9, atol=1e-9)
            if variant == BEST:
                assert [int(np.argmax(trees[:, p])) for p in range(nparts)] == [int(np.argmax(trees1[:, p])) for p in range(nparts)]
            case.update({"lnl_single_gpu": lnl1, "rel_diff": abs(lnl - lnl1) / abs(lnl1), "d1": d[0], "d1_single_gpu": d1[0]})
            s.close()
        # every rank must hold the same global value (the all-reduce is in-engine)
        t = torch.tensor([lnl, -lnl], dtype=torch.float64, device=f"cuda:{lr}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        assert float(t[0]) == lnl and float(-t[1]) == lnl, "ranks disagree on the network lnL"
        g.close()
        report["cases"].append(case)
    # ---- the callers under sharding: alpha optimisation (every Brent iterate all-reduces per-partition lnLs; the
    # "all converged" flag travels through the same reduction), pseudo-likelihood, batched scoring of two networks
    from netrax_b200._capi import SARAH_PSEUDO
    from netrax_b200.engine import compute_loglikelihood_batch
    net = random_network(16, 2, seed=91)
    m, w = simulate_alignment(net, 2001, seed=91)
    full = Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w)
    shard = full.slice(rank * full.sites // world, (rank + 1) * full.sites // world)
    g = NetraxB200(net, [shard], device=lr, comm=(fresh_uid(), rank, world))
    g.set_alpha(0, 2.0)
    la = g.optimize_alpha()
    alpha = g.get_alpha(0)
    gp = NetraxB200(net, [shard], variant=SARAH_PSEUDO, device=lr, comm=(fresh_uid(), rank, world))
    lp = gp.computeLoglikelihood(0, 1)
    net2 = random_network(16, 3, seed=92)
    g2 = NetraxB200(net2, [shard], device=lr, comm=(fresh_uid(), rank, world))
    lb2 = compute_loglikelihood_batch([g, g2], 0, 1)
    case = {"variant": "callers", "alpha": alpha}
    if rank == 0:
        s = NetraxB200(net, [full], device=lr)
        s.set_alpha(0, 2.0)
        la1 = s.optimize_alpha()
        assert abs(la - la1) <= 1e-9 * abs(la1) and abs(alpha - s.get_alpha(0)) <= 1e-5 * alpha, (la, la1, alpha, s.get_alpha(0))
        sp = NetraxB200(net, [full], variant=SARAH_PSEUDO, device=lr)
        lp1 = sp.computeLoglikelihood(0, 1)
        assert abs(lp - lp1) <= 1e-12 * abs(lp1), (lp, lp1)
        s2 = NetraxB200(net2, [full], device=lr)
        want = [s.computeLoglikelihood(0, 1), s2.computeLoglikelihood(0, 1)]
        np.testing.assert_allclose(lb2, want, rtol=1e-12)
        case.update({"lnl": la, "lnl_single_gpu": la1, "rel_diff": abs(la - la1) / abs(la1), "pseudo": lp, "batched": [float(x) for x in lb2]})
        for x in (s, sp, s2):
            x.close()
    for x in (g, gp, g2):
        x.close()
    # optimize_scalers under sharding: the normalisation sums scaler x pattern_weight_sum over the shards (host values,
    # all-reduced through the engine's communicator)
    from netrax_b200._capi import SCALED
    net = random_network(9, 2, seed=5)
    fulls = []
    for i in range(3):
        m, w = simulate_alignment(net, 150 + 40 * i, seed=5 + i)
        fulls.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    total_sites = int(sum(int(p.pattern_weights.sum()) for p in fulls))

    def run_scalers(eng):
        for p, sc in enumerate([3.0, 0.3, 150.0]):
            eng.set_brlen_scaler(p, sc)
        eng.set_scoring_sizes(9, total_sites)
        return eng.optimize_scalers(), eng.brlen_scalers(), eng.branch_lengths()

    gs = NetraxB200(net, [p.slice(rank * p.sites // world, (rank + 1) * p.sites // world) for p in fulls], linkage=SCALED, device=lr,
                    comm=(fresh_uid(), rank, world))
    bic, scal, brl = run_scalers(gs)
    if rank == 0:
        s = NetraxB200(net, fulls, linkage=SCALED, device=lr)
        bic1, scal1, brl1 = run_scalers(s)
        assert abs(bic - bic1) <= 1e-9 * abs(bic1), (bic, bic1)
        np.testing.assert_allclose(scal, scal1, rtol=1e-5)
        np.testing.assert_allclose(brl, brl1, rtol=1e-5)
        case.update({"scalers": [float(x) for x in scal], "scalers_single_gpu": [float(x) for x in scal1], "bic": bic, "bic_single_gpu": bic1})
        s.close()
    gs.close()
    report["cases"].append(case)
    if rank == 0:
        report["ok"] = True
        print(json.dumps(report))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
