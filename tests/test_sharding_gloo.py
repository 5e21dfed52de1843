"""N>1 site sharding on CPU: world_size-2 `gloo` process group, each rank owns a contiguous pattern slice of every
partition and runs the (CPU) oracle engine with the reference's parallel_reduce_cb hooked to an all-reduce — the
host-side contract the GPU engine's in-engine NCCL all-reduce implements.  Checks: the sharded result
equals the unsharded one (lnL, per-partition lnL, branch-length derivatives), all ranks agree bit-for-bit, and the
mixing happens per PARTITION after the reduce, for AVERAGE/linked and BEST/unlinked."""
import os
import sys

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _case(variant, linkage, nparts):
    from netrax_b200._capi import UNLINKED, Partition
    from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment
    net = random_network(12, 2, seed=31)
    rng = np.random.default_rng(1)
    parts, brl = [], []
    for p in range(nparts):
        m, w = simulate_alignment(net, 301 + 64 * p, seed=40 + p)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES * (1 + 0.2 * p), GAMMA4_ALPHA05, pattern_weights=w))
        brl.append(net.edge_length * rng.uniform(0.5, 2.0, net.num_edges))
    return net, parts, (brl if linkage == UNLINKED else None)


def _evaluate(eng, net):
    lnl = eng.computeLoglikelihood(0, 1)
    e = int(net.ret_first_edge[0])
    eng.brlen_prepare(e)
    lb = eng.computeLoglikelihoodBrlenOpt(e)
    eng.computePartitionSumtables(e)
    d = eng.computeLoglikelihoodDerivatives(e)
    lf = eng.brlen_finish(e)
    return np.concatenate([[lnl, lb, lf, d[0], d[1]], eng.partition_loglh(), np.ravel(d[2]), np.ravel(d[3])])


def _scalers_case():
    from netrax_b200._capi import Partition
    from netrax_b200.synth import DNA_FREQS, GAMMA4_ALPHA05, GTR_RATES, random_network, simulate_alignment
    net = random_network(9, 2, seed=5)
    parts = []
    for i in range(3):
        m, w = simulate_alignment(net, 150 + 40 * i, seed=5 + i)
        parts.append(Partition(4, 4, m, DNA_FREQS, GTR_RATES, GAMMA4_ALPHA05, pattern_weights=w))
    return net, parts


def _optimize_scalers(eng, parts_full):
    """optimize_scalers + the alpha and pinv steps under scaled linkage: the callers whose site-sharded form needs more than the lnL
    all-reduce (pllmod_treeinfo_normalize_brlen_scalers sums scaler x pattern_weight_sum over the shards; the Brent
    drivers all-reduce their convergence flag)."""
    for p, s in enumerate([3.0, 0.3, 150.0]):
        eng.set_brlen_scaler(p, s)
    eng.set_alpha(1, 1.3)
    eng.set_scoring_sizes(9, int(sum(int(p.pattern_weights.sum()) for p in parts_full)))
    bic = eng.optimize_scalers()
    la = eng.optimize_alpha()
    eng.set_pinv(2, 0.1)    # +I: every shard finds the invariant patterns of its own slice
    lp = eng.optimize_pinv()
    return np.concatenate([[bic, la, lp, eng.get_alpha(1), eng.get_pinv(2)], eng.brlen_scalers(), eng.branch_lengths()])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    from netrax_b200._capi import AVERAGE, BEST, LINKED, UNLINKED
    from oracle import oracle
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)

    def reduce(arr):
        t = torch.from_numpy(arr)   # shares memory: the all-reduce writes the sums in place
        dist.all_reduce(t)

    out = {}
    for name, variant, linkage, nparts in (("avg", AVERAGE, LINKED, 1), ("best", BEST, UNLINKED, 3)):
        net, parts, brl = _case(variant, linkage, nparts)
        shard = [p.slice(rank * p.sites // world, (rank + 1) * p.sites // world) for p in parts]
        eng = oracle.make_engine("port", net, shard, variant=variant, linkage=linkage, partition_brlens=brl)
        eng.set_reduce(reduce)
        out[name] = _evaluate(eng, net)
        eng.close()
    from netrax_b200._capi import SCALED
    net, parts = _scalers_case()
    shard = [p.slice(rank * p.sites // world, (rank + 1) * p.sites // world) for p in parts]
    eng = oracle.make_engine("port", net, shard, linkage=SCALED)
    eng.set_reduce(reduce)
    out["scalers"] = _optimize_scalers(eng, parts)
    eng.close()
    q.put((rank, out))
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_sharded_equals_unsharded():
    from netrax_b200._capi import AVERAGE, BEST, LINKED, UNLINKED
    from oracle import oracle
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=240) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for name, variant, linkage, nparts in (("avg", AVERAGE, LINKED, 1), ("best", BEST, UNLINKED, 3)):
        net, parts, brl = _case(variant, linkage, nparts)
        eng = oracle.make_engine("port", net, parts, variant=variant, linkage=linkage, partition_brlens=brl)
        want = _evaluate(eng, net)
        assert np.array_equal(res[0][name], res[1][name])             # all ranks hold identical global values
        np.testing.assert_allclose(res[0][name][:3 + 2 + nparts], want[:3 + 2 + nparts], rtol=1e-12)   # lnLs: sum order only
        np.testing.assert_allclose(res[0][name], want, rtol=1e-8, atol=1e-8)                         # derivatives
    from netrax_b200._capi import SCALED
    net, parts = _scalers_case()
    eng = oracle.make_engine("port", net, parts, linkage=SCALED)
    want = _optimize_scalers(eng, parts)
    assert np.array_equal(res[0]["scalers"], res[1]["scalers"])
    np.testing.assert_allclose(res[0]["scalers"][:3], want[:3], rtol=1e-9)      # BIC, lnL after the alpha step, lnL after the pinv step
    np.testing.assert_allclose(res[0]["scalers"][3:], want[3:], rtol=1e-4, atol=1e-6)   # alpha, pinv, scalers (mean 1 over ALL sites), branch lengths


def test_partition_slice_covers_every_pattern_once():
    from netrax_b200._capi import AVERAGE, LINKED
    net, parts, _ = _case(AVERAGE, LINKED, 1)
    p = parts[0]
    for world in (2, 3, 8):
        sl = [p.slice(r * p.sites // world, (r + 1) * p.sites // world) for r in range(world)]
        assert sum(s.sites for s in sl) == p.sites
        assert np.array_equal(np.concatenate([s.tip_masks for s in sl], axis=1), p.tip_masks)
        assert np.array_equal(np.concatenate([s.pattern_weights for s in sl]), p.pattern_weights)
