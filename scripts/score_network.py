#!/usr/bin/env python
"""netrax --score_only on the B200 engine: ``python scripts/score_network.py --msa aln.fasta --start_network net.nw
--model GTR+G`` (option names of src/main.cpp:40-62; --model also takes a partition file)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--msa", required=True)
    ap.add_argument("--start_network", required=True)
    ap.add_argument("--model", default="GTR{1/2.5/0.8/1.2/3.0/1}+FC+G", help="model string or partition file (ML-estimated rates / frequencies, e.g. plain GTR+G, need --allow-unoptimised)")
    ap.add_argument("--best_displayed_tree_variant", action="store_true", help="LikelihoodVariant::BEST_DISPLAYED_TREE")
    ap.add_argument("--brlen", default="linked", choices=["linked", "scaled", "unlinked"])
    ap.add_argument("--no_optimize", action="store_true", help="score the network as given (model parameters are still optimised)")
    ap.add_argument("--allow-unoptimised", action="store_true",
                    help="score although the model has ML-estimated rates / frequencies, which stay at their start values here "
                         "(the reference fits them with L-BFGS-B: lnL / BIC then differ from ./netrax)")
    ap.add_argument("--device", type=int, default=0)
    ap.add_argument("--json", action="store_true")
    args = ap.parse_args()
    from netrax_b200._capi import AVERAGE, BEST, LINKED, SCALED, UNLINKED
    from netrax_b200.engine import NetraxB200
    from netrax_b200.score import score_only
    model = open(args.model).read() if os.path.exists(args.model) else args.model
    res = score_only(lambda net, parts, **kw: NetraxB200(net, parts, device=args.device, **kw), open(args.start_network).read(),
                     open(args.msa).read(), model, variant=BEST if args.best_displayed_tree_variant else AVERAGE,
                     linkage={"linked": LINKED, "scaled": SCALED, "unlinked": UNLINKED}[args.brlen], optimize=not args.no_optimize,
                     log=None if args.json else print, allow_unoptimised_ml_params=args.allow_unoptimised)
    if args.json:
        print(json.dumps(res))


if __name__ == "__main__":
    main()
