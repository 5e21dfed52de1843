"""``--score_only`` of the reference's CLI (src/main.cpp:287-326) on top of the likelihood API: read a network and an
alignment from files, optimise the model (optimizeModel), report BIC / lnL, run optimizeAllNonTopology(SLOW), report BIC /
lnL / AIC / AICc and the optimised network.  The caller supplies the engine factory — ``scripts/score_network.py`` passes
``NetraxB200`` (the CUDA engine; there is no CPU path in this package)."""
from __future__ import annotations

import math
from typing import Callable, Dict, Optional

from ._capi import AVERAGE, LINKED, SCALED, UNLINKED
from .msa_io import apply_model_state, build_partitions, read_msa
from .network_io import parse_extended_newick


def param_count(eng, model_params: int, linkage: int) -> int:
    """get_param_count (src/likelihood/ComplexityScoring.cpp:32-47)."""
    net = eng.net
    k = model_params + net.num_reticulations
    if linkage == UNLINKED:
        return k + eng.P * net.num_edges
    return k + net.num_edges + (eng.P - 1 if linkage == SCALED else 0)


class UnoptimisedModelError(ValueError):
    """The model string asks for ML-estimated substitution rates / frequencies, which the reference's optimizeModel fits with
    pll-modules' L-BFGS-B (out of scope here: SURVEY §2) — scoring with the start values would silently differ from ./netrax."""


def ml_estimated_params(specs) -> list:
    """Per partition, the parameter groups the reference would estimate by L-BFGS-B and this package leaves at their start values."""
    out = []
    for i, ms in enumerate(specs):
        what = [w for w, on in (("substitution rates", getattr(ms, "rate_mode", "") == "ML"), ("base frequencies", getattr(ms, "freq_mode", "") == "ML")) if on]
        if what:
            out.append((i, getattr(ms, "name", "?"), what))
    return out


def score_only(engine_factory: Callable, network_text: str, msa_text: str, model: str, variant: int = AVERAGE,
               linkage: int = LINKED, optimize: bool = True, log: Optional[Callable[[str], None]] = print,
               allow_unoptimised_ml_params: bool = False) -> Dict[str, object]:
    """`allow_unoptimised_ml_params`: models with ML-estimated rates / frequencies (e.g. plain ``GTR+G``) raise
    UnoptimisedModelError unless this is set; when set, the result carries ``unoptimised_ml_params`` and a loud warning, because
    lnL / BIC / AIC then differ from the reference's --score_only (which optimises those parameters).  Give the rates and
    frequencies explicitly (``GTR{...}+FU{...}``), use empirical / equal frequencies, or a fixed-matrix model to score like ./netrax."""
    say = log or (lambda s: None)
    net = parse_extended_newick(network_text)
    msa = read_msa(msa_text)
    parts, specs = build_partitions(msa, net.tip_labels, model)
    unopt = ml_estimated_params(specs)
    if unopt:
        msg = ("model parameters the reference would estimate by ML (L-BFGS-B in pll-modules) stay at their start values here: " +
               "; ".join(f"partition {i} ({name}): {' + '.join(what)}" for i, name, what in unopt))
        if not allow_unoptimised_ml_params:
            raise UnoptimisedModelError(msg + " — pass allow_unoptimised_ml_params=True (--allow-unoptimised) to score anyway")
        say("WARNING: " + msg + "; lnL / BIC / AIC / AICc will differ from ./netrax --score_only")
    eng = engine_factory(net, parts, variant=variant, linkage=linkage)
    try:
        k_model = apply_model_state(eng, specs)
        out: Dict[str, object] = {"taxa": net.num_tips, "reticulations": net.num_reticulations, "partitions": len(parts),
                                  "patterns": [p.sites for p in parts], "model_params": k_model,
                                  "unoptimised_ml_params": [f"partition {i} ({name}): {' + '.join(what)}" for i, name, what in unopt]}
        # optimizeModel (src/optimization/Optimization.cpp:72-84) with the built-in optimize_params steps
        eng.optimize_alpha()
        eng.optimize_pinv()
        say("Initial, given network:")
        say(eng.toExtendedNewick(6))
        out["start_bic"], out["start_logl"] = eng.scoreNetwork(), eng.computeLoglikelihood(1, 1)
        say(f"Initial (before brlen and reticulation opt) BIC Score: {out['start_bic']:.6f}")
        say(f"Initial (before brlen and reticulation opt) loglikelihood: {out['start_logl']:.6f}")
        if optimize:
            eng.optimizeAllNonTopology(2)   # OptimizeAllNonTopologyType::SLOW
            say("Network after optimization of brlens and reticulation probs:")
            say(eng.toExtendedNewick(6))
        bic, logl = eng.scoreNetwork(), eng.computeLoglikelihood(1, 1)
        k = param_count(eng, k_model, linkage)
        n = float(sum(int(p.pattern_weights.sum()) if p.pattern_weights is not None else p.sites for p in parts)) * net.num_tips
        aic = -2 * logl + 2 * k                                   # ComplexityScoring.cpp:7-18
        aicc = aic + (2.0 * k * k + 2 * k) / (n - k - 1)
        out.update({"bic": bic, "logl": logl, "aic": aic, "aicc": aicc, "param_count": k, "network": eng.toExtendedNewick(),
                    "alphas": [eng.get_alpha(p) for p in range(eng.P)], "pinvs": [eng.get_pinv(p) for p in range(eng.P)]})
        say(f"Number of reticulations: {net.num_reticulations}")
        say(f"BIC Score: {bic:.6f}")
        say(f"Loglikelihood: {logl:.6f}")
        say(f"AIC Score: {aic:.6f}")
        say(f"AICc Score: {aicc:.6f}")
        assert abs(bic - (-2 * logl + k * math.log(n))) <= 1e-9 * abs(bic), "BIC bookkeeping differs from scoreNetwork"
        return out
    finally:
        eng.close()
