"""Alignment + model input for the likelihood API ("data formats either side of the path").

What the reference gets from raxml-ng before it can call ``computeLoglikelihood`` (src/RaxmlWrapper.cpp:686-760 builds
one pll_partition_t per raxml-ng PartitionInfo): an MSA file (FASTA or PHYLIP), an evolutionary model given either as one
model string for the whole alignment or as a partition file (``MODEL, name = 1-500, 700-900\\3``), pattern compression,
and the starting values of the model parameters.  This module restates that input side for the model families the engine
evaluates, following raxml-ng (RAXML = /root/reference/libs/raxml-ng/src) and pll-modules (PLLMOD = .../libs/pll-modules/src):

  parse_model           RAXML/Model.cpp:206-296 (name -> substitution model), :378-870 (the ``+`` options)
  DNA_MODELS            PLLMOD/util/models_dna.c:37-100 (rate / frequency symmetries of the 22 named DNA models + aliases)
  msa_stats             pllmod_msa_compute_features, PLLMOD/msa/pll_msa.c:640-880 (empirical frequencies ignoring gap cells,
                        proportion of invariant columns)
  starting values       assign(Model&, PartitionStats&), RAXML/PartitionInfo.cpp:144-199; Model::init_model_opts (:378-384)
  free_params           Model::num_free_params, RAXML/Model.cpp:1014-1055 (what NetRAX sums into the BIC's k,
                        src/RaxmlWrapper.cpp:682-684)

Not covered (a ValueError names the option): FreeRate (+R), ascertainment bias (+ASC), custom character maps (+M), mixture
multistate models, LG4X (FreeRate mixture), PROTGTR, PAML files.  The 20 empirical protein matrices and the LG4M components
come from ``aa_models.json`` (libpll's tables, extracted by tests/golden/make_aa_models.py).
Nothing here touches the GPU; the Partition objects it returns are what ``NetraxB200`` takes.
"""
from __future__ import annotations

import json
import os
import re
from dataclasses import dataclass, field
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np

from ._capi import Partition
from .network_io import _DNA, compress_patterns, read_fasta

_EQUAL, _FREE = (0, 0, 0, 0), (0, 1, 2, 3)
#             name      (rate symmetries AC AG AT CG CT GT, frequency symmetries)
DNA_MODELS: Dict[str, Tuple[Tuple[int, ...], Tuple[int, ...]]] = {
    "JC": ((0, 0, 0, 0, 0, 0), _EQUAL), "K80": ((0, 1, 0, 0, 1, 0), _EQUAL), "F81": ((0, 0, 0, 0, 0, 0), _FREE),
    "HKY": ((0, 1, 0, 0, 1, 0), _FREE), "TN93ef": ((0, 1, 0, 0, 2, 0), _EQUAL), "TN93": ((0, 1, 0, 0, 2, 0), _FREE),
    "K81": ((0, 1, 2, 2, 1, 0), _EQUAL), "K81uf": ((0, 1, 2, 2, 1, 0), _FREE), "TPM2": ((0, 1, 0, 2, 1, 2), _EQUAL),
    "TPM2uf": ((0, 1, 0, 2, 1, 2), _FREE), "TPM3": ((0, 1, 2, 0, 1, 2), _EQUAL), "TPM3uf": ((0, 1, 2, 0, 1, 2), _FREE),
    "TIM1": ((0, 1, 2, 2, 3, 0), _EQUAL), "TIM1uf": ((0, 1, 2, 2, 3, 0), _FREE), "TIM2": ((0, 1, 0, 2, 3, 2), _EQUAL),
    "TIM2uf": ((0, 1, 0, 2, 3, 2), _FREE), "TIM3": ((0, 1, 2, 0, 3, 2), _EQUAL), "TIM3uf": ((0, 1, 2, 0, 3, 2), _FREE),
    "TVMef": ((0, 1, 2, 3, 1, 4), _EQUAL), "TVM": ((0, 1, 2, 3, 1, 4), _FREE), "SYM": ((0, 1, 2, 3, 4, 5), _EQUAL),
    "GTR": ((0, 1, 2, 3, 4, 5), _FREE),
}
_DNA_ALIASES = {"TrNef": "TN93ef", "TrN": "TN93", "TPM1": "K81", "TPM1uf": "K81uf", "TPM2ef": "TPM2", "TPM3ef": "TPM3",
                "TIM1ef": "TIM1", "TIM2ef": "TIM2", "TIM3ef": "TIM3", "DNA": "GTR"}   # "DNA": RAXML/Model.cpp:239-243
_AA_ORDER = "ARNDCQEGHILKMFPSTWYV"
_AA_AMBIG = {"B": "ND", "Z": "QE", "J": "IL"}


_AA_MODELS: Dict[str, Dict[str, List[float]]] = {}


def aa_model(name: str) -> Tuple[np.ndarray, np.ndarray]:
    """(190 exchangeabilities, 20 frequencies) of one of libpll's empirical protein matrices, by pll-modules' name
    (PLLMOD/util/models_aa.c:28-59; data file written by tests/golden/make_aa_models.py from the compiled reference library)."""
    if not _AA_MODELS:
        d = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "aa_models.json")))
        _AA_MODELS.update({k.upper(): v for k, v in d["models"].items()})
    m = _AA_MODELS.get(name.upper())
    if m is None:
        raise ValueError(f"Invalid model name: {name}")
    return np.asarray(m["rates"], float), np.asarray(m["freqs"], float)


def aa_model_names() -> List[str]:
    aa_model("LG")
    return sorted(_AA_MODELS)


@dataclass
class ModelSpec:
    """One partition's model as raxml-ng's Model object holds it after parsing (before optimisation)."""
    name: str
    data_type: str                      # "DNA" | "AA"
    states: int
    rate_sym: Tuple[int, ...] = ()      # () = every rate its own parameter (protein)
    freq_mode: str = "ML"               # model | ML | empirical | equal | user     (ParamValue, RAXML/Model.hpp)
    rate_mode: str = "ML"               # model | ML | user
    subst_rates: Optional[np.ndarray] = None    # full upper-triangle rates
    freqs: Optional[np.ndarray] = None
    rate_cats: int = 1
    gamma_mode: int = 0                 # 0 mean (PLL_GAMMA_RATES_MEAN), 1 median
    alpha: float = 1.0
    alpha_mode: str = "undefined"       # undefined (no +G) | ML | user
    pinv: float = 0.0
    pinv_mode: str = "undefined"        # undefined (no +I) | ML | empirical | user
    brlen_scaler: float = 1.0
    brlen_scaler_mode: str = "undefined"
    # mixture with one matrix per rate category (LG4M): [(rates, freqs)] and the matrix of each category
    submodels: Optional[List[Tuple[np.ndarray, np.ndarray]]] = None
    ratecat_submodels: Optional[List[int]] = None

    @property
    def num_uniq_rates(self) -> int:
        return (max(self.rate_sym) + 1) if self.rate_sym else self.states * (self.states - 1) // 2

    def free_params(self) -> int:
        """Model::num_free_params (RAXML/Model.cpp:1014-1055)."""
        k = 0
        if self.freq_mode in ("ML", "empirical"):
            k += self.states - 1
        if self.rate_mode == "empirical":
            k += self.states * (self.states - 1) // 2 - 1
        elif self.rate_mode == "ML":
            k += self.num_uniq_rates - 1
        if self.pinv_mode in ("ML", "empirical"):
            k += 1
        if self.rate_cats > 1 and self.alpha_mode in ("ML", "empirical"):
            k += 1
        return k


def _read_braces(s: str, i: int) -> Tuple[Optional[str], int]:
    if i < len(s) and s[i] == "{":
        j = s.find("}", i)
        if j < 0:
            raise ValueError(f"unterminated '{{' in model string {s!r}")
        return s[i + 1: j], j + 1
    return None, i


def _floats(text: str) -> List[float]:
    return [float(x) for x in re.split(r"[/,]", text) if x.strip()]


def parse_model(spec: str) -> ModelSpec:
    """``GTR+G4{0.7}+I``, ``HKY{1/2.5}+FC``, ``LG+G+F`` ... -> ModelSpec (RAXML/Model.cpp:206-296, 378-870)."""
    spec = spec.strip()
    m = re.match(r"[A-Za-z0-9]+(?:-[A-Za-z]+)?", spec)   # "JTT-DCMUT" carries a dash
    if not m:
        raise ValueError(f"Invalid model name: {spec!r}")
    name = m.group(0)
    i = m.end()
    canon = {k.upper(): k for k in DNA_MODELS}
    alias = {k.upper(): v for k, v in _DNA_ALIASES.items()}
    up = name.upper()
    if up in alias or up in canon:
        real = alias.get(up) or canon[up]
        sym, fsym = DNA_MODELS[real]
        ms = ModelSpec(real, "DNA", 4, rate_sym=sym)
        ms.freq_mode = "model" if fsym == _EQUAL else "ML"
        ms.freqs = np.full(4, 0.25)
        ms.rate_mode = "model" if max(sym) == 0 else "ML"
        ms.subst_rates = np.ones(6)
    elif up == "LG4M":   # M_LG4M (PLLMOD/util/models_aa.c:103-105): four LG4M matrices, one per GAMMA category
        comps = [aa_model(f"LG4M{k + 1}") for k in range(4)]
        ms = ModelSpec("LG4M", "AA", 20, freq_mode="model", rate_mode="model", subst_rates=comps[0][0], freqs=comps[0][1],
                       rate_cats=4, alpha_mode="ML", submodels=comps, ratecat_submodels=[0, 1, 2, 3])
    elif up == "LG4X":
        raise ValueError("LG4X is a FreeRate mixture (PLLMOD_UTIL_MIXTYPE_FREE): free rates / weights are not supported by this input layer")
    elif up in ("PROTGTR", "PROT"):
        raise ValueError(f"protein model {name}: a 189-parameter GTR needs pll-modules' BFGS optimiser (optimize_params_cb)")
    elif up in [n for n in aa_model_names() if not n.startswith("LG4")]:
        r, f = aa_model(up)
        ms = ModelSpec(up, "AA", 20, freq_mode="model", rate_mode="model", subst_rates=r, freqs=f)
    else:
        raise ValueError(f"Invalid model name: {name}")
    user, i = _read_braces(spec, i)
    if user is not None:   # user-defined rates: one value per free rate, normalised by the last one (set_user_srates, :337-353)
        vals = _floats(user)
        if len(vals) != ms.num_uniq_rates:
            raise ValueError(f"Invalid number of substitution rates specified: expected {ms.num_uniq_rates}, found {len(vals)}")
        sym = ms.rate_sym or tuple(range(ms.num_uniq_rates))
        last = vals[sym[-1]]
        ms.subst_rates = np.array([vals[k] / last for k in sym])
        ms.rate_mode = "user"
    while i < len(spec):
        if spec[i] != "+":
            raise ValueError(f"Invalid model options: {spec[i:]!r}")
        i += 1
        if i >= len(spec):
            raise ValueError("Invalid model options: trailing '+'")
        opt = spec[i].upper()
        i += 1
        if ms.submodels is not None and opt in ("F", "R"):
            raise ValueError(f"{ms.name}: +{opt} is not defined for a mixture with one matrix per category")
        if opt == "G":
            mm = re.match(r"\d+", spec[i:])
            if mm:
                ms.rate_cats = int(mm.group(0))
                i += mm.end()
                if ms.submodels is not None and ms.rate_cats != len(ms.submodels):
                    raise ValueError(f"{ms.name} has {len(ms.submodels)} matrices: the number of rate categories must match")
            elif ms.rate_cats == 1:
                ms.rate_cats = 4
            if i < len(spec) and spec[i] in "aA":
                ms.gamma_mode, i = 1, i + 1
            elif i < len(spec) and spec[i] in "mM":
                ms.gamma_mode, i = 0, i + 1
            val, i = _read_braces(spec, i)
            ms.alpha_mode = "ML"
            if val is not None:
                ms.alpha, ms.alpha_mode = float(val), "user"
        elif opt == "I":
            val, j = _read_braces(spec, i)
            if val is not None:
                ms.pinv, ms.pinv_mode, i = float(val), "user", j
            elif i < len(spec) and spec[i].upper() == "O":
                ms.pinv_mode, i = "ML", i + 1
            elif i < len(spec) and spec[i].upper() == "C":
                ms.pinv_mode, i = "empirical", i + 1
            elif i >= len(spec) or spec[i] == "+":
                ms.pinv_mode = "ML"
            else:
                raise ValueError(f"Invalid p-inv specification: {spec}")
        elif opt == "F":
            if spec[i: i + 2].upper() == "U{":
                i += 1
            val, j = _read_braces(spec, i)
            if val is not None:
                f = _floats(val)
                if len(f) != ms.states:
                    raise ValueError(f"Invalid number of user frequencies specified: {len(f)}")
                if any(v < 0.0 or v >= 1.0 for v in f):
                    raise ValueError("Invalid base frequencies specified! Frequencies must be positive numbers between 0. and 1.")
                ms.freqs, ms.freq_mode, i = np.asarray(f, float), "user", j
            elif i >= len(spec) or spec[i] == "+":
                ms.freq_mode = "empirical"
            else:
                mode = spec[i].upper()
                i += 1
                if mode == "C":
                    ms.freq_mode = "empirical"
                elif mode == "O":
                    ms.freq_mode = "ML"
                elif mode == "E":
                    ms.freq_mode, ms.freqs = "equal", np.full(ms.states, 1.0 / ms.states)
                else:
                    raise ValueError(f"Invalid frequencies specification: {spec}")
        elif opt == "B":
            val, j = _read_braces(spec, i)
            if val is not None:
                ms.brlen_scaler, ms.brlen_scaler_mode, i = float(val), "user", j
            else:
                if i < len(spec) and spec[i].upper() == "O":
                    i += 1
                ms.brlen_scaler_mode = "ML"
        elif opt == "R":
            raise ValueError("FreeRate models (+R) are not supported by this engine's input layer")
        elif opt == "A":
            raise ValueError("ascertainment bias correction (+ASC) is rejected by the engine (DESIGN.md §7)")
        elif opt == "M":
            raise ValueError("custom character maps (+M) are not supported")
        else:
            raise ValueError(f"Invalid model options: +{opt}{spec[i:]}")
    return ms


# ---- alignment files ------------------------------------------------------------------------------------------------
def read_phylip(text: str) -> Dict[str, str]:
    """Relaxed PHYLIP, sequential or interleaved (what pll_phylip_parse_* accept through raxml-ng's MSA loader)."""
    lines = [l.rstrip() for l in text.splitlines() if l.strip()]
    if not lines:
        raise ValueError("empty alignment file")
    head = lines[0].split()
    if len(head) < 2 or not head[0].isdigit() or not head[1].isdigit():
        raise ValueError("PHYLIP header must be '<taxa> <sites>'")
    ntax, nchar = int(head[0]), int(head[1])
    body = lines[1:]
    if len(body) < ntax:
        raise ValueError(f"PHYLIP: {ntax} taxa announced, {len(body)} lines found")

    def split_named(line: str) -> Tuple[str, str]:
        parts = line.split(None, 1)
        return parts[0], (parts[1].replace(" ", "").replace("\t", "") if len(parts) > 1 else "")

    # interleaved (or one line per taxon): names on the first ntax lines, further blocks continue in the same order
    names, seqs = [], []
    for l in body[:ntax]:
        n, s = split_named(l)
        names.append(n); seqs.append(s)
    rest = body[ntax:]
    if len(rest) % ntax == 0 and len(set(names)) == ntax:
        trial = list(seqs)
        for k, l in enumerate(rest):
            trial[k % ntax] += l.replace(" ", "").replace("\t", "")
        if all(len(s) == nchar for s in trial):
            return dict(zip(names, trial))
    # sequential: a name line, then continuation lines until the announced length is reached
    out: Dict[str, str] = {}
    it = iter(body)
    for _ in range(ntax):
        n, s = split_named(next(it))
        while len(s) < nchar:
            try:
                s += next(it).replace(" ", "").replace("\t", "")
            except StopIteration:
                raise ValueError(f"PHYLIP: sequence {n} is shorter than {nchar}")
        if len(s) != nchar:
            raise ValueError(f"PHYLIP: sequence {n} has {len(s)} characters, expected {nchar}")
        if n in out:
            raise ValueError(f"duplicate sequence name {n}")
        out[n] = s
    return out


def read_msa(text: str) -> Dict[str, str]:
    """FASTA if the first non-blank character is '>', PHYLIP otherwise (raxml-ng probes the formats in turn)."""
    seqs = read_fasta(text) if text.lstrip().startswith(">") else read_phylip(text)
    lens = set(len(s) for s in seqs.values())
    if len(lens) != 1:
        raise ValueError("sequences of an alignment must have the same length")
    return seqs


def encode(seq: str, data_type: str) -> np.ndarray:
    """Characters -> state bit masks with libpll's maps (pll_map_nt / pll_map_aa, LIBPLL/maps.c); unknown characters raise
    like pllmod_msa_compute_features does ("Unknown state")."""
    out = np.zeros(len(seq), np.uint32)
    if data_type == "DNA":
        for k, c in enumerate(seq.upper()):
            v = _DNA.get(c)
            if v is None:
                raise ValueError(f"Unknown state {c} at position {k + 1}")
            out[k] = v
        return out
    full = (1 << 20) - 1
    for k, c in enumerate(seq.upper()):
        if c in _AA_ORDER:
            out[k] = 1 << _AA_ORDER.index(c)
        elif c in _AA_AMBIG:
            out[k] = sum(1 << _AA_ORDER.index(x) for x in _AA_AMBIG[c])
        elif c in "-X?*":
            out[k] = full
        else:
            raise ValueError(f"Unknown state {c} at position {k + 1}")
    return out


def msa_stats(masks: np.ndarray, weights: Optional[np.ndarray], states: int) -> Tuple[np.ndarray, float]:
    """(empirical frequencies, proportion of invariant columns) as pllmod_msa_compute_features computes them
    (PLLMOD/msa/pll_msa.c:742-830): a cell spreads its weight evenly over its states, gap cells (all states set) are left out of
    the frequencies and of the invariance test, a column is invariant when the union of its non-gap cells is one state."""
    masks = np.asarray(masks, np.uint32)
    w = np.ones(masks.shape[1]) if weights is None else np.asarray(weights, float)
    pop = np.zeros(masks.shape, np.int64)
    for k in range(states):
        pop += (masks >> k) & 1
    gap = pop == states
    freqs = np.zeros(states)
    share = np.where(gap, 0.0, w[None, :] / np.maximum(pop, 1))
    for k in range(states):
        freqs[k] = float((share * ((masks >> k) & 1)).sum())
    total = float(w.sum()) * masks.shape[0] - float((gap * w[None, :]).sum())
    freqs /= total
    union = np.bitwise_or.reduce(np.where(gap, 0, masks), axis=0)
    upop = np.zeros(masks.shape[1], np.int64)
    for k in range(states):
        upop += (union >> k) & 1
    inv_prop = float(w[upop == 1].sum() / w.sum())
    return freqs, inv_prop


# ---- partition files --------------------------------------------------------------------------------------------------
@dataclass
class PartitionRange:
    model: ModelSpec
    name: str
    ranges: List[Tuple[int, int, int]] = field(default_factory=list)   # 1-based inclusive start, end, stride

    def columns(self, nsites: int) -> np.ndarray:
        cols = []
        for a, b, s in self.ranges:
            if a < 1 or b > nsites or a > b:
                raise ValueError(f"partition {self.name}: range {a}-{b} outside the alignment (1-{nsites})")
            cols.append(np.arange(a - 1, b, s))
        return np.concatenate(cols) if cols else np.zeros(0, np.int64)


def parse_partition_file(text: str) -> List[PartitionRange]:
    """RAxML-style partition file: ``MODEL, name = 1-100, 250-300\\3`` per line (RAXML/io/part_info.cpp)."""
    out = []
    for raw in text.splitlines():
        line = raw.split("#", 1)[0].strip()
        if not line:
            continue
        if "," not in line or "=" not in line:
            raise ValueError(f"partition file: cannot parse line {raw!r}")
        model, rest = line.split(",", 1)
        name, rng = rest.split("=", 1)
        pr = PartitionRange(parse_model(model.strip()), name.strip())
        for tok in rng.split(","):
            tok = tok.strip()
            mm = re.fullmatch(r"(\d+)(?:\s*-\s*(\d+))?(?:\s*[\\/]\s*(\d+))?", tok)
            if not mm:
                raise ValueError(f"partition {pr.name}: invalid range {tok!r}")
            a = int(mm.group(1))
            pr.ranges.append((a, int(mm.group(2)) if mm.group(2) else a, int(mm.group(3)) if mm.group(3) else 1))
        out.append(pr)
    if not out:
        raise ValueError("partition file defines no partitions")
    return out


def build_partitions(msa: Dict[str, str], tip_labels: Sequence[str], model: str,
                     gamma_rates: Optional[Callable[[float, int, int], np.ndarray]] = None,
                     compress: bool = True) -> Tuple[List[Partition], List[ModelSpec]]:
    """MSA + (model string | partition-file text) -> the engine's Partition inputs, rows ordered like ``tip_labels`` (the
    network's tips), columns compressed to patterns with multiplicities as weights, parameters at raxml-ng's starting values:
    empirical frequencies for +F/+FC, equal ones for ML-estimated frequencies, all free rates 1, alpha 1 (or the user's),
    pinv = empirical / half the empirical proportion for +IC / +I (RAXML/PartitionInfo.cpp:144-199).  The returned specs carry
    those values; ``apply_model_state`` pushes alpha / pinv / scalers into an engine."""
    missing = [t for t in tip_labels if t not in msa]
    if missing:
        raise ValueError(f"taxa of the network missing from the alignment: {missing[:5]}")
    nsites = len(next(iter(msa.values())))
    if "=" in model:
        prs = parse_partition_file(model)
    else:
        prs = [PartitionRange(parse_model(model), "noname", [(1, nsites, 1)])]
    if gamma_rates is None:
        from . import engine
        gamma_rates = engine.load().gamma_rates
    parts, specs = [], []
    for pr in prs:
        ms = pr.model
        cols = pr.columns(nsites)
        if cols.size == 0:
            raise ValueError(f"partition {pr.name} is empty")
        chars = {t: np.frombuffer(msa[t].encode("ascii"), dtype="S1") for t in tip_labels}
        masks = np.stack([encode(b"".join(chars[t][cols]).decode("ascii"), ms.data_type) for t in tip_labels])
        weights = None
        if compress:
            masks, weights = compress_patterns(masks)
        emp_freqs, inv_prop = msa_stats(masks, weights, ms.states)
        if ms.freq_mode == "empirical":
            ms.freqs = emp_freqs
        if ms.pinv_mode == "empirical":
            ms.pinv = inv_prop
        elif ms.pinv_mode == "ML":
            ms.pinv = inv_prop / 2   # "use half of empirical pinv as a starting value"
        rates = gamma_rates(ms.alpha, ms.rate_cats, ms.gamma_mode) if ms.rate_cats > 1 else np.ones(1)
        parts.append(Partition(ms.states, ms.rate_cats, masks, ms.freqs, ms.subst_rates, rates, pattern_weights=weights))
        specs.append(ms)
    return parts, specs


def apply_model_state(eng, specs: Sequence[ModelSpec]) -> int:
    """Attach the Gamma shapes, +I proportions and branch-length scalers of ``specs`` to an engine (so that optimize_alpha /
    optimize_pinv / optimize_scalers treat them as free parameters) and hand the BIC its parameter count; returns that count
    (sum of Model::num_free_params, src/RaxmlWrapper.cpp:682-684)."""
    for p, ms in enumerate(specs):
        if ms.rate_cats > 1 and ms.alpha_mode != "undefined":
            if ms.gamma_mode != 0:
                raise ValueError("the engine's setAlpha uses the mean discretisation (raxml-ng's default); +Ga is not supported")
            eng.set_alpha(p, ms.alpha)
        if ms.pinv_mode != "undefined" and ms.pinv > 0.0:
            eng.set_pinv(p, ms.pinv)
        # which of the two are FREE parameters comes from the model spec, not from the current value: a +I partition whose
        # empirical (or optimised) proportion is 0 must stay in optimize_pinv (pll-modules' params_to_optimize)
        eng.set_params_to_optimize(p, alpha=(ms.rate_cats > 1 and ms.alpha_mode == "ML"), pinv=(ms.pinv_mode == "ML"))
        if ms.brlen_scaler_mode == "user":
            eng.set_brlen_scaler(p, ms.brlen_scaler)
        if ms.submodels is not None:   # raxml-ng's ratecat_submodels -> libpll params_indices (src/RaxmlWrapper.cpp:199-203)
            eng.set_submodels(p, ms.ratecat_submodels, np.stack([f for _, f in ms.submodels]), np.stack([r for r, _ in ms.submodels]))
    k = sum(ms.free_params() for ms in specs)
    eng.set_scoring_sizes(k)
    return k
