#!/usr/bin/env python
"""Extract the empirical amino-acid model constants (190 exchangeabilities + 20 frequencies each) of the 20 single-matrix
protein models and the LG4M / LG4X components from the reference's compiled libpll (oracle/_ref/libpll_ref.so, symbols
pll_aa_rates_* / pll_aa_freqs_*, LIBPLL/pll.h:545-590), under the names pll-modules gives them (PLLMOD/util/models_aa.c:28-59),
into netrax_b200/aa_models.json — model DATA shipped with the package (msa_io.parse_model).  Run in the build container."""
import ctypes as C
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libpll_ref.so"))
NAMES = {"DAYHOFF": "dayhoff", "LG": "lg", "DCMUT": "dcmut", "JTT": "jtt", "MTREV": "mtrev", "WAG": "wag", "RTREV": "rtrev",
         "CPREV": "cprev", "VT": "vt", "BLOSUM62": "blosum62", "MTMAM": "mtmam", "MTART": "mtart", "MTZOA": "mtzoa", "PMB": "pmb",
         "HIVB": "hivb", "HIVW": "hivw", "JTT-DCMUT": "jttdcmut", "FLU": "flu", "STMTREV": "stmtrev", "DEN": "den"}
out = {"source": "pll_aa_rates_* / pll_aa_freqs_* of the reference's forked libpll; names of PLLMOD/util/models_aa.c", "models": {}}
for name, sym in NAMES.items():
    out["models"][name] = {"rates": list((C.c_double * 190).in_dll(lib, "pll_aa_rates_" + sym)),
                           "freqs": list((C.c_double * 20).in_dll(lib, "pll_aa_freqs_" + sym))}
for mix in ("lg4m", "lg4x"):
    r = ((C.c_double * 190) * 4).in_dll(lib, "pll_aa_rates_" + mix)
    f = ((C.c_double * 20) * 4).in_dll(lib, "pll_aa_freqs_" + mix)
    for k in range(4):
        out["models"][f"{mix.upper()}{k + 1}"] = {"rates": list(r[k]), "freqs": list(f[k])}
json.dump(out, open(os.path.join(HERE, "..", "..", "netrax_b200", "aa_models.json"), "w"))
print(len(out["models"]), "models;", "LG freq sum", sum(out["models"]["LG"]["freqs"]))
