#!/usr/bin/env python
"""Extract the LG amino-acid model constants (190 exchangeabilities, 20 frequencies; Le & Gascuel 2008) from the
reference's compiled libpll (oracle/_ref/libpll_ref.so, symbols pll_aa_rates_lg / pll_aa_freqs_lg declared at
LIBPLL/pll.h:555,578) into netrax_b200/lg_model.json (model DATA shipped with the package; bench config 4 and the protein tests use it).  Run in the build container (needs oracle/_ref)."""
import ctypes as C
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
lib = C.CDLL(os.path.join(HERE, "..", "..", "oracle", "_ref", "libpll_ref.so"))
rates = (C.c_double * 190).in_dll(lib, "pll_aa_rates_lg")
freqs = (C.c_double * 20).in_dll(lib, "pll_aa_freqs_lg")
json.dump({"source": "pll_aa_rates_lg / pll_aa_freqs_lg of the reference's forked libpll", "rates": list(rates), "freqs": list(freqs)},
          open(os.path.join(HERE, "..", "..", "netrax_b200", "lg_model.json"), "w"))
print(sum(freqs), len(list(rates)))
