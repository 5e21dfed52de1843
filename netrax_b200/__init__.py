"""netrax_b200 — B200-native network-likelihood engine behind NetRAX's likelihood API.

Only what the hot path needs lives here: csrc/ (sm_100a CUDA kernels + the C-ABI), the C++ host
mirror of the reference's likelihood layer, and thin ctypes bindings.  Nothing in this package
imports, links or executes anything under oracle/ (test infrastructure).
"""
__version__ = "0.1.0"
