"""fsgm_b200 — B200-native (sm_100a) implementation of fSGM's hot path.

The product is the C-ABI shared library built from fsgm_b200/csrc (see include/fsgm.h).
This Python package is only the ctypes harness used by tests/ and bench.py.
"""
