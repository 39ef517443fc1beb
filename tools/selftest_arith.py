#!/usr/bin/env python
"""Mismatch counts of the FAST Roe units' branch-free division / reciprocal / square root against div.rn / sqrt.rn (GPU box)."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pluto_b200 import load_library
L = load_library()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000_000
for seed in (1, 2, 20240607):
    bad = (C.c_ulonglong * 3)()
    rc = L.pluto_gpu_selftest_arith(0, n, seed, C.byref(bad))
    print(f"seed {seed}: rc {rc}, samples {n}: quotient {bad[0]}, reciprocal {bad[1]}, root {bad[2]} mismatches", flush=True)
