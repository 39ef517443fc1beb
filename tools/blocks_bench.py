#!/usr/bin/env python
"""Several blocks of ONE domain on ONE device (pluto_gpu_multi_*, every block on its own stream): do the HBM-bound stage
completion kernels of one block overlap the FP64-bound sweeps of another?
    python tools/blocks_bench.py [n=256] [steps=20] [grids="1,1,1 1,1,2 1,1,4"] [problem=blast]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pluto_b200 import MultiGpuStepper, problems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
grids = [tuple(int(x) for x in g.split(",")) for g in (sys.argv[3] if len(sys.argv) > 3 else "1,1,1 1,1,2 1,1,4").split()]
prob = sys.argv[4] if len(sys.argv) > 4 else "blast"
gn = (n, n, n)
st0, meta = problems.make(prob, 3, gn)
for grid in grids:
    nb = grid[0] * grid[1] * grid[2]
    s = MultiGpuStepper(3, gn, meta["dx"], grid, devices=[0] * nb, lib_path=os.environ.get("BLOCKS_LIB"), bc=meta["bc"], gamma=meta["gamma"], arith="fast")
    s.set_state(st0)
    dt = 1e-4
    for _ in range(3):
        info = s.advance(dt)
    if torch.cuda.is_available():
        torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        info = s.advance(dt)
        dt = min(1.1 * dt, 0.3 / info.inv_dt_hyp)
    wall = time.perf_counter() - t0
    print(f"blocks {grid} on one device, {prob} {n}^3: {n**3*steps/wall:.4e} zone-updates/s, {1e3*wall/steps:.3f} ms/step, nan {info.nan_events}", flush=True)
    s.close()
