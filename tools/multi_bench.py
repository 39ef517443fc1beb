#!/usr/bin/env python
"""Throughput of the single-thread multi-device stepper (pluto_gpu_multi_advance, the path the C shim takes with
PLUTO_GPU_NDEV): NDEV blocks of n^3 zones of a periodic turbulence box, one per visible GPU.
    python tools/multi_bench.py [n=256] [ndev=all] [steps=20]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pluto_b200 import MultiGpuStepper, problems

n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
ndev = int(sys.argv[2]) if len(sys.argv) > 2 else torch.cuda.device_count()
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
grid = {1: (1, 1, 1), 2: (1, 1, 2), 4: (1, 2, 2), 8: (2, 2, 2)}[ndev]
gn = tuple(n * g for g in grid)
st0, meta = problems.make("turb", 3, gn)
for arith in ("fast",):
    s = MultiGpuStepper(3, gn, meta["dx"], grid, devices=list(range(ndev)), bc=meta["bc"], gamma=meta["gamma"], arith=arith)
    s.set_state(st0)
    dt = 1e-3
    for _ in range(3):
        info = s.advance(dt)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        info = s.advance(dt)                       # step_end of every block synchronises: host-driven NextTimeStep
        dt = min(1.1 * dt, 0.3 / info.inv_dt_hyp)
    wall = time.perf_counter() - t0
    zones = float(np.prod(gn))
    print(f"multi C path: {ndev} device(s), {n}^3 per device, {arith}: {zones*steps/wall:.4e} zone-updates/s, "
          f"{1e3*wall/steps:.3f} ms/step (wall, host-driven dt), nan {info.nan_events}", flush=True)
    s.close()
