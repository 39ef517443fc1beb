#!/usr/bin/env python
"""One GPU: cost of the shell / interior split of the stage completion (final_kernel) for a block whose six sides are
SHARED, as on 8 GPUs -- PLUTO_GPU_SHELL_W = width of the x1 slabs.  Times K steps issued as stage_shell + stage_interior
against the unsplit stage (ghost zones are not exchanged: timing only)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from pluto_b200 import GpuStepper, problems
n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
_, meta = problems.make("turb", 3, (8, 8, 8))
ng = 2
# the periodic field INCLUDING its ghost zones (the generators are analytic: zones -ng .. n+ng-1), so that a block with
# SHARED sides and no exchange still sees valid neighbours; dt = 0 keeps it that way
ext, _ = problems.make("turb", 3, (n, n, n), offset=(-ng, -ng, -ng), count=(n + 2 * ng,) * 3)
meta["dx"] = [6.28318530717959 / n] * 3
for w in (os.environ.get("WIDTHS", "32,8,4,2")).split(","):
    os.environ["PLUTO_GPU_SHELL_W"] = w
    s = GpuStepper(3, (n, n, n), meta["dx"], bc=("shared",) * 6, gamma=meta["gamma"], arith="fast")
    Vc, s1, s2, s3 = s.data_buffers()
    for iv, nm in enumerate(["rho", "vx1", "vx2", "vx3", "Bx1", "Bx2", "Bx3", "prs"]):
        Vc[iv] = ext[nm]
    s1[...] = ext["Bx1s"]; s2[...] = ext["Bx2s"]; s3[...] = ext["Bx3s"]
    s.upload_data(Vc, s1, s2, s3)
    res = {}
    for mode in ("split", "whole"):
        for rep in range(2):
            torch.cuda.synchronize()
            s.timing(True)
            for _ in range(4):
                s.step_begin()
                for stage in (1, 1):      # stage 1 only: its input (buffer 0) keeps valid ghost zones without an exchange
                    if mode == "split":
                        s.stage_shell(stage, 0.0); s.stage_interior(stage)
                    else:
                        s.stage(stage, 0.0)
                info = s.step_end()
            rep_ = s.timing_report()
            s.timing(False)
        res[mode] = {k: round(v[0] / 4, 3) for k, v in rep_.items() if v[1]}
    print(f"n {n} shell_w {w}: final split {res['split'].get('final')} ms/step, whole {res['whole'].get('final')} ms/step; "
          f"events {info.floor_events} {info.nan_events}; split {res['split']}; whole {res['whole']}", flush=True)
    s.close()
