"""Debug aid (GPU box): where does FAST arithmetic deviate from EXACT after one step?"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import Golden, rel_l1
from pluto_b200 import GpuStepper
name = sys.argv[1] if len(sys.argv) > 1 else "rotor2d_ppm_roe"
g = Golden(name)
mk = lambda ar: GpuStepper(g.dims, g.n, g.dx, recon=g.recon, solver=g.solver, rk_order=g.rk_order, bc=g.bc, gamma=g.gamma, arith=ar)
e, f = mk("exact"), mk("fast")
e.set_state(g.states[0]); f.set_state(g.states[0])
e.advance(g.first_dt); f.advance(g.first_dt)
a, b = e.get_state(), f.get_state()
for k in a:
    d = np.abs(a[k] - b[k])
    idx = np.unravel_index(d.argmax(), d.shape)
    print(f"{k:5s} relL1 {rel_l1(b[k], a[k]):.3e} max|d| {d.max():.3e} at {idx} exact {a[k][idx]:.17g} fast {b[k][idx]:.17g}  n(d>1e-14)={int((d>1e-14).sum())}")
