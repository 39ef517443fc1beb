"""Debug aid (GPU box): compare every intermediate array of RK stage 1 between the
CUDA path and the CPU oracle.  usage: python tools/debug_stage1.py <golden-name>"""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests.util import Golden
from oracle.oracle_lib import Oracle
from pluto_b200 import GpuStepper

name = sys.argv[1] if len(sys.argv) > 1 else "blast2d_plm_hlld"
g = Golden(name)
o = Oracle(g.dims, g.n, g.dx, recon=g.recon, solver=g.solver, rk_order=g.rk_order, bc=g.bc, gamma=g.gamma)
s = GpuStepper(g.dims, g.n, g.dx, recon=g.recon, solver=g.solver, rk_order=g.rk_order, bc=g.bc, gamma=g.gamma)
o.set_state(g.states[0]); s.set_state(g.states[0])
o.debug_stop_after(1)
dt = g.first_dt
o.advance(dt)
s.step_begin()
for d in range(g.dims):
    s.boundary_dim(1, d)
s.stage(1, dt)
info = s.step_end()
ng = o.ng
n1, n2, n3 = g.n
def interior(a, ext=0):
    # oracle arrays are padded by 1 in all 3 dims; GPU arrays by 1 in active dims only
    if a.shape[0] == 3 and g.dims == 2:
        a = a[1:2]
    kk = slice(1 + ng - ext, 1 + ng + n3 + ext) if g.dims == 3 else slice(0, 1)
    return a[kk, 1 + ng - ext:1 + ng + n2 + ext, 1 + ng - ext:1 + ng + n1 + ext]
pairs = [("rho", "0:rho", 2), ("bx1", "0:bx1", 2), ("bx1s", "0:bx1s", 1), ("bx2s", "0:bx2s", 1),
         ("C_dt", "cdt", 0), ("ezi", "ezi", 0), ("ezj", "ezj", 0), ("ez", "ez", 0),
         ("u_rho", "u_rho", 0), ("u_mx1", "u_mx1", 0), ("u_mx2", "u_mx2", 0), ("u_eng", "u_eng", 0),
         ("bx1s", "1:bx1s", 0), ("bx2s", "1:bx2s", 0),
         ("rho", "1:rho", 0), ("vx1", "1:vx1", 0), ("vx2", "1:vx2", 0), ("bx1", "1:bx1", 0), ("prs", "1:prs", 0)]
if g.dims == 3:
    pairs += [("eyi", "eyi", 0), ("exj", "exj", 0), ("exk", "exk", 0), ("eyk", "eyk", 0), ("ex", "ex", 0), ("ey", "ey", 0),
              ("u_mx3", "u_mx3", 0), ("bx3s", "1:bx3s", 0), ("vx3", "1:vx3", 0)]
for on, gn, ext in pairs:
    if on in ("rho", "bx1", "bx1s", "bx2s") and gn.startswith("0:"):
        # the oracle's Vc/Vs now hold stage-1 OUTPUT; ghost inputs cannot be compared here
        continue
    a = interior(o.tap(on), ext); b = interior(s.read_field(gn), ext)
    bad = np.argwhere(a != b)
    print(f"{on:6s} vs {gn:8s}: {len(bad):6d} / {a.size} differ; max abs {np.abs(a-b).max():.3e}",
          ("first at kji=%s  oracle %.17g gpu %.17g" % (tuple(bad[0]), a[tuple(bad[0])], b[tuple(bad[0])])) if len(bad) else "")
print("inv_dt", info.inv_dt_hyp)
