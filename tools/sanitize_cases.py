#!/usr/bin/env python
"""Small cases of every kernel family for compute-sanitizer (memcheck / racecheck / initcheck):

    compute-sanitizer --tool racecheck python tools/sanitize_cases.py [family ...]

families: rk (fused x1+x2 sweep, x3 march, CT, final, bc), exact (one kernel per direction), ppm_roe, hll_uct_hll,
ctu, bc (outflow / reflective / eqtsymmetric fills incl. div B), halo (two blocks in one process: pack / unpack tables),
io (dbl writer / analysis), grid (non-uniform grids, grid-dependent weights, PLUTO_GPU_R3), schemes2 (CHARACTERISTIC_TRACING, CHAR_LIMITING + CTU, MULTID + PARABOLIC), schemes3 (hllc / tvdlf, BODY_FORCE combinations, scheme options and PARABOLIC on non-uniform grids).  Launches run without graph capture so that a report names the kernel.
"""
import os
import sys

os.environ["PLUTO_GPU_NO_GRAPH"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np                                   # noqa: E402
from pluto_b200 import GpuStepper, problems          # noqa: E402


def run(problem, dims, n, steps=2, dt=1e-4, **kw):
    st0, meta = problems.make(problem, dims, n)
    bc = kw.pop("bc", meta["bc"])
    s = GpuStepper(dims, n, meta["dx"], bc=bc, gamma=meta["gamma"], **kw)
    s.set_state(st0)
    for _ in range(steps):
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
    st = s.get_state()
    assert all(np.isfinite(v).all() for v in st.values())
    s.close()


def fam_rk():
    run("blast", 3, (40, 24, 20), arith="fast")
    run("ot", 2, (48, 40, 1), arith="fast", dt=1e-3)
    run("ot", 2, (48, 40, 1), arith="fast", dt=1e-3, char_lim=True)
    run("rotor", 2, (40, 32, 1), arith="exact", dt=1e-3, char_lim=True, solver="roe")
    run("turb", 3, (24, 24, 24), arith="fast", rk_order=3, dt=1e-3)


def fam_exact():
    run("blast", 3, (40, 24, 20), arith="exact")
    run("ot", 2, (48, 40, 1), arith="exact", dt=1e-3)


def fam_ppm_roe():
    run("rotor", 2, (48, 40, 1), arith="fast", recon="ppm", solver="roe", dt=1e-4)
    run("ot", 3, (24, 20, 16), arith="fast", recon="ppm", solver="roe", dt=1e-3)
    run("ot", 3, (24, 20, 16), arith="exact", recon="ppm", solver="roe", dt=1e-3)


def fam_hll_uct_hll():
    run("turb", 3, (24, 20, 16), arith="fast", solver="hll", emf="uct_hll", dt=1e-3)
    run("blast", 3, (24, 20, 16), arith="fast", flatten=True, emf="uct0")
    run("blast", 3, (24, 20, 16), arith="fast", en_corr=True, emf="arith", limiter="vl")


def fam_ctu():
    run("blast", 3, (24, 20, 16), arith="fast", ctu=True)
    run("blast", 3, (24, 20, 16), arith="exact", ctu=True)
    run("ot", 2, (48, 40, 1), arith="fast", ctu=True, dt=1e-3)


def fam_bc():
    for b in ("outflow", "reflective", "eqtsymmetric"):
        run("blast", 3, (20, 16, 24), arith="fast", bc=(b,) * 6)
        run("blast", 2, (32, 24, 1), arith="exact", bc=(b,) * 4 + ("outflow",) * 2)


def fam_halo():
    from pluto_b200.parallel import BlockLayout, LocalMultiBlock
    for problem, dims, n in (("turb", 3, (24, 20, 16)), ("blast", 3, (24, 20, 16))):
        st0, meta = problems.make(problem, dims, n)
        lay = BlockLayout.strong(dims, n, 4, periodic=meta["bc"][0] == "periodic")
        for split in (False, True):
            many = LocalMultiBlock(lay, meta["dx"], meta["bc"], gamma=meta["gamma"], arith="fast", exchange="all", split=split)
            many.set_state(st0)
            many.advance(1e-4)
            many.advance(1e-4)
            many.get_state()
            for blk in many.blocks:
                blk.close()


def fam_grid():
    """Non-uniform grids (pluto_gpu_set_grid), grid-dependent reconstruction weights (RECON_PLMW variants), and the x3 sweep with its
    flux difference kept apart (PLUTO_GPU_R3=1); the cell-centred EMF arrays of the fused sweep are part of every FAST case."""
    rng = np.random.default_rng(3)

    def weights(dx):
        xl = np.concatenate([[0.0], np.cumsum(dx)[:-1]]); xr = xl + dx; x = 0.5*(xl + xr)
        six = [np.zeros(dx.size) for _ in range(6)]
        for i in range(1, dx.size - 1):
            six[0][i] = (x[i+1] - x[i])/(xr[i] - x[i]); six[1][i] = (x[i] - x[i-1])/(x[i] - xr[i-1])
            six[2][i] = dx[i]/(x[i+1] - x[i]); six[3][i] = dx[i]/(x[i] - x[i-1])
            six[4][i] = (xr[i] - x[i])/dx[i]; six[5][i] = (x[i] - xr[i-1])/dx[i]
        return six
    for dims, n, arith, w in ((3, (40, 24, 20), "fast", False), (3, (24, 20, 16), "fast", True), (2, (48, 40, 1), "exact", True),
                              (3, (24, 20, 16), "exact", False)):
        st0, meta = problems.make("blast", dims, n)
        s = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith=arith)
        dxs = [meta["dx"][d]*(0.7 + 0.6*rng.random(n[d] + 4)) for d in range(dims)]
        s.set_grid(*dxs)
        if w:
            s.set_plm_coeffs([weights(d) for d in dxs])
        s.set_state(st0)
        dt = 1e-4
        for _ in range(2):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
        assert all(np.isfinite(v).all() for v in s.get_state().values())
        s.close()
    os.environ["PLUTO_GPU_R3"] = "1"
    run("blast", 3, (40, 24, 20), arith="fast")
    del os.environ["PLUTO_GPU_R3"]


def fam_schemes2():
    """Scheme options added late in round 2: CHARACTERISTIC_TRACING (2-D), CHAR_LIMITING with the CTU steps, MULTID + PARABOLIC."""
    run("ot", 2, (48, 40, 1), arith="exact", ctu="chtr", dt=1e-3)
    run("rotor", 2, (40, 32, 1), arith="exact", ctu="chtr", dt=1e-3, solver="roe", limiter="mc", emf="uct0")
    run("ot", 2, (48, 40, 1), arith="fast", ctu=True, char_lim=True, dt=1e-3, limiter="mc", emf="arith")
    run("blast", 2, (40, 32, 1), arith="exact", ctu="chtr", char_lim=True)
    for dims, n, arith in ((3, (24, 20, 16), "fast"), (2, (48, 40, 1), "exact")):
        st0, meta = problems.make("blast", dims, n)
        s = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith=arith, recon="ppm", flatten=True)
        s.set_plm_coeffs([[np.full(n[d] + 6, c) for c in (2.0, 2.0, 1.0, 1.0, 0.5, 0.5)] for d in range(dims)])
        s.set_state(st0)
        dt = 1e-4
        for _ in range(6):                  # the blast has to start moving before zones are flagged
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
        assert all(np.isfinite(v).all() for v in s.get_state().values())
        s.close()


def fam_schemes3():
    """Last part of round 2: hllc / tvdlf in every kernel family, BODY_FORCE with UCT_HLL and with flattening, non-uniform grids with
    flattening / energy correction / body force inside the CTU step, PARABOLIC with per-zone interface weights."""
    run("blast", 3, (40, 24, 20), arith="fast", solver="hllc")
    run("blast", 3, (24, 20, 16), arith="exact", solver="hllc", recon="ppm", rk_order=3)
    run("ot", 2, (48, 40, 1), arith="fast", solver="tvdlf", emf="uct_hll", dt=1e-3)
    run("turb", 3, (24, 20, 16), arith="fast", solver="tvdlf", ctu=True, limiter="um", emf="uct0", dt=1e-3)
    run("blast", 3, (24, 20, 16), arith="fast", solver="hllc", ctu=True)
    run("blast", 3, (24, 20, 16), arith="fast", emf="uct_hll", grav=(0.3, -1.0, 0.5))
    run("blast", 3, (24, 20, 16), arith="fast", flatten=True, grav=(0.3, -1.0, 0.5), steps=6)
    run("blast", 3, (24, 20, 16), arith="fast", flatten=True, emf="uct_hll", grav=(0.3, -1.0, 0.5), steps=6)
    run("ot", 2, (48, 40, 1), arith="fast", char_lim=True, emf="uct_hll", dt=1e-3)
    run("blast", 2, (40, 32, 1), arith="fast", char_lim=True, emf="uct_hll", grav=(0.5, 0.25, 0.0), solver="roe")
    run("blast", 2, (40, 32, 1), arith="exact", char_lim=True, grav=(0.5, 0.25, 0.0), ctu=True)
    for dims, n, arith in ((3, (24, 20, 16), "fast"), (2, (48, 40, 1), "fast")):      # PARABOLIC + MULTID + UCT_HLL (+ body force)
        st0, meta = problems.make("blast", dims, n)
        s = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith=arith, recon="ppm", flatten=True, emf="uct_hll",
                       grav=(0.3, -1.0, 0.5)[:dims] + (0.0,)*(3 - dims))
        s.set_plm_coeffs([[np.full(n[d] + 2*s.ng, c) for c in (2.0, 2.0, 1.0, 1.0, 0.5, 0.5)] for d in range(dims)])
        s.set_state(st0)
        dt = 1e-4
        for _ in range(6):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
        assert all(np.isfinite(v).all() for v in s.get_state().values())
        s.close()
    rng = np.random.default_rng(5)
    for dims, n, kw in ((3, (24, 20, 16), dict(arith="fast", flatten=True)), (2, (48, 40, 1), dict(arith="fast", en_corr=True)),
                        (3, (24, 20, 16), dict(arith="fast", ctu=True, grav=(0.3, -1.0, 0.5))),
                        (3, (24, 20, 16), dict(arith="fast", recon="ppm")), (2, (48, 40, 1), dict(arith="exact", recon="ppm", solver="roe"))):
        st0, meta = problems.make("blast", dims, n)
        s = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], **kw)
        dxs = [meta["dx"][d]*(0.7 + 0.6*rng.random(n[d] + 2*s.ng)) for d in range(dims)]
        s.set_grid(*dxs)
        if kw.get("recon") == "ppm":
            s.set_ppm_coeffs([[w*(0.9 + 0.2*rng.random(a.size)) for w in (-1/12, 7/12, 7/12, -1/12)] for a in dxs])
        s.set_state(st0)
        dt = 1e-4
        for _ in range(6 if kw.get("flatten") else 2):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
        assert all(np.isfinite(v).all() for v in s.get_state().values())
        s.close()


def fam_io():
    import tempfile
    st0, meta = problems.make("ot", 3, (24, 20, 16))
    s = GpuStepper(3, (24, 20, 16), meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith="fast")
    s.set_state(st0)
    s.advance(1e-3)
    with tempfile.TemporaryDirectory() as d:
        if hasattr(s, "write_dbl"):
            s.write_dbl(d, 0, 0.0, 1e-3, 1)
            s.read_dbl(os.path.join(d, "data.0000.dbl"))
    if hasattr(s, "analysis"):
        s.analysis()
    s.close()


FAMILIES = {"rk": fam_rk, "exact": fam_exact, "ppm_roe": fam_ppm_roe, "hll_uct_hll": fam_hll_uct_hll, "ctu": fam_ctu,
            "bc": fam_bc, "halo": fam_halo, "io": fam_io, "grid": fam_grid, "schemes2": fam_schemes2, "schemes3": fam_schemes3}

if __name__ == "__main__":
    for name in (sys.argv[1:] or list(FAMILIES)):
        FAMILIES[name]()
        print("family", name, "done", flush=True)
