"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI
against (a) the golden fixtures the compiled reference generated and (b) the
CPU restatement on seeded inputs at sizes beyond the fixtures.

Bars:  EXACT arithmetic  -> bit-identical states, dt sequence and max Mach.
       FAST  arithmetic  -> BASELINE.json tolerances: per-variable relative L1
                            <= 1e-12 after one step, <= 1e-9 after the run
                            (<= 100 steps), dt to 1e-12, div B at round-off.
"""
import numpy as np
import pytest

from tests.util import (Golden, golden_names, divb_max, rel_l1, apply_force_field, TOL_ONE_STEP, TOL_100_STEPS, TOL_DT)

pytestmark = pytest.mark.gpu


def _stepper(g, arith):
    from pluto_b200 import GpuStepper
    s = GpuStepper(g.dims, g.n, g.dx, recon=g.recon, solver=g.solver, rk_order=g.rk_order,
                   bc=g.bc, gamma=g.gamma, arith=arith, limiter=g.limiter, emf=g.emf, flatten=g.flatten, ctu=g.ctu, en_corr=g.en_corr, grav=g.force, potential=g.potential, char_lim=g.char_lim)
    apply_force_field(s, g)
    return s


@pytest.mark.parametrize("name", golden_names())
def test_exact_bit_identical_to_reference_golden(name):
    g = Golden(name)
    s = _stepper(g, "exact")
    s.set_state(g.states[0])
    dt = g.first_dt
    for step in range(1, g.nsteps + 1):
        assert dt == g.dt[step - 1], f"dt used for step {step-1} differs from the reference tap"
        info = s.advance(dt)
        assert info.nan_events == 0
        dt = s.next_dt(info.inv_dt_hyp, g.cfl, g.cfl_max_var, dt)
        if step in g.states:
            st = s.get_state()
            for k, ref in g.states[step].items():
                assert np.array_equal(st[k], ref), f"{name}: {k} differs after {step} steps " \
                    f"(max abs diff {np.abs(st[k]-ref).max():.3e})"
    assert dt == g.dt[g.nsteps]
    st = s.get_state()
    bscale = max(np.abs(st["Bx1s"]).max(), 1e-30) / min(g.dx)
    assert divb_max(st, g.dims, g.dx_zones) < 1e-12 * bscale
    s.close()


# Roe's eigenvector normalisation switches discontinuously on `cf <= a` / `a <= cs`
# (reference Src/MHD/roe.c:336-364).  Where the transverse field vanishes EXACTLY
# (the rotor's and the 2-D blast's initial By = 0) cf2 equals a2 or b2 up to the last rounding, the
# branch taken is decided by round-off and the two branches differ by sqrt(ulp) ~ 1e-8 in
# alpha_s / alpha_f.  The FAST Roe kernels therefore evaluate everything upstream of those
# comparisons in the reference's IEEE arithmetic (mhd_device.cuh, riemann_roe): the Roe fixtures
# are held to the same 1e-12 / 1e-9 / dt 1e-12 bars as every other scheme.


@pytest.mark.parametrize("name", golden_names())
def test_fast_within_tolerance_of_reference_golden(name):
    g = Golden(name)
    if g.ctu == "chtr":
        pytest.skip("TIME_STEPPING CHARACTERISTIC_TRACING is offered with EXACT arithmetic only (test_characteristic_tracing_is_2d_only)")
    s = _stepper(g, "fast")
    s.set_state(g.states[0])
    dt = g.first_dt
    tol1, tolN, tol_dt = TOL_ONE_STEP, TOL_100_STEPS, TOL_DT
    worst = {}
    for step in range(1, g.nsteps + 1):
        assert abs(dt - g.dt[step - 1]) <= tol_dt * g.dt[step - 1], f"dt of step {step-1}"
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, g.cfl, g.cfl_max_var, dt)
        if step in g.states:
            st = s.get_state()
            tol = tol1 if step == 1 else tolN
            for k, ref in g.states[step].items():
                err = rel_l1(st[k], ref)
                worst[step] = max(worst.get(step, 0.0), err)
                assert err <= tol, f"{name}: {k} after {step} steps: {err:.3e}"
    print(f"FAST {name}: worst rel L1 " + ", ".join(f"step {k}: {v:.2e}" for k, v in worst.items()))
    st = s.get_state()
    bscale = max(np.abs(st["Bx1s"]).max(), 1e-30) / min(g.dx)
    assert divb_max(st, g.dims, g.dx_zones) < 1e-12 * bscale
    s.close()


@pytest.mark.parametrize("problem,dims,n,recon,solver", [("blast", 3, (40, 24, 20), "plm", "hlld"), ("ot", 2, (64, 48, 1), "ppm", "roe"),
                                                         ("turb", 3, (62, 20, 16), "ppm", "hll"), ("ot", 2, (31, 40, 1), "plm", "hlld")])
def test_tma_staging_bit_identical_to_cp_async(monkeypatch, problem, dims, n, recon, solver):
    """PLUTO_GPU_TMA=1: the ring rows of the fused x1+x2 sweep arrive by cp.async.bulk.tensor (one elected lane, mbarrier)
    instead of per-lane cp.async -- same values in the same shared-memory entries, so the FAST results are identical bit
    for bit (ragged last segments read zeros instead of the next row: those lanes are masked).  n1 = 31 has odd rows: the
    tensor map cannot be built (strides must be multiples of 16 bytes) and the library stays with cp.async."""
    from pluto_b200 import GpuStepper, problems
    st0, meta = problems.make(problem, dims, n)
    out = []
    for tma in ("0", "1"):
        monkeypatch.setenv("PLUTO_GPU_TMA", tma)
        s = GpuStepper(dims, n, meta["dx"], recon=recon, solver=solver, bc=meta["bc"], gamma=meta["gamma"], arith="fast")
        s.set_state(st0)
        dt = 1e-3 if problem != "blast" else 1e-4
        for _ in range(3):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
        out.append((s.get_state(), dt))
        s.close()
    assert out[0][1] == out[1][1]
    for k, v in out[0][0].items():
        assert np.array_equal(v, out[1][0][k]), k


def _plm_weights(dx):
    """PLM_CoefficientsSet (plm_coeffs.c:56-75) for a Cartesian direction with zone widths dx: xgc = zone centres."""
    xl = np.concatenate([[0.0], np.cumsum(dx)[:-1]])
    xr = xl + dx
    x = 0.5 * (xl + xr)
    six = [np.zeros(dx.size) for _ in range(6)]
    for i in range(1, dx.size - 1):
        six[0][i] = (x[i + 1] - x[i]) / (xr[i] - x[i])
        six[1][i] = (x[i] - x[i - 1]) / (x[i] - xr[i - 1])
        six[2][i] = dx[i] / (x[i + 1] - x[i])
        six[3][i] = dx[i] / (x[i] - x[i - 1])
        six[4][i] = (xr[i] - x[i]) / dx[i]
        six[5][i] = (x[i] - xr[i - 1]) / dx[i]
    return six


@pytest.mark.parametrize("problem,dims,n,solver,rk,weights", [("blast", 3, (37, 12, 10), "hlld", 2, None), ("rotor", 2, (33, 70, 1), "hlld", 3, None),
                                                             ("turb", 3, (10, 9, 12), "hll", 2, None), ("ot", 2, (64, 31, 1), "roe", 2, None),
                                                             # UNIFORM_CARTESIAN_GRID NO: grid-dependent weights, LIMITER as named
                                                             ("blast", 3, (33, 12, 10), "hlld", 2, "default"), ("rotor", 2, (33, 40, 1), "hlld", 3, "vl"),
                                                             ("turb", 3, (10, 9, 12), "roe", 2, "mc"), ("ot", 2, (40, 31, 1), "hll", 2, "os"),
                                                             # corner transport upwind (rk = 0: Hancock, -1: characteristic tracing)
                                                             ("blast", 3, (33, 12, 10), "hlld", 0, None), ("rotor", 2, (33, 40, 1), "roe", 0, None),
                                                             ("ot", 2, (40, 31, 1), "hlld", -1, None)])
def test_nonuniform_grid_random_widths(problem, dims, n, solver, rk, weights):
    """pluto_gpu_set_grid with zone widths drawn at random (0.7 .. 1.3 of the uniform one, every direction): dt/dx[i] of the
    flux differences, 1/dx[i] of the inverse time step, dt/dx2[j] ... of CT_Update and the face areas of the div B fill all
    differ from zone to zone.  EXACT: bit-identical to the oracle (pinned against the live reference on stretched grids,
    tests/test_oracle_vs_ref.py) incl. the inverse time step; FAST: within the one-step tolerance; div B at round-off."""
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    st0, meta = problems.make(problem, dims, n)
    rng = np.random.default_rng(11)
    ctu = {0: True, -1: "chtr"}.get(rk, False)
    rk = rk if rk > 0 else 2
    ng = 3 if ctu else 2
    dxs = [meta["dx"][d] * (0.7 + 0.6 * rng.random(n[d] + 2 * ng)) for d in range(dims)]
    if meta["bc"][0] == "periodic":            # the ghost zones of a periodic side repeat the interior widths
        for d in range(dims):
            a = dxs[d]
            a[:ng] = a[n[d]:n[d] + ng]
            a[n[d] + ng:] = a[ng:2 * ng]
    for arith in (("exact",) if ctu == "chtr" else ("exact", "fast")):
        lim = weights or "default"
        o = Oracle(dims, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], rk_order=rk, limiter=lim, ctu=ctu)
        s = GpuStepper(dims, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], arith=arith, rk_order=rk, limiter=lim, ctu=ctu)
        o.set_grid(*dxs)
        s.set_grid(*dxs)
        if weights:
            o.set_plm_coeffs([_plm_weights(d) for d in dxs])
            s.set_plm_coeffs([_plm_weights(d) for d in dxs])
        o.set_state(st0)
        s.set_state(st0)
        dt = 1e-4 if problem == "blast" else 1e-3
        for _ in range(3):
            inv, mach, _ = o.advance(dt)
            info = s.advance(dt)
            assert info.nan_events == 0
            if arith == "exact":
                assert info.inv_dt_hyp == inv and info.max_mach == mach
            else:
                assert abs(info.inv_dt_hyp - inv) <= 1e-12 * inv
            dt = next_dt(inv, meta["cfl"], 1.1, dt)
        a, b = s.get_state(), o.get_state()
        for k in b:
            if arith == "exact":
                assert np.array_equal(a[k], b[k]), k
            else:
                assert rel_l1(a[k], b[k]) <= TOL_ONE_STEP, k
        # constrained transport keeps the divergence of THIS metric (the initial field was built on the uniform one, so
        # it is the change that must vanish)
        zones = [dxs[d][ng:ng + n[d]] for d in range(dims)]
        bscale = max(np.abs(a["Bx1s"]).max(), 1e-30) / min(z.min() for z in zones)
        assert divb_max({k: a[k] - st0[k] for k in ("Bx1s", "Bx2s", "Bx3s") if k in a}, dims, zones) < 1e-12 * bscale
        s.close()


def test_characteristic_tracing_is_2d_only():
    """TIME_STEPPING CHARACTERISTIC_TRACING: in 3-D the reference's own result depends on never-cleared eigenvector scratch
    (as with CHAR_LIMITING), so pluto_gpu_create refuses it with that reason; PARABOLIC and UCT_HLL as for the Hancock step."""
    from pluto_b200 import GpuStepper
    from pluto_b200.stepper import PlutoGpuError
    with pytest.raises(PlutoGpuError, match="2-D only"):
        GpuStepper(3, (8, 8, 8), (0.1, 0.1, 0.1), ctu="chtr")
    with pytest.raises(PlutoGpuError, match="LINEAR"):
        GpuStepper(2, (16, 16, 1), (0.1, 0.1), ctu="chtr", recon="ppm")
    with pytest.raises(PlutoGpuError, match="SHOCK_FLATTENING"):
        GpuStepper(2, (16, 16, 1), (0.1, 0.1), ctu="chtr", flatten=True)
    # the predictor carries sqrt(cf^2 - a^2) at first order: round-off where the transverse field vanishes, so only the
    # reference's operation order reproduces the reference -- FAST arithmetic is refused instead of missing the tolerance
    with pytest.raises(PlutoGpuError, match="EXACT arithmetic only"):
        GpuStepper(2, (16, 16, 1), (0.1, 0.1), ctu="chtr", arith="fast")


@pytest.mark.parametrize("problem,dims,n,solver,kw", [
    ("blast", 3, (33, 12, 10), "hlld", dict(flatten=True)), ("blast", 2, (33, 28, 1), "roe", dict(flatten=True, emf="uct_hll")),
    ("blast", 3, (12, 14, 10), "hlld", dict(flatten=True, ctu=True, emf="uct0")),
    ("blast", 2, (28, 24, 1), "roe", dict(char_lim=True)), ("ot", 2, (32, 28, 1), "hlld", dict(char_lim=True, ctu=True, limiter="mc", emf="arith")),
    ("blast", 2, (28, 24, 1), "hlld", dict(en_corr=True)), ("blast", 3, (10, 14, 12), "hlld", dict(en_corr=True, ctu=True)),
    ("turb", 3, (10, 8, 12), "hll", dict(en_corr=True, rk_order=3, emf="uct0")),
    ("turb", 3, (10, 12, 8), "hllc", dict(ctu=True, grav=(0.3, -1.0, 0.5))), ("blast", 2, (28, 24, 1), "roe", dict(ctu=True, grav=(0.5, 0.25, 0.0), flatten=True)),
    # PARABOLIC: per-zone interface weights (any four numbers per zone serve the comparison; here the uniform ones +- 10 %)
    ("rotor", 2, (33, 40, 1), "roe", dict(recon="ppm")), ("blast", 3, (33, 12, 10), "hlld", dict(recon="ppm", rk_order=3)),
    ("turb", 3, (10, 9, 12), "hll", dict(recon="ppm", emf="uct_hll", grav=(0.3, -1.0, 0.5)))])
def test_nonuniform_grid_with_scheme_options(problem, dims, n, solver, kw):
    """Random zone widths together with the options that read the grid elsewhere: SHOCK_FLATTENING MULTID (flag_shock.c:143-145 divides
    the velocity differences by dx1[i], dx2[j], dx3[k]), CT_EN_CORRECTION (the cell-centred field of the sweeps is rebuilt with dt/dx of
    the zone), CHAR_LIMITING (uniform weights: nothing changes), BODY_FORCE inside the corner-transport-upwind steps.  The oracle is
    pinned on these combinations against the live reference (tests/test_oracle_vs_ref.py: *_nug_sfl*, *_nug_cl*, *_nug_en, *_nug_ctu_b*)."""
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    st0, meta = problems.make(problem, dims, n)
    rng = np.random.default_rng(7)
    for arith in ("exact", "fast"):
        o = Oracle(dims, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], **kw)
        s = GpuStepper(dims, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], arith=arith, **kw)
        ng = s.ng
        if arith == "exact":
            dxs = [meta["dx"][d] * (0.7 + 0.6 * rng.random(n[d] + 2 * ng)) for d in range(dims)]
            if meta["bc"][0] == "periodic":
                for d in range(dims):
                    a = dxs[d]
                    a[:ng] = a[n[d]:n[d] + ng]
                    a[n[d] + ng:] = a[ng:2 * ng]
        o.set_grid(*dxs)
        s.set_grid(*dxs)
        if kw.get("recon") == "ppm":
            if arith == "exact":
                uni = (-1.0 / 12.0, 7.0 / 12.0, 7.0 / 12.0, -1.0 / 12.0)
                qcs = [[w * (0.9 + 0.2 * rng.random(a.size)) for w in uni] for a in dxs]
            o.set_ppm_coeffs(qcs)
            s.set_ppm_coeffs(qcs)
        o.set_state(st0)
        s.set_state(st0)
        dt = 1e-4 if problem == "blast" else 1e-3
        for _ in range(4):
            inv, mach, _ = o.advance(dt)
            info = s.advance(dt)
            assert info.nan_events == 0
            if arith == "exact":
                assert info.inv_dt_hyp == inv and info.max_mach == mach
            else:
                assert abs(info.inv_dt_hyp - inv) <= 1e-12 * inv
            dt = next_dt(inv, meta["cfl"], 1.1, dt)
        a, b = s.get_state(), o.get_state()
        for k in b:
            if arith == "exact":
                assert np.array_equal(a[k], b[k]), k
            else:
                assert rel_l1(a[k], b[k]) <= TOL_ONE_STEP, k
        s.close()


def test_nonuniform_grid_is_refused_where_the_weights_would_change():
    """PARABOLIC reconstruction takes its weights from the grid (ppm_coeffs.c): on a non-uniform grid the step asks for them."""
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.stepper import PlutoGpuError
    st0, meta = problems.make("ot", 2, (16, 16, 1))
    s = GpuStepper(2, (16, 16, 1), meta["dx"], recon="ppm", bc=meta["bc"], gamma=meta["gamma"])
    s.set_grid(np.full(16 + 2 * s.ng, meta["dx"][0]), np.full(16 + 2 * s.ng, meta["dx"][1]))
    s.set_state(st0)
    with pytest.raises(PlutoGpuError, match="PPM_CoefficientsGet"):
        s.advance(1e-3)
    s.close()
    s = GpuStepper(2, (16, 16, 1), (0.1, 0.1))
    with pytest.raises(PlutoGpuError, match="dx1"):
        s.set_grid(np.zeros(20), np.full(20, 0.1))
    s.close()


@pytest.mark.parametrize("problem,n,solver", [("blast", (40, 24, 20), "hlld"), ("turb", (31, 17, 70), "roe")])
def test_x3_flux_difference_kept_apart(monkeypatch, problem, n, solver):
    """PLUTO_GPU_R3=1 (measured opt-in): the x3 sweep stores its flux difference instead of adding it to U (no U staged: 38
    shared-memory slots per thread, four blocks per SM) and the stage completion forms U + R3 -- the same sum up to the FMA
    contraction of the last addition, so the results agree to round-off, the dt sequence included."""
    from pluto_b200 import GpuStepper, problems
    st0, meta = problems.make(problem, 3, n)
    out = []
    for r3 in ("0", "1"):
        monkeypatch.setenv("PLUTO_GPU_R3", r3)
        s = GpuStepper(3, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], arith="fast")
        s.set_state(st0)
        dt = 1e-3 if problem != "blast" else 1e-4
        for _ in range(4):
            info = s.advance(dt)
            assert info.nan_events == 0
            dt = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
        out.append((s.get_state(), dt, s.device_bytes))
        s.close()
    assert out[1][2] > out[0][2]                     # the five extra arrays: the path was taken
    assert abs(out[0][1] - out[1][1]) <= 1e-12 * out[0][1]
    for k, v in out[0][0].items():
        assert rel_l1(out[1][0][k], v) <= TOL_ONE_STEP, k


def test_fast_roe_division_and_square_root_are_ieee():
    """The FAST Roe kernels keep IEEE arithmetic upstream of the solver's eigenvector switches (mhd_device.cuh) with a
    branch-free correctly rounded division / reciprocal / square root.  2e9 operand pairs (random and adversarial
    mantissas, 2^-40 .. 2^40) against div.rn.f64 / sqrt.rn.f64 on the device: no mismatch."""
    import ctypes as C
    from pluto_b200 import load_library
    bad = (C.c_ulonglong * 3)()
    for seed in (1, 20240607):
        assert load_library().pluto_gpu_selftest_arith(0, 1_000_000_000, seed, C.byref(bad)) == 0
        assert list(bad) == [0, 0, 0], f"quotient / reciprocal / root mismatches: {list(bad)}"


def test_fast_unfused_sweeps_within_tolerance(monkeypatch):
    """FAST with one kernel per direction (PLUTO_GPU_NO_FUSE_XY): the path EXACT uses, with FAST arithmetic."""
    monkeypatch.setenv("PLUTO_GPU_NO_FUSE_XY", "1")
    for name in ("blast3d_plm_hlld", "ot2d_plm_hlld", "ot3d_ppm_roe"):
        g = Golden(name)
        s = _stepper(g, "fast")
        s.set_state(g.states[0])
        dt = g.first_dt
        for step in range(1, g.nsteps + 1):
            info = s.advance(dt)
            dt = s.next_dt(info.inv_dt_hyp, g.cfl, g.cfl_max_var, dt)
            if step in g.states:
                st = s.get_state()
                tol = TOL_ONE_STEP if step == 1 else TOL_100_STEPS
                for k, ref in g.states[step].items():
                    assert rel_l1(st[k], ref) <= tol, f"{name}: {k} after {step} steps"
        s.close()


@pytest.mark.parametrize("problem,dims,n,arith", [("blast", 3, (24, 20, 28), "exact"), ("ot", 2, (48, 40, 1), "exact"),
                                                   ("turb", 3, (20, 24, 16), "fast")])
def test_device_next_dt_matches_host_loop(problem, dims, n, arith):
    """NextTimeStep evaluated on the device (steps enqueued back to back) gives the dt sequence,
    the step scalars and the state of the host-driven loop, bit for bit."""
    from pluto_b200 import GpuStepper, problems
    st0, meta = problems.make(problem, dims, n)
    mk = lambda: GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith=arith)
    a, b = mk(), mk()
    a.set_state(st0)
    b.set_state(st0)
    nsteps, dt0 = 12, 1e-4
    dts, infos, dt = [], [], dt0
    for _ in range(nsteps):
        dts.append(dt)
        info = a.advance(dt)
        infos.append(info)
        dt = a.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt)
    b.set_dt(dt0)
    for _ in range(nsteps // 2):
        b.advance_async(meta["cfl"], 1.1)
    d1, i1, _ = b.sync_results()
    for _ in range(nsteps - nsteps // 2):
        b.advance_async(meta["cfl"], 1.1)
    d2, i2, dt_next = b.sync_results()
    assert d1 + d2 == dts
    assert dt_next == dt
    for x, y in zip(i1 + i2, infos):
        assert (x.inv_dt_hyp, x.max_mach, x.floor_events, x.nan_events) == (y.inv_dt_hyp, y.max_mach, y.floor_events, y.nan_events)
    sa, sb = a.get_state(), b.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), k
    a.close()
    b.close()


# ---- against the CPU restatement at sizes beyond the fixtures ------------------
ORACLE_CASES = [
    # (problem, dims, n, recon, solver, rk_order, nsteps, first_dt)
    ("ot", 2, (96, 80, 1), "plm", "hlld", 2, 12, 5e-3),
    ("ot", 3, (40, 36, 33), "plm", "hlld", 2, 6, 1e-2),
    ("blast", 3, (33, 40, 36), "plm", "hlld", 2, 8, 2e-4),
    ("blast", 3, (24, 28, 20), "plm", "hll", 2, 5, 2e-4),
    ("blast", 3, (24, 20, 28), "plm", "roe", 2, 5, 2e-4),
    ("rotor", 2, (100, 90, 1), "ppm", "roe", 2, 10, 1e-3),
    ("turb", 3, (36, 33, 40), "ppm", "hlld", 2, 5, 1e-2),
    ("turb", 3, (30, 28, 26), "plm", "hlld", 3, 5, 1e-2),
    ("ot", 2, (70, 64, 1), "ppm", "hll", 3, 8, 5e-3),
    ("turb", 2, (64, 72, 1), "plm", "roe", 2, 8, 8e-3),
    # smallest legal blocks (n = 2*nghost), ragged extents, one-segment rows, chunk remainders
    ("turb", 3, (4, 4, 4), "plm", "hlld", 2, 4, 2e-2),
    ("turb", 3, (5, 4, 7), "plm", "hll", 2, 4, 2e-2),
    ("turb", 3, (6, 7, 6), "ppm", "hlld", 2, 4, 2e-2),
    ("ot", 2, (4, 5, 1), "plm", "hlld", 2, 4, 2e-2),
    ("blast", 2, (31, 6, 1), "ppm", "roe", 3, 4, 2e-4),
    ("ot", 2, (29, 300, 1), "plm", "hlld", 2, 3, 2e-3),
    ("ot", 3, (61, 5, 130), "plm", "hlld", 2, 3, 2e-3),
]


@pytest.mark.parametrize("case", ORACLE_CASES, ids=lambda c: f"{c[0]}{c[1]}d_{c[3]}_{c[4]}_rk{c[5]}")
def test_exact_bit_identical_to_oracle(case):
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    problem, dims, n, recon, solver, rk, nsteps, first_dt = case
    st0, meta = problems.make(problem, dims, n)
    o = Oracle(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"])
    s = GpuStepper(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"],
                   gamma=meta["gamma"], arith="exact")
    o.set_state(st0)
    s.set_state(st0)
    dt_o = dt_g = first_dt
    for step in range(nsteps):
        inv, mach, nfl = o.advance(dt_o)
        info = s.advance(dt_g)
        assert info.inv_dt_hyp == inv, f"step {step}: inv_dt_hyp {info.inv_dt_hyp!r} vs {inv!r}"
        assert info.max_mach == mach, f"step {step}: max_mach {info.max_mach!r} vs {mach!r}"
        assert info.floor_events == nfl
        dt_o = next_dt(inv, meta["cfl"], 1.1, dt_o)
        dt_g = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt_g)
        assert dt_o == dt_g
    a, b = s.get_state(), o.get_state()
    for k in b:
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k]-b[k]).max():.3e}"
    s.close()


# corner transport upwind (TIME_STEPPING HANCOCK): ctu_kernels.cuh against the restatement of
# ctu_step.c / hancock.c, which is itself pinned bit for bit by the compiled reference (tests/test_oracle_vs_ref.py)
CTU_ORACLE_CASES = [
    # (problem, dims, n, solver, nsteps, first_dt, scheme options)
    ("ot", 2, (70, 64, 1), "hlld", 10, 5e-3, {}),
    ("blast", 3, (33, 28, 24), "hlld", 6, 2e-4, {}),
    ("turb", 3, (24, 20, 28), "roe", 5, 1e-2, {}),
    ("blast", 3, (20, 24, 16), "hll", 5, 2e-4, {"limiter": "vl", "emf": "arith"}),
    ("blast", 3, (18, 16, 20), "hlld", 6, 2e-4, {"flatten": True, "emf": "uct0"}),
    ("turb", 3, (14, 12, 16), "hlld", 5, 1e-2, {"en_corr": True}),                 # CT_EN_CORRECTION on Uh and Uc
    ("blast", 2, (36, 30, 1), "roe", 6, 2e-4, {"en_corr": True, "emf": "uct0"}),
    ("blast", 3, (20, 14, 18), "hlld", 6, 2e-4, {"grav": (0.3, -1.0, 0.5)}),         # BODY_FORCE VECTOR, uniform acceleration
    ("ot", 2, (40, 33, 1), "hll", 6, 5e-3, {"grav": (-0.5, 2.0, 0.0), "en_corr": True, "emf": "arith"}),
    # smallest legal blocks (n = 2*nghost), ragged segments, chunk remainders
    ("turb", 3, (6, 6, 6), "hlld", 4, 2e-2, {}),
    ("ot", 2, (6, 7, 1), "hlld", 4, 2e-2, {}),
    ("ot", 2, (31, 200, 1), "hlld", 3, 2e-3, {}),
    ("ot", 3, (61, 7, 70), "hlld", 3, 2e-3, {}),
]


@pytest.mark.parametrize("case", CTU_ORACLE_CASES, ids=lambda c: f"ctu_{c[0]}{c[1]}d_{c[3]}_{'x'.join(map(str, c[2]))}" + ("_en" if c[6].get("en_corr") else "") + ("_bf" if c[6].get("grav") else ""))
def test_ctu_exact_bit_identical_to_oracle(case):
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    problem, dims, n, solver, nsteps, first_dt, opt = case
    st0, meta = problems.make(problem, dims, n)
    o = Oracle(dims, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], ctu=True, **opt)
    s = GpuStepper(dims, n, meta["dx"], solver=solver, bc=meta["bc"], gamma=meta["gamma"], arith="exact", ctu=True, **opt)
    assert s.nstages == 1 and s.ng == (4 if opt.get("flatten") else 3)
    
    o.set_state(st0)
    s.set_state(st0)
    dt_o = dt_g = first_dt
    for step in range(nsteps):
        inv, mach, nfl = o.advance(dt_o)
        info = s.advance(dt_g)
        assert info.inv_dt_hyp == inv, f"step {step}: inv_dt_hyp {info.inv_dt_hyp!r} vs {inv!r}"
        assert info.max_mach == mach, f"step {step}: max_mach {info.max_mach!r} vs {mach!r}"
        assert info.floor_events == nfl
        dt_o = next_dt(inv, meta["cfl"], 1.1, dt_o)
        dt_g = s.next_dt(info.inv_dt_hyp, meta["cfl"], 1.1, dt_g)
        assert dt_o == dt_g
    a, b = s.get_state(), o.get_state()
    for k in b:
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k]-b[k]).max():.3e}"
    s.close()


def test_ctu_fast_within_tolerance_of_oracle_and_device_next_dt():
    """FAST arithmetic on a seeded 3-D case beyond the fixtures, stepped with NextTimeStep on the device."""
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    n = (28, 24, 20)
    st0, meta = problems.make("turb", 3, n)
    o = Oracle(3, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], ctu=True)
    s = GpuStepper(3, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith="fast", ctu=True)
    o.set_state(st0)
    s.set_state(st0)
    dt, dts = 1e-2, []
    s.set_dt(dt)
    for _ in range(8):
        dts.append(dt)
        inv, _, _ = o.advance(dt)
        dt = next_dt(inv, meta["cfl"], 1.1, dt)
        s.advance_async(meta["cfl"], 1.1)
    d, infos, dt_next = s.sync_results()
    assert len(d) == 8 and all(abs(x - y) <= TOL_DT*y for x, y in zip(d, dts)) and abs(dt_next - dt) <= TOL_DT*dt
    a, b = s.get_state(), o.get_state()
    for k in b:
        assert rel_l1(a[k], b[k]) <= TOL_100_STEPS, k
    bscale = max(np.abs(a["Bx1s"]).max(), 1e-30) / min(meta["dx"])
    assert divb_max(a, 3, meta["dx"]) < 1e-12 * bscale
    s.close()


@pytest.mark.parametrize("problem,dims,n,recon,solver,rk,emf", [
    ("turb", 3, (20, 24, 16), "ppm", "hlld", 2, "uct_contact"), ("blast", 3, (24, 20, 28), "plm", "roe", 3, "uct0"),
    ("ot", 2, (70, 64, 1), "plm", "hll", 2, "arith"), ("blast", 2, (40, 36, 1), "plm", "hlld", 2, "uct_contact")])
def test_en_correction_bit_identical_to_oracle(problem, dims, n, recon, solver, rk, emf):
    """CT_EN_CORRECTION YES (ct_field_average.c:116-129): the energy correction needs the cell-centred field the
    reference's sweeps carry in Uc; final_kernel rebuilds it from the face EMFs (+ the round-off residue of the
    normal-component flux), bit for bit.  Single block and four blocks."""
    import os
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.parallel import BlockLayout, LocalMultiBlock
    st0, meta = problems.make(problem, dims, n)
    o = Oracle(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], emf=emf, en_corr=True)
    s = GpuStepper(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], emf=emf,
                   arith="exact", en_corr=True)
    lay = BlockLayout.strong(dims, n, 4 if dims == 3 else 2, periodic=meta["bc"][0] == "periodic")
    many = LocalMultiBlock(lay, meta["dx"], meta["bc"], recon=recon, solver=solver, rk_order=rk, gamma=meta["gamma"], emf=emf,
                           en_corr=True, exchange="all", split=True,
                           host_buffers=os.environ.get("PLUTO_GPU_LIB", "").endswith("_emu.so"))
    o.set_state(st0); s.set_state(st0); many.set_state(st0)
    dt = {"ot": 5e-3, "blast": 2e-4, "turb": 1e-2}[problem]
    for step in range(5):
        inv, mach, nfl = o.advance(dt)
        info = s.advance(dt)
        many.advance(dt)
        assert (info.inv_dt_hyp, info.max_mach, info.floor_events) == (inv, mach, nfl), step
        dt = next_dt(inv, meta["cfl"], 1.1, dt)
    a, b, c = s.get_state(), o.get_state(), many.get_state()
    for k in b:
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k]-b[k]).max():.3e}"
        assert np.array_equal(c[k], b[k]), f"{k} (4 blocks): max abs diff {np.abs(c[k]-b[k]).max():.3e}"
    s.close()
    for blk in many.blocks:
        blk.close()


@pytest.mark.parametrize("problem,dims,n,recon,solver,rk", [
    ("blast", 3, (24, 20, 28), "plm", "hlld", 2), ("turb", 3, (16, 20, 12), "ppm", "roe", 3), ("rotor", 2, (50, 44, 1), "ppm", "hll", 2),
    ("ot", 2, (61, 40, 1), "plm", "roe", 3)])
def test_body_force_bit_identical_to_oracle(problem, dims, n, recon, solver, rk):
    """BODY_FORCE VECTOR with a uniform acceleration (rhs_source.c:214-217, 277-280, 342-345), RK2 / RK3, single
    block and decomposed (every exchange mode of the decomposition test is covered by the goldens' schemes)."""
    import os
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.parallel import BlockLayout, LocalMultiBlock
    grav = (0.7, -1.3, 0.4) if dims == 3 else (0.7, -1.3, 0.0)
    st0, meta = problems.make(problem, dims, n)
    o = Oracle(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], grav=grav)
    s = GpuStepper(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], arith="exact",
                   grav=grav)
    lay = BlockLayout.strong(dims, n, 4 if dims == 3 else 2, periodic=meta["bc"][0] == "periodic")
    many = LocalMultiBlock(lay, meta["dx"], meta["bc"], recon=recon, solver=solver, rk_order=rk, gamma=meta["gamma"], grav=grav,
                           exchange="all", split=True, host_buffers=os.environ.get("PLUTO_GPU_LIB", "").endswith("_emu.so"))
    o.set_state(st0); s.set_state(st0); many.set_state(st0)
    dt = {"ot": 5e-3, "blast": 2e-4, "turb": 1e-2, "rotor": 1e-3}[problem]
    for step in range(5):
        inv, mach, nfl = o.advance(dt)
        info = s.advance(dt)
        many.advance(dt)
        assert (info.inv_dt_hyp, info.max_mach, info.floor_events) == (inv, mach, nfl), step
        dt = next_dt(inv, meta["cfl"], 1.1, dt)
    a, b, c = s.get_state(), o.get_state(), many.get_state()
    for k in b:
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k]-b[k]).max():.3e}"
        assert np.array_equal(c[k], b[k]), f"{k} (blocks): max abs diff {np.abs(c[k]-b[k]).max():.3e}"
    s.close()
    for blk in many.blocks:
        blk.close()
    # FAST arithmetic (the fused x1+x2 sweep carries the source as well) within the one-step tolerance
    f = GpuStepper(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], arith="fast",
                   grav=grav)
    o2 = Oracle(dims, n, meta["dx"], recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], grav=grav)
    f.set_state(st0); o2.set_state(st0)
    dt = {"ot": 5e-3, "blast": 2e-4, "turb": 1e-2, "rotor": 1e-3}[problem]
    f.advance(dt); o2.advance(dt)
    a, b = f.get_state(), o2.get_state()
    tol = TOL_ONE_STEP
    for k in b:
        assert rel_l1(a[k], b[k]) <= tol, k
    f.close()


@pytest.mark.parametrize("problem,dims,n,recon,solver,rk,ctu,vector", [
    ("blast", 3, (20, 16, 24), "plm", "hlld", 2, False, False), ("rotor", 2, (44, 40, 1), "ppm", "hll", 3, False, True),
    ("blast", 3, (16, 20, 12), "plm", "roe", 2, True, True), ("blast", 2, (40, 30, 1), "plm", "hlld", 2, True, False)])
def test_body_potential_bit_identical_to_oracle(problem, dims, n, recon, solver, rk, ctu, vector):
    """BODY_FORCE POTENTIAL (gravitational energy flux rhs.c:388-392, sources rhs_source.c:233-237, 316-320, 358-362, predictor
    source prim_eqn.c:304-307), alone and together with a uniform VECTOR force; RK2 / RK3 / CTU; one block and decomposed."""
    import os
    from oracle.oracle_lib import Oracle, next_dt
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.parallel import BlockLayout, LocalMultiBlock
    from tests.util import step_potential_arrays
    st0, meta = problems.make(problem, dims, n)
    dom = ((-0.5, 0.5),) * 3
    steps = (0.05, -0.03, 0.04 if dims == 3 else 0.0)
    grav = ((0.4, -0.9, 0.3) if dims == 3 else (0.4, -0.9, 0.0)) if vector else None
    kw = dict(recon=recon, solver=solver, rk_order=rk, gamma=meta["gamma"], ctu=ctu, grav=grav)
    o = Oracle(dims, n, meta["dx"], bc=meta["bc"], **kw)
    o.set_body_potential(*step_potential_arrays(dims, n, o.ng, dom, steps))
    s = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], arith="exact", potential=True, **kw)
    s.set_body_potential(*step_potential_arrays(dims, n, s.ng, dom, steps))
    o.set_state(st0); s.set_state(st0)
    dt = {"blast": 2e-4, "rotor": 1e-3}[problem]
    for step in range(5):
        inv, mach, nfl = o.advance(dt)
        info = s.advance(dt)
        assert (info.inv_dt_hyp, info.max_mach, info.floor_events) == (inv, mach, nfl), step
        dt = next_dt(inv, meta["cfl"], 1.1, dt)
    a, b = s.get_state(), o.get_state()
    for k in b:
        assert np.array_equal(a[k], b[k]), f"{k}: max abs diff {np.abs(a[k]-b[k]).max():.3e}"
    s.close()
    f = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], arith="fast", potential=True, **kw)
    f.set_body_potential(*step_potential_arrays(dims, n, f.ng, dom, steps))
    o2 = Oracle(dims, n, meta["dx"], bc=meta["bc"], **kw)
    o2.set_body_potential(*step_potential_arrays(dims, n, o2.ng, dom, steps))
    f.set_state(st0); o2.set_state(st0)
    f.advance(dt); o2.advance(dt)
    a, b = f.get_state(), o2.get_state()
    for k in b:
        assert rel_l1(a[k], b[k]) <= TOL_ONE_STEP, k
    f.close()


def test_reflective_boundaries_match_oracle():
    from oracle.oracle_lib import Oracle
    from pluto_b200 import GpuStepper, problems
    n = (24, 20, 16)
    st0, meta = problems.make("blast", 3, n, radius=0.3)
    bc = ("reflective", "outflow", "outflow", "reflective", "reflective", "outflow")
    o = Oracle(3, n, meta["dx"], bc=bc, gamma=meta["gamma"])
    s = GpuStepper(3, n, meta["dx"], bc=bc, gamma=meta["gamma"])
    o.set_state(st0)
    s.set_state(st0)
    for _ in range(6):
        o.advance(3e-4)
        s.advance(3e-4)
    a, b = s.get_state(), o.get_state()
    for k in b:
        assert np.array_equal(a[k], b[k]), k
    s.close()


def test_data_layout_roundtrip_and_advance_data():
    """The literal AdvanceStep contract on host Data arrays (with ghost zones)."""
    from pluto_b200 import GpuStepper, problems
    n = (20, 16, 12)
    st0, meta = problems.make("ot", 3, n)
    s = GpuStepper(3, n, meta["dx"], gamma=meta["gamma"])
    s.set_state(st0)
    ref = GpuStepper(3, n, meta["dx"], gamma=meta["gamma"])
    ref.set_state(st0)
    Vc, s1, s2, s3 = s.data_buffers()
    s.download_data(Vc, s1, s2, s3)
    g = s.ng
    assert np.array_equal(Vc[0, g:-g, g:-g, g:-g], st0["rho"])
    assert np.array_equal(Vc[7, g:-g, g:-g, g:-g], st0["prs"])
    assert np.array_equal(s1[g:-g, g:-g, g:g + n[0] + 1], st0["Bx1s"])
    assert np.array_equal(s3[g:g + n[2] + 1, g:-g, g:-g], st0["Bx3s"])
    info = s.advance_data(1e-2, Vc, s1, s2, s3)
    info_ref = ref.advance(1e-2)
    assert info.inv_dt_hyp == info_ref.inv_dt_hyp
    want = ref.get_state()
    assert np.array_equal(Vc[0, g:-g, g:-g, g:-g], want["rho"])
    assert np.array_equal(Vc[3, g:-g, g:-g, g:-g], want["vx3"])
    assert np.array_equal(s2[g:-g, g:g + n[1] + 1, g:-g], want["Bx2s"])
    s.close(); ref.close()


def test_data_layout_2d_has_six_variables():
    from pluto_b200 import GpuStepper, problems
    n = (24, 20, 1)
    st0, meta = problems.make("ot", 2, n)
    s = GpuStepper(2, n, meta["dx"])
    s.set_state(st0)
    Vc, s1, s2, s3 = s.data_buffers()
    assert Vc.shape[0] == 6 and s3 is None
    s.download_data(Vc, s1, s2)
    g = s.ng
    for q, nm in enumerate(["rho", "vx1", "vx2", "Bx1", "Bx2", "prs"]):
        assert np.array_equal(Vc[q, 0, g:-g, g:-g], st0[nm][0]), nm
    s.close()


def test_full_size_properties_blast_256():
    """BASELINE config 2 at full size: size-independent properties."""
    from pluto_b200 import GpuStepper, Integrator, problems
    n = (256, 256, 256)
    st0, meta = problems.make("blast", 3, n)
    s = GpuStepper(3, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"], arith="exact")
    s.set_state(st0)
    it = Integrator(s, cfl=0.3, first_dt=1e-5)
    it.run(4)
    st = s.get_state()
    # div B stays at round-off
    bscale = np.abs(st["Bx1s"]).max() / min(meta["dx"])
    assert divb_max(st, 3, meta["dx"]) < 1e-12 * bscale
    # the blast is centred and B lies in the x-z plane: y -> -y mirror symmetry
    # of rho holds to round-off (symmetric arithmetic is not guaranteed bitwise)
    rho = st["rho"]
    assert np.abs(rho - rho[:, ::-1, :]).max() < 1e-9
    # mass is conserved while nothing has reached the outflow boundaries
    assert abs(rho.sum() - st0["rho"].sum()) < 1e-9 * st0["rho"].sum()
    assert np.isfinite(st["prs"]).all() and st["prs"].min() > 0
    # dt history: ramp by cfl_max_var from first_dt (main.c:532)
    assert it.dt_history[1] == pytest.approx(1.1e-5, rel=1e-12)
    s.close()


# ---- decomposition: several blocks + halo exchange == one block, bit for bit -----
@pytest.mark.parametrize("problem,dims,gn,world,recon,solver", [
    ("ot", 3, (32, 24, 32), 8, "plm", "hlld"),
    ("blast", 3, (24, 32, 24), 4, "plm", "hlld"),
    ("turb", 3, (16, 24, 32), 2, "ppm", "roe"),
    ("ot", 2, (48, 64, 1), 4, "plm", "hlld"),
    ("rotor", 2, (40, 48, 1), 2, "ppm", "roe"),
    # corner transport upwind: ONE exchange per step, three ghost layers
    ("ot", 3, (24, 32, 24), 8, "ctu", "hlld"),
    ("blast", 3, (24, 16, 32), 4, "ctu", "roe"),
    ("ot", 2, (48, 64, 1), 4, "ctu", "hlld"),
])
@pytest.mark.parametrize("exchange", ["dims", "all", "all+split"])
def test_decomposed_blocks_match_single_block(problem, dims, gn, world, recon, solver, exchange):
    import os
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.parallel import BlockLayout, LocalMultiBlock
    st0, meta = problems.make(problem, dims, gn)
    periodic = meta["bc"][0] == "periodic"
    lay = BlockLayout.strong(dims, gn, world, periodic=periodic)
    ctu = recon == "ctu"
    recon = "plm" if ctu else recon
    one = GpuStepper(dims, gn, meta["dx"], recon=recon, solver=solver, bc=meta["bc"], gamma=meta["gamma"], ctu=ctu)
    # "all+split": every stage issued as shell + interior, the form the overlapped NCCL exchange uses
    many = LocalMultiBlock(lay, meta["dx"], meta["bc"], recon=recon, solver=solver, gamma=meta["gamma"], ctu=ctu,
                           exchange=exchange.split("+")[0], split=exchange.endswith("+split"),
                           host_buffers=os.environ.get("PLUTO_GPU_LIB", "").endswith("_emu.so"))
    one.set_state(st0)
    many.set_state(st0)
    dt = {"ot": 5e-3, "blast": 2e-4, "turb": 5e-3, "rotor": 1e-3}[problem]
    for step in range(4):
        a = one.advance(dt)
        b = many.advance(dt)
        assert a.inv_dt_hyp == b.inv_dt_hyp and a.max_mach == b.max_mach, step
        dt = one.next_dt(a.inv_dt_hyp, meta["cfl"], 1.1, dt)
    sa, sb = one.get_state(), many.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), f"{k}: max abs diff {np.abs(sa[k]-sb[k]).max():.3e}"
    one.close()
    for blk in many.blocks:
        blk.close()


@pytest.mark.parametrize("problem,dims,gn,world,weights", [("blast", 3, (24, 32, 24), 4, False), ("ot", 3, (32, 24, 32), 8, True),
                                                          ("ot", 2, (48, 64, 1), 4, True)])
def test_decomposed_blocks_on_a_nonuniform_grid(problem, dims, gn, world, weights):
    """Blocks of a decomposed domain on a non-uniform grid (each takes its slice of the zone widths and of the reconstruction
    weights; the exchange itself knows nothing of the grid): bit-identical to the single block, overlapped form included."""
    import os
    from pluto_b200 import GpuStepper, problems
    from pluto_b200.parallel import BlockLayout, LocalMultiBlock
    st0, meta = problems.make(problem, dims, gn)
    periodic = meta["bc"][0] == "periodic"
    lay = BlockLayout.strong(dims, gn, world, periodic=periodic)
    rng = np.random.default_rng(17)
    ng = 2
    dxs = [meta["dx"][d] * (0.7 + 0.6 * rng.random(gn[d] + 2 * ng)) for d in range(dims)]
    if periodic:
        for d in range(dims):
            a = dxs[d]
            a[:ng] = a[gn[d]:gn[d] + ng]
            a[gn[d] + ng:] = a[ng:2 * ng]
    one = GpuStepper(dims, gn, meta["dx"], bc=meta["bc"], gamma=meta["gamma"])
    many = LocalMultiBlock(lay, meta["dx"], meta["bc"], gamma=meta["gamma"], exchange="all", split=True,
                           host_buffers=os.environ.get("PLUTO_GPU_LIB", "").endswith("_emu.so"))
    one.set_grid(*dxs)
    many.set_grid(*dxs)
    if weights:
        cs = [_plm_weights(d) for d in dxs]
        if periodic:                       # the weights of the ghost zones repeat the interior ones as well
            for d in range(dims):
                for a in cs[d]:
                    a[:ng] = a[gn[d]:gn[d] + ng]
                    a[gn[d] + ng:] = a[ng:2 * ng]
        one.set_plm_coeffs(cs)
        many.set_plm_coeffs(cs)
    one.set_state(st0)
    many.set_state(st0)
    dt = {"ot": 5e-3, "blast": 2e-4}[problem]
    for step in range(3):
        a = one.advance(dt)
        b = many.advance(dt)
        assert a.inv_dt_hyp == b.inv_dt_hyp and a.max_mach == b.max_mach, step
        dt = one.next_dt(a.inv_dt_hyp, meta["cfl"], 1.1, dt)
    sa, sb = one.get_state(), many.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), f"{k}: max abs diff {np.abs(sa[k]-sb[k]).max():.3e}"
    one.close()
    for blk in many.blocks:
        blk.close()


def test_dbl_output_restart_and_analysis(tmp_path):
    """Device-side output in the reference's .dbl format (read back with the reader used for the reference's
    own dumps), restart from it, and the diagnostics reductions against numpy."""
    from oracle.refrun import read_dbl
    from pluto_b200 import GpuStepper, problems
    for dims, n in ((3, (20, 16, 12)), (2, (24, 20, 1))):
        st0, meta = problems.make("ot", dims, n)
        s = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"])
        s.set_state(st0)
        for _ in range(3):
            s.advance(5e-3)
        st = s.get_state()
        s.write_dbl(str(tmp_path), 0, 0.015, 5e-3, 3)
        s.write_dbl(str(tmp_path), 1, 0.015, 5e-3, 3)
        back = read_dbl(str(tmp_path / "data.0001.dbl"), dims, n)
        for k, v in st.items():
            assert np.array_equal(back[k], v), k
        lines = open(tmp_path / "dbl.out").read().splitlines()
        assert len(lines) == 2 and lines[1].split()[0] == "1" and lines[1].split()[4:6] == ["single_file", "little"]
        assert lines[1].split()[6:] == (["rho", "vx1", "vx2", "vx3", "Bx1", "Bx2", "Bx3", "prs", "Bx1s", "Bx2s", "Bx3s"]
                                        if dims == 3 else ["rho", "vx1", "vx2", "Bx1", "Bx2", "prs", "Bx1s", "Bx2s"])
        # a run restarted from file 0 writes file 1 again: its line REPLACES line 1 and drops what followed
        # (write_data.c:369-376 opens r+ and skips nfile lines), it is not appended
        s.write_dbl(str(tmp_path), 2, 0.020, 5e-3, 4)
        s.write_dbl(str(tmp_path), 1, 0.017, 5e-3, 7)
        lines = open(tmp_path / "dbl.out").read().splitlines()
        assert [l.split()[0] for l in lines] == ["0", "1"] and lines[1].split()[3] == "7"
        r = GpuStepper(dims, n, meta["dx"], bc=meta["bc"], gamma=meta["gamma"])
        r.read_dbl(str(tmp_path / "data.0000.dbl"))
        a, b = s.advance(5e-3), r.advance(5e-3)
        assert (a.inv_dt_hyp, a.max_mach) == (b.inv_dt_hyp, b.max_mach)
        sa, sb = s.get_state(), r.get_state()
        for k in sa:
            assert np.array_equal(sa[k], sb[k]), k
        an = s.analysis()
        vol = float(np.prod(meta["dx"][:dims]))
        v2 = sa["vx1"]**2 + sa["vx2"]**2 + (sa["vx3"]**2 if dims == 3 else 0.0)
        b2 = sa["Bx1"]**2 + sa["Bx2"]**2 + (sa["Bx3"]**2 if dims == 3 else 0.0)
        assert abs(an["mass"] - sa["rho"].sum()*vol) <= 1e-12*abs(an["mass"])
        assert abs(an["e_kin"] - (0.5*sa["rho"]*v2).sum()*vol) <= 1e-12*abs(an["e_kin"])
        assert abs(an["e_mag"] - (0.5*b2).sum()*vol) <= 1e-12*abs(an["e_mag"])
        assert abs(an["e_th"] - (sa["prs"]/(meta["gamma"] - 1.0)).sum()*vol) <= 1e-12*abs(an["e_th"])
        assert an["max_divb"] <= 1.01*divb_max(sa, dims, meta["dx"]) + 1e-300 and an["max_divb"] < 1e-10
        s.close()
        r.close()


@pytest.mark.parametrize("dims,n", [(3, (12, 10, 8)), (2, (16, 12, 1))])
def test_flt_and_vtk_output_match_the_reference_files(tmp_path, dims, n):
    """Device-side single-precision output: data.NNNN.flt and the legacy-VTK file written from the device state are the
    reference's own files BYTE FOR BYTE (header text, big-endian node coordinates, SCALARS blocks; flt.out / vtk.out
    lines), for the state after 3 steps of the same run (EXACT arithmetic, the reference's dt sequence)."""
    from oracle.refrun import RefConfig, have_ref, run_reference, read_dbl
    from pluto_b200 import GpuStepper
    cfg = RefConfig(problem="ot", dims=dims, n=n, first_dt=1e-3, cfl=0.3)
    if not have_ref(cfg):
        pytest.skip("oracle/_ref not built")
    ref = run_reference(cfg, maxsteps=2, dump_every=1, flt_every=1, vtk_every=1, workdir=str(tmp_path / "ref"), keep=True)
    dom = cfg.resolved_domain()
    dx = [(dom[d][1] - dom[d][0]) / n[d] for d in range(dims)]
    s = GpuStepper(dims, n, dx, bc=cfg.resolved_bc(), gamma=cfg.resolved_gamma(), arith="exact")
    s.read_dbl(str(tmp_path / "ref" / "data.0000.dbl"))
    dt = cfg.first_dt
    for step in range(3):
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, cfg.cfl, cfg.cfl_max_var, dt)
    st, last = s.get_state(), read_dbl(str(tmp_path / "ref" / "data.0002.dbl"), dims, n)
    for k, v in last.items():
        assert np.array_equal(st[k], v), k          # same state as the reference's last dump, bit for bit
    out = tmp_path / "gpu"
    out.mkdir()
    # node coordinates as the reference builds them (set_grid.c:400-402: xlft = xL + (i - iL)*dx)
    xl = [dom[d][0] + np.arange(n[d] + 1) * dx[d] for d in range(dims)]
    lines = {e: open(tmp_path / "ref" / f"{e}.out").read().splitlines() for e in ("flt", "vtk")}
    for nfile in range(3):          # the .out lists grow line by line; only file 2 holds this state
        w = lines["flt"][nfile].split()
        s.write_flt(str(out), nfile, float(w[1]), float(w[2]), int(w[3]))
        s.write_vtk(str(out), nfile, float(w[1]), float(w[2]), int(w[3]), xl)
    for ext in ("flt", "vtk"):
        a = open(out / f"data.0002.{ext}", "rb").read()
        b = open(tmp_path / "ref" / f"data.0002.{ext}", "rb").read()
        assert a == b, f"data.0002.{ext}: {len(a)} vs {len(b)} bytes, first difference at " \
            f"{next((q for q in range(min(len(a), len(b))) if a[q] != b[q]), -1)}"
        assert open(out / f"{ext}.out").read() == open(tmp_path / "ref" / f"{ext}.out").read()
    s.close()


@pytest.mark.parametrize("problem,dims,n,grid,recon,solver,rk,ctu", [
    ("turb", 3, (16, 16, 16), (2, 2, 2), "plm", "hlld", 2, False), ("blast", 3, (16, 12, 24), (1, 2, 2), "plm", "hlld", 2, False),
    ("ot", 2, (32, 24, 1), (2, 2, 1), "ppm", "roe", 3, False), ("rotor", 2, (32, 24, 1), (1, 2, 1), "plm", "hll", 2, False),
    ("ot", 3, (16, 16, 12), (2, 1, 2), "plm", "hlld", 2, True)])
def test_single_thread_multi_block_stepper_matches_single_block(problem, dims, n, grid, recon, solver, rk, ctu):
    """pluto_gpu_multi_*: the domain cut into blocks that ONE host thread drives through the C ABI (peer stores between the
    blocks inside the pack launch, events between the streams) reproduces the single-block run bit for bit, from the
    reference's Data arrays of the whole domain and back.  On a one-GPU box all blocks share device 0; on several GPUs the
    blocks are spread over them (tests/test_gpu_dist.py)."""
    import torch
    from pluto_b200 import GpuStepper, MultiGpuStepper, problems
    st0, meta = problems.make(problem, dims, n)
    kw = dict(recon=recon, solver=solver, rk_order=rk, bc=meta["bc"], gamma=meta["gamma"], arith="exact", ctu=ctu)
    one = GpuStepper(dims, n, meta["dx"], **kw)
    nb = grid[0] * grid[1] * grid[2]
    ndev = torch.cuda.device_count() if torch.cuda.is_available() else 1
    many = MultiGpuStepper(dims, n, meta["dx"], grid, devices=[b % max(ndev, 1) for b in range(nb)], **kw)
    assert many.nblocks == nb
    one.set_state(st0)
    many.set_state(st0)
    dt = {"ot": 5e-3, "blast": 2e-4, "turb": 5e-3, "rotor": 1e-3}[problem]
    for step in range(4):
        a, b = one.advance(dt), many.advance(dt)
        assert (a.inv_dt_hyp, a.max_mach, a.floor_events) == (b.inv_dt_hyp, b.max_mach, b.floor_events), step
        dt = one.next_dt(a.inv_dt_hyp, meta["cfl"], 1.1, dt)
    sa, sb = one.get_state(), many.get_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k]), f"{k}: max abs diff {np.abs(sa[k] - sb[k]).max():.3e}"
    one.close()
    many.close()
