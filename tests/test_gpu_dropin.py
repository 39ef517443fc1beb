"""The drop-in proof: the reference's OWN driver (main loop, pluto.ini parser,
Init(), .dbl writer -- compiled from the unmodified sources) linked with
integration/advance_step_gpu.c + libpluto_gpu.so instead of rk_step.o /
update_stage.o must reproduce the dumps of the all-CPU reference build:
bit-identical states and dt sequence with EXACT arithmetic, within the
BASELINE.json tolerances with FAST arithmetic.  The binaries are built in the
development container (integration/build_shim.sh) and travel with the snapshot."""
import numpy as np
import pytest

from oracle.refrun import RefConfig, have_ref, run_reference
from tests.util import Golden, rel_l1, TOL_ONE_STEP, TOL_100_STEPS

pytestmark = pytest.mark.gpu

CASES = ["ot2d_plm_hlld", "blast3d_plm_hlld", "rotor2d_ppm_roe", "ot3d_ppm_roe", "ot2d_plm_hll", "turb3d_plm_hlld",
         # LIMITER / CT_EMF_AVERAGE read from definitions.h by the shim (Blast #02's and Rotor #01's scheme options)
         "blast3d_vl_arith", "rotor2d_mc_arith", "blast3d_mc_uct_hll_roe", "blast3d_sfl",
         # TIME_STEPPING HANCOCK: the shim replaces ctu_step.o
         "ot2d_ctu", "blast3d_ctu", "turb3d_ctu_roe",
         # CT_EN_CORRECTION YES: the complete scheme of the shipped Blast #02 (definitions_02.h, pluto_02.ini)
         "blast3d_blast02_en", "blast2d_en", "blast3d_ctu_en",
         # BODY_FORCE VECTOR: the shim samples init.c's BodyForceVector and passes the uniform acceleration
         "blast3d_bf", "rotor2d_ppm_rk3_bf", "turb3d_ctu_bf",
         # position-dependent force: tabulated per zone by the shim (pluto_gpu_set_body_force)
         "blast3d_bfx", "blast2d_ctu_bfx_roe",
         # BODY_FORCE POTENTIAL: the shim tabulates BodyForcePotential at the zone centres and faces
         "blast3d_bp", "blast2d_ctu_bp",
         # CHAR_LIMITING YES (2-D): the shim reads it from definitions.h
         "ot2d_cl", "rotor2d_cl_vl_rk3",
         # non-uniform grids (uniform + stretched patches in pluto.ini): the shim hands grid->dx to pluto_gpu_set_grid
         "blast3d_nug", "rotor2d_nug_roe_rk3", "blast2d_nug_mc_arith_reflective",
         # UNIFORM_CARTESIAN_GRID NO: the shim hands the arrays of PLM_CoefficientsGet to pluto_gpu_set_plm_coeffs
         "blast3d_nuw", "blast2d_nuw_mc_arith",
         # TIME_STEPPING CHARACTERISTIC_TRACING (2-D): the shim replaces ctu_step.o, plm_states.o keeps char_tracing.o
         "ot2d_chtr", "rotor2d_chtr_mc_uct0", "blast2d_chtr_mc_roe",
         # CHAR_LIMITING YES with the corner-transport-upwind steps (Orszag_Tang #09's scheme; with characteristic tracing)
         "ot2d_ctu_cl_mc_arith", "blast2d_chtr_cl",
         # SHOCK_FLATTENING MULTID with PARABOLIC reconstruction: the shim hands over the weights of PLM_CoefficientsGet
         "blast2d_ppm_sfl_roe", "blast3d_ppm_sfl",
         # body forces on non-uniform grids: potential and position-dependent force tabulated by the shim at grid->x / xr
         "blast3d_nug_bp", "blast2d_nug_bfx_roe",
         # the corner-transport-upwind steps on non-uniform grids
         "blast3d_nug_ctu", "rotor2d_nug_chtr_mc",
         # [Solver] hllc / tvdlf
         "blast3d_hllc", "rotor2d_ppm_rk3_hllc", "turb3d_ctu_hllc", "ot2d_tvdlf",
         # BODY_FORCE with UCT_HLL
         "blast3d_bf_uct_hll", "rotor2d_ppm_rk3_bp_uct_hll_roe",
         # BODY_FORCE with SHOCK_FLATTENING MULTID
         "blast3d_sfl_bf", "blast3d_ctu_sfl_bf", "blast2d_ppm_sfl_bp",
         # non-uniform grids with SHOCK_FLATTENING, CHAR_LIMITING, CT_EN_CORRECTION, BODY_FORCE inside the corner-transport-upwind step
         "blast3d_nug_sfl", "blast2d_nug_cl_roe", "blast2d_nug_en", "blast2d_nug_ctu_bp",
         # PARABOLIC on non-uniform grids: the shim hands over the weights of PPM_CoefficientsGet
         "rotor2d_nug_ppm", "blast3d_nug_ppm_roe", "blast2d_nug_ppm_sfl_rk3",
         # PARABOLIC + SHOCK_FLATTENING MULTID with the default average UCT_HLL
         "blast3d_ppm_sfl_uct_hll", "blast2d_ppm_sfl_uct_hll_roe",
         # CHAR_LIMITING with UCT_HLL / BODY_FORCE
         "ot2d_cl_uct_hll", "blast2d_cl_bf_roe", "blast2d_ctu_cl_bf"]


def _blast_params(g):
    """The blast fixtures on walls use a larger sphere (tools/make_golden.py); everything else the default parameters."""
    b = dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.125)
    if g.name == "blast2d_nug_mc_arith_reflective":
        b["RADIUS"] = 0.3
    return b


def _cfg(g):
    return RefConfig(problem=g.problem, dims=g.dims, n=g.n, recon=g.recon, solver=g.solver, tstep=g.tstep,
                     cfl=g.cfl, cfl_max_var=g.cfl_max_var, first_dt=g.first_dt, gamma=g.gamma,
                     limiter=g.limiter, emf=g.emf, flatten=g.flatten, en_corr=g.en_corr, grav=g.grav, grav_mode=g.grav_mode, potential=g.potential,
                     char_lim=g.char_lim, grid=g.grid, grid_weights=g.grid_weights, bc=g.bc, blast=_blast_params(g), prefix="pluto_gpu_")


@pytest.mark.parametrize("name", CASES)
def test_reference_driver_with_gpu_step_is_bit_identical(name):
    g = Golden(name)
    cfg = _cfg(g)
    if not have_ref(cfg):
        pytest.skip("oracle/_ref/pluto_gpu_* not built (integration/build_shim.sh)")
    r = run_reference(cfg, maxsteps=g.nsteps + 1, dump_every=1, env={"PLUTO_GPU_ARITH": "exact"})
    tap = {int(a): c for a, b, c in r.dt_tap}
    for s in range(1, g.nsteps + 1):
        assert tap[s] == g.dt[s], f"dt after step {s}"
    for s, ref in g.states.items():
        for k, v in ref.items():
            assert np.array_equal(r.dumps[s][k], v), f"{name}: {k} after {s} steps"


@pytest.mark.parametrize("name", ["ot2d_plm_hlld_100", "blast3d_plm_hlld_100"])
def test_reference_driver_with_gpu_step_fast_100_steps(name):
    g = Golden(name)
    cfg = _cfg(g)
    if not have_ref(cfg):
        pytest.skip("oracle/_ref/pluto_gpu_* not built (integration/build_shim.sh)")
    r = run_reference(cfg, maxsteps=g.nsteps + 1, dump_every=1, env={"PLUTO_GPU_ARITH": "fast"})
    tap = {int(a): c for a, b, c in r.dt_tap}
    for s in range(1, g.nsteps + 1):
        assert abs(tap[s] - g.dt[s]) <= 1e-12 * g.dt[s], f"dt after step {s}"
    for s, ref in g.states.items():
        tol = TOL_ONE_STEP if s <= 1 else TOL_100_STEPS
        for k, v in ref.items():
            assert rel_l1(r.dumps[s][k], v) <= tol, f"{name}: {k} after {s} steps"


@pytest.mark.parametrize("name,arith", [("blast3d_plm_hlld_100", "exact"), ("ot2d_ctu_100", "exact"), ("rotor2d_ppm_roe_100", "exact"),
                                        ("turb3d_plm_hlld_100", "fast")])
def test_resident_state_drop_in(name, arith):
    """PLUTO_GPU_RESIDENT=1: the state is uploaded once, advanced in HBM by pluto_gpu_advance and copied back only where
    the driver reads it (the two one-line hooks of INTEGRATION.md in WriteData and CheckForAnalysis).  With a dump and an
    analysis call every 25 steps the host arrays are refreshed 5 times in 100 steps, and every dump is the all-CPU
    reference's, bit for bit (EXACT) / within tolerance (FAST)."""
    g = Golden(name)
    cfg = _cfg(g)
    if not have_ref(cfg):
        pytest.skip("oracle/_ref/pluto_gpu_* not built (integration/build_shim.sh)")
    every = 25
    r = run_reference(cfg, maxsteps=g.nsteps - 1, dump_every=every, analysis_every=every,
                      env={"PLUTO_GPU_ARITH": arith, "PLUTO_GPU_RESIDENT": "1"})
    assert "resident in HBM" in r.stdout
    syncs = r.stdout.count("PlutoGpuSyncHost: state copied")
    assert 1 <= syncs <= g.nsteps // every + 1, r.stdout[-2000:]      # not once per step
    cpu = run_reference(RefConfig(**{**cfg.__dict__, "prefix": "pluto_"}), maxsteps=g.nsteps - 1, dump_every=every,
                        analysis_every=every)
    assert sorted(r.dumps) == sorted(cpu.dumps) and max(r.dumps) == g.nsteps
    for s in sorted(cpu.dumps):
        for k, v in cpu.dumps[s].items():
            if arith == "exact":
                assert np.array_equal(r.dumps[s][k], v), f"{name}: {k} after {s} steps"
            else:
                assert rel_l1(r.dumps[s][k], v) <= (TOL_ONE_STEP if s <= 1 else TOL_100_STEPS), f"{name}: {k} after {s} steps"
    # the final dump is also the golden fixture's last state
    for k, v in g.states[g.nsteps].items():
        if arith == "exact":
            assert np.array_equal(r.dumps[g.nsteps][k], v), f"{name}: {k} (golden)"


@pytest.mark.parametrize("name,ndev,resident", [("turb3d_plm_hlld", 8, False), ("blast3d_plm_hlld_100", 4, True), ("rotor2d_ppm_roe", 2, False),
                                                ("ot2d_ctu", 4, True), ("ot2d_cl", 2, False),
                                                # non-uniform grid / grid-dependent weights: every block takes its slice
                                                ("blast3d_nug", 4, False), ("blast2d_nuw_mc_arith", 2, True),
                                                # BODY_FORCE: the force / potential arrays of the whole domain are cut into the blocks' pieces
                                                ("blast3d_bfx", 4, False), ("blast3d_bp", 2, True), ("blast3d_nug_bp", 4, False),
                                                ("blast2d_ctu_bfx_roe", 2, False),
                                                # PARABOLIC on a non-uniform grid: slices of the interface weights
                                                ("rotor2d_nug_ppm", 4, False)])
def test_reference_driver_on_several_blocks(name, ndev, resident):
    """PLUTO_GPU_NDEV: the reference's serial, single-threaded driver with the domain cut into 2 / 4 / 8 blocks (one per GPU where
    the box has them, round robin otherwise), driven through pluto_gpu_multi_* -- no MPI, no Python.  Dumps bit-identical to the
    all-CPU reference (= the golden fixtures), also with the state resident on the devices."""
    g = Golden(name)
    cfg = _cfg(g)
    if not have_ref(cfg):
        pytest.skip("oracle/_ref/pluto_gpu_* not built (integration/build_shim.sh)")
    env = {"PLUTO_GPU_ARITH": "exact", "PLUTO_GPU_NDEV": str(ndev)}
    if resident:
        env["PLUTO_GPU_RESIDENT"] = "1"
    every = 1 if g.nsteps <= 30 else 50
    r = run_reference(cfg, maxsteps=g.nsteps - 1 if every > 1 else g.nsteps + 1, dump_every=every, analysis_every=every, env=env)
    assert f"{ndev} block(s)" in r.stdout
    for s_, ref in g.states.items():
        if s_ in r.dumps:
            for k, v in ref.items():
                assert np.array_equal(r.dumps[s_][k], v), f"{name}: {k} after {s_} steps on {ndev} blocks"
    assert g.nsteps in r.dumps
