"""CPU-side check of the CUDA kernels' SOURCE (no GPU needed): pluto_b200/csrc is compiled
unchanged by g++ against the lock-step interpreter of tests/emu (test infrastructure, see
tests/emu/include/cuda_runtime.h) and driven through the same C ABI and the same Python host
as the GPU library.  What this pins before GPU time is spent: every index range, launch
geometry, shuffle/shared-memory exchange and the operation order of the kernels -- EXACT
arithmetic bit-identical to the reference goldens, FAST arithmetic (FMA contraction; IEEE
division and square root stand in for the MUFU-seeded iterations) within the BASELINE
tolerances.  The GPU parity tests (tests/test_gpu_parity.py) remain the parity gate.
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from tests.util import Golden, golden_names, divb_max, rel_l1, apply_force_field, TOL_ONE_STEP, TOL_100_STEPS

SHORT = [n for n in golden_names() if not n.endswith("_100")]


@pytest.fixture(scope="module")
def emu_lib():
    from tests.emu.build_emu import build
    os.environ["PLUTO_GPU_NO_GRAPH"] = "1"      # launches run at once: nothing to capture
    return build()


def _stepper(g, arith, lib):
    from pluto_b200 import GpuStepper
    kw = dict(ctu=g.ctu, en_corr=g.en_corr, grav=g.force, potential=g.potential, char_lim=g.char_lim)
    s = GpuStepper(g.dims, g.n, g.dx, recon=g.recon, solver=g.solver, rk_order=g.rk_order, bc=g.bc, gamma=g.gamma,
                   arith=arith, limiter=g.limiter, emf=g.emf, flatten=g.flatten, lib_path=lib, **kw)
    apply_force_field(s, g)
    return s


@pytest.mark.parametrize("name", SHORT)
def test_emulated_exact_kernels_bit_identical_to_golden(name, emu_lib):
    g = Golden(name)
    s = _stepper(g, "exact", emu_lib)
    s.set_state(g.states[0])
    dt = g.first_dt
    for step in range(1, g.nsteps + 1):
        assert dt == g.dt[step - 1]
        info = s.advance(dt)
        assert info.nan_events == 0
        dt = s.next_dt(info.inv_dt_hyp, g.cfl, g.cfl_max_var, dt)
        if step in g.states:
            st = s.get_state()
            for k, ref in g.states[step].items():
                assert np.array_equal(st[k], ref), f"{name}: {k} differs after {step} steps"
    assert dt == g.dt[g.nsteps]
    s.close()


FAST_SUBSET = ["blast3d_plm_hlld", "ot2d_plm_hlld", "rotor2d_ppm_roe", "ot3d_ppm_roe", "turb3d_uct_hll", "blast3d_sfl",
               "blast3d_ctu", "ot2d_ctu", "turb3d_ctu_roe", "blast3d_blast02_en", "blast3d_bf", "turb3d_ctu_bf", "blast2d_ctu_bfx_roe",
               "ot2d_cl", "blast2d_cl_roe", "rotor2d_cl_vl_rk3", "blast3d_nug", "blast2d_nug_mc_arith_reflective", "blast3d_ppm_sfl",
               "blast3d_hllc", "ot3d_ctu_um_uct0_tvdlf", "blast3d_nug_ctu"]


@pytest.mark.parametrize("name", FAST_SUBSET)
def test_emulated_fast_kernels_within_tolerance(name, emu_lib):
    g = Golden(name)
    s = _stepper(g, "fast", emu_lib)
    s.set_state(g.states[0])
    dt = g.first_dt
    tol1, tolN = TOL_ONE_STEP, TOL_100_STEPS
    for step in range(1, g.nsteps + 1):
        info = s.advance(dt)
        dt = s.next_dt(info.inv_dt_hyp, g.cfl, g.cfl_max_var, dt)
        if step in g.states:
            st = s.get_state()
            for k, ref in g.states[step].items():
                assert rel_l1(st[k], ref) <= (tol1 if step == 1 else tolN), f"{name}: {k} after {step} steps"
    st = s.get_state()
    bscale = max(np.abs(st["Bx1s"]).max(), 1e-30) / min(g.dx)
    assert divb_max(st, g.dims, g.dx_zones) < 1e-12 * bscale
    s.close()


# A slice of the GPU test files themselves (same test code, same Python host) run against the interpreted
# kernels: smallest and ragged blocks, multi-block decomposition with the halo pack/unpack kernels
# (RK and corner transport upwind), NextTimeStep on the device, reflective walls, the reference Data layout.
GPU_SUITE_SLICE = ("6x6x6 or 6x7x1 or 31x200x1 or (decomposed and gn7 and split) or (decomposed and gn4 and all) "
                   "or (decomposed and gn2 and dims) or (device_next_dt and ot-2) or reflective or data_layout "
                   "or turb3d_plm_hll_rk2 or ot2d_plm_hlld_rk2_1 or blast2d_ppm_roe_rk3 or (en_correction and ot-2) or (body_force and rotor) or tma_staging or (flux_difference_kept_apart and blast) or (nonuniform_grid and (rotor or turb or refused)) or characteristic_tracing_is_2d or (decomposed_blocks_on_a_nonuniform and ot-2)")


def test_gpu_test_files_through_the_interpreter(emu_lib):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PLUTO_GPU_LIB=emu_lib, PLUTO_GPU_NO_GRAPH="1")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-m", "gpu", "-p", "no:cacheprovider",
                        "tests/test_gpu_parity.py", "-k", GPU_SUITE_SLICE], cwd=root, env=env, capture_output=True, text=True)
    tail = r.stdout[-1500:] + r.stderr[-500:]
    assert r.returncode == 0, tail
    assert " passed" in r.stdout and "failed" not in r.stdout, tail


def test_reference_driver_with_interpreted_kernels(emu_lib, tmp_path):
    """The drop-in path of tests/test_gpu_dropin.py on the CPU: the reference's unmodified driver + the shim
    (integration/advance_step_gpu.c), with the interpreted kernel library found first on the library path."""
    from oracle.refrun import RefConfig, have_ref, run_reference
    libdir = tmp_path / "lib"
    libdir.mkdir()
    os.symlink(emu_lib, libdir / "libpluto_gpu.so")
    for name in ("ot2d_plm_hlld", "ot2d_ctu", "rotor2d_ppm_rk3_bf", "blast2d_ctu_bfx_roe", "blast2d_ctu_bp", "rotor2d_cl_vl_rk3",
                 "rotor2d_nug_roe_rk3", "blast2d_nuw_mc_arith", "rotor2d_chtr_mc_uct0", "blast2d_ppm_sfl_roe"):
        g = Golden(name)
        cfg = RefConfig(problem=g.problem, dims=g.dims, n=g.n, recon=g.recon, solver=g.solver, tstep=g.tstep, cfl=g.cfl,
                        cfl_max_var=g.cfl_max_var, first_dt=g.first_dt, gamma=g.gamma, limiter=g.limiter, emf=g.emf,
                        flatten=g.flatten, en_corr=g.en_corr, grav=g.grav, grav_mode=g.grav_mode, potential=g.potential, char_lim=g.char_lim,
                        grid=g.grid, grid_weights=g.grid_weights, prefix="pluto_gpu_")
        if not have_ref(cfg):
            pytest.skip("oracle/_ref/pluto_gpu_* not built (integration/build_shim.sh)")
        r = run_reference(cfg, maxsteps=g.nsteps + 1, dump_every=1,
                          env={"PLUTO_GPU_ARITH": "exact", "PLUTO_GPU_NO_GRAPH": "1", "LD_LIBRARY_PATH": str(libdir)})
        tap = {int(a): c for a, b, c in r.dt_tap}
        for s in range(1, g.nsteps + 1):
            assert tap[s] == g.dt[s], f"{name}: dt after step {s}"
        for s, ref in g.states.items():
            for k, v in ref.items():
                assert np.array_equal(r.dumps[s][k], v), f"{name}: {k} after {s} steps"


def test_create_refuses_unsupported_combinations(emu_lib):
    """pluto_gpu_create validates the scheme options (same code in the GPU library) and reports why."""
    from pluto_b200 import GpuStepper
    from pluto_b200.stepper import PlutoGpuError
    ok = dict(dims=3, n=(8, 8, 8), dx=(0.1, 0.1, 0.1), lib_path=emu_lib)
    for bad, msg in [(dict(recon="ppm", ctu=True), "LINEAR"), (dict(emf="uct_hll", ctu=True), "UCT_HLL"),
                     (dict(emf="uct_hll", en_corr=True), "CT_EN_CORRECTION"), (dict(ctu="chtr"), "2-D only"),
                     (dict(rk_order=4), "rk_order"), (dict(n=(8, 3, 8)), "nghost")]:
        kw = dict(ok); kw.update(bad)
        with pytest.raises(PlutoGpuError, match=msg):
            GpuStepper(kw.pop("dims"), kw.pop("n"), kw.pop("dx"), **kw)
    s = GpuStepper(3, (8, 8, 8), (0.1, 0.1, 0.1), lib_path=emu_lib, ctu=True, flatten=True)
    assert (s.ng, s.nstages) == (4, 1)
    s.close()
    # PARABOLIC + MULTID: the minmod fallback of flagged zones takes the weights of PLM_CoefficientsGet -- a step without them says so
    s = GpuStepper(3, (8, 8, 8), (0.1, 0.1, 0.1), lib_path=emu_lib, recon="ppm", flatten=True)
    from pluto_b200 import problems
    s.set_state(problems.make("blast", 3, (8, 8, 8))[0])
    with pytest.raises(PlutoGpuError, match="pluto_gpu_set_plm_coeffs"):
        s.advance(1e-4)
    s.close()


def test_results_do_not_depend_on_lane_or_block_order(emu_lib):
    """PG_EMU_REVERSE=1 runs the lanes of every warp (between two collectives) and the blocks of every launch in
    descending order.  Bit-identical goldens under both schedules: no kernel relies on an execution order that the
    CUDA model does not promise (a missing __syncwarp / cp.async wait shows up here as a wrong result)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, PG_EMU_REVERSE="1")
    sel = ("(exact or fast) and (blast3d_plm_hlld or ot3d_ppm_roe or rotor2d_ppm_uct_hll_hll "
           "or blast3d_ctu_sfl_uct0 or ot2d_ctu_arith_en_roe or blast3d_ctu\\])")
    r = subprocess.run([sys.executable, "-m", "pytest", "-x", "-q", "-p", "no:cacheprovider", "tests/test_emu_kernels.py", "-k", sel],
                       cwd=root, env=env, capture_output=True, text=True)
    assert r.returncode == 0 and " passed" in r.stdout, r.stdout[-1500:] + r.stderr[-500:]
