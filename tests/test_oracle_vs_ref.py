"""The CPU restatement against the compiled reference run live
(oracle/_ref/pluto_*), at sizes/configs beyond the committed fixtures.
Skipped when the binaries are absent (they are git-ignored build products;
`python -c 'import __graft_entry__ as g; g.build()'` creates them when
/root/reference is present)."""
import numpy as np
import pytest

from oracle.oracle_lib import Oracle, next_dt
from oracle.refrun import RefConfig, have_ref, run_reference

CASES = [
    ("ot2d_hlld", RefConfig(problem="ot", dims=2, n=(48, 40, 1), first_dt=1.5e-2), 12),
    ("ot3d_hll", RefConfig(problem="ot", dims=3, n=(12, 16, 10), first_dt=3e-2, cfl=0.3, solver="hll"), 6),
    ("blast3d_roe", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, solver="roe"), 6),
    ("rotor2d_ppm_hlld", RefConfig(problem="rotor", dims=2, n=(40, 36, 1), recon="ppm", first_dt=2e-3), 10),
    ("turb2d_hlld", RefConfig(problem="turb", dims=2, n=(24, 20, 1), first_dt=2e-2), 8),
    # single limiters (plm_coeffs.h:72-123) and the other EMF averages (ct_emf.c:241-283): the options the
    # shipped Test_Problems/MHD configurations use (OT #03, Blast #02 #05, Rotor #01)
    ("blast3d_vl_arith", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, limiter="vl", emf="arith"), 6),
    ("ot2d_arith_roe", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, solver="roe", emf="arith"), 8),
    ("rotor2d_mc_arith", RefConfig(problem="rotor", dims=2, n=(36, 32, 1), first_dt=2e-3, limiter="mc", emf="arith"), 8),
    ("blast3d_va_arith", RefConfig(problem="blast", dims=3, n=(10, 12, 14), first_dt=3e-4, cfl=0.3, limiter="va", emf="arith"), 6),
    ("turb3d_uct0", RefConfig(problem="turb", dims=3, n=(10, 12, 8), first_dt=2e-2, cfl=0.3, emf="uct0"), 6),
    ("ot2d_mm", RefConfig(problem="ot", dims=2, n=(28, 24, 1), first_dt=1.5e-2, limiter="mm"), 8),
    ("turb3d_um", RefConfig(problem="turb", dims=3, n=(8, 10, 12), first_dt=2e-2, cfl=0.3, limiter="um"), 5),
    ("blast2d_os", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, limiter="os"), 8),
    # UCT_HLL, the reference's default average (ct.h:43-45): Blast #10 = MC_LIM + roe
    ("blast3d_mc_uct_hll_roe", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, limiter="mc", emf="uct_hll", solver="roe"), 6),
    ("ot2d_uct_hll", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, emf="uct_hll"), 8),
    ("turb3d_uct_hll", RefConfig(problem="turb", dims=3, n=(10, 12, 8), first_dt=2e-2, cfl=0.3, emf="uct_hll"), 6),
    ("rotor2d_ppm_uct_hll_hll", RefConfig(problem="rotor", dims=2, n=(36, 32, 1), recon="ppm", first_dt=2e-3, emf="uct_hll", solver="hll"), 8),
    # SHOCK_FLATTENING MULTID (flag_shock.c): minmod + HLL in shocked zones
    ("blast3d_sfl", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, flatten=True), 10),
    ("blast2d_sfl_roe", RefConfig(problem="blast", dims=2, n=(36, 32, 1), first_dt=3e-4, solver="roe", flatten=True), 12),
    ("blast3d_sfl_uct_hll", RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=3e-4, cfl=0.3, emf="uct_hll", flatten=True), 8),
    # corner-transport upwind with the MUSCL-Hancock predictor (ctu_step.c, hancock.c): SURVEY 8f row 1
    ("ot2d_ctu", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, tstep="hancock", cfl=0.4), 8),
    ("blast3d_ctu", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, tstep="hancock"), 6),
    ("turb3d_ctu_roe", RefConfig(problem="turb", dims=3, n=(10, 12, 8), first_dt=2e-2, cfl=0.3, tstep="hancock", solver="roe"), 6),
    ("blast2d_ctu_hll_arith_vl", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="hancock", solver="hll",
                                           emf="arith", limiter="vl"), 10),
    ("blast3d_ctu_sfl_uct0", RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=3e-4, cfl=0.3, tstep="hancock",
                                       emf="uct0", flatten=True), 8),
    ("blast2d_ctu_reflective", RefConfig(problem="blast", dims=2, n=(24, 20, 1), first_dt=3e-4, tstep="hancock",
                                         bc=("reflective", "outflow", "reflective", "reflective", "outflow", "outflow"),
                                         blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.3)), 12),
    # CT_EN_CORRECTION YES (ct_field_average.c:116-129): RK2 (the scheme of the shipped Blast #02), RK3 + UCT0, CTU
    ("blast3d_blast02_en", RefConfig(problem="blast", dims=3, n=(14, 10, 12), first_dt=3e-4, cfl=0.2, limiter="vl", emf="arith",
                                     solver="roe", en_corr=True), 8),
    ("blast2d_en", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, en_corr=True), 10),
    ("turb3d_rk3_uct0_en", RefConfig(problem="turb", dims=3, n=(10, 8, 12), first_dt=2e-2, cfl=0.3, tstep="rk3", emf="uct0",
                                     en_corr=True), 5),
    ("blast3d_ctu_en", RefConfig(problem="blast", dims=3, n=(10, 14, 12), first_dt=3e-4, cfl=0.3, tstep="hancock", en_corr=True), 8),
    ("ot2d_ctu_arith_en_roe", RefConfig(problem="ot", dims=2, n=(28, 32, 1), first_dt=1.5e-2, tstep="hancock", emf="arith",
                                        solver="roe", en_corr=True), 8),
    # BODY_FORCE VECTOR, uniform acceleration (rhs_source.c:214-217, 277-280, 342-345; prim_eqn.c:289-360 in the Hancock predictor)
    ("blast3d_bf", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, grav=(0.3, -1.0, 0.5)), 8),
    ("ot2d_bf_roe", RefConfig(problem="ot", dims=2, n=(28, 24, 1), first_dt=1.5e-2, solver="roe", grav=(0.0, -0.8, 0.0)), 8),
    ("rotor2d_ppm_rk3_bf", RefConfig(problem="rotor", dims=2, n=(28, 24, 1), recon="ppm", tstep="rk3", first_dt=2e-3, grav=(0.5, 0.25, 0.0)), 6),
    ("turb3d_ctu_bf", RefConfig(problem="turb", dims=3, n=(10, 12, 8), first_dt=2e-2, cfl=0.3, tstep="hancock", grav=(0.3, -1.0, 0.5)), 6),
    ("blast2d_ctu_bf_hll", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="hancock", solver="hll",
                                     grav=(-2.0, 1.0, 0.0)), 10),
    # static position-dependent force (GRAV_MODE 1 of the problem file: component d = grav[d]*sign(x_d))
    ("blast3d_bfx", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, grav=(-3.0, -1.0, 2.0), grav_mode=1), 8),
    ("rotor2d_ppm_bfx", RefConfig(problem="rotor", dims=2, n=(28, 24, 1), recon="ppm", first_dt=2e-3, grav=(-1.5, -2.5, 0.0), grav_mode=1), 6),
    ("blast3d_ctu_bfx", RefConfig(problem="blast", dims=3, n=(10, 12, 8), first_dt=3e-4, cfl=0.3, tstep="hancock",
                                  grav=(-3.0, -1.0, 2.0), grav_mode=1), 6),
    # BODY_FORCE POTENTIAL (rhs.c:388-392, rhs_source.c:233-237, prim_eqn.c:304-307): step potential of the problem file
    ("blast3d_bp", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, grav=(0.05, -0.03, 0.04), potential=True), 8),
    ("rotor2d_ppm_rk3_bp", RefConfig(problem="rotor", dims=2, n=(28, 24, 1), recon="ppm", tstep="rk3", first_dt=2e-3,
                                     grav=(0.02, -0.03, 0.0), potential=True), 6),
    ("blast3d_ctu_bp", RefConfig(problem="blast", dims=3, n=(10, 12, 8), first_dt=3e-4, cfl=0.3, tstep="hancock",
                                 grav=(0.05, -0.03, 0.04), potential=True), 6),
    # EQTSYMMETRIC boundaries (boundary.c:333-336, 423-427), the condition of the shipped Blast #02
    ("blast3d_eqtsym", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3,
                                 bc=("reflective", "outflow", "eqtsymmetric", "outflow", "eqtsymmetric", "reflective"),
                                 blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=30.0, RADIUS=0.3)), 10),
    ("blast2d_ctu_eqtsym", RefConfig(problem="blast", dims=2, n=(24, 20, 1), first_dt=3e-4, tstep="hancock",
                                     bc=("eqtsymmetric", "eqtsymmetric", "outflow", "eqtsymmetric", "outflow", "outflow"),
                                     blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.3)), 12),
    # CHAR_LIMITING YES (plm_states.c:448-706, PrimEigenvectors eigenv.c:190-560): slopes limited on the characteristic variables.
    # Pinned in 2-D.  In 3-D the reference's right-eigenvector scratch (stateC->Rp, only non-zero entries are ever written) keeps
    # the Alfven entries of the previous sweep direction in the row of the normal velocity, and they enter the slopes: the
    # reference's own 3-D result depends on the order in which the pencils were swept, and nothing can be pinned against it
    ("ot2d_cl", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, char_lim=True), 10),
    ("blast2d_cl_roe", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, solver="roe", char_lim=True), 10),
    ("rotor2d_cl_vl", RefConfig(problem="rotor", dims=2, n=(36, 32, 1), first_dt=2e-3, limiter="vl", char_lim=True), 8),
    ("turb2d_cl_rk3_mc", RefConfig(problem="turb", dims=2, n=(24, 20, 1), first_dt=2e-2, tstep="rk3", limiter="mc", char_lim=True), 6),
    ("blast3d_bfp", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, grav=(0.05, -0.03, 0.04), potential=True,
                              vector_too=True), 6),
    ("blast2d_ctu_bfp", RefConfig(problem="blast", dims=2, n=(24, 20, 1), first_dt=3e-4, tstep="hancock", grav=(0.05, -0.03, 0.0),
                                  potential=True, vector_too=True), 8),
    # non-uniform Cartesian grids: uniform + stretched patches of pluto.ini's [Grid] block (set_grid.c:330-560).
    # With the reference's default UNIFORM_CARTESIAN_GRID YES the reconstruction keeps its uniform weights; the zone widths
    # enter rhs.c:195, the inverse time step, CT_Update and the face areas of FillMagneticField
    ("blast3d_nug", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3,
                              grid=("2  -0.5  8  u  0.1  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                    "3  -0.5  4  s  -0.2  8  u  0.2  4  s  0.5")), 10),
    ("rotor2d_nug_roe", RefConfig(problem="rotor", dims=2, n=(36, 30, 1), first_dt=2e-3, solver="roe",
                                  grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  10  s  0.5", None)), 10),
    ("blast2d_nug_mc_arith_reflective", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, limiter="mc", emf="arith",
                                                  bc=("reflective", "outflow", "outflow", "reflective", "outflow", "outflow"),
                                                  blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.3),
                                                  grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 12),
    # body forces on a non-uniform grid (stratified set-ups on stretched grids): uniform acceleration, position-dependent
    # force, potential (its momentum source takes dt/dx[i] of the zone)
    ("blast3d_nug_bf", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, grav=(0.3, -1.0, 0.5),
                                 grid=("2  -0.5  8  u  0.1  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                       "3  -0.5  4  s  -0.2  8  u  0.2  4  s  0.5")), 8),
    ("blast2d_nug_bp", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, grav=(0.05, -0.03, 0.0), potential=True,
                                 grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 10),
    ("blast3d_nug_bfx", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, grav=(-3.0, -1.0, 2.0), grav_mode=1,
                                  grid=("2  -0.5  8  u  0.1  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                        "3  -0.5  4  s  -0.2  8  u  0.2  4  s  0.5")), 8),
    # the corner-transport-upwind steps on non-uniform grids: d_dl[i] of the Hancock predictor (hancock.c:83), dt/dx[i] of the
    # characteristic tracing (char_tracing.c:346-347), dt2_dx[i] of CTU_CT_Source and of both right-hand sides (ctu_step.c:310)
    ("blast3d_nug_ctu", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, tstep="hancock",
                                  grid=("2  -0.5  8  u  0.1  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                        "3  -0.5  4  s  -0.2  8  u  0.2  4  s  0.5")), 8),
    ("rotor2d_nug_ctu_roe", RefConfig(problem="rotor", dims=2, n=(36, 30, 1), first_dt=2e-3, solver="roe", tstep="hancock",
                                      grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  10  s  0.5", None)), 10),
    ("blast2d_nug_chtr_mc", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="chtr", limiter="mc",
                                      grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 12),
    # UNIFORM_CARTESIAN_GRID NO: the reconstruction takes the grid-dependent weights of PLM_CoefficientsGet (plm_coeffs.c:30-104)
    # and the limiters "on irregular grids" (plm_coeffs.h:130-152)
    ("blast3d_nuw", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, grid_weights=True,
                              grid=("2  -0.5  8  u  0.1  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                    "3  -0.5  4  s  -0.2  8  u  0.2  4  s  0.5")), 10),
    ("rotor2d_nuw_roe", RefConfig(problem="rotor", dims=2, n=(36, 30, 1), first_dt=2e-3, solver="roe", grid_weights=True,
                                  grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  10  s  0.5", None)), 10),
    ("blast2d_nuw_mc_arith", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, limiter="mc", emf="arith", grid_weights=True,
                                       grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 12),
    ("ot2d_nuw_uniform", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, grid_weights=True), 8),
    # TIME_STEPPING CHARACTERISTIC_TRACING (States/char_tracing.c:278-560): the corner-transport-upwind step with the
    # characteristic-tracing predictor, the scheme of the shipped Field_Loop #01 / #02 (LINEAR, MC_LIM, UCT_CONTACT / UCT0).
    # 2 components (3: the reference's eigenvector scratch keeps entries of the previous sweep direction, as with CHAR_LIMITING)
    ("ot2d_chtr", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, tstep="chtr", cfl=0.4), 10),
    ("blast2d_chtr_mc_roe", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="chtr", limiter="mc", solver="roe"), 12),
    # CHAR_LIMITING YES with the corner-transport-upwind steps: ot2d_ctu_cl_mc_arith = the scheme of the shipped Orszag_Tang #09
    ("ot2d_ctu_cl_mc_arith", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, tstep="hancock", char_lim=True, limiter="mc",
                                       emf="arith"), 10),
    ("blast2d_chtr_cl", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="chtr", char_lim=True), 10),
    # SHOCK_FLATTENING MULTID with PARABOLIC reconstruction: flagged zones fall back to minmod-limited linear states with the
    # weights of PLM_CoefficientsGet (ppm_states.c:167-181), handed over as the reference built them
    ("blast2d_ppm_sfl_roe", RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", first_dt=3e-4, solver="roe", flatten=True), 12),
    ("blast3d_ppm_sfl", RefConfig(problem="blast", dims=3, n=(14, 12, 16), recon="ppm", first_dt=3e-4, cfl=0.3, flatten=True), 10),
    ("rotor2d_chtr_mc_uct0_hll", RefConfig(problem="rotor", dims=2, n=(36, 32, 1), first_dt=2e-3, tstep="chtr", limiter="mc", emf="uct0",
                                           solver="hll"), 10),
    # the HLLC (hllc.c) and Lax-Friedrichs (tvdlf.c) solvers: pluto.ini's [Solver] block picks them at run time
    ("ot2d_hllc", RefConfig(problem="ot", dims=2, n=(48, 40, 1), first_dt=1.5e-2, solver="hllc"), 12),
    ("blast3d_hllc", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, solver="hllc"), 8),
    ("rotor2d_ppm_rk3_hllc", RefConfig(problem="rotor", dims=2, n=(40, 36, 1), recon="ppm", tstep="rk3", first_dt=2e-3, solver="hllc"), 8),
    ("blast3d_sfl_uct_hll_hllc", RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=3e-4, cfl=0.3, emf="uct_hll", flatten=True,
                                           solver="hllc"), 8),
    ("turb3d_ctu_hllc", RefConfig(problem="turb", dims=3, n=(10, 12, 8), first_dt=2e-2, cfl=0.3, tstep="hancock", solver="hllc"), 6),
    ("ot2d_tvdlf", RefConfig(problem="ot", dims=2, n=(48, 40, 1), first_dt=1.5e-2, solver="tvdlf"), 12),
    ("blast3d_tvdlf_uct_hll", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, solver="tvdlf", emf="uct_hll"), 8),
    ("ot3d_ctu_um_uct0_tvdlf", RefConfig(problem="ot", dims=3, n=(12, 16, 10), first_dt=3e-2, cfl=0.3, tstep="hancock", limiter="um",
                                         emf="uct0", solver="tvdlf"), 6),
    ("blast2d_sfl_tvdlf", RefConfig(problem="blast", dims=2, n=(36, 32, 1), first_dt=3e-4, solver="tvdlf", flatten=True), 10),
    # BODY_FORCE with CT_EMF_AVERAGE UCT_HLL (the reference's DEFAULT average, CT/ct.h: what a configuration with gravity and no
    # explicit CT_EMF_AVERAGE runs, e.g. the shipped Rayleigh_Taylor #07)
    ("blast3d_bf_uct_hll", RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=3e-4, cfl=0.3, grav=(0.3, -1.0, 0.5), emf="uct_hll"), 8),
    ("rotor2d_ppm_rk3_bp_uct_hll_roe", RefConfig(problem="rotor", dims=2, n=(28, 24, 1), recon="ppm", tstep="rk3", first_dt=2e-3,
                                                 grav=(0.05, -0.03, 0.0), potential=True, emf="uct_hll", solver="roe"), 6),
    # BODY_FORCE with SHOCK_FLATTENING MULTID
    ("blast3d_sfl_bf", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, flatten=True, grav=(0.3, -1.0, 0.5)), 10),
    ("blast3d_sfl_uct_hll_bf_roe", RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=3e-4, cfl=0.3, flatten=True, emf="uct_hll",
                                             solver="roe", grav=(0.3, -1.0, 0.5)), 8),
    ("blast3d_ctu_sfl_bf", RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=3e-4, cfl=0.3, tstep="hancock", flatten=True,
                                     grav=(0.3, -1.0, 0.5)), 8),
    ("blast2d_ppm_sfl_bp", RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", first_dt=3e-4, flatten=True, grav=(0.05, -0.03, 0.0),
                                     potential=True), 10),
    # non-uniform grids with SHOCK_FLATTENING MULTID (flag_shock.c:143-145 divides by the zone's widths) and CHAR_LIMITING
    ("blast3d_nug_sfl", RefConfig(problem="blast", dims=3, n=(14, 12, 16), first_dt=3e-4, cfl=0.3, flatten=True,
                                  grid=("2  -0.5  8  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5", "3  -0.5  4  s  -0.3  8  u  0.3  4  s  0.5")), 10),
    ("blast2d_nug_sfl_roe", RefConfig(problem="blast", dims=2, n=(36, 32, 1), first_dt=3e-4, solver="roe", flatten=True,
                                      grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  24  u  0.1  8  s  0.5", None)), 12),
    ("blast3d_nug_ctu_sfl_uct0", RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=3e-4, cfl=0.3, tstep="hancock", flatten=True, emf="uct0",
                                  grid=("2  -0.5  6  u  0.0  6  s  0.5", "2  -0.5  6  s  -0.2  8  u  0.5", "2  -0.5  4  s  -0.3  6  u  0.5")), 8),
    ("blast2d_nug_cl_roe", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, solver="roe", char_lim=True,
                                     grid=("2  -0.5  20  u  0.2  8  s  0.5", "3  -0.5  6  s  -0.3  12  u  0.3  6  s  0.5", None)), 10),
    ("ot2d_nug_ctu_cl_mc_arith", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, tstep="hancock", char_lim=True, limiter="mc", emf="arith",
                                           grid=("2  0.0  20  u  4.0  12  s  6.283185307179586", "2  0.0  8  s  1.5  20  u  6.283185307179586", None)), 8),
    ("blast2d_nug_en", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, en_corr=True,
                                 grid=("2  -0.5  20  u  0.2  8  s  0.5", "3  -0.5  6  s  -0.3  12  u  0.3  6  s  0.5", None)), 10),
    ("blast3d_nug_ctu_en", RefConfig(problem="blast", dims=3, n=(10, 14, 12), first_dt=3e-4, cfl=0.3, tstep="hancock", en_corr=True,
                                     grid=("2  -0.5  4  u  -0.1  6  s  0.5", "2  -0.5  6  s  -0.2  8  u  0.5", "2  -0.5  4  s  -0.3  8  u  0.5")), 8),
    # corner transport upwind + BODY_FORCE on non-uniform grids (prim_eqn.c:304-307, 335-338 divide the potential difference by dx[i])
    ("turb3d_nug_ctu_bf", RefConfig(problem="turb", dims=3, n=(10, 12, 8), first_dt=2e-2, cfl=0.3, tstep="hancock", grav=(0.3, -1.0, 0.5),
                                    grid=("2  0.0  6  u  0.6  4  s  1.0", "2  0.0  4  s  0.3  8  u  1.0", "2  0.0  4  u  0.5  4  s  1.0")), 6),
    ("blast2d_nug_ctu_bp", RefConfig(problem="blast", dims=2, n=(24, 20, 1), first_dt=3e-4, tstep="hancock", grav=(0.05, -0.03, 0.0), potential=True,
                                     grid=("2  -0.5  16  u  0.2  8  s  0.5", "3  -0.5  4  s  -0.3  12  u  0.3  4  s  0.5", None)), 8),
    # PARABOLIC on non-uniform grids: the interface weights of PPM_FindWeights (ppm_coeffs.c:300-480), handed over as the reference built them
    ("rotor2d_nug_ppm", RefConfig(problem="rotor", dims=2, n=(40, 36, 1), recon="ppm", first_dt=2e-3,
                                  grid=("3  -0.5  10  s  -0.2  20  u  0.2  10  s  0.5", "2  -0.5  24  u  0.1  12  s  0.5", None)), 10),
    ("blast3d_nug_ppm_roe", RefConfig(problem="blast", dims=3, n=(14, 12, 16), recon="ppm", first_dt=3e-4, cfl=0.3, solver="roe",
                                      grid=("2  -0.5  8  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5", "3  -0.5  4  s  -0.3  8  u  0.3  4  s  0.5")), 8),
    ("rotor2d_nug_ppm_rk3_bp", RefConfig(problem="rotor", dims=2, n=(28, 24, 1), recon="ppm", tstep="rk3", first_dt=2e-3,
                                         grav=(0.05, -0.03, 0.0), potential=True,
                                         grid=("2  -0.5  20  u  0.2  8  s  0.5", "3  -0.5  6  s  -0.3  12  u  0.3  6  s  0.5", None)), 6),
    ("blast2d_nug_ppm_sfl_roe", RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", first_dt=3e-4, solver="roe", flatten=True,
                                          grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  24  u  0.1  8  s  0.5", None)), 10),
    # PARABOLIC + SHOCK_FLATTENING MULTID with the default average UCT_HLL
    ("blast2d_ppm_sfl_uct_hll_roe", RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", first_dt=3e-4, solver="roe", flatten=True,
                                              emf="uct_hll"), 12),
    ("blast3d_ppm_sfl_uct_hll", RefConfig(problem="blast", dims=3, n=(14, 12, 16), recon="ppm", first_dt=3e-4, cfl=0.3, flatten=True,
                                          emf="uct_hll"), 10),
    # CHAR_LIMITING with the default average UCT_HLL and with BODY_FORCE (RK and Hancock)
    ("ot2d_cl_uct_hll", RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, char_lim=True, emf="uct_hll"), 10),
    ("blast2d_cl_bf_roe", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, solver="roe", char_lim=True, grav=(0.5, 0.25, 0.0)), 10),
    ("blast2d_ctu_cl_bf", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="hancock", char_lim=True, grav=(0.5, 0.25, 0.0)), 10),
    ("blast2d_chtr_mc_hllc", RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=3e-4, tstep="chtr", limiter="mc", solver="hllc"), 10),
]


@pytest.mark.parametrize("label,cfg,nsteps", CASES, ids=[c[0] for c in CASES])
def test_oracle_bit_exact_vs_live_reference(label, cfg, nsteps):
    if not have_ref(cfg):
        pytest.skip("oracle/_ref binary not built")
    r = run_reference(cfg, maxsteps=nsteps + 1, dump_every=1)
    n = list(cfg.n)
    if cfg.dims == 2:
        n[2] = 1
    dom = cfg.resolved_domain()
    dx = [(dom[d][1] - dom[d][0]) / n[d] for d in range(cfg.dims)]
    o = Oracle(cfg.dims, n, dx, recon=cfg.recon, solver=cfg.solver, bc=cfg.resolved_bc(),
               gamma=cfg.resolved_gamma(), limiter=cfg.limiter, emf=cfg.emf, flatten=cfg.flatten, ctu=("chtr" if cfg.tstep == "chtr" else cfg.tstep == "hancock"),
               rk_order=(3 if cfg.tstep == "rk3" else 2), en_corr=cfg.en_corr, grav=(None if cfg.potential and not cfg.vector_too else cfg.grav),
               char_lim=cfg.char_lim)
    if cfg.potential:
        from tests.util import step_potential_arrays
        o.set_body_potential(*step_potential_arrays(cfg.dims, n, o.ng, dom, cfg.grav, widths=(r.dx if cfg.grid is not None else None)))
    if cfg.grav_mode == 1:
        from tests.util import sign_force_arrays
        o.set_body_force(*sign_force_arrays(cfg.dims, n, o.ng, dom, cfg.grav, widths=(r.dx if cfg.grid is not None else None)))
    if cfg.grid is not None:
        assert r.dx is not None and len(r.dx) == cfg.dims and max(np.ptp(a) for a in r.dx) > 0.0
        o.set_grid(*r.dx)
    if cfg.grid is not None and cfg.recon == "ppm":
        assert r.ppm_coeffs is not None and len(r.ppm_coeffs) == cfg.dims
        o.set_ppm_coeffs(r.ppm_coeffs)
    if cfg.grid_weights or (cfg.flatten and cfg.recon == "ppm"):
        assert r.plm_coeffs is not None and len(r.plm_coeffs) == cfg.dims
        o.set_plm_coeffs(r.plm_coeffs)
    o.set_state(r.dumps[0])
    tap = {int(a): c for a, b, c in r.dt_tap}
    dt = cfg.first_dt
    for s in range(1, nsteps + 1):
        inv, mach, _ = o.advance(dt)
        dt = next_dt(inv, cfg.cfl, cfg.cfl_max_var, dt)
        assert dt == tap[s], f"dt after step {s}"
        st = o.get_state()
        for k, ref in r.dumps[s].items():
            assert np.array_equal(st[k], ref), f"{label}: {k} differs after {s} steps"
