#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the compiled, unmodified reference
(oracle/_ref/pluto_*, built from /root/reference by
oracle/ref_build/build_ref.sh: gcc -O3, serial, no FMA).

Each fixture holds, for one small configuration:
  cfg_*            the run parameters (problem, grid, solver, CFL, ...)
  s0_<name>        the initial interior state as dumped in data.0000.dbl
  s<K>_<name>      the state after K steps (K = 1 and K = nsteps)
  dt               dt[s] used for step s (dt[0] = first_dt), full precision
                   from the Analysis() tap in oracle/ref_build/problem/init.c

Run here (container with /root/reference); the fixtures are committed so
that the GPU box, which has no /root/reference, can check against them.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.refrun import RefConfig, run_reference  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")

CASES = {
    # name: (RefConfig, nsteps)
    "ot2d_plm_hlld": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4), 20),
    "ot2d_plm_hll": (RefConfig(problem="ot", dims=2, n=(24, 32, 1), first_dt=2.5e-2, cfl=0.4, solver="hll"), 10),
    "ot2d_plm_roe": (RefConfig(problem="ot", dims=2, n=(24, 32, 1), first_dt=2.5e-2, cfl=0.4, solver="roe"), 10),
    "ot3d_plm_hlld": (RefConfig(problem="ot", dims=3, n=(16, 12, 8), first_dt=3e-2, cfl=0.3), 10),
    "blast3d_plm_hlld": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3), 12),
    "blast2d_plm_hlld": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4), 20),
    "rotor2d_ppm_roe": (RefConfig(problem="rotor", dims=2, n=(32, 24, 1), recon="ppm", solver="roe",
                                  first_dt=2.5e-3, cfl=0.4), 20),
    "turb3d_plm_hlld": (RefConfig(problem="turb", dims=3, n=(8, 12, 16), first_dt=3e-2, cfl=0.3), 10),
    "ot3d_ppm_roe": (RefConfig(problem="ot", dims=3, n=(12, 8, 16), recon="ppm", solver="roe",
                               first_dt=4.5e-2, cfl=0.3), 8),
    # reflective walls (FlipSign + FillMagneticField on reflective sides) and mixed conditions
    "blast2d_reflective": (RefConfig(problem="blast", dims=2, n=(24, 20, 1), first_dt=4e-4, cfl=0.4,
                                     bc=("reflective", "outflow", "reflective", "reflective", "outflow", "outflow"),
                                     blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.3)), 30),
    "blast3d_reflective": (RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=6e-4, cfl=0.3,
                                     bc=("reflective", "reflective", "outflow", "reflective", "reflective", "outflow"),
                                     blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=30.0, RADIUS=0.35)), 25),
    # 100-step runs for the BASELINE.json "1e-9 after 100 steps" bar
    "ot2d_plm_hlld_100": (RefConfig(problem="ot", dims=2, n=(64, 48, 1), first_dt=1e-2, cfl=0.4), 100),
    "blast3d_plm_hlld_100": (RefConfig(problem="blast", dims=3, n=(24, 20, 16), first_dt=5e-4, cfl=0.3), 100),
    "turb3d_plm_hlld_100": (RefConfig(problem="turb", dims=3, n=(16, 20, 24), first_dt=2e-2, cfl=0.3), 100),
    "rotor2d_ppm_roe_100": (RefConfig(problem="rotor", dims=2, n=(48, 40, 1), recon="ppm", solver="roe",
                                      first_dt=1e-3, cfl=0.4), 100),
    # the scheme options of the shipped Test_Problems/MHD configurations (single limiters, other EMF averages):
    # Orszag_Tang #03 (ARITHMETIC, roe), Blast #02 (VANLEER_LIM, ARITHMETIC), Blast #05 (VANALBADA_LIM,
    # ARITHMETIC), Rotor #01 (MC_LIM, ARITHMETIC), plus UCT0 and the remaining limiters
    "ot2d_arith_roe": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, solver="roe", emf="arith"), 20),
    "blast3d_vl_arith": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3, limiter="vl", emf="arith"), 12),
    "blast3d_va_arith": (RefConfig(problem="blast", dims=3, n=(12, 16, 8), first_dt=6e-4, cfl=0.3, limiter="va", emf="arith"), 12),
    "rotor2d_mc_arith": (RefConfig(problem="rotor", dims=2, n=(32, 24, 1), first_dt=2.5e-3, cfl=0.4, limiter="mc", emf="arith"), 20),
    "turb3d_uct0": (RefConfig(problem="turb", dims=3, n=(8, 12, 16), first_dt=3e-2, cfl=0.3, emf="uct0"), 10),
    "ot2d_mm": (RefConfig(problem="ot", dims=2, n=(24, 32, 1), first_dt=2.5e-2, cfl=0.4, limiter="mm"), 10),
    "turb3d_um": (RefConfig(problem="turb", dims=3, n=(12, 8, 16), first_dt=3e-2, cfl=0.3, limiter="um"), 10),
    "blast2d_os": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, limiter="os"), 20),
    # UCT_HLL, the reference's default CT_EMF_AVERAGE (ct.h:43-45); Blast #10 = MC_LIM + roe + UCT_HLL
    "blast3d_mc_uct_hll_roe": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3, limiter="mc",
                                         emf="uct_hll", solver="roe"), 12),
    "ot2d_uct_hll": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, emf="uct_hll"), 20),
    "turb3d_uct_hll": (RefConfig(problem="turb", dims=3, n=(8, 12, 16), first_dt=3e-2, cfl=0.3, emf="uct_hll"), 10),
    "rotor2d_ppm_uct_hll_hll": (RefConfig(problem="rotor", dims=2, n=(32, 24, 1), recon="ppm", solver="hll",
                                          first_dt=2.5e-3, cfl=0.4, emf="uct_hll"), 20),
    # SHOCK_FLATTENING MULTID (flag_shock.c): minmod + HLL in shocked zones (Blast #07-#09's option)
    "blast3d_sfl": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, flatten=True), 15),
    "blast2d_sfl_roe": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, solver="roe", flatten=True), 25),
    "blast3d_sfl_uct_hll": (RefConfig(problem="blast", dims=3, n=(12, 16, 12), first_dt=6e-4, cfl=0.3, emf="uct_hll",
                                      flatten=True), 15),
    # corner-transport upwind, MUSCL-Hancock predictor (TIME_STEPPING HANCOCK: ctu_step.c, hancock.c)
    "ot2d_ctu": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, tstep="hancock"), 20),
    "blast3d_ctu": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3, tstep="hancock"), 12),
    "turb3d_ctu_roe": (RefConfig(problem="turb", dims=3, n=(8, 12, 16), first_dt=3e-2, cfl=0.3, tstep="hancock", solver="roe"), 10),
    "blast3d_ctu_sfl_uct0": (RefConfig(problem="blast", dims=3, n=(12, 16, 12), first_dt=6e-4, cfl=0.3, tstep="hancock",
                                       emf="uct0", flatten=True), 15),
    # CT_EN_CORRECTION YES (ct_field_average.c:116-129); blast3d_blast02 = the scheme of the shipped Blast #02
    # (LINEAR, RK2, VANLEER_LIM, ARITHMETIC, CT_EN_CORRECTION YES, roe, CFL 0.2)
    "blast3d_blast02_en": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.2, limiter="vl", emf="arith",
                                     solver="roe", en_corr=True), 12),
    "blast2d_en": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, en_corr=True), 20),
    "turb3d_rk3_uct0_en": (RefConfig(problem="turb", dims=3, n=(8, 12, 16), first_dt=3e-2, cfl=0.3, tstep="rk3", emf="uct0",
                                     en_corr=True), 10),
    "blast3d_ctu_en": (RefConfig(problem="blast", dims=3, n=(12, 16, 8), first_dt=6e-4, cfl=0.3, tstep="hancock", en_corr=True), 12),
    "ot2d_ctu_arith_en_roe": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, tstep="hancock", emf="arith",
                                        solver="roe", en_corr=True), 20),
    # BODY_FORCE VECTOR with a uniform acceleration (first slice)
    "blast3d_bf": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3, grav=(0.3, -1.0, 0.5)), 12),
    "rotor2d_ppm_rk3_bf": (RefConfig(problem="rotor", dims=2, n=(32, 24, 1), recon="ppm", tstep="rk3", first_dt=2.5e-3, cfl=0.4,
                                     grav=(0.5, 0.25, 0.0)), 15),
    "turb3d_ctu_bf": (RefConfig(problem="turb", dims=3, n=(8, 12, 16), first_dt=3e-2, cfl=0.3, tstep="hancock", grav=(0.3, -1.0, 0.5)), 10),
    # static position-dependent force (GRAV_MODE 1: component d = grav[d]*sign(x_d)) through pluto_gpu_set_body_force
    "blast3d_bfx": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3, grav=(-3.0, -1.0, 2.0), grav_mode=1), 12),
    "blast2d_ctu_bfx_roe": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, tstep="hancock", solver="roe",
                                      grav=(-3.0, -1.0, 0.0), grav_mode=1), 20),
    # BODY_FORCE POTENTIAL: step potential of the problem file (pluto_gpu_set_body_potential)
    "blast3d_bp": (RefConfig(problem="blast", dims=3, n=(16, 12, 8), first_dt=6e-4, cfl=0.3, grav=(0.05, -0.03, 0.04), potential=True), 12),
    "blast2d_ctu_bp": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, tstep="hancock",
                                 grav=(0.05, -0.03, 0.0), potential=True), 20),
    # EQTSYMMETRIC boundaries (the condition of the shipped Blast #02), mixed with reflective and outflow sides
    "blast3d_eqtsym": (RefConfig(problem="blast", dims=3, n=(12, 10, 14), first_dt=6e-4, cfl=0.3,
                                 bc=("reflective", "outflow", "eqtsymmetric", "outflow", "eqtsymmetric", "reflective"),
                                 blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=30.0, RADIUS=0.3)), 25),
    "blast2d_ctu_eqtsym": (RefConfig(problem="blast", dims=2, n=(24, 20, 1), first_dt=4e-4, cfl=0.4, tstep="hancock",
                                     bc=("eqtsymmetric", "eqtsymmetric", "outflow", "eqtsymmetric", "outflow", "outflow"),
                                     blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.3)), 30),
    # CHAR_LIMITING YES (plm_states.c:448-706): slopes limited on the characteristic variables, 2-D (see tests/test_oracle_vs_ref.py)
    "ot2d_cl": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, char_lim=True), 20),
    "blast2d_cl_roe": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, solver="roe", char_lim=True), 20),
    "rotor2d_cl_vl_rk3": (RefConfig(problem="rotor", dims=2, n=(32, 24, 1), first_dt=2.5e-3, cfl=0.4, limiter="vl", tstep="rk3",
                                    char_lim=True), 15),
    "ot2d_ctu_100": (RefConfig(problem="ot", dims=2, n=(64, 48, 1), first_dt=1e-2, cfl=0.4, tstep="hancock"), 100),
    # non-uniform Cartesian grids: uniform + stretched patches (set_grid.c:330-560); the fixtures carry the zone
    # widths grid->dx[d] the reference built (grid_tap.bin of the problem file's Analysis)
    "blast3d_nug": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3,
                              grid=("2  -0.5  10  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                    "3  -0.5  3  s  -0.3  8  u  0.3  3  s  0.5")), 15),
    "rotor2d_nug_roe_rk3": (RefConfig(problem="rotor", dims=2, n=(36, 28, 1), first_dt=2.5e-3, cfl=0.4, solver="roe", tstep="rk3",
                                      grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  8  s  0.5", None)), 15),
    "blast2d_nug_mc_arith_reflective": (RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=4e-4, cfl=0.4, limiter="mc", emf="arith",
                                                  bc=("reflective", "outflow", "outflow", "reflective", "outflow", "outflow"),
                                                  blast=dict(P_IN=100.0, P_OUT=1.0, BMAG=10.0, THETA=45.0, PHI=0.0, RADIUS=0.3),
                                                  grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 25),
    # body forces on non-uniform grids: potential (momentum source with dt/dx[i]), position-dependent force
    "blast3d_nug_bp": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, grav=(0.05, -0.03, 0.04), potential=True,
                                 grid=("2  -0.5  10  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                                         "3  -0.5  3  s  -0.3  8  u  0.3  3  s  0.5")), 12),
    "blast2d_nug_bfx_roe": (RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=4e-4, cfl=0.4, solver="roe", grav=(-3.0, -1.0, 0.0),
                                      grav_mode=1, grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 20),
    # the HLLC and Lax-Friedrichs solvers (hllc.c, tvdlf.c)
    "blast3d_hllc": (RefConfig(problem="blast", dims=3, n=(20, 16, 12), first_dt=6e-4, cfl=0.3, solver="hllc"), 12),
    "rotor2d_ppm_rk3_hllc": (RefConfig(problem="rotor", dims=2, n=(40, 36, 1), recon="ppm", tstep="rk3", first_dt=2.5e-3, solver="hllc"), 12),
    "turb3d_ctu_hllc": (RefConfig(problem="turb", dims=3, n=(12, 14, 10), first_dt=2e-2, cfl=0.3, tstep="hancock", solver="hllc"), 10),
    "ot2d_tvdlf": (RefConfig(problem="ot", dims=2, n=(48, 40, 1), first_dt=1.5e-2, solver="tvdlf"), 15),
    "ot3d_ctu_um_uct0_tvdlf": (RefConfig(problem="ot", dims=3, n=(12, 16, 10), first_dt=3e-2, cfl=0.3, tstep="hancock", limiter="um",
                                         emf="uct0", solver="tvdlf"), 10),
    "blast3d_sfl_uct_hll_hllc": (RefConfig(problem="blast", dims=3, n=(14, 16, 12), first_dt=6e-4, cfl=0.3, emf="uct_hll", flatten=True,
                                           solver="hllc"), 10),
    # BODY_FORCE with the reference's default EMF average, UCT_HLL
    "blast3d_bf_uct_hll": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, grav=(0.3, -1.0, 0.5), emf="uct_hll"), 12),
    "rotor2d_ppm_rk3_bp_uct_hll_roe": (RefConfig(problem="rotor", dims=2, n=(36, 28, 1), recon="ppm", tstep="rk3", first_dt=2.5e-3,
                                                 grav=(0.05, -0.03, 0.0), potential=True, emf="uct_hll", solver="roe"), 10),
    # PARABOLIC + SHOCK_FLATTENING MULTID with the default average UCT_HLL
    "blast3d_ppm_sfl_uct_hll": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), recon="ppm", first_dt=6e-4, cfl=0.3, flatten=True,
                                          emf="uct_hll"), 12),
    "blast2d_ppm_sfl_uct_hll_roe": (RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", first_dt=6e-4, solver="roe", flatten=True,
                                              emf="uct_hll"), 12),
    # CHAR_LIMITING with the default average UCT_HLL and with BODY_FORCE (RK, Hancock)
    "ot2d_cl_uct_hll": (RefConfig(problem="ot", dims=2, n=(32, 28, 1), first_dt=1.5e-2, char_lim=True, emf="uct_hll"), 12),
    "blast2d_cl_bf_roe": (RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=6e-4, solver="roe", char_lim=True, grav=(0.5, 0.25, 0.0)), 12),
    "blast2d_ctu_cl_bf": (RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=6e-4, tstep="hancock", char_lim=True, grav=(0.5, 0.25, 0.0)), 12),
    # BODY_FORCE with SHOCK_FLATTENING MULTID
    "blast3d_sfl_bf": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, flatten=True, grav=(0.3, -1.0, 0.5)), 12),
    "blast3d_ctu_sfl_bf": (RefConfig(problem="blast", dims=3, n=(12, 14, 10), first_dt=6e-4, cfl=0.3, tstep="hancock", flatten=True,
                                     grav=(0.3, -1.0, 0.5)), 10),
    "blast2d_ppm_sfl_bp": (RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", first_dt=6e-4, flatten=True, grav=(0.05, -0.03, 0.0),
                                     potential=True), 12),
    # the corner-transport-upwind steps on non-uniform grids (Hancock 3-D; characteristic tracing 2-D, MC_LIM)
    "blast3d_nug_ctu": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, tstep="hancock",
                                  grid=("2  -0.5  10  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                        "3  -0.5  3  s  -0.3  8  u  0.3  3  s  0.5")), 12),
    "rotor2d_nug_chtr_mc": (RefConfig(problem="rotor", dims=2, n=(36, 28, 1), first_dt=2.5e-3, cfl=0.4, tstep="chtr", limiter="mc",
                                      grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  8  s  0.5", None)), 15),
    # ... with SHOCK_FLATTENING MULTID, CHAR_LIMITING, CT_EN_CORRECTION and BODY_FORCE inside the corner-transport-upwind step
    "blast3d_nug_sfl": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, flatten=True,
                                  grid=("2  -0.5  10  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                        "3  -0.5  3  s  -0.3  8  u  0.3  3  s  0.5")), 12),
    "blast2d_nug_cl_roe": (RefConfig(problem="blast", dims=2, n=(36, 28, 1), first_dt=6e-4, solver="roe", char_lim=True,
                                     grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  8  s  0.5", None)), 12),
    "blast2d_nug_en": (RefConfig(problem="blast", dims=2, n=(36, 28, 1), first_dt=6e-4, en_corr=True,
                                 grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  8  s  0.5", None)), 12),
    "blast2d_nug_ctu_bp": (RefConfig(problem="blast", dims=2, n=(36, 28, 1), first_dt=6e-4, tstep="hancock", grav=(0.05, -0.03, 0.0), potential=True,
                                     grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  20  u  0.1  8  s  0.5", None)), 12),
    # PARABOLIC on non-uniform grids (the fixtures carry the interface weights of PPM_CoefficientsGet)
    "rotor2d_nug_ppm": (RefConfig(problem="rotor", dims=2, n=(40, 36, 1), recon="ppm", first_dt=2.5e-3,
                                  grid=("3  -0.5  10  s  -0.2  20  u  0.2  10  s  0.5", "2  -0.5  24  u  0.1  12  s  0.5", None)), 12),
    "blast3d_nug_ppm_roe": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), recon="ppm", first_dt=6e-4, cfl=0.3, solver="roe",
                                      grid=("2  -0.5  10  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                            "3  -0.5  3  s  -0.3  8  u  0.3  3  s  0.5")), 10),
    "blast2d_nug_ppm_sfl_rk3": (RefConfig(problem="blast", dims=2, n=(36, 32, 1), recon="ppm", tstep="rk3", first_dt=6e-4, flatten=True,
                                          grid=("3  -0.5  8  s  -0.25  20  u  0.25  8  s  0.5", "2  -0.5  24  u  0.1  8  s  0.5", None)), 10),
    # UNIFORM_CARTESIAN_GRID NO: grid-dependent reconstruction weights (plm_coeffs.c) -- the fixtures carry the arrays of PLM_CoefficientsGet
    "blast3d_nuw": (RefConfig(problem="blast", dims=3, n=(16, 12, 14), first_dt=6e-4, cfl=0.3, grid_weights=True,
                              grid=("2  -0.5  10  u  0.0  6  s  0.5", "2  -0.5  4  s  -0.2  8  u  0.5",
                                    "3  -0.5  3  s  -0.3  8  u  0.3  3  s  0.5")), 15),
    # TIME_STEPPING CHARACTERISTIC_TRACING (char_tracing.c): CTU with the characteristic-tracing predictor, 2 components;
    # rotor2d_chtr_mc_uct0 = the scheme of the shipped Field_Loop #02 (LINEAR, MC_LIM, UCT0), ot2d_chtr_mc_roe #01's with roe
    "ot2d_chtr": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, tstep="chtr"), 20),
    "rotor2d_chtr_mc_uct0": (RefConfig(problem="rotor", dims=2, n=(32, 24, 1), first_dt=2.5e-3, cfl=0.4, tstep="chtr", limiter="mc",
                                       emf="uct0"), 20),
    "blast2d_chtr_mc_roe": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, tstep="chtr", limiter="mc",
                                      solver="roe"), 20),
    # CHAR_LIMITING YES with the corner-transport-upwind steps; ot2d_ctu_cl_mc_arith = the scheme of the shipped Orszag_Tang #09
    "ot2d_ctu_cl_mc_arith": (RefConfig(problem="ot", dims=2, n=(32, 24, 1), first_dt=2e-2, cfl=0.4, tstep="hancock", char_lim=True,
                                       limiter="mc", emf="arith"), 20),
    "blast2d_chtr_cl": (RefConfig(problem="blast", dims=2, n=(32, 24, 1), first_dt=4e-4, cfl=0.4, tstep="chtr", char_lim=True), 20),
    # SHOCK_FLATTENING MULTID with PARABOLIC reconstruction (ppm_states.c:167-181; fixtures carry the weights of PLM_CoefficientsGet)
    "blast2d_ppm_sfl_roe": (RefConfig(problem="blast", dims=2, n=(36, 24, 1), recon="ppm", first_dt=4e-4, cfl=0.4, solver="roe",
                                      flatten=True), 25),
    "blast3d_ppm_sfl": (RefConfig(problem="blast", dims=3, n=(18, 12, 14), recon="ppm", first_dt=6e-4, cfl=0.3, flatten=True), 15),
    "blast2d_nuw_mc_arith": (RefConfig(problem="blast", dims=2, n=(28, 24, 1), first_dt=4e-4, cfl=0.4, limiter="mc", emf="arith",
                                       grid_weights=True,
                                       grid=("2  -0.5  20  u  0.2  8  s  0.5", "2  -0.5  8  s  -0.1  16  u  0.5", None)), 25),
}


def make(name):
    cfg, nsteps = CASES[name]
    r = run_reference(cfg, maxsteps=nsteps + 1, dump_every=1)
    out = {
        "cfg_problem": cfg.problem, "cfg_dims": cfg.dims, "cfg_n": np.array(cfg.n),
        "cfg_recon": cfg.recon, "cfg_solver": cfg.solver, "cfg_tstep": cfg.tstep,
        "cfg_cfl": cfg.cfl, "cfg_cfl_max_var": cfg.cfl_max_var,
        "cfg_first_dt": cfg.first_dt, "cfg_gamma": cfg.resolved_gamma(),
        "cfg_domain": np.array(cfg.resolved_domain()),
        "cfg_bc": np.array(cfg.resolved_bc()), "cfg_nsteps": nsteps,
        "cfg_limiter": cfg.limiter, "cfg_emf": cfg.emf, "cfg_flatten": int(cfg.flatten),
        "cfg_en_corr": int(cfg.en_corr), "cfg_char_lim": int(cfg.char_lim),
    }
    if cfg.grav is not None:
        out["cfg_grav"] = np.array(cfg.grav, dtype=float)
        out["cfg_grav_mode"] = int(cfg.grav_mode)
        out["cfg_potential"] = int(cfg.potential)
    if cfg.grid is not None:
        out["cfg_grid"] = np.array([g or "" for g in cfg.grid])
        for d in range(cfg.dims):
            out[f"grid_dx{d+1}"] = r.dx[d]
    if cfg.grid_weights:
        out["cfg_grid_weights"] = 1
    if r.plm_coeffs is not None:           # UNIFORM_CARTESIAN_GRID NO, or PARABOLIC + MULTID (the minmod fallback's weights)
        for d in range(cfg.dims):
            out[f"plm_coeffs{d+1}"] = np.array(r.plm_coeffs[d])
    if r.ppm_coeffs is not None and cfg.grid is not None:       # PARABOLIC on a non-uniform grid: the interface weights of PPM_CoefficientsGet
        for d in range(cfg.dims):
            out[f"ppm_coeffs{d+1}"] = np.array(r.ppm_coeffs[d])
    for s in (0, 1, nsteps):
        for k, v in r.dumps[s].items():
            out[f"s{s}_{k}"] = v
    dt = np.zeros(nsteps + 1)
    dt[0] = cfg.first_dt
    tap = {int(a): c for a, b, c in r.dt_tap}
    for s in range(1, nsteps + 1):
        dt[s] = tap[s]
    out["dt"] = dt
    os.makedirs(GOLDEN, exist_ok=True)
    path = os.path.join(GOLDEN, name + ".npz")
    np.savez_compressed(path, **out)
    print(f"{name}: {os.path.getsize(path)/1024:.0f} KiB, dt[-1]/dt[-2] = {dt[-1]/dt[-2]:.4f}")


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        make(nm)
