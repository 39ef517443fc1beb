#!/usr/bin/env bash
# Build the UNMODIFIED reference (PLUTO 4.3, serial, gcc -O3, no FMA) from the
# sources where they lie under $PLUTO_DIR (default /root/reference) for one
# compile-time scheme variant, with oracle/ref_build/problem/init.c as the
# user problem file.  Output: oracle/_ref/pluto_<variant>  (git-ignored).
#
# Mirrors Src/Templates/makefile (base OBJ lists :22-37) + the module
# makefiles Src/{Math_Tools,MHD,MHD/CT,EOS/Ideal}/makefile + the objects
# Tools/Python/define_problem.py:506-550 would add.  setup.py itself is
# interactive (curses) and is bypassed.  Nothing is copied from the
# reference: the makefile VPATHs into it.
#
# usage: build_ref.sh <variant> [<variant> ...]
#   variants:  2d_plm 3d_plm 2d_ppm 3d_ppm 2d_plm_rk3 3d_plm_rk3 2d_plm_hancock 3d_plm_hancock (CTU) 2d_plm_chtr (CTU, characteristic tracing)
#   optional suffixes:  _l{fl,mm,va,os,um,vl,mc}  single LIMITER for all variables
#                                                 (Src/States/plm_coeffs.h:72-123)
#                       _e{arith,uct0,uct_hll}    CT_EMF_AVERAGE
#                       _en                       CT_EN_CORRECTION YES
#                       _bf                       BODY_FORCE VECTOR (uniform acceleration)
#                       _sfl                      SHOCK_FLATTENING MULTID (Src/flag_shock.c) (Src/MHD/CT/ct_emf.c:241-283)
#                       _cl                       CHAR_LIMITING YES (Src/States/plm_states.c:448-706, Src/MHD/eigenv.c:190)
#                       _nuw                      UNIFORM_CARTESIAN_GRID NO (Src/States/plm_coeffs.c: weights of a non-uniform grid)
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ORACLE="$(cd "$HERE/.." && pwd)"
PLUTO_DIR="${PLUTO_DIR:-/root/reference}"
if [ ! -d "$PLUTO_DIR/Src" ]; then
  echo "build_ref.sh: reference sources not found at $PLUTO_DIR (skipping)" >&2
  exit 0
fi
mkdir -p "$ORACLE/_ref" "$ORACLE/_build"

for VARIANT in "$@"; do
  case "$VARIANT" in
    2d_*) DIMS=2 ;;
    3d_*) DIMS=3 ;;
    *) echo "unknown variant $VARIANT" >&2; exit 1 ;;
  esac
  case "$VARIANT" in
    *_ppm*) RECON=PARABOLIC; STATES_OBJ="ppm_states.o ppm_coeffs.o"; EXTRA_HDR="ppm_coeffs.h" ;;
    *)      RECON=LINEAR;    STATES_OBJ="plm_states.o";               EXTRA_HDR="" ;;
  esac
  case "$VARIANT" in
    *_rk3*)     TSTEP=RK3;     STEP_OBJ="rk_step.o update_stage.o" ;;
    *_hancock*) TSTEP=HANCOCK; STEP_OBJ="ctu_step.o hancock.o" ;;     # define_problem.py:542-553
    *_chtr*)    TSTEP=CHARACTERISTIC_TRACING; STEP_OBJ="ctu_step.o char_tracing.o" ;;   # define_problem.py:555-557
    *)          TSTEP=RK2;     STEP_OBJ="rk_step.o update_stage.o" ;;
  esac
  case "$VARIANT" in
    *_lfl*) LIMITER=FLAT_LIM ;;  *_lmm*) LIMITER=MINMOD_LIM ;;  *_lva*) LIMITER=VANALBADA_LIM ;;
    *_los*) LIMITER=OSPRE_LIM ;; *_lum*) LIMITER=UMIST_LIM ;;   *_lvl*) LIMITER=VANLEER_LIM ;;
    *_lmc*) LIMITER=MC_LIM ;;    *)      LIMITER=DEFAULT ;;
  esac
  case "$VARIANT" in
    *_earith*) EMFAVG=ARITHMETIC ;; *_euct0*) EMFAVG=UCT0 ;; *_euct_hll*) EMFAVG=UCT_HLL ;; *) EMFAVG=UCT_CONTACT ;;
  esac
  case "$VARIANT" in
    *_sfl*) SHOCKFLAT=MULTID ;; *) SHOCKFLAT=NO ;;
  esac
  case "$VARIANT" in
    *_en*) ENCORR=YES ;; *) ENCORR=NO ;;         # CT_EN_CORRECTION (ct_field_average.c:116-129)
  esac
  case "$VARIANT" in
    *_cl*) CHARLIM=YES ;; *) CHARLIM=NO ;;       # limiting on characteristic variables (plm_states.c:448-706)
  esac
  case "$VARIANT" in
    *_nuw*) UCG=NO ;; *) UCG=YES ;;              # UNIFORM_CARTESIAN_GRID NO: grid-dependent reconstruction weights (plm_coeffs.c)
  esac
  case "$VARIANT" in
    *_bfp*) BODYF="(VECTOR+POTENTIAL)" ;;        # both (uniform acceleration and step potential from the same GRAV1..3)
    *_bf*) BODYF=VECTOR ;;                       # BODY_FORCE VECTOR: acceleration GRAV1..3 (init.c BodyForceVector)
    *_bp*) BODYF=POTENTIAL ;;                    # BODY_FORCE POTENTIAL: step potential of init.c BodyForcePotential
    *) BODYF=NO ;;
  esac
  B="$ORACLE/_build/$VARIANT"
  mkdir -p "$B"
  cp "$HERE/problem/init.c" "$B/init.c"

  cat > "$B/definitions.h" <<EOF
#define  PHYSICS                        MHD
#define  DIMENSIONS                     $DIMS
#define  COMPONENTS                     $DIMS
#define  GEOMETRY                       CARTESIAN
#define  BODY_FORCE                     $BODYF
#define  FORCED_TURB                    NO
#define  COOLING                        NO
#define  RECONSTRUCTION                 $RECON
#define  TIME_STEPPING                  $TSTEP
#define  DIMENSIONAL_SPLITTING          NO
#define  NTRACER                        0
#define  USER_DEF_PARAMETERS            13

/* -- physics dependent declarations -- */

#define  EOS                            IDEAL
#define  ENTROPY_SWITCH                 NO
#define  DIVB_CONTROL                   CONSTRAINED_TRANSPORT
#define  BACKGROUND_FIELD               NO
#define  AMBIPOLAR_DIFFUSION            NO
#define  RESISTIVITY                    NO
#define  HALL_MHD                       NO
#define  THERMAL_CONDUCTION             NO
#define  VISCOSITY                      NO
#define  ROTATING_FRAME                 NO

/* -- user-defined parameters (labels) -- */

#define  PROBLEM                        0
#define  GAMMA_EOS                      1
#define  P_IN                           2
#define  P_OUT                          3
#define  BMAG                           4
#define  THETA                          5
#define  PHI                            6
#define  RADIUS                         7
#define  SEED                           8
#define  GRAV1                          9
#define  GRAV2                          10
#define  GRAV3                          11
#define  GRAV_MODE                      12

/* [Beg] user-defined constants (do not change this line) */

#define  LIMITER                        $LIMITER
#define  CHAR_LIMITING                  $CHARLIM
#define  UNIFORM_CARTESIAN_GRID         $UCG
#define  SHOCK_FLATTENING               $SHOCKFLAT
#define  CT_EMF_AVERAGE                 $EMFAVG
#define  CT_EN_CORRECTION               $ENCORR
#define  ASSIGN_VECTOR_POTENTIAL        YES
#define  CHECK_DIVB_CONDITION           NO
#define  WARNING_MESSAGES               NO

/* [End] user-defined constants (do not change this line) */
EOF

  cat > "$B/makefile" <<EOF
pluto:
PLUTO_DIR = $PLUTO_DIR
SRC = \$(PLUTO_DIR)/Src
INCLUDE_DIRS = -I. -I\$(SRC)
VPATH = ./:\$(SRC):\$(SRC)/Time_Stepping:\$(SRC)/States
CC = gcc
CFLAGS = -c -O3
LDFLAGS = -lm
HEADERS = pluto.h prototypes.h structs.h definitions.h macros.h mod_defs.h plm_coeffs.h $EXTRA_HDR
OBJ = adv_flux.o arrays.o boundary.o check_states.o cmd_line_opt.o entropy_switch.o \\
      flag_shock.o flatten.o get_nghost.o init.o int_bound_reset.o input_data.o \\
      mappers3D.o mean_mol_weight.o parse_file.o plm_coeffs.o rbox.o set_indexes.o \\
      set_geometry.o set_output.o tools.o var_names.o
OBJ += bin_io.o colortable.o initialize.o jet_domain.o main.o output_log.o restart.o \\
       runtime_setup.o set_image.o show_config.o set_grid.o startup.o split_source.o \\
       userdef_output.o write_data.o write_tab.o write_img.o write_vtk.o write_vtk_proc.o
include \$(SRC)/Math_Tools/makefile
OBJ += $STATES_OBJ vec_pot_diff.o vec_pot_update.o $STEP_OBJ
include \$(SRC)/MHD/makefile
include \$(SRC)/MHD/CT/makefile
include \$(SRC)/EOS/Ideal/makefile
pluto: \$(OBJ)
	\$(CC) \$(OBJ) \$(LDFLAGS) -o \$@
.c.o:
	\$(CC) \$(CFLAGS) \$(INCLUDE_DIRS) \$<
\$(OBJ): definitions.h
EOF
  ( cd "$B" && make -j"$(nproc)" pluto >make.log 2>&1 ) || { tail -30 "$B/make.log"; exit 1; }
  cp "$B/pluto" "$ORACLE/_ref/pluto_$VARIANT"
  echo "built oracle/_ref/pluto_$VARIANT"
done
