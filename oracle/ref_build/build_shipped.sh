#!/usr/bin/env bash
# Build a configuration SHIPPED with the reference (Test_Problems/MHD/<Problem>/definitions_NN.h + init.c, UNMODIFIED,
# used where they lie -- nothing is copied into the repository) twice: all-CPU, and with integration/advance_step_gpu.c +
# libpluto_gpu.so in place of the time-stepping objects.  Outputs (git-ignored, they travel to the GPU box):
#   oracle/_ref/shipped/<problem>_<NN>        the reference
#   oracle/_ref/shipped/<problem>_<NN>_gpu    the reference's driver + the GPU step
#   oracle/_ref/shipped/<problem>_<NN>.ini    pluto_NN.ini with the output cadence changed to one .dbl per step
# usage: build_shipped.sh Orszag_Tang 03 [Rotor 01 ...]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ORACLE="$(cd "$HERE/.." && pwd)"
ROOT="$(cd "$ORACLE/.." && pwd)"
PLUTO_DIR="${PLUTO_DIR:-/root/reference}"
[ -d "$PLUTO_DIR/Src" ] || { echo "build_shipped.sh: no reference sources (skipping)" >&2; exit 0; }
[ -f "$ORACLE/_build/2d_plm/makefile" ] || "$HERE/build_ref.sh" 2d_plm
mkdir -p "$ORACLE/_ref/shipped"
while [ $# -ge 2 ]; do
  PROB="$1"; NN="$2"; shift 2
  TP="$PLUTO_DIR/Test_Problems/MHD/$PROB"
  TAG="$(echo "$PROB" | tr 'A-Z' 'a-z')_$NN"
  # base makefile by the configuration's time stepping: RK2 / RK3 (rk_step.o update_stage.o) or the corner-transport-upwind step
  # with the characteristic-tracing predictor (ctu_step.o char_tracing.o; plm_states.o still calls CharTracingStep in the GPU build)
  BASE=2d_plm
  if grep -q "RECONSTRUCTION *PARABOLIC" "$TP/definitions_$NN.h"; then BASE=2d_ppm; fi            # ppm_states.o ppm_coeffs.o (RK2 / RK3)
  if grep -q "TIME_STEPPING *CHARACTERISTIC_TRACING" "$TP/definitions_$NN.h"; then BASE=2d_plm_chtr; fi
  if grep -q "TIME_STEPPING *HANCOCK" "$TP/definitions_$NN.h"; then BASE=2d_plm_hancock; fi      # ctu_step.o hancock.o
  [ -f "$ORACLE/_build/$BASE/makefile" ] || "$HERE/build_ref.sh" "$BASE"
  for KIND in cpu gpu; do
    B="$ORACLE/_build/shipped_${TAG}_$KIND"
    mkdir -p "$B"
    cp "$TP/definitions_$NN.h" "$B/definitions.h"          # build directory only (git-ignored)
    cp "$TP/init.c" "$B/init.c"
    [ "$KIND" = gpu ] && cp "$ROOT/integration/advance_step_gpu.c" "$B/"
    # the object lists of the LINEAR + RK2/RK3 build (oracle/_build/2d_plm/makefile = Src/Templates/makefile + module lists)
    if [ "$KIND" = gpu ]; then
      sed -e 's/rk_step.o update_stage.o/advance_step_gpu.o/' -e 's/ctu_step.o char_tracing.o/advance_step_gpu.o char_tracing.o/' \
          -e 's/ctu_step.o hancock.o/advance_step_gpu.o hancock.o/' \
          -e "s#^INCLUDE_DIRS = .*#INCLUDE_DIRS = -I. -I\$(SRC) -I$ROOT/include#" \
          -e "s#^LDFLAGS = .*#LDFLAGS = -lm -L$ROOT/pluto_b200/lib -lpluto_gpu -Wl,-rpath,'\$\$ORIGIN/../../../pluto_b200/lib'#" \
          "$ORACLE/_build/$BASE/makefile" > "$B/makefile"
    else
      cp "$ORACLE/_build/$BASE/makefile" "$B/makefile"
    fi
    ( cd "$B" && make -j"$(nproc)" pluto >make.log 2>&1 ) || { tail -30 "$B/make.log"; exit 1; }
    if [ "$KIND" = gpu ]; then cp "$B/pluto" "$ORACLE/_ref/shipped/${TAG}_gpu"; else cp "$B/pluto" "$ORACLE/_ref/shipped/$TAG"; fi
  done
  # one single-file .dbl per step, nothing else (the scheme, grid, boundaries, parameters stay as shipped)
  sed -E -e 's/^dbl .*/dbl       -1.0  1   single_file/' -e 's/^(flt|vtk|tab|ppm|png|dbl\.h5|flt\.h5) .*/\1       -1.0  -1/' \
      -e 's/^log .*/log        1000/' -e 's/^analysis .*/analysis  -1.0  -1/' "$TP/pluto_$NN.ini" > "$ORACLE/_ref/shipped/$TAG.ini"
  echo "built oracle/_ref/shipped/$TAG (+ _gpu, .ini)"
done
