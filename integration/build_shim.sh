#!/usr/bin/env bash
# Build the UNMODIFIED reference driver (main loop, ini parser, Init, output)
# with integration/advance_step_gpu.c in place of rk_step.o + update_stage.o,
# linked against pluto_b200/lib/libpluto_gpu.so.  Output:
# oracle/_ref/pluto_gpu_<variant> (git-ignored; travels to the GPU box).
# Needs the reference sources (this container only) and a built libpluto_gpu.so.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
ROOT="$(cd "$HERE/.." && pwd)"
PLUTO_DIR="${PLUTO_DIR:-/root/reference}"
[ -d "$PLUTO_DIR/Src" ] || { echo "build_shim.sh: no reference sources (skipping)" >&2; exit 0; }
[ -f "$ROOT/pluto_b200/lib/libpluto_gpu.so" ] || { echo "build_shim.sh: build libpluto_gpu.so first" >&2; exit 1; }
VARIANTS="${*:-2d_plm 3d_plm 2d_ppm 3d_ppm 3d_plm_lvl_earith 2d_plm_lmc_earith 3d_plm_lmc_euct_hll 3d_plm_sfl 2d_plm_hancock 3d_plm_hancock 3d_plm_lvl_earith_en 2d_plm_en 3d_plm_hancock_en 3d_plm_bf 2d_ppm_rk3_bf 3d_plm_hancock_bf 2d_plm_hancock_bf 3d_plm_bp 2d_plm_hancock_bp 2d_plm_cl 2d_plm_rk3_lvl_cl 2d_plm_rk3 3d_plm_nuw 2d_plm_lmc_earith_nuw 2d_plm_chtr 2d_plm_chtr_lmc 2d_plm_chtr_lmc_euct0 2d_plm_hancock_lmc_earith_cl 2d_plm_chtr_cl 2d_ppm_sfl 3d_ppm_sfl 2d_ppm_rk3 3d_plm_euct_hll_bf 2d_ppm_rk3_euct_hll_bp 3d_plm_sfl_bf 3d_plm_hancock_sfl_bf 2d_ppm_sfl_bp 2d_ppm_rk3_sfl 2d_ppm_euct_hll_sfl 3d_ppm_euct_hll_sfl 2d_plm_euct_hll_cl 2d_plm_cl_bf 2d_plm_hancock_cl_bf}"
for VARIANT in $VARIANTS; do
  B="$ROOT/oracle/_build/$VARIANT"
  [ -f "$B/definitions.h" ] || "$ROOT/oracle/ref_build/build_ref.sh" "$VARIANT"
  G="$ROOT/oracle/_build/gpu_$VARIANT"
  mkdir -p "$G"
  cp "$B/definitions.h" "$B/init.c" "$G/"
  cp "$HERE/advance_step_gpu.c" "$G/"
  # the two one-line hooks of INTEGRATION.md ("Keeping the state on the device"): patched copies of the reference's
  # write_data.c and main.c in the (git-ignored) build directory, found first through the makefile's VPATH = ./:...
  sed -e 's|^  output->nfile++;|  { void PlutoGpuSyncHost (Data *); PlutoGpuSyncHost ((Data *)d); }   /* libpluto_gpu: resident state */\n  output->nfile++;|' \
      "$PLUTO_DIR/Src/write_data.c" > "$G/write_data.c"
  sed -e 's|^  if (check_dt \|\| check_dn) Analysis (d, grid);|  if (check_dt \|\| check_dn){ void PlutoGpuSyncHost (Data *); PlutoGpuSyncHost (d); Analysis (d, grid); }|' \
      "$PLUTO_DIR/Src/main.c" > "$G/main.c"
  grep -q PlutoGpuSyncHost "$G/write_data.c" && grep -q PlutoGpuSyncHost "$G/main.c" || { echo "build_shim.sh: hook patch did not apply" >&2; exit 1; }
  # same makefile as the CPU reference build, with the two time-stepping objects
  # replaced by the shim and the GPU library added to the link line
  sed -e 's/rk_step.o update_stage.o/advance_step_gpu.o/' -e 's/ctu_step.o hancock.o/advance_step_gpu.o hancock.o/' -e 's/ctu_step.o char_tracing.o/advance_step_gpu.o char_tracing.o/' \
      -e "s#^INCLUDE_DIRS = .*#INCLUDE_DIRS = -I. -I\$(SRC) -I$ROOT/include#" \
      -e "s#^LDFLAGS = .*#LDFLAGS = -lm -L$ROOT/pluto_b200/lib -lpluto_gpu -Wl,-rpath,'\$\$ORIGIN/../../pluto_b200/lib'#" \
      "$B/makefile" > "$G/makefile"
  ( cd "$G" && make -j"$(nproc)" pluto >make.log 2>&1 ) || { tail -30 "$G/make.log"; exit 1; }
  cp "$G/pluto" "$ROOT/oracle/_ref/pluto_gpu_$VARIANT"
  echo "built oracle/_ref/pluto_gpu_$VARIANT"
done
