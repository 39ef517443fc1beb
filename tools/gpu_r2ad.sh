#!/bin/bash
# fused x1+x2 sweep: __launch_bounds__(128, 3) (168 registers; PARABOLIC and Roe variants spill) against (128, 2) (up to 255 registers,
# no spills) per reconstruction / solver pair
mkdir -p gpurun_out
{
for w in blast3d_256 blast3d_256_ppm blast3d_256_roe blast3d_256_ppm_roe blast3d_256_hllc rotor2d_4096 rotor2d_4096_plm_roe rotor2d_4096_ppm_hlld ot2d_512; do
  echo "## $w: shipped library / fused sweeps with __launch_bounds__(128, 2)"
  BENCH_ARGS="--workload $w" STEPS=${STEPS:-12} tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so pluto_b200/lib/variants/libpluto_gpu_xy2.so
done
} > gpurun_out/r2ad_ab.log 2>&1
cat gpurun_out/r2ad_ab.log
