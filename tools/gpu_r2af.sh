#!/bin/bash
# Roe, 3-D: marching (x3) sweep with three against two resident blocks per SM; the shipped library has xy_min_blocks applied
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "roe and fast" 2>&1 | tail -3) > gpurun_out/r2ag_pytest.log
{
for w in ot3d_256_roe ot3d_256_ppm_roe; do
  echo "## $w: shipped library / Roe marching sweeps with __launch_bounds__(128, 2)"
  BENCH_ARGS="--workload $w" STEPS=${STEPS:-12} tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so pluto_b200/lib/variants/libpluto_gpu_m2roe.so
done
} > gpurun_out/r2ag_ab.log 2>&1
cat gpurun_out/r2ag_pytest.log gpurun_out/r2ag_ab.log
