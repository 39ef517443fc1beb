#!/bin/bash
# PARABOLIC + hlld marching sweep with two blocks per SM (variant) against three (shipped); sanitizers on the families added last
mkdir -p gpurun_out
{
for w in blast3d_256_ppm ot3d_256_ppm_roe ot3d_256_roe; do
  echo "## $w: shipped library / hlld marching sweeps with __launch_bounds__(128, 2)"
  BENCH_ARGS="--workload $w" STEPS=${STEPS:-12} tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so pluto_b200/lib/variants/libpluto_gpu_m2hlld.so
done
} > gpurun_out/r2ah_ab.log 2>&1
for tool in memcheck racecheck initcheck; do
  (timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_cases.py schemes3 2>&1 | grep -v "^$" | cut -c1-300 | tail -12) > gpurun_out/r2ah_sanitizer_${tool}.log
done
cat gpurun_out/r2ah_ab.log; tail -4 gpurun_out/r2ah_sanitizer_*.log
