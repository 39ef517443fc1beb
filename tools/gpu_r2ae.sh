#!/bin/bash
# after xy_min_blocks: PARABOLIC parity on the B200 and the shipped library on the PARABOLIC workloads
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider -k "ppm" 2>&1 | tail -3) > gpurun_out/r2ae_pytest.log
{
for w in rotor2d_4096 rotor2d_4096_ppm_hlld blast3d_256_ppm blast3d_256; do
  echo "## $w"
  BENCH_ARGS="--workload $w" STEPS=${STEPS:-12} tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
done
} > gpurun_out/r2ae_bench.log 2>&1
cat gpurun_out/r2ae_pytest.log gpurun_out/r2ae_bench.log
