#!/bin/bash
# R3 (x3 flux difference kept apart, four march blocks per SM) and the 4-blocks-per-SM fused sweep: parity subset + A/B bench
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast or decomposed or full_size" 2>&1 | tail -4) > gpurun_out/r2r_pytest.log
{
echo "## default (R3 march, 4 blocks per SM)"; STEPS=20 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
echo "## PLUTO_GPU_NO_R3=1"; PLUTO_GPU_NO_R3=1 STEPS=20 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
echo "## xy4: fused sweep 128 registers, 3 ring rows, 4 blocks per SM"; STEPS=20 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu_xy4.so
echo "## turb3d_512 default / NO_R3 / xy4"
BENCH_ARGS="--workload turb3d_512" STEPS=6 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
PLUTO_GPU_NO_R3=1 BENCH_ARGS="--workload turb3d_512" STEPS=6 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
BENCH_ARGS="--workload turb3d_512" STEPS=6 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu_xy4.so
} > gpurun_out/r2r_ab.log 2>&1
cat gpurun_out/r2r_pytest.log gpurun_out/r2r_ab.log
