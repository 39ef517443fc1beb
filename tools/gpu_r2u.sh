#!/bin/bash
# cell-centred EMFs stored by the fused sweep (default) against PLUTO_GPU_NO_EC=1; parity subset first
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "fast or decomposed or full_size or nonuniform" 2>&1 | tail -4) > gpurun_out/r2u_pytest.log
{
echo "## default (cell-centred EMFs from the fused sweep)"; STEPS=20 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
echo "## PLUTO_GPU_NO_EC=1"; PLUTO_GPU_NO_EC=1 STEPS=20 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
echo "## turb3d_512 default / NO_EC"
BENCH_ARGS="--workload turb3d_512" STEPS=6 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
PLUTO_GPU_NO_EC=1 BENCH_ARGS="--workload turb3d_512" STEPS=6 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
echo "## ot2d_512, rotor2d_4096 default / NO_EC"
BENCH_ARGS="--workload ot2d_512" STEPS=200 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
PLUTO_GPU_NO_EC=1 BENCH_ARGS="--workload ot2d_512" STEPS=200 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
BENCH_ARGS="--workload rotor2d_4096" STEPS=10 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
PLUTO_GPU_NO_EC=1 BENCH_ARGS="--workload rotor2d_4096" STEPS=10 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so
} > gpurun_out/r2u_ab.log 2>&1
cat gpurun_out/r2u_pytest.log gpurun_out/r2u_ab.log
