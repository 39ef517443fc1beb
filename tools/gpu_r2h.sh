#!/bin/bash
# usage (GPU box): tools/gpu_r2h.sh <tag>  -- full suite + Roe (branch-free IEEE division) bench
tag=$1
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_pytest.log
tools/workload_bench.sh fast rotor2d_4096 blast3d_256 > gpurun_out/${tag}_workloads.log 2>&1
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_workloads.log
