#!/bin/bash
# usage (GPU box): tools/gpu_iter.sh <tag> [variant libs...]
# FAST-tolerance tests with the worst errors printed, then the default bench of the
# in-tree library and of every variant library given.
tag=$1; shift
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q -s -k "fast or smoke or drop" 2>&1 | grep -E "FAST|passed|failed|Error|error|assert" | tail -40) > gpurun_out/${tag}_fasttests.log
tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so "$@" > gpurun_out/${tag}_variants.log 2>&1
cat gpurun_out/${tag}_fasttests.log | tail -25
cat gpurun_out/${tag}_variants.log
