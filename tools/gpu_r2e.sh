#!/bin/bash
tag=$1
mkdir -p gpurun_out
for t in 1 2; do
(echo "== PLUTO_GPU_TMA=$t"; PLUTO_GPU_TMA=$t PLUTO_GPU_NO_GRAPH=1 timeout 300 python tools/sanitize_cases.py rk 2>&1 | tail -3) >> gpurun_out/${tag}_tma_err.log
done
cat gpurun_out/${tag}_tma_err.log
