#!/bin/bash
# usage (N-GPU box): tools/gpu_r2p.sh <tag> <ngpu>
tag=$1; n=$2
mkdir -p gpurun_out
(timeout 1200 python -m pytest tests -m gpu -x -q -k "single_thread_multi_block or several_blocks or nccl_decomposition" 2>&1 | tail -6) > gpurun_out/${tag}_pytest.log
(timeout 600 python tools/multi_bench.py 256 $n 20 2>&1 | tail -3; timeout 600 python tools/multi_bench.py 256 1 20 2>&1 | tail -1) > gpurun_out/${tag}_multi.log
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_multi.log
