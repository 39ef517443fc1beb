#!/bin/bash
# usage (GPU box): tools/gpu_r2c.sh <tag>  -- the bench contract end to end: default run (other workloads included), reference arm
tag=$1
mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 3 ) > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
( time timeout 600 python bench.py --steps 20 --warmup 3 --no-extras ) > gpurun_out/${tag}_bench_k20.json 2> gpurun_out/${tag}_bench_k20.err
tail -3 gpurun_out/${tag}_bench_default.err; cut -c1-3000 gpurun_out/${tag}_bench_default.json
tail -3 gpurun_out/${tag}_bench_reference.err; cut -c1-1500 gpurun_out/${tag}_bench_reference.json
tail -3 gpurun_out/${tag}_bench_k20.err; cut -c1-800 gpurun_out/${tag}_bench_k20.json
