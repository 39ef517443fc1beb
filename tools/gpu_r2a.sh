#!/bin/bash
# usage (GPU box): tools/gpu_r2a.sh <tag>
# round 2, first call: the full GPU suite (all tests, no -x), default bench + the other BASELINE workloads,
# compute-sanitizer (memcheck, racecheck, initcheck) on small cases of every kernel family
tag=$1
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -s 2>&1 | grep -E "FAST|passed|failed|FAILED|Error|error|assert" | tail -80) > gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 50 > gpurun_out/${tag}_bench_fast.json 2> gpurun_out/${tag}_bench_fast.err
tools/workload_bench.sh fast rotor2d_4096 ot2d_512 turb3d_512 > gpurun_out/${tag}_workloads.log 2>&1
for tool in memcheck racecheck initcheck; do
  (timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_cases.py 2>&1 | grep -v "^$" | tail -60) > gpurun_out/${tag}_sanitizer_${tool}.log
done
cat gpurun_out/${tag}_pytest.log | tail -30
cat gpurun_out/${tag}_bench_fast.json | cut -c1-600
cat gpurun_out/${tag}_workloads.log
tail -5 gpurun_out/${tag}_sanitizer_*.log
