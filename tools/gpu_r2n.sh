#!/bin/bash
# usage (GPU box): tools/gpu_r2n.sh <tag>  -- full suite, sanitizers on the families touched last, default bench line, launch list
tag=$1
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/${tag}_pytest.log
for tool in memcheck racecheck initcheck; do
  (timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_cases.py rk ppm_roe io 2>&1 | grep -v "^$" | cut -c1-300 | tail -30) > gpurun_out/${tag}_sanitizer_${tool}.log
done
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
timeout 300 python bench.py --steps 20 --no-extras > gpurun_out/${tag}_bench_k20.json 2> gpurun_out/${tag}_bench_k20.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_fast.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
cat gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_sanitizer_*.log
cut -c1-400 gpurun_out/${tag}_bench_default.json; cut -c1-300 gpurun_out/${tag}_bench_k20.json
