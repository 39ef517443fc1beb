#!/bin/bash
# usage (GPU box): tools/gpu_r2x.sh <tag>  -- evidence of the final code: full suite, sanitizers on the families touched last,
# default bench line + burst line, ncu launch list, ncu --set full of one launch of every stage kernel
tag=$1
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -8) > gpurun_out/${tag}_pytest.log
for tool in memcheck racecheck initcheck; do
  (timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_cases.py rk grid 2>&1 | grep -v "^$" | cut -c1-300 | tail -30) > gpurun_out/${tag}_sanitizer_${tool}.log
done
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
timeout 300 python bench.py --steps 20 --no-extras > gpurun_out/${tag}_bench_k20.json 2> gpurun_out/${tag}_bench_k20.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_fast.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep|ct_|final|bc_" -s 27 -c 9 -f -o gpurun_out/${tag}_ncu python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_pytest.log; tail -3 gpurun_out/${tag}_sanitizer_*.log
cut -c1-400 gpurun_out/${tag}_bench_default.json; cut -c1-300 gpurun_out/${tag}_bench_k20.json; cut -c1-600 gpurun_out/${tag}_bench_reference.json
