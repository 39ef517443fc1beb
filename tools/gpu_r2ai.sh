#!/bin/bash
# usage (GPU box): tools/gpu_r2ai.sh <tag>  -- evidence of the final code: full suite, smoke, default bench line + burst line + reference arm,
# ncu launch list of the bench command
tag=$1
mkdir -p gpurun_out
(timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider 2>&1 | tail -8) > gpurun_out/${tag}_pytest.log
(timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3) > gpurun_out/${tag}_smoke.log
( time timeout 900 python bench.py ) > gpurun_out/${tag}_bench_default.json 2> gpurun_out/${tag}_bench_default.err
timeout 300 python bench.py --steps 20 --no-extras > gpurun_out/${tag}_bench_k20.json 2> gpurun_out/${tag}_bench_k20.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/${tag}_bench_reference.json 2> gpurun_out/${tag}_bench_reference.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_fast.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > /dev/null 2>&1
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_smoke.log
cut -c1-400 gpurun_out/${tag}_bench_default.json; cut -c1-300 gpurun_out/${tag}_bench_k20.json; cut -c1-600 gpurun_out/${tag}_bench_reference.json
