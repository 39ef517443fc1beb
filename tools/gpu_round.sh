#!/bin/bash
# usage (GPU box): tools/gpu_round.sh <tag>
# the round's evidence in one call: GPU tests, the default bench line (RK2) and the CTU bench line,
# ncu launch lists of both, one ncu --set full launch of every CTU kernel
tag=$1
mkdir -p gpurun_out
(timeout 400 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) > gpurun_out/${tag}_pytest.log
timeout 200 python bench.py > gpurun_out/${tag}_bench_fast.json 2> gpurun_out/${tag}_bench_fast.err
timeout 120 python bench.py --steps 20 --no-cpu-baseline --time-stepping hancock > gpurun_out/${tag}_bench_ctu.json 2> gpurun_out/${tag}_bench_ctu.err
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_fast.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > /dev/null 2>&1
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches_ctu.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --time-stepping hancock > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ctu_" -s 7 -c 7 -f -o gpurun_out/${tag}_ctu_ncu python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --time-stepping hancock > gpurun_out/${tag}_ctu_ncu.log 2>&1
cat gpurun_out/${tag}_pytest.log
