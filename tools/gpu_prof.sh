#!/bin/bash
# usage (GPU box): tools/gpu_prof.sh <tag> [bench args...]
# one default bench (no e2e/cpu) + ncu --set full of one launch of each sweep kernel
tag=$1; shift
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --set full --clock-control none --import-source on -k regex:"sweep|ct_|final" -s 27 -c 6 -f -o gpurun_out/${tag}_ncu python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline "$@" > gpurun_out/${tag}_ncu.log 2>&1
python - <<PY
import json
for line in open("gpurun_out/${tag}_bench.json"):
    if line.startswith("{"):
        d = json.loads(line)
        print("value %.3e ms/step %.3f" % (d["value"], d["ms_per_step"]), {k: round(v["ms_per_step"], 3) for k, v in d["kernels"].items()})
PY
