#!/bin/bash
# usage (GPU box): tools/gpu_ab_env.sh <tag> <ENVVAR> <workload> ...  -- bench.py --steps 20 with ENVVAR=0 and =1, two repetitions
tag=$1; var=$2; shift; shift
mkdir -p gpurun_out
for rep in 1 2; do for v in 0 1; do for w in "$@"; do
  echo "== $var=$v $w (rep $rep)"; env $var=$v python bench.py --workload $w --steps 20 --no-e2e --no-cpu-baseline 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('  value %.3e  ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
"
done; done; done > gpurun_out/${tag}_ab.log 2>&1
cat gpurun_out/${tag}_ab.log
