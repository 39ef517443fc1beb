#!/bin/bash
# usage: tools/variant_bench.sh <arith> <lib1> <lib2> ...   (run on the GPU box)
arith=$1; shift
for lib in "$@"; do
  echo "== $lib ($arith)"
  PLUTO_GPU_LIB=$lib python bench.py --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --arith $arith 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('  value %.3e  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 2) for k, v in d['kernels'].items()})
    else:
        print(line.rstrip())
"
done
