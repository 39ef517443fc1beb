#!/bin/bash
# usage: tools/variant_bench.sh <arith> <lib1> <lib2> ...   (run on the GPU box)
arith=$1; shift
for lib in "$@"; do
  echo "== $lib ($arith)"
  python bench.py --dev-lib $lib --steps ${STEPS:-10} --warmup 3 --no-e2e --no-cpu-baseline --no-extras --arith $arith ${BENCH_ARGS:-} 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('  value %.3e  ms/step %.2f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 2) for k, v in d['kernels'].items()})
    else:
        print(line.rstrip())
"
done
