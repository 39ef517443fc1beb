#!/bin/bash
# usage: tools/workload_bench.sh <arith> <workload> ...   (GPU box)
arith=$1; shift
for w in "$@"; do
  echo "== $w ($arith)"
  python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-e2e --workload $w --arith $arith 2>&1 | python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('  value %.3e  ms/step %.3f  e2e %.3e  launches %d  roofline_frac %.3f' % (d['value'], d['ms_per_step'], (d['e2e'] or {}).get('value', 0), d['gpu_launches'], d['step_roofline']['stencil_roofline_frac']), {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
    else:
        print(line.rstrip()[:300])
"
done
