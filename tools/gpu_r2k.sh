#!/bin/bash
# usage (N-GPU box): tools/gpu_r2k.sh <tag> <ngpu> [bench args]  -- the default multi-GPU bench line (dist_check, turb3d_512 weak, 1024^3 strong)
tag=$1; n=$2; shift; shift
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"
( time timeout 1500 $RUN bench.py --gpus $n "$@" ) > gpurun_out/${tag}_bench${n}.json 2> gpurun_out/${tag}_bench${n}.err
grep "^{" gpurun_out/${tag}_bench${n}.json | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line)
    print('value %.4e ms/step %.3f steps %d n_gpus %d workload %s halo %s dist_check %s' % (d['value'], d['ms_per_step'], d['steps'], d['n_gpus'], d['config']['workload'], d['run']['halo'], d.get('dist_check')))
    print(' kernels', {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
    print(' e2e', d['e2e'] and d['e2e']['value'], 'strong', {k: v for k, v in (d.get('strong') or {}).items() if k != 'kernels'})
    print(' clocks', d['clocks'])
"
tail -5 gpurun_out/${tag}_bench${n}.err
