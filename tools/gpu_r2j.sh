#!/bin/bash
# usage (2-GPU box): tools/gpu_r2j.sh <tag> <ngpu>  -- NCCL vs peer-store halo exchange: bit-identity check + bench
tag=$1; n=${2:-2}
mkdir -p gpurun_out
RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517"
for halo in nccl peer; do
  (echo "== check_dist halo=$halo"; PLUTO_GPU_HALO=$halo timeout 600 $RUN tools/check_dist.py 2>&1 | grep -v "^W\|^\*\*\*\|Setting OMP" | tail -8) >> gpurun_out/${tag}_dist.log
done
for rep in 1 2; do for halo in nccl peer; do
  (echo "== bench blast3d_256 x$n halo=$halo (rep $rep)"; PLUTO_GPU_HALO=$halo timeout 600 $RUN bench.py --gpus $n --workload blast3d_256 --steps 40 --no-e2e --no-extras 2>&1 | grep "^{" | python -c "
import sys, json
for line in sys.stdin:
    d = json.loads(line); print('  value %.4e ms/step %.3f halo %s' % (d['value'], d['ms_per_step'], d['run']['halo']), {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
") >> gpurun_out/${tag}_dist.log
done; done
cat gpurun_out/${tag}_dist.log
