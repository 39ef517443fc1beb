#!/bin/bash
# usage (GPU box): tools/gpu_r2d.sh <tag>  -- TMA A/B of the fused x1+x2 sweep (north_star: "each choice evidenced by ncu")
tag=$1
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -x -q -k "tma_staging or (fast_within and (blast3d_plm_hlld or rotor2d_ppm_roe or ot2d_plm_hlld))" 2>&1 | tail -4) > gpurun_out/${tag}_pytest.log
summ () { python -c "
import sys, json
for line in sys.stdin:
    if line.startswith('{'):
        d = json.loads(line)
        print('  value %.3e  ms/step %.3f' % (d['value'], d['ms_per_step']), {k: round(v['ms_per_step'], 3) for k, v in d['kernels'].items()})
"; }
for rep in 1 2; do for tma in 0 1; do for w in blast3d_256 turb3d_256 rotor2d_4096; do
  echo "== TMA=$tma $w (rep $rep)"; PLUTO_GPU_TMA=$tma python bench.py --workload $w --steps 20 --no-e2e --no-cpu-baseline 2>&1 | summ
done; done; done > gpurun_out/${tag}_ab.log 2>&1
for tma in 0 1; do
  PLUTO_GPU_TMA=$tma timeout 300 ncu --set full --clock-control none --import-source on -k regex:"sweep_xy" -s 8 -c 1 -f -o gpurun_out/${tag}_ncu_tma$tma python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-extras > gpurun_out/${tag}_ncu_tma$tma.log 2>&1
done
cat gpurun_out/${tag}_pytest.log gpurun_out/${tag}_ab.log
