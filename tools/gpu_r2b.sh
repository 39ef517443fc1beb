#!/bin/bash
# usage (GPU box): tools/gpu_r2b.sh <tag>
tag=$1
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) > gpurun_out/${tag}_pytest.log
tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so pluto_b200/lib/libpluto_gpu_m2.so pluto_b200/lib/libpluto_gpu_ct12.so > gpurun_out/${tag}_variants.log 2>&1
(echo "== no plan"; PLUTO_GPU_NO_PLAN=1 tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so) >> gpurun_out/${tag}_variants.log 2>&1
(echo "== turb3d_256"; BENCH_ARGS="--workload turb3d_256" tools/variant_bench.sh fast pluto_b200/lib/libpluto_gpu.so pluto_b200/lib/libpluto_gpu_m2.so) >> gpurun_out/${tag}_variants.log 2>&1
for fam in rk exact ppm_roe hll_uct_hll ctu bc halo io; do
  (timeout 600 compute-sanitizer --tool racecheck --print-limit 6 python tools/sanitize_cases.py $fam 2>&1 | grep -v "^$" | cut -c1-400) > gpurun_out/${tag}_racecheck_${fam}.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"sweep|ct_|final|bc_" -s 27 -c 9 -f -o gpurun_out/${tag}_ncu python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/${tag}_ncu.log 2>&1
cat gpurun_out/${tag}_pytest.log
cat gpurun_out/${tag}_variants.log
grep -h "SUMMARY" gpurun_out/${tag}_racecheck_*.log
