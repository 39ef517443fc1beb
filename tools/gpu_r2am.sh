#!/bin/bash
# compute-sanitizer on the final code: the families added in the last part of the round and the PARABOLIC + roe kernels whose launch bounds changed
mkdir -p gpurun_out
for tool in memcheck racecheck initcheck; do
  (timeout 900 compute-sanitizer --tool $tool --print-limit 10 python tools/sanitize_cases.py schemes3 ppm_roe 2>&1 | grep -v "^$" | cut -c1-300 | tail -12) > gpurun_out/r2am_sanitizer_${tool}.log
done
for t in memcheck racecheck initcheck; do echo "== $t"; tail -n 4 gpurun_out/r2am_sanitizer_$t.log; done
