#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tests/first_light_v2.py > gpurun_out/c3_first_light.log 2>&1
tail -3 gpurun_out/c3_first_light.log
for v in 0 1 2; do
  echo "=== PSE_SPREAD_VAR=$v" >> gpurun_out/c3_variants.log
  PSE_SPREAD_VAR=$v timeout 300 python tests/prof_step.py 1000000 0.3 6 >> gpurun_out/c3_variants.log 2>&1
done
for v in 0 1; do
echo "=== P=8 v2 VAR=$v" >> gpurun_out/c3_variants.log
PSE_SPREAD_VAR=$v PSE_ERROR=1e-4 PSE_XI=0.45 timeout 300 python tests/prof_step.py 1000000 0.4 4 >> gpurun_out/c3_variants.log 2>&1
done
grep -E "===|wall|spread|interp|wave_bin" gpurun_out/c3_variants.log
