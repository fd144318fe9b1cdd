#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tests/first_light_v2.py > gpurun_out/c2_first_light.log 2>&1
tail -12 gpurun_out/c2_first_light.log
for v in "PSE_WAVE=v1" "PSE_WAVE=v2"; do
  echo "=== $v" >> gpurun_out/c2_variants.log
  env $v timeout 300 python tests/prof_step.py 1000000 0.3 6 >> gpurun_out/c2_variants.log 2>&1
done
echo "=== P=8 v2" >> gpurun_out/c2_variants.log
PSE_ERROR=1e-4 PSE_XI=0.45 timeout 300 python tests/prof_step.py 1000000 0.4 4 >> gpurun_out/c2_variants.log 2>&1
grep -E "===|wall|spread|interp|wave_bin" gpurun_out/c2_variants.log
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "dense_ewald or lanczos_matches or tilt_flip or bitwise" > gpurun_out/c2_newtests.log 2>&1
tail -15 gpurun_out/c2_newtests.log
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/c2_alltests.log 2>&1
tail -15 gpurun_out/c2_alltests.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spread2_kernel|interp2_kernel" -s 2 -c 2 -o gpurun_out/c2_wave -f python tests/prof_step.py 1000000 0.3 2 > gpurun_out/c2_ncu_wave.log 2>&1
tail -3 gpurun_out/c2_ncu_wave.log
