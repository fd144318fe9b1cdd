#!/bin/bash
# round-2 GPU call 1: first light of spread2/interp2, new parity tests, variant timings, ncu of the new kernels
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv > gpurun_out/c1_gpu.txt
BULK=all
timeout 300 python tests/first_light_v2.py all > gpurun_out/c1_first_light.log 2>&1
rc=$?
if [ $rc -ne 0 ]; then
  echo "first light with bulk failed rc=$rc; retrying without bulk" >> gpurun_out/c1_first_light.log
  timeout 300 python tests/first_light_v2.py nobulk >> gpurun_out/c1_first_light.log 2>&1
  export PSE_SPREAD_BULK=0
fi
tail -30 gpurun_out/c1_first_light.log
# per-phase timings of variants at the headline config
for v in "PSE_WAVE=v1" "PSE_SPREAD_BULK=0" "PSE_SPREAD_BULK=1" "PSE_TILE_ALT=1 PSE_SPREAD_BULK=0" "PSE_TILE_ALT=1 PSE_SPREAD_BULK=1"; do
  echo "=== $v" >> gpurun_out/c1_variants.log
  env $v timeout 300 python tests/prof_step.py 1000000 0.3 6 >> gpurun_out/c1_variants.log 2>&1
done
PSE_ERROR=1e-4 PSE_XI=0.45 timeout 300 python tests/prof_step.py 1000000 0.4 4 >> gpurun_out/c1_variants.log 2>&1
echo "=== P=8 v1" >> gpurun_out/c1_variants.log
PSE_WAVE=v1 PSE_ERROR=1e-4 PSE_XI=0.45 timeout 300 python tests/prof_step.py 1000000 0.4 4 >> gpurun_out/c1_variants.log 2>&1
cat gpurun_out/c1_variants.log
# the parity suite (new tests first)
timeout 1500 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -s -k "dense_ewald or lanczos_matches or tilt_flip or bitwise" > gpurun_out/c1_newtests.log 2>&1
tail -15 gpurun_out/c1_newtests.log
timeout 2400 python -m pytest tests -q -m gpu -x > gpurun_out/c1_alltests.log 2>&1
tail -15 gpurun_out/c1_alltests.log
# ncu: the new kernels
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"spread2_kernel|interp2_kernel|wweights|wgather|wbin" -s 8 -c 10 -o gpurun_out/c1_wave -f python tests/prof_step.py 1000000 0.3 2 > gpurun_out/c1_ncu_wave.log 2>&1
tail -3 gpurun_out/c1_ncu_wave.log
