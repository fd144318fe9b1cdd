#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 2 4 3 5 6 7; do
  echo "=== PSE_SPREAD_DBG=$v" >> gpurun_out/c4_dbg.log
  PSE_SPREAD_DBG=$v timeout 300 python tests/prof_step.py 1000000 0.3 4 2>&1 | grep -E "spread " >> gpurun_out/c4_dbg.log
done
cat gpurun_out/c4_dbg.log
