#!/usr/bin/env bash
mkdir -p gpurun_out
for v in "" "MOOLIB_B200_NO_AR_STREAM=1" "MOOLIB_B200_NO_MID_EVENT=1"; do
  echo "=== variant: $v"
  env MB_AR_DEBUG=1 $v timeout 120 python -m pytest tests/test_accumulator_gpu.py -x -q -k single_learner 2>&1 | grep -E "mb_ar dbg|passed|failed|Error" | head -20
done
