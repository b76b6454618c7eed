#!/bin/bash
# one gpurun call: timing of the TMEM ping-pong kernel (+ experiment without epilogue arithmetic), A/B, quick parity
mkdir -p gpurun_out
{
  for t in 0 2; do echo "== tp MN_TC_TRACE=$t"; MN_TC_TP=1 MN_TC_TRACE=$t timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep "TFLOP\|err"; done
  echo "== pp"; MN_TC_TP=0 timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep TFLOP
} > gpurun_out/tp_check.txt 2>&1
cat gpurun_out/tp_check.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
