#!/bin/bash
# one gpurun call: timing of the TMEM ping-pong kernel, A/B against the shared-memory ping-pong kernel, quick parity, ncu
mkdir -p gpurun_out
{
  echo "== tp"; MN_TC_TP=1 timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep "TFLOP\|err"
  echo "== pp"; MN_TC_TP=0 timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep TFLOP
} > gpurun_out/tp_check.txt 2>&1
cat gpurun_out/tp_check.txt
MN_TC_TP=1 timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
MN_TC_TP=1 timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_mlp_tp_kernel -s 4 -c 1 -f -o gpurun_out/tc_mlp_tp_kernel python scripts/mlp_time.py 256 32 > gpurun_out/ncu_tp.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -2
