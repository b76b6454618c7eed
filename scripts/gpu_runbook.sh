#!/bin/bash
# GPU runbook: the gpurun calls that (re)produce every number and capture under profiles/.
# Usage (from the build container):  scripts/gpu_runbook.sh <step>      e.g.  scripts/gpu_runbook.sh tests
# Each step is ONE gpurun call (one at a time; see `gpurun --status` for the budget).  Outputs land in gpurun_out/;
# copy what should be judged into profiles/ and commit it.
set -euo pipefail
G=/usr/local/graft/bin/gpurun
case "${1:-help}" in
  tests)        # the whole GPU suite, most-verified files first (the driver runs the same with -x)
    $G --timeout 900 -- 'python -m pytest tests -q -m gpu --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -40 gpurun_out/gpu_tests.log' ;;
  staging)      # (historic name) the four GPU test files first run on hardware in round 2 - now part of the default suite
    $G --timeout 900 -- 'python -m pytest tests/test_gpu_zea_train_mode.py tests/test_gpu_zf_fused_render.py tests/test_gpu_zg_expert_parallel.py tests/test_gpu_zh_kernel_variants.py -q --tb=short -p no:cacheprovider > gpurun_out/staging_tests.log 2>&1; tail -60 gpurun_out/staging_tests.log' ;;
  bench)        # graded line (+ gpu_incumbent), reference arm, training and cluster diagnostics
    $G --timeout 900 -- 'python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json;
                         python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null;
                         python bench.py --mode train --steps 5 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -1 gpurun_out/bench_train.json;
                         python bench.py --mode cluster --steps 10 > gpurun_out/bench_cluster.json 2>/dev/null; tail -1 gpurun_out/bench_cluster.json' ;;
  launches)     # ncu launch list of one render step and one training step (shares only: cold caches, serialised)
    $G --timeout 900 -- 'ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_render.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-incumbent > gpurun_out/l1.log 2>&1;
                         ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_train.csv python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/l2.log 2>&1; tail -3 gpurun_out/l2.log' ;;
  ncu-mlp)      # full capture of the fine-pass launch of the forward MLP kernel
    $G --timeout 900 -- 'ncu --set full --import-source on --clock-control none -k regex:tc_mlp_tp_kernel -s 5 -c 1 -f -o gpurun_out/tc_mlp_tp_kernel python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-incumbent > gpurun_out/n1.log 2>&1; ls -la gpurun_out/*.ncu-rep' ;;
  ncu-bwd)      # full capture of the two backward kernels
    $G --timeout 900 -- 'ncu --set full --import-source on --clock-control none -k regex:mlp_bwd -s 4 -c 2 -o gpurun_out/mlp_bwd_kernels python bench.py --mode train --steps 1 --warmup 3 > gpurun_out/n3.log 2>&1; ls -la gpurun_out/*.ncu-rep' ;;
  scale2)       # 2-GPU lines: ray-sharded (graded form) and owner-computes sub-modules
    $G --gpus 2 --timeout 900 -- 'python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; tail -1 gpurun_out/bench_n2.json;
                         python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --parallelism experts --no-cpu-baseline > gpurun_out/bench_n2_experts.json 2> gpurun_out/bench_n2_experts.err; tail -1 gpurun_out/bench_n2_experts.json;
                         python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --gather peer --no-cpu-baseline > gpurun_out/bench_n2_peer.json 2> gpurun_out/bench_n2_peer.err; tail -1 gpurun_out/bench_n2_peer.json' ;;
  probe)        # the 256-wide kernel alone on a 606k-row batch: default (TMEM ping-pong) vs MN_TC_TP=0 (shared-memory ping-pong), + the three probes
    $G --timeout 600 -- 'timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep "TFLOP\|err"; MN_TC_TP=0 timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep TFLOP;
                         timeout 150 python scripts/mlp_time.py 512 8 2>&1 | grep TFLOP;
                         timeout 120 scripts/probes/mma_chain_probe; timeout 120 scripts/probes/tmem_ld_probe; timeout 120 scripts/probes/l2_stream_probe' ;;
  ncu-probe)    # full capture of the default kernel on the probe batch (source-level stall samples per role: producer / issuers / epilogue)
    $G --timeout 600 -- 'timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_mlp_tp_kernel -s 4 -c 1 -f -o gpurun_out/tc_mlp_tp_kernel_probe python scripts/mlp_time.py 256 32 > gpurun_out/ncu_tp.log 2>&1; ls -la gpurun_out/*.ncu-rep' ;;
  shapes)       # the other BASELINE shapes and the parity-grade tensor mode, each with its full-size parity sample
    $G --timeout 900 -- 'for w in c4 c5; do timeout 250 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_$w.json 2>/dev/null; tail -c 600 gpurun_out/bench_$w.json; done;
                         timeout 250 python bench.py --precision tc_f16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_f16x3.json 2>/dev/null; tail -c 600 gpurun_out/bench_f16x3.json' ;;
  *) sed -n 2,6p "$0"; grep -E '^  [a-z0-9-]+\)' "$0" ;;
esac
