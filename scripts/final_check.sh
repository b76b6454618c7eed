#!/bin/bash
# one gpurun call: the whole GPU suite, the graded bench line, the training line, the ncu capture + launch list of the default build
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -x --tb=short -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -4 gpurun_out/gpu_tests.log
timeout 300 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; tail -1 gpurun_out/bench_n1.json | cut -c1-600
timeout 200 python bench.py --mode train --steps 8 --warmup 3 > gpurun_out/bench_train.json 2> gpurun_out/bench_train.err; tail -1 gpurun_out/bench_train.json | cut -c1-300
timeout 300 ncu --set full --import-source on --clock-control none -k regex:tc_mlp_tp_kernel -s 5 -c 1 -f -o gpurun_out/tc_mlp_tp_kernel python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-incumbent > gpurun_out/n1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_render.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-incumbent > gpurun_out/l1.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/launches_render.csv | tail -3
