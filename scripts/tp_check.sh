#!/bin/bash
# one gpurun call: probe timing of the default kernel + the parity tests that exercise it
mkdir -p gpurun_out
{ echo "== tp"; timeout 150 python scripts/mlp_time.py 256 32 2>&1 | grep "TFLOP\|err"; } > gpurun_out/tp_check.txt 2>&1
cat gpurun_out/tp_check.txt
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_zh_kernel_variants.py tests/test_gpu_ze_properties.py -x -q 2>&1 | tail -3
timeout 100 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-gpu-incumbent 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('bench', d['ms_per_step'], d['e2e']['ms_per_step'], round(d['roofline']['frac'],4), d['parity']['max_rel_rgb_vs_oracle'])"
