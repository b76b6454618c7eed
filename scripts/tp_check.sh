#!/bin/bash
# one gpurun call: the other BASELINE shapes and the parity-grade tensor mode, each with its full-size parity sample
mkdir -p gpurun_out
for w in c4 c5; do
  timeout 250 python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err
  tail -1 gpurun_out/bench_$w.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w', d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],4), d['parity']['max_rel_rgb_vs_oracle'], d['parity']['pass'])"
done
timeout 250 python bench.py --precision tc_f16x3 --steps 10 --warmup 3 --no-cpu-baseline --no-gpu-incumbent > gpurun_out/bench_f16x3.json 2> gpurun_out/bench_f16x3.err
tail -1 gpurun_out/bench_f16x3.json | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('f16x3', d['ms_per_step'], d['roofline']['kernel'], round(d['roofline']['frac'],4), d['parity']['max_rel_rgb_vs_oracle'], d['parity']['pass'])"
echo "== wide kernel probe"; timeout 150 python scripts/mlp_time.py 512 8 2>&1 | grep "TFLOP\|err"
