#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for v in 1,1 1,0 1,2 0,1; do
  timeout 600 python bench.py --steps 3 --warmup 3 --log2-states 24 --no-cpu-baseline --no-e2e --variant $v > gpurun_out/bench_v$v.json 2>> gpurun_out/bench.err
  python -c "
import json,sys
d=json.load(open('gpurun_out/bench_v$v.json'))
print('$v', d['value'], d['roofline']['frac'], d['roofline']['peak'], d['kernel_info'], d['clocks'])
"
done
