#!/bin/bash
# usage: VARIANTS="1,0 1,4" bash tools/gpu_variants.sh   -- quick A/B of kernel variants (2^24 states)
for v in ${VARIANTS:-1,0}; do
  python bench.py --steps 3 --warmup 3 --log2-states ${LOG2:-24} --no-cpu-baseline --no-e2e --variant $v > gpurun_out/bench_v$v.json 2>gpurun_out/bench_v$v.err || tail -3 gpurun_out/bench_v$v.err
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_v$v.json'))
    print('$v', '%.4g perms/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], d['kernel_info'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['digest'][0])
except Exception as e: print('$v failed', e)
PY
done
