#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for v in ${VARIANTS:-1,1 1,0 0,1}; do
  timeout 600 python bench.py --steps 3 --warmup 3 --log2-states 24 --no-cpu-baseline --no-e2e --variant $v > gpurun_out/bench_v$v.json 2>> gpurun_out/bench.err
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_v$v.json'))
print('$v', '%.4g perms/s' % d['value'], 'frac %.3f' % d['roofline']['frac'], 'peak %.3f' % d['roofline']['peak'], d['kernel_info'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
PY
done
V=${NCU_VARIANT:-1,1}
timeout 900 ncu --set full --clock-control none --import-source on -k regex:perm_batch -s 1 -c 1 -f -o gpurun_out/prof_perm5 \
    python bench.py --steps 1 --warmup 1 --log2-states 22 --no-cpu-baseline --no-e2e --variant $V > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/bench.err
