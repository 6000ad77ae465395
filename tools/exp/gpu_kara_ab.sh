#!/bin/bash
# A/B of the Karatsuba partial-round dot products (kara.cuh, -DHADES_KARA tagged build) against the default kernel
for lib in "" kara; do
  for v in 2,6 2,9; do
    if [ -n "$lib" ]; then export HADES_B200_LIB=$PWD/hades252_b200/lib/libhades_b200_$lib.so; else unset HADES_B200_LIB; fi
    python bench.py --steps 3 --warmup 3 --log2-states 24 --no-cpu-baseline --no-e2e --no-checks --variant $v > gpurun_out/kara_${lib:-default}_$v.json 2>gpurun_out/kara_${lib:-default}_$v.err || tail -3 gpurun_out/kara_${lib:-default}_$v.err
    python - <<PY
import json
try:
    d=json.load(open('gpurun_out/kara_${lib:-default}_$v.json'))
    print('${lib:-default}', '$v', '%.4g perms/s' % d['value'], d.get('kernel_info'), 'oracle', d.get('oracle_sample_match'), d['clocks']['sm_mhz'])
except Exception as e: print('${lib:-default} $v failed', e)
PY
  done
done
