#!/bin/bash
# usage: LIBS="tag1 tag2" [VARIANTS="2,6"] [LOG2=24] bash tools/exp/gpu_libs_ab.sh  -- A/B of tagged builds (lib/libhades_b200_<tag>.so) vs the default
for lib in "" $LIBS; do
  for v in ${VARIANTS:-2,6}; do
    if [ -n "$lib" ]; then export HADES_B200_LIB=$PWD/hades252_b200/lib/libhades_b200_$lib.so; else unset HADES_B200_LIB; fi
    f=gpurun_out/ab_${lib:-default}_$v
    python bench.py --steps 3 --warmup 3 --log2-states ${LOG2:-24} --no-cpu-baseline --no-e2e --no-checks --variant $v > $f.json 2>$f.err || tail -3 $f.err
    python - <<PY
import json
try:
    d=json.load(open('$f.json'))
    print('${lib:-default}', '$v', '%.4g perms/s' % d['value'], d.get('kernel_info'), 'oracle', d.get('extras',{}).get('oracle_sample_match', d.get('oracle_sample_match')), d['clocks']['sm_mhz'])
except Exception as e: print('${lib:-default} $v failed', e)
PY
  done
done
