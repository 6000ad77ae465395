#!/bin/bash
# one full ncu capture of the perm kernel (2^22 states) for variant $V (default: library default)
mkdir -p gpurun_out
VARG=""; [ -n "$V" ] && VARG="--variant $V"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:perm_batch -s 1 -c 1 -f -o gpurun_out/prof_perm5 \
    python bench.py --steps 1 --warmup 1 --log2-states ${LOG2:-22} --no-cpu-baseline --no-e2e $VARG > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
