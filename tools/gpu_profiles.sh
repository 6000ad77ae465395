#!/bin/bash
# ncu captures of the round: cooperative kernel (2048 states), full-size default perm kernel (2^26 states) with source
# page, launch list of the default bench command, integer peak with its launch list
mkdir -p gpurun_out
N=2048 bash tools/gpu_ncu_coop.sh
LOG2=26 bash tools/gpu_ncu.sh
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pageable > gpurun_out/r02_launches_bench.log 2>&1
tail -1 gpurun_out/r02_launches_bench.log | cut -c1-200
cat > gpurun_out/imad_peak.py <<PY
import sys, json; sys.path.insert(0, ".")
from hades252_b200 import CudaStrategy
s = CudaStrategy([0])
names = {0: "imad_wide_x_carry_chain4", 1: "imad_wide_carry_out_only", 2: "imad_lo32_half_product_context_only", 3: "imad_lo_plus_imad_hi_pair", 4: "imad_wide_x_carry_chain16"}
print(json.dumps({names[v]: s.imad_peak(v) / 1e12 for v in range(5)}))
PY
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_imad_peak.csv python gpurun_out/imad_peak.py | tail -1 > gpurun_out/r02_imad_peak_under_ncu.json
python gpurun_out/imad_peak.py | tail -1 > gpurun_out/r02_imad_peak.json; cat gpurun_out/r02_imad_peak.json
