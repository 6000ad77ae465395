#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -2 gpurun_out/bench.err; cat gpurun_out/bench.json | cut -c1-600
timeout 600 python bench.py --workload merkle --steps 5 --warmup 3 --verify > gpurun_out/bench_merkle.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_merkle.json
timeout 600 python bench.py --workload sponge --steps 5 --warmup 3 --verify > gpurun_out/bench_sponge.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_sponge.json
timeout 900 python bench.py --workload sweep --log2-states 26 > gpurun_out/bench_sweep.json 2>> gpurun_out/bench.err; python -c "
import json; d=json.load(open('gpurun_out/bench_sweep.json'))
for r in d['rows']: print(r['width'], r['log2_states_per_gpu'], '%.3f ms' % r['ms'], '%.4g' % r['perms_per_s'])
"
tail -3 gpurun_out/bench.err
