#!/bin/bash
# usage: N=2 bash tools/gpu_multi.sh            (perm, merkle, sponge, sweep, engine, tests)
#        N=8 ONLY="perm merkle" bash tools/gpu_multi.sh
# Results land in gpurun_out/r02_*_n$N.json (copied to profiles/ by hand).
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
ONLY=${ONLY:-perm merkle sponge sweep engine tests}
has() { [[ " $ONLY " == *" $1 "* ]]; }
has tests && python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "multi_device or restore_current or multi_context" 2>&1 | tail -2
has perm && {
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r02_bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_n$N.json')); e=d['e2e']
print('perm N=$N', '%.4g' % d['value'], 'e2e pinned %.4g' % e['value'], 'pageable %.4g' % e.get('pageable',{}).get('value',0), 'probe GB/s pinned %.1f pageable %.1f' % (e['copy_probe_pinned']['GBps_each_direction_all_ranks'], e.get('pageable',{}).get('copy_probe',{}).get('GBps_each_direction_all_ranks',0)))
print('checks', json.dumps(d['checks']['merkle'])[:400], d['checks']['perm'], d['clocks'])"
}
has merkle && {
timeout 600 $TR bench.py --gpus $N --workload merkle --steps 5 --warmup 3 > gpurun_out/r02_bench_merkle_n$N.json 2>> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
cut -c1-300 gpurun_out/r02_bench_merkle_n$N.json; echo
}
has sponge && {
timeout 600 $TR bench.py --gpus $N --workload sponge --steps 5 --warmup 3 > gpurun_out/r02_bench_sponge_n$N.json 2>> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
cut -c1-300 gpurun_out/r02_bench_sponge_n$N.json; echo
}
has sweep && {
timeout 900 $TR bench.py --gpus $N --workload sweep --log2-states ${SWEEP_LOG2:-30} > gpurun_out/r02_bench_sweep_n$N.json 2>> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_sweep_n$N.json'))
for r in d['rows']:
    if r['log2_states_per_gpu'] in (16, 22, 30): print(r)"
}
has engine && { python tools/gpu_engine_multi.py $N 2>&1 | tail -1 | tee gpurun_out/r02_engine_multi_n$N.json; }
has reference && timeout 300 $TR bench.py --gpus $N --impl reference --steps 1 --warmup 1 | cut -c1-200
true
