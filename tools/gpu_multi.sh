#!/bin/bash
# usage: N=2 bash tools/gpu_multi.sh            (perm, merkle, sponge, reference arm)
#        N=8 ONLY="perm merkle" bash tools/gpu_multi.sh
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
ONLY=${ONLY:-perm merkle sponge reference}
has() { [[ " $ONLY " == *" $1 "* ]]; }
has perm && {
timeout 900 $TR bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
python -c "
import json; d=json.load(open('gpurun_out/bench_n$N.json')); print('perm N=$N', '%.4g' % d['value'], 'e2e %.4g' % d['e2e']['value'], d['e2e'].get('numa_node_rank0'), d['clocks'])"
}
has merkle && {
timeout 600 $TR bench.py --gpus $N --workload merkle --steps 5 --warmup 3 > gpurun_out/bench_merkle_n$N.json 2>> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
cut -c1-500 gpurun_out/bench_merkle_n$N.json
}
has sponge && {
timeout 600 $TR bench.py --gpus $N --workload sponge --steps 5 --warmup 3 > gpurun_out/bench_sponge_n$N.json 2>> gpurun_out/bench_n$N.err || tail -5 gpurun_out/bench_n$N.err
cut -c1-400 gpurun_out/bench_sponge_n$N.json
}
has reference && \
timeout 300 $TR bench.py --gpus $N --impl reference --steps 1 --warmup 1 | cut -c1-200
