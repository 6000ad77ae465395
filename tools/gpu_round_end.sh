#!/bin/bash
# Round-end GPU pass in one call: ncu captures (default perm kernel at 2^26 with source page; both cooperative kernels),
# launch list of the default bench command, integer peak, latency workload, then what the driver does (tests, smoke,
# reference arm, default bench).  Summaries are produced afterwards on the CPU side (tools/update_kernel_profile.py,
# tools/ncu_summary.py) and copied to profiles/.
mkdir -p gpurun_out
rm -f gpurun_out/*.ncu-rep
LOG2=26 bash tools/gpu_ncu.sh
for n in 64 2048; do
cat > gpurun_out/coop_run.py <<PY
import sys; sys.path.insert(0, ".")
import torch
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); sp = torch.cuda.current_stream().cuda_stream
n = $n
buf = torch.empty(max(n, 8) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, max(n, 8) * 5, 7, sp)
for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none -k regex:coop -s 1 -c 1 -f -o gpurun_out/prof_coop_n$n python gpurun_out/coop_run.py > gpurun_out/ncu_coop_n$n.log 2>&1
tail -1 gpurun_out/ncu_coop_n$n.log
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_launches_default.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-pageable > gpurun_out/r02_launches_bench.log 2>&1
tail -1 gpurun_out/r02_launches_bench.log | cut -c1-200
python bench.py --workload latency > gpurun_out/r02_bench_latency.json 2> gpurun_out/bench_latency.err || tail -3 gpurun_out/bench_latency.err
cut -c1-400 gpurun_out/r02_bench_latency.json; echo
python bench.py --workload merkle --steps 5 --warmup 3 > gpurun_out/r02_bench_merkle_2p24.json 2>> gpurun_out/bench_latency.err; cut -c1-300 gpurun_out/r02_bench_merkle_2p24.json; echo
bash tools/gpu_final.sh
ls -la gpurun_out/*.ncu-rep
