#!/bin/bash
# one full ncu capture of the cooperative perm kernel on a small batch (N states, default 2048: one warp per scheduler)
mkdir -p gpurun_out
mkdir -p gpurun_out; cat > gpurun_out/coop_run.py <<PY
import sys; sys.path.insert(0, ".")
import torch
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); sp = torch.cuda.current_stream().cuda_stream
n = int("${N:-2048}")
buf = torch.empty(max(n, 8) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, max(n, 8) * 5, 7, sp)
for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
torch.cuda.synchronize()
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:coop -s 1 -c 1 -f -o gpurun_out/prof_coop python gpurun_out/coop_run.py > gpurun_out/ncu_coop.log 2>&1
tail -2 gpurun_out/ncu_coop.log
