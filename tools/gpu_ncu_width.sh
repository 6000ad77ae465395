#!/bin/bash
# one full ncu capture of the default perm kernel of width $W (3 or 9), 2^${LOG2:-22} states
W=${W:-9}
mkdir -p gpurun_out
cat > /tmp/ncu_width.py <<PY
import torch
from hades252_b200 import CudaStrategy
w, n = $W, 1 << ${LOG2:-22}
s = CudaStrategy([0], width=w)
sp = torch.cuda.current_stream().cuda_stream
buf = torch.empty(n * w * 4, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, n * w, 1234, sp)
for _ in range(2):
    s.perm_batch_device(buf.data_ptr(), n, sp)
torch.cuda.synchronize()
print("done", s.kernel_info("perm"))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:perm_batch -s 1 -c 1 -f -o gpurun_out/prof_perm$W \
    env PYTHONPATH=$PWD python /tmp/ncu_width.py > gpurun_out/ncu_w$W.log 2>&1
tail -2 gpurun_out/ncu_w$W.log
