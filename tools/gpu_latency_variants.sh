#!/bin/bash
# Small-batch latency of every kernel variant (device-resident launches, CUDA events):
# which register budget / schedule is fastest when only a few warps run (lone-warp regime)?
python - <<'PY'
import torch
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
buf = torch.empty((1 << 16) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, (1 << 16) * 5, 7, sp)
for algo, regs in ((2, 6), (2, 3), (2, 0), (2, 1), (2, 2), (1, 2), (0, 2), (0, 0)):
    s.set_variant(algo, regs)
    info = s.kernel_info("perm")
    line = f"algo {algo} regs {regs} ({info['regs_per_thread']:3d} regs, {info['local_bytes']} B local):"
    for n in (1, 2048, 8192, 16384, 32768, 65536):
        for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(10): s.perm_batch_device(buf.data_ptr(), n, sp)
        b.record(stream); torch.cuda.synchronize()
        line += f"  n={n}: {a.elapsed_time(b) / 10 * 1e3:7.1f} us"
    print(line)
PY
