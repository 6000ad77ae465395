#!/bin/bash
python - <<'PY'
import torch, time, numpy as np
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
buf = torch.empty((1 << 20) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, (1 << 20) * 5, 7, sp)
host = np.zeros((1 << 16, 5, 4), dtype=np.uint64); host[:] = buf[: (1 << 16) * 20].cpu().numpy().view(np.uint64).reshape(-1, 5, 4)
for n in (1, 32, 128, 1024, 4096, 1 << 14, 1 << 16, 1 << 18, 1 << 20):
    for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 20 if n <= 4096 else 5
    a.record(stream)
    for _ in range(reps): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    dev_us = a.elapsed_time(b) / reps * 1e3
    line = f"n={n:8d}  device-resident {dev_us:9.1f} us  ({n / dev_us:8.3f} perms/us)"
    if n <= (1 << 16):
        h = host[:n].copy()
        s.perm_batch(h)
        t = time.perf_counter()
        for _ in range(reps): s.perm_batch(h)
        line += f"   host call {(time.perf_counter() - t) / reps * 1e6:9.1f} us"
    print(line)
PY
