#!/bin/bash
# Small-batch latency table: device-resident launch and host call (pageable numpy memory), default context
# (cooperative kernels up to 4736 states) and with the cooperative kernels switched off
python - <<'PY'
import torch, time, numpy as np
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
buf = torch.empty((1 << 20) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, (1 << 20) * 5, 7, sp)
host = np.zeros((1 << 16, 5, 4), dtype=np.uint64); host[:] = buf[: (1 << 16) * 20].cpu().numpy().view(np.uint64).reshape(-1, 5, 4)
def dev(n, reps):
    for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
def hostcall(n, reps):
    h = host[:n].copy()
    s.perm_batch(h)
    t = time.perf_counter()
    for _ in range(reps): s.perm_batch(h)
    return (time.perf_counter() - t) / reps * 1e6
print("states   | cooperative (default, <= 4736 states): device us / host call us | one thread per state: device us / host call us")
for n in (1, 32, 128, 1024, 2368, 4096, 4736, 8192, 1 << 14, 1 << 15, 1 << 16, 1 << 18, 1 << 20):
    reps = 20 if n <= 8192 else 5
    s.set_coop_threshold(4736); d1 = dev(n, reps); h1 = hostcall(n, reps) if n <= (1 << 16) else float("nan")
    s.set_coop_threshold(0); d0 = dev(n, reps); h0 = hostcall(n, reps) if n <= (1 << 16) else float("nan")
    print(f"n={n:8d} | {d1:9.1f} / {h1:9.1f} | {d0:9.1f} / {h0:9.1f} | {n / d1:8.3f} perms/us")
PY
