#!/bin/bash
# Small-batch latency: cooperative 8-lanes-per-state kernel vs the one-thread-per-state kernel (device-resident
# launches, CUDA events) -> profiles/r02_latency_small_batches.txt
python - <<'PY'
import torch
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
buf = torch.empty((1 << 16) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, (1 << 16) * 5, 7, sp)
print("perm_coop", s.kernel_info("perm_coop"), "merkle_coop", s.kernel_info("merkle_coop"))
def t(n, reps=20):
    for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for n in (1, 8, 64, 512, 1024, 2048, 2368, 4096, 4736, 8192, 16384, 32768, 65536):
    s.set_coop_threshold(1 << 20); c = t(n)
    s.set_coop_threshold(0); o = t(n)
    print(f"n={n:6d}  cooperative {c:8.1f} us   one-thread {o:8.1f} us   ratio {o / c:5.2f}")
PY
