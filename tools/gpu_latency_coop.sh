#!/bin/bash
# Small-batch latency: cooperative warp-per-state and 8-lanes-per-state kernels vs the one-thread-per-state kernel
# (device-resident launches, CUDA events) -> profiles/r02_latency_small_batches.txt
python - <<'PY'
import torch
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
buf = torch.empty((1 << 16) * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, (1 << 16) * 5, 7, sp)
print("perm_coop", s.kernel_info("perm_coop"), "merkle_coop", s.kernel_info("merkle_coop"))
print("perm_coop_wide", s.kernel_info("perm_coop_wide"), "merkle_coop_wide", s.kernel_info("merkle_coop_wide"))
def t(n, reps=20):
    for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for n in (1, 8, 64, 512, 592, 593, 1024, 1184, 2048, 2368, 4096, 4736, 8192, 16384, 32768, 65536):
    s.set_coop_threshold(1 << 20); s.set_coop_wide_threshold(0); c = t(n)
    s.set_coop_wide_threshold(1 << 20); w = t(n) if n <= 8192 else float("nan")
    s.set_coop_threshold(0); o = t(n)
    print(f"n={n:6d}  warp per state {w:8.1f} us   8 lanes per state {c:8.1f} us   one thread per state {o:8.1f} us")
PY
python - <<'PY'
# a lone sponge hash (9 elements = 3 permutations) through the host call, per kernel choice
import time
import numpy as np
from hades252_b200 import CudaStrategy
from oracle import cpu_oracle as C
s = CudaStrategy([0])
elems = C.gen_elems(5, 9); offsets = np.array([0, 9], dtype=np.uint64)
def t(reps=200):
    s.sponge_batch(elems, offsets)
    t0 = time.perf_counter()
    for _ in range(reps): s.sponge_batch(elems, offsets)
    return (time.perf_counter() - t0) / reps * 1e6
s.set_coop_threshold(4736); s.set_coop_wide_threshold(592); w = t()
s.set_coop_wide_threshold(0); c = t()
s.set_coop_threshold(0); o = t()
print(f"lone sponge hash of 9 elements (3 permutations), host call: warp per message {w:7.1f} us   8 lanes {c:7.1f} us   one thread {o:7.1f} us")
PY
