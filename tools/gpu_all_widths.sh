#!/bin/bash
# perms/s of every supported width 2..14 (2^20 states): tuned kernels at 3, 5, 9, dense per-width kernels elsewhere
python - <<'PY'
import torch
from hades252_b200 import CudaStrategy
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
n = 1 << 20
for w in range(2, 15):
    s = CudaStrategy([0], width=w)
    buf = torch.empty(n * w * 4, dtype=torch.int64, device="cuda")
    s.gen_elems_device(buf.data_ptr(), 0, n * w, 1234, sp)
    for _ in range(2): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    rate = 3 * n / (a.elapsed_time(b) * 1e-3)
    dense_products = 8 * (w * 280 + w * (64 * w + 48)) + 59 * (280 + w * (64 * w + 48))
    print(f"W={w:2d}  {rate:10.4g} perms/s  {s.kernel_info('perm')}  dense-schedule products/perm {dense_products}  -> {rate * dense_products / 1e12:5.2f} Tprod/s if dense")
    s.close()
PY
