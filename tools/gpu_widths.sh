#!/bin/bash
# A/B launch shapes for widths 3 and 9 (2^22 states)
python - <<'PY'
import torch, time
from hades252_b200 import CudaStrategy
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
for w, variants in ((3, [(1,6),(2,6),(2,7)]), (9, [(1,7),(2,7),(2,6),(2,2)])):
    s = CudaStrategy([0], width=w)
    n = 1 << 22
    buf = torch.empty(n * w * 4, dtype=torch.int64, device="cuda")
    for v in variants:
        s.set_variant(*v)
        s.gen_elems_device(buf.data_ptr(), 0, n * w, 1234, sp)
        s.perm_batch_device(buf.data_ptr(), n, sp)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
        b.record(stream); torch.cuda.synchronize()
        dig = torch.zeros(4, dtype=torch.int64, device="cuda")
        s.digest_device(buf.data_ptr(), 0, n * w * 4, dig.data_ptr(), sp); torch.cuda.synchronize()
        print(w, v, "%.4g perms/s" % (3 * n / (a.elapsed_time(b) * 1e-3)), s.kernel_info("perm"), hex(int(dig[0].item()) & (2**64-1)))
    s.close()
PY
