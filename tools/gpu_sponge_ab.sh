#!/bin/bash
python - <<'PY'
import numpy as np, torch
from hades252_b200 import CudaStrategy
from oracle import cpu_oracle
n = 1 << 22
rng = np.random.default_rng(1)
lens = rng.integers(1, 33, size=n).astype(np.uint64)
offsets = np.concatenate([[np.uint64(0)], np.cumsum(lens, dtype=np.uint64)])
s = CudaStrategy([0])
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
elems = torch.empty(int(offsets[-1]) * 4, dtype=torch.int64, device="cuda")
s.gen_elems_device(elems.data_ptr(), 0, int(offsets[-1]), 99, sp)
d_off = torch.from_numpy(offsets.view(np.int64)).cuda()
out = torch.empty(n * 4, dtype=torch.int64, device="cuda")
res = {}
for rep in range(3):
    for v in ((1, 6), (1, 3), (1, 0)):
        s.set_variant(*v)
        for _ in range(2): s.sponge_batch_device(elems.data_ptr(), d_off.data_ptr(), n, out.data_ptr(), sp)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream)
        for _ in range(5): s.sponge_batch_device(elems.data_ptr(), d_off.data_ptr(), n, out.data_ptr(), sp)
        b.record(stream); torch.cuda.synchronize()
        res.setdefault(v, []).append(a.elapsed_time(b) / 5)
        dig = out.cpu().numpy().view(np.uint64)[:8].tolist()
    print(rep, {k: round(v[-1], 1) for k, v in res.items()})
print({k: [round(x, 1) for x in v] for k, v in res.items()})
PY
