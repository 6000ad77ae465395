#!/bin/bash
# host-call latency of hades_perm_batch for medium batches, pageable and pinned memory, with the chunk-splitting
# threshold at 1 MB per chunk (default) and at 8 MB per chunk (round 1's)
for thr in 1048576 8388608; do
echo "== HADES_MIN_SPLIT_BYTES=$thr"
HADES_MIN_SPLIT_BYTES=$thr python - <<'PY'
import torch, time, numpy as np, sys
sys.path.insert(0, ".")
from hades252_b200 import CudaStrategy
s = CudaStrategy([0])
for l2 in (13, 14, 15, 16, 17, 18, 19, 20):
    n = 1 << l2
    pageable = np.zeros((n, 5, 4), dtype=np.uint64)
    pinned_t = torch.zeros(n * 20, dtype=torch.int64, pin_memory=True)
    out = []
    for name, ptr in (("pageable", pageable.ctypes.data), ("pinned", pinned_t.data_ptr())):
        for _ in range(2): s.perm_batch_ptr(ptr, n)
        reps = 10 if l2 <= 17 else 4
        t = time.perf_counter()
        for _ in range(reps): s.perm_batch_ptr(ptr, n)
        out.append("%s %8.1f us" % (name, (time.perf_counter() - t) / reps * 1e6))
    print("n=2^%d (%5.1f MB)  " % (l2, n * 160 / 1e6) + "   ".join(out) + "   [" + s.last_host_path.split(":")[0] + "]")
PY
done
