#!/bin/bash
# timing experiments on the cooperative kernel (libs built with HADES_COOP_EXP; results are WRONG by design)
for tag in "" exp1 exp3; do
lib=""; [ -n "$tag" ] && lib="$PWD/hades252_b200/lib/libhades_b200_$tag.so"
HADES_B200_LIB=$lib python - <<PY
import torch, sys
sys.path.insert(0, ".")
from hades252_b200 import CudaStrategy
s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
buf = torch.zeros(4096 * 20, dtype=torch.int64, device="cuda")
s.gen_elems_device(buf.data_ptr(), 0, 4096 * 5, 7, sp)
out = []
for n in (1, 2048):
    for _ in range(3): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(20): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    out.append("n=%d %.1f us" % (n, a.elapsed_time(b) / 20 * 1e3))
print("${tag:-default}", s.kernel_info("perm_coop")["regs_per_thread"], "regs:", ", ".join(out))
PY
done
