#!/bin/bash
# A/B of the dot-product operand roles (hades.cuh DotRoles): default build (constants as `sca` from W = 9) vs the
# build with -DHADES_SWAP_FROM_W=3 (constants as `sca` at every width), 2^22 states, default launch shapes
for lib in "" "$PWD/hades252_b200/lib/libhades_b200_swap3.so"; do
echo "== ${lib:-default build}"
HADES_B200_LIB=$lib python - <<'PY'
import torch
from hades252_b200 import CudaStrategy
stream = torch.cuda.current_stream(); sp = stream.cuda_stream
for w in (3, 5, 9):
    s = CudaStrategy([0], width=w)
    n = 1 << 22
    buf = torch.empty(n * w * 4, dtype=torch.int64, device="cuda")
    s.gen_elems_device(buf.data_ptr(), 0, n * w, 1234, sp)
    for _ in range(2): s.perm_batch_device(buf.data_ptr(), n, sp)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(5): s.perm_batch_device(buf.data_ptr(), n, sp)
    b.record(stream); torch.cuda.synchronize()
    print(w, "%.4g perms/s" % (5 * n / (a.elapsed_time(b) * 1e-3)), s.kernel_info("perm"))
    s.close()
PY
done
