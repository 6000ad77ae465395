"""Single-process multi-device context (the C library owns the communicator): 2^24-leaf Merkle root with resident
leaves through hades_merkle_root_sharded_dev (ncclAllGather of the subtree roots inside the library), the host entry
point, and perm_batch on pageable memory.  usage: python tools/gpu_engine_multi.py [n_devices]"""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from hades252_b200 import CudaStrategy  # noqa: E402

G = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
SEED = 0x4861646573323532
n = 1 << 24
out = {"devices": G}
with CudaStrategy(list(range(G))) as s:
    out["collective"] = s.collective
    per = n // G
    bufs = []
    for g in range(G):
        with torch.cuda.device(g):
            b = torch.empty(per * 4, dtype=torch.int64, device=f"cuda:{g}")
            s.gen_elems_device(b.data_ptr(), g * per, per, SEED, 0, g)
            torch.cuda.synchronize(g)
            bufs.append(b)
    ptrs = [b.data_ptr() for b in bufs]
    for _ in range(3):
        root = s.merkle_root_sharded_device(ptrs, n)
    reps = 10
    t = time.perf_counter()
    for _ in range(reps):
        root = s.merkle_root_sharded_device(ptrs, n)
    out["merkle_2p24_resident_ms"] = (time.perf_counter() - t) / reps * 1e3
    out["root_mont_limbs"] = [hex(int(x)) for x in root]
    out["root_equals_committed_value"] = out["root_mont_limbs"] == ["0x7695019b62c48e7e", "0xa403f682e9373c0", "0xd57e20ff7fb97d67", "0x3eab808a8f6b96a3"]
    host = np.concatenate([b.cpu().numpy().view(np.uint64).reshape(per, 4) for b in bufs])
    s.merkle_root(host)
    t = time.perf_counter()
    r2 = s.merkle_root(host)
    out["merkle_2p24_host_leaves_ms"] = (time.perf_counter() - t) * 1e3
    out["host_root_equal"] = bool(np.array_equal(r2, root))
print(json.dumps(out))
