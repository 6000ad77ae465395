"""What ONE rank of the 8-GPU Merkle run does (2^21 leaves -> 2 subtree roots, then 16 gathered roots -> root), timed on
one GPU as a whole chain and level by level: how much of the chain is kernel time, how much launch gaps."""
import sys
import torch
sys.path.insert(0, ".")
from hades252_b200 import CudaStrategy

s = CudaStrategy([0]); stream = torch.cuda.current_stream(); sp = stream.cuda_stream
n = 1 << 21
leaves = torch.empty(n * 4, dtype=torch.int64, device="cuda")
s.gen_elems_device(leaves.data_ptr(), 0, n, 7, sp)
scratch = torch.empty((n // 4 + n // 16 + 8) * 4, dtype=torch.int64, device="cuda")
roots = torch.empty(16 * 4, dtype=torch.int64, device="cuda")
top = torch.empty(4, dtype=torch.int64, device="cuda")


def chain():
    s.merkle_reduce_device(leaves.data_ptr(), n, 10, scratch.data_ptr(), roots.data_ptr(), sp)
    s.merkle_reduce_device(roots.data_ptr(), 16, 2, scratch.data_ptr(), top.data_ptr(), sp)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    for _ in range(reps):
        fn()
    b.record(stream)
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


whole = timed(chain)
print(f"whole chain (12 launches, no gather): {whole:9.1f} us")
total = 0.0
for m in [n >> (2 * l) for l in range(10)] + [16, 4]:   # timing only: every level reads the first m leaves
    t = timed(lambda: s.merkle_reduce_device(leaves.data_ptr(), m, 1, 0, scratch.data_ptr(), sp))
    print(f"  {m >> 2:8d} nodes  {t:8.1f} us")
    total += t
print(f"sum of the levels, each timed on its own back to back: {total:9.1f} us")
