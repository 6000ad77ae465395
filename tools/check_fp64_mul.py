#!/usr/bin/env python
"""Offline check of the DFMA Montgomery product of tools/microbench_fp64.cu: reads its `CHECK a[5] b[5] r[5]` lines
(52-bit limbs) from stdin and verifies r == (a * b / 2^260) * b / 2^260 (mod p) with r < 2 p."""
import sys

P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
RI = pow(1 << 260, -1, P)
n = bad = 0
for line in sys.stdin:
    if not line.startswith("CHECK"):
        continue
    v = [int(x) for x in line.split()[1:]]
    a, b, r = (sum(x << (52 * k) for k, x in enumerate(v[i:i + 5])) for i in (0, 5, 10))
    want = a * b * RI % P * b * RI % P
    n += 1
    if r % P != want or r >= 2 * P or any(x >= 1 << 52 for x in v[10:14]):
        bad += 1
print(f"fp64 montgomery product check: {n} products, {bad} wrong")
sys.exit(1 if bad or not n else 0)
