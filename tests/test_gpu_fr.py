"""Directed tests of the DEVICE field arithmetic (hades252_b200/csrc/fr.cuh) through the test-only entry point
hades_fr_op_dev: the real PTX carry chains on caller-supplied operands against Python big integers.

The permutation tests only ever see these routines through whole permutations, where some paths are hit with
probability 2^-32 per reduction step (a zero low limb under non-zero upper limbs).  Here they are constructed.
Reference call sites of the operations: src/strategies/scalar.rs:28 (`+=`), :33 (`square`, `*`), :44 (`*`, `+=`).
Bit-exact everywhere: a Montgomery reduction's un-normalised result is the unique (T + M p) / R with M = -T/p mod R."""
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

P = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
R = 1 << 256
NP = (-pow(P, -1, R)) % R          # -p^-1 mod 2^256
RINV = pow(R, -1, P)


def limbs(v, n):
    return [(v >> (32 * k)) & 0xFFFFFFFF for k in range(n)]


def value(row):
    return sum(int(x) << (32 * k) for k, x in enumerate(row))


def redc(T, bits=256):
    """the exact un-normalised Montgomery result: (T + M p) / 2^bits, M = T * (-1/p) mod 2^bits"""
    m = ((T % (1 << bits)) * (NP % (1 << bits))) % (1 << bits)
    assert (T + m * P) % (1 << bits) == 0
    return (T + m * P) >> bits


def run(strategy, op, rows):
    return strategy.fr_op(op, np.array(rows, dtype=np.uint32))


rng = random.Random(20261017)
EDGE = [0, 1, 2, P - 1, P - 2, (1 << 255) % P, (1 << 256) % P, (1 << 32) - 1, (1 << 64) - 1, (1 << 128) - 1, P >> 1,
        (P >> 1) + 1, 1 << 32, 1 << 224, ((1 << 256) - 1) % P, 0xFFFFFFFF << 224]


def rand_fp():
    r = rng.random()
    if r < 0.15:
        return rng.choice(EDGE) % P
    if r < 0.25:  # sparse limb patterns: zero / all-ones limbs
        return value([rng.choice([0, 0xFFFFFFFF, rng.getrandbits(32)]) for _ in range(8)]) % P
    return rng.randrange(P)


def test_fr_mul_add_sbox_exact(cuda_strategy):
    a = [rand_fp() for _ in range(20000)] + [x for x in EDGE for _ in EDGE]
    b = [rand_fp() for _ in range(20000)] + [y % P for _ in EDGE for y in EDGE]
    a = [x % P for x in a]
    rows = [limbs(x, 8) + limbs(y, 8) for x, y in zip(a, b)]
    got = run(cuda_strategy, 0, rows)
    for x, y, g in zip(a, b, got):
        assert value(g) == x * y * RINV % P
    got = run(cuda_strategy, 1, rows)
    for x, y, g in zip(a, b, got):
        assert value(g) == (x + y) % P
    got = run(cuda_strategy, 2, [limbs(x, 8) for x in a])
    for x, g in zip(a, got):
        x2 = x * x * RINV % P
        assert value(g) == x2 * x2 * RINV * x * RINV % P   # Montgomery form of the fifth power


def test_raw_products_and_squares_exact(cuda_strategy):
    """dot_mont<1>, fr_mul_lazy, sqr_mont and mul_wide on operands up to the lazy bounds (x^4 < 1.956 p; any a, b
    with a*b/R + p < 2^256 for the 8-limb forms; anything below 2^256 for the raw 9-limb forms)"""
    full = [rng.getrandbits(256) for _ in range(6000)] + [(1 << 256) - 1, (1 << 256) - (1 << 32), P, 2 * P - 1, 0, 1]
    pairs = [(rng.choice(full), rng.choice(full)) for _ in range(12000)] + [((1 << 256) - 1, (1 << 256) - 1), (P - 1, P - 1)]
    got = run(cuda_strategy, 9, [limbs(x, 8) + limbs(y, 8) for x, y in pairs])
    for (x, y), g in zip(pairs, got):
        assert value(g) == redc(x * y)
    got = run(cuda_strategy, 4, [limbs(x, 8) + limbs(y, 8) for x, y in pairs])
    for (x, y), g in zip(pairs, got):
        assert value(g) == x * y
    lazy = [(x % (2 * P), y % (2 * P)) for x, y in pairs]          # a*b/R + p < 4p^2/R + p = 2.81p ... keep below 2^256
    lazy = [(x, y) for x, y in lazy if redc(x * y) < R]
    got = run(cuda_strategy, 16, [limbs(x, 8) + limbs(y, 8) for x, y in lazy])
    for (x, y), g in zip(lazy, got):
        assert value(g) == redc(x * y)
    sq = [x for x in full if x * x < P * R] + [P - 1, int(1.45 * P), 0, 1, (1 << 255) + 12345]
    got = run(cuda_strategy, 3, [limbs(x, 8) for x in sq])
    for x, g in zip(sq, got):
        assert value(g) == redc(x * x)


def _t_with_zero_quotient_limbs(zero_steps):
    """512-bit T whose Montgomery quotient M has zero limbs exactly at `zero_steps` (the running low limb is zero at
    those reduction steps while the limbs above it are not): T = -M p mod R in the low half, random upper half"""
    m = [rng.getrandbits(32) | 1 for _ in range(8)]
    for s in zero_steps:
        m[s] = 0
    M = value(m)
    lo = (-M * P) % R
    hi = rng.getrandbits(254)
    T = lo + (hi << 256)
    assert ((T % R) * NP) % R == M
    return T


def test_reduction_steps_with_zero_low_limb(cuda_strategy):
    """`nz = 0` under non-zero upper limbs, at each of the 8 reduction steps, alone and combined"""
    cases = []
    for s in range(8):
        cases += [_t_with_zero_quotient_limbs([s]) for _ in range(40)]
    for _ in range(400):
        k = rng.randrange(2, 8)
        cases.append(_t_with_zero_quotient_limbs(rng.sample(range(8), k)))
    cases.append(_t_with_zero_quotient_limbs(list(range(8))))
    cases += [0, 1 << 256, ((1 << 254) - 1) << 256, (P - 1) * (P - 1), P * R - 1]
    got = run(cuda_strategy, 5, [limbs(t, 16) for t in cases])
    for t, g in zip(cases, got):
        assert value(g) == redc(t)
    # the same through products: a * b with b chosen so that the quotient of a*b has a zero limb at step s
    rows, want = [], []
    for s in range(8):
        for _ in range(60):
            a = rng.randrange(1, P) | 1                       # odd: invertible mod 2^256
            t = _t_with_zero_quotient_limbs([s]) % R           # low half only: b = t / a mod R gives a*b = t (mod R)
            b = (t * pow(a, -1, R)) % R
            rows.append(limbs(a, 8) + limbs(b, 8))
            want.append(redc(a * b))
    got = run(cuda_strategy, 9, rows)
    for w, g in zip(want, got):
        assert value(g) == w


def test_dot_products_at_their_lazy_bounds(cuda_strategy):
    """dot_mont<4>, dot_mont<5>, dot_mont_plus<4> (512-bit addend injected into the reduction) and the short reduction
    of constant products, random and with every operand at its asserted maximum (constants and words p - 1, the S-box
    product x^4 * x with x^4 < 1.956 p): exact values, and the 9-limb results stay within their documented bounds"""
    def rows_for(n_terms, extreme):
        A = [(P - 1) if extreme else rand_fp() for _ in range(n_terms)]
        B = [(P - 1) if extreme else rand_fp() for _ in range(n_terms)]
        return A, B
    for op, n_terms, bound in ((6, 4, 4), (8, 5, 4)):
        rows, want = [], []
        for i in range(3000):
            A, B = rows_for(n_terms, extreme=(i < 3))
            rows.append(sum((limbs(x, 8) for x in A), []) + sum((limbs(y, 8) for y in B), []))
            want.append(redc(sum(x * y for x, y in zip(A, B))))
        got = run(cuda_strategy, op, rows)
        for w, g in zip(want, got):
            assert value(g) == w and w < bound * P
    rows, want = [], []
    for i in range(4000):
        A, B = rows_for(4, extreme=(i < 3))
        x4 = int(1.956 * P) - i if i < 3 else rng.randrange(int(1.956 * P))
        x = (P - 1) if i < 3 else rand_fp()
        t = x4 * x
        rows.append(sum((limbs(a, 8) for a in A), []) + sum((limbs(b, 8) for b in B), []) + limbs(t, 16))
        want.append(redc(t + sum(a * b for a, b in zip(A, B))))
    got = run(cuda_strategy, 7, rows)
    for w, g in zip(want, got):
        assert value(g) == w and w < 3.7 * P + 1
    # mul_const_short<4>: r = (sum_j X_j * y[64 j .. 64 j + 63]) / 2^64 with a two-step reduction
    rows, want = [], []
    for i in range(3000):
        X = [(P - 1) if i < 2 else rand_fp() for _ in range(4)]
        y = (P - 1) if i < 2 else rng.getrandbits(256)
        rows.append(sum((limbs(x, 8) for x in X), []) + limbs(y, 8) + [0] * 24)
        T = sum(X[j] * ((y >> (64 * j)) & ((1 << 64) - 1)) for j in range(4))
        want.append(redc(T, 64))
    got = run(cuda_strategy, 15, rows)
    for w, g in zip(want, got):
        assert value(g) == w


@pytest.mark.parametrize("log2", [0, 1, 2, 3, 4])
def test_canonicalisation(cuda_strategy, log2):
    top = (2 * P) << log2
    vals = [rng.randrange(top) for _ in range(5000)]
    for k in range(1, (2 << log2) + 1):   # around every multiple of p inside the range
        vals += [k * P - 1, k * P, k * P + 1] if k * P + 1 < top else [k * P - 1]
    vals += [0, 1, top - 1]
    vals = [v for v in vals if 0 <= v < top]
    got = run(cuda_strategy, 10 + log2, [limbs(v, 9) for v in vals])
    for v, g in zip(vals, got):
        assert value(g) == v % P
