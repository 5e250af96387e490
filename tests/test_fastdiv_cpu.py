"""CPU model of csrc/common.cuh: make_fastdiv / fdiv -- the multiply-shift division the single-thread TMA producers and
MMA issuers of the tensor-core convolution use for their tile indices (a runtime `/` costs them ~25 dependent
instructions).  Exactness is needed for every divisor the launcher can produce and every numerator below 2^31."""
import random


def make_fastdiv(d):
    sh = 0
    while (1 << sh) < d:
        sh += 1
    mul = ((1 << 32) * ((1 << sh) - d)) // d + 1
    assert 0 < mul < (1 << 32)
    return mul, sh


def fdiv(n, f):
    mul, sh = f
    hi = (mul * n) >> 32                      # __umulhi
    t = hi + n
    assert t < (1 << 32)                      # the 32-bit add in the kernel may not wrap
    return t >> sh


def test_fastdiv_exact_for_launcher_divisors():
    rnd = random.Random(7)
    divisors = list(range(1, 1200)) + [2047, 2048, 4095, 4096, 65535, 65536, 999983, (1 << 20) + 1]
    for d in divisors:
        f = make_fastdiv(d)
        ns = list(range(0, 600)) + [d * k + e for k in (1, 7, 4097) for e in (-1, 0, 1)]
        ns += [rnd.randrange(1 << 31) for _ in range(64)] + [(1 << 31) - 1]
        for n in ns:
            if 0 <= n < (1 << 31):
                assert fdiv(n, f) == n // d, (n, d)


def test_fastdiv_split_ranges_partition_the_k_loop():
    """it0 / it1 of split s are fdiv(s * k_iters, S), fdiv((s + 1) * k_iters, S): contiguous, complete, balanced."""
    for k_iters in (9, 36, 72, 73, 144, 160, 304):
        for S in (1, 2, 3, 4, 5, 8):
            f = make_fastdiv(S)
            edges = [fdiv(s * k_iters, f) for s in range(S + 1)]
            assert edges[0] == 0 and edges[-1] == k_iters
            sizes = [b - a for a, b in zip(edges, edges[1:])]
            assert min(sizes) >= k_iters // S and max(sizes) <= -(-k_iters // S)
