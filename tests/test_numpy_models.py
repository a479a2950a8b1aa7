"""The two numpy behaviours the bit-exact CUDA path restates, re-verified on whatever
numpy / CPU this suite runs on: if either model stops matching, GPU parity for thresholded or complex inputs
would silently degrade from "identical bits" to "1e-16 apart, then chaotic" -- so they are pinned here.

* ``np.mean`` / ``np.sum`` of a contiguous float64 array: pairwise summation with 8 accumulators per block of
  <= 128 elements and halves split at multiples of 8 (k_stats in csrc/kernels_generic.cu walks the same tree).
* ``np.abs`` of a contiguous complex128 array: max * sqrt(fma(q, q, 1)) with q = min / max (x_cabs in
  csrc/exact.cuh).
"""
from fractions import Fraction

import numpy as np


def pairwise(a, lo, n):
    if n < 8:
        r = 0.0
        for i in range(n):
            r += a[lo + i]
        return r
    if n <= 128:
        r = [a[lo + j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] += a[lo + i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res += a[lo + i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise(a, lo, n2) + pairwise(a, lo + n2, n - n2)


def test_numpy_mean_is_the_pairwise_tree():
    rng = np.random.default_rng(0)
    for shape in [(1, 1), (1, 5), (3, 3), (1, 17), (7, 17), (66, 17), (97, 33), (253, 257), (40, 513), (5, 1025)]:
        a = np.abs(rng.standard_normal(shape)) * 10.0 ** rng.uniform(-3, 3)
        flat = a.ravel().tolist()
        s = pairwise(flat, 0, len(flat))
        assert s == float(np.sum(a)), shape
        assert s / a.size == float(np.mean(a)), shape


def cabs_model(x, y):
    a, b = abs(x), abs(y)
    mx, mn = max(a, b), min(a, b)
    if mx == 0.0:
        return 0.0
    q = mn / mx
    f = float(Fraction(q) * Fraction(q) + 1)  # fma(q, q, 1): one rounding of the exact value
    return mx * float(np.sqrt(np.float64(f)))


def test_numpy_complex_abs_model():
    rng = np.random.default_rng(1)
    n = 20000
    xs = rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)
    ys = rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)
    z = xs + 1j * ys
    got = np.abs(z)                       # contiguous complex128: the loop np.abs(ExtS) runs (lws.pyx:239)
    bad = sum(cabs_model(float(x), float(y)) != float(g) for x, y, g in zip(xs, ys, got))
    assert bad == 0, "%d of %d magnitudes differ from max*sqrt(fma(q,q,1))" % (bad, n)
    assert np.array_equal(np.abs(np.array([0j, 3 + 0j, -4j, 3 + 4j])), [0.0, 3.0, 4.0, 5.0])
