"""The arithmetic behind csrc/i8emu.cuh, restated with Python integers (CPU; no device needed): seven signed 7-bit slices per element under one
power-of-two scale per row are an exact representation up to a residual below 2^-49 of the row maximum, every slice product is an exact integer,
the 28 products with i + j <= 6 stay below 2^31 for the k-ranges the engine is given, and the recombined sum differs from the exact product by
less than the stated normwise bound.  (The device engine itself is checked by tools/i8emu_test.cu -> tests/test_gpu_engines.py.)"""
from fractions import Fraction

import numpy as np

S = 7


def pow2_scale(mx: float) -> float:
    if not mx > 0.0:
        return 1.0
    _, e = np.frexp(mx)
    return float(np.ldexp(1.0, e))


def slice7(x: float, inv_scale: float):
    """slice7() of csrc/i8emu.cuh, operation by operation in float64."""
    t = x * inv_scale
    q = []
    for _ in range(S):
        t = t * 128.0
        qi = float(np.trunc(t))
        q.append(int(qi))
        t = t - qi
    return q, t


def test_slices_are_exact_and_bounded():
    rng = np.random.default_rng(7)
    for spread in (0.0, 6.0, 12.0):
        row = rng.uniform(-1, 1, size=400) * 10.0 ** (-spread * rng.random(400))
        sc = pow2_scale(float(np.max(np.abs(row))))
        assert 0.5 <= np.max(np.abs(row)) / sc < 1.0
        for x in row:
            q, res = slice7(float(x), 1.0 / sc)
            assert all(-127 <= v <= 127 for v in q)
            # exact identity in rational arithmetic: x = sc * (sum_i q_i 128^-(i+1) + res * 128^-S)
            recon = sum(Fraction(v, 128 ** (i + 1)) for i, v in enumerate(q)) + Fraction(res) / 128**S
            assert recon * Fraction(sc) == Fraction(float(x))
            assert abs(res) < 1.0  # i.e. the dropped part is below 2^-49 of the row scale


def test_group_sums_fit_int32_and_the_product_meets_its_bound():
    rng = np.random.default_rng(11)
    K = 1024
    a = rng.uniform(-1, 1, size=K) * 10.0 ** (-3.0 * rng.random(K))
    b = rng.uniform(-1, 1, size=K) * 10.0 ** (-3.0 * rng.random(K))
    sa, sb = pow2_scale(float(np.max(np.abs(a)))), pow2_scale(float(np.max(np.abs(b))))
    qa = np.array([slice7(float(x), 1.0 / sa)[0] for x in a], dtype=np.int64)  # [K][S]
    qb = np.array([slice7(float(x), 1.0 / sb)[0] for x in b], dtype=np.int64)
    groups = [0] * S
    for i in range(S):
        for j in range(S - i):
            groups[i + j] += int(np.dot(qa[:, i], qb[:, j]))  # exact integer accumulation = what the INT32 TMEM accumulator holds
    assert all(abs(g) < 2**31 for g in groups)
    # worst case the engine may meet: 7 pairs per group, 127^2 per term, k <= 16384 (the KM_SPLIT cap)
    assert 7 * 127 * 127 * 16384 < 2**31
    # the epilogue: sum_g acc_g 2^-7g in float64, then 2^-14 * scales
    v = 0.0
    for g in range(S):
        v = float(np.float64(groups[g]) * np.float64(128.0 ** (-g)) + np.float64(v))
    got = v * sa * (1.0 / 16384.0) * sb
    exact = sum(Fraction(float(x)) * Fraction(float(y)) for x, y in zip(a, b))
    # normwise bound: dropped residuals and dropped products (i + j >= S) are below S 2^-49 of (row max) x (row max) per term, plus the
    # rounding of the seven float64 additions
    bound = K * (S + 2) * 2.0**-49 * sa * sb + 8 * np.finfo(np.float64).eps * float(abs(exact))
    assert abs(Fraction(got) - exact) <= Fraction(bound)
