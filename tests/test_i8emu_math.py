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


def slice_rn(x: float, inv_scale: float):
    """slice_rn() of csrc/i8emu.cuh (the Float64-tolerance mode's operands): signed digits by round-to-nearest."""
    t = x * inv_scale
    q = []
    for _ in range(S):
        t = t * 128.0
        qi = float(np.rint(t))
        q.append(int(qi))
        t = t - qi
    return q, t


def test_round_to_nearest_slices_are_exact_bounded_and_unbiased():
    """AGP_COMPUTE_F64_EMU slices by rounding (scale = 2 x pow2_scale so that the first digit fits): every digit lies in [-64, 64], the
    representation is exact up to a residual of at most half a unit of the last digit, and -- unlike truncation, whose residual carries the sign
    of its element -- the residual has zero mean, which is what keeps a sum over 1.5e5 points from accumulating it linearly."""
    rng = np.random.default_rng(5)
    row = rng.uniform(0.0, 1.0, size=4000) * 10.0 ** (-4.0 * rng.random(4000))  # one sign: the worst case for a truncation bias
    sc = 2.0 * pow2_scale(float(np.max(np.abs(row))))
    res_rn, res_tr = [], []
    for x in row:
        q, res = slice_rn(float(x), 1.0 / sc)
        assert all(-64 <= v <= 64 for v in q)
        recon = sum(Fraction(v, 128 ** (i + 1)) for i, v in enumerate(q)) + Fraction(res) / 128**S
        assert recon * Fraction(sc) == Fraction(float(x))
        assert abs(res) <= 0.5
        res_rn.append(res)
        res_tr.append(slice7(float(x), 2.0 / sc)[1])
    res_rn, res_tr = np.array(res_rn), np.array(res_tr)
    assert np.all(res_tr >= 0.0) and np.mean(res_tr) > 0.3  # truncation: a bias of about half a unit
    assert abs(np.mean(res_rn)) < 4.0 * 0.29 / np.sqrt(len(row))  # rounding: zero mean (uniform on [-1/2, 1/2] has sigma 0.29)
    # exact accumulation over a slab of 16384 points: seven pairs per group at most, 64^2 per term
    assert 7 * 64 * 64 * 16384 < 2**31


def test_round_to_nearest_product_error_grows_like_sqrt_n():
    """Sum over n of a_n b_n with all-positive operands (the shape of G += (dv A) A^T on its diagonal): the recombined INT8 result with rounded
    slices stays within a few sqrt(n) units of 2^-49 (scale_a scale_b), the truncated one drifts by ~n such units."""
    rng = np.random.default_rng(3)
    n = 4096
    a = rng.uniform(0.05, 1.0, size=n)
    b = rng.uniform(0.05, 1.0, size=n)
    exact = sum(Fraction(float(x)) * Fraction(float(y)) for x, y in zip(a, b))

    def emulate(slicer, mult):
        sa, sb = mult * pow2_scale(float(a.max())), mult * pow2_scale(float(b.max()))
        qa = np.array([slicer(float(x), 1.0 / sa)[0] for x in a], dtype=np.int64)
        qb = np.array([slicer(float(x), 1.0 / sb)[0] for x in b], dtype=np.int64)
        tot = Fraction(0)
        for i in range(S):
            for j in range(S - i):
                tot += Fraction(int(np.dot(qa[:, i], qb[:, j])), 128 ** (i + j + 2))
        return tot * Fraction(sa) * Fraction(sb), sa * sb

    got_rn, unit_rn = emulate(slice_rn, 2.0)
    got_tr, unit_tr = emulate(slice7, 1.0)
    err_rn = abs(float(got_rn - exact)) / (2.0**-49 * unit_rn)
    err_tr = abs(float(got_tr - exact)) / (2.0**-49 * unit_tr)
    assert err_rn < 12.0 * np.sqrt(n)
    assert err_tr > 0.25 * n > err_rn
