"""Arithmetic of k_stats_tones_imma restated in numpy (the kernel itself is held to the reference's tone levels at 1e-9 by
the -m gpu fixtures): int16 samples as two 8-bit slices, phasors as six signed base-256 digits of round(p 2^45), slice x digit
products accumulated as exact integers and combined with weights 256^e 2^-45.  The result is the exact sum of x[n] P[n] / 2^45:
it differs from the double-precision sum only by the phasor quantisation."""
import math

import numpy as np

DIGITS, SHIFT, TB = 6, 45, 256


def digits_of(P):
    out, q = [], P.copy()
    for _ in range(DIGITS):
        d = ((q & 0xFF) ^ 0x80) - 0x80          # low byte as a signed digit
        q = (q - d) >> 8
        out.append(d)
    assert (q == 0).all()                        # |p| <= 1 fits six signed digits at 2^45
    return out


def block_sum_int8(x, p):
    P = np.round(p * 2.0 ** SHIFT).astype(np.int64)
    d = digits_of(P)
    xh, xl = x >> 8, x & 0xFF                    # x = 256 xh + xl, xh signed, xl unsigned
    acc = [0] * (DIGITS + 1)
    for j in range(DIGITS):
        acc[j] += int((xl * d[j]).sum())         # IMMA u8 x s8, weight 256^j
        acc[j + 1] += int((xh * d[j]).sum())     # IMMA s8 x s8, weight 256^(j+1)
    assert max(abs(a) for a in acc) < 2 ** 31    # int32 accumulators never overflow on a 256-sample block
    v = 0.0
    for e in range(DIGITS, -1, -1):
        v = float(acc[e]) * 2.0 ** (8 * e - SHIFT) + v
    exact = sum(int(a) * int(b) for a, b in zip(x.tolist(), P.tolist()))
    return v, exact / 2.0 ** SHIFT


def test_sliced_integer_block_sums_match_double_precision():
    rng = np.random.default_rng(5)
    m = np.arange(TB)
    for fs in (44100.0, 48000.0):
        for f in (400.0, 7500.0, 3000.0, 2500.0):
            for trig in (np.cos, np.sin):
                p = trig(2 * np.pi * m / fs * f)
                for x in (rng.integers(-32768, 32768, TB), np.full(TB, -32768), np.full(TB, 32767),
                          (rng.normal(0, 160, TB) + 9000 * np.cos(2 * np.pi * m / fs * 400 + 0.7)).astype(np.int64)):
                    x = x.astype(np.int64)
                    v, exact_q = block_sum_int8(x, p)
                    ref = math.fsum((x.astype(np.float64) * p).tolist())
                    assert abs(v - exact_q) <= 1e-9 * max(1.0, abs(exact_q))          # the combination in double is exact to round-off
                    assert abs(v - ref) <= TB * 32768 * 2.0 ** -(SHIFT + 1) + 1e-9    # quantisation bound: 2^-46 per phasor
