"""Known-answer material from the reference tree pinned on
the oracle restatement.  CPU only."""
import math

import numpy as np
import pytest

import synth
from oracle import axctd_oracle as ao
from oracle import pss78

README_FRAME = [int(c) for c in "10" "011100100001" "001000011111" "011110"]   # reference README.md:87


def test_readme_frame_known_answer():
    assert ao.check_crc(README_FRAME)
    assert ao.bits_to_hex(README_FRAME) == "9c8487de"
    cint = ao.bits_to_int(README_FRAME[2:14])
    tint = ao.bits_to_int(README_FRAME[14:26])
    assert (cint, tint) == (1825, 543)
    lut = ao.load_temp_lut()
    assert lut[543] == 0.32812604
    assert cint * 60 / 4096 == 26.7333984375


def test_crc_single_bit_flips_fail():
    for i in range(32):
        f = list(README_FRAME)
        f[i] ^= 1
        assert not ao.check_crc(f)


def test_crc_generator_matches_synth_and_is_linear():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.integers(0, 2, 26).tolist()
        b = rng.integers(0, 2, 26).tolist()
        fa, fb = a + synth.crc6(a), b + synth.crc6(b)
        assert ao.check_crc(fa) and ao.check_crc(fb)
        assert ao.check_crc([x ^ y for x, y in zip(fa, fb)])
    v = ao.crc_valid_positions(([1, 0] + [0] * 30) * 3)
    assert len(v) == 96 - 31


def test_vectorised_crc_equals_scalar():
    rng = np.random.default_rng(1)
    bits = rng.integers(0, 2, 3000).tolist()
    v = ao.crc_valid_positions(bits)
    for s in range(len(v)):
        assert bool(v[s]) == (bits[s:s + 2] == [1, 0] and ao.check_crc(bits[s:s + 32]))


def test_temp_lut_layout():
    lut = np.asarray(ao.load_temp_lut())
    assert lut.shape == (4096,)
    assert lut[0] == lut[4094] == lut[4095] == -99.0
    assert lut[1] == -5.7246472 and lut[4093] == 35.606299
    assert np.all(np.diff(lut[1:4094]) > 0)


def test_coefficient_text_format():
    assert ao.coefficient_from_hex("b72000000d01") == 0.7200000000000001
    assert ao.coefficient_from_hex("b27612400b00") == 2.76124
    assert ao.coefficient_from_hex("d23800700d04") == -0.000238007
    assert ao.coefficient_from_hex("b00000000b00") == 0.0
    with pytest.raises(ValueError):
        ao.coefficient_from_hex("a72000000d01")
    for v in (0.72, 2.76124, -0.000238007, -0.053328, 0.994372, 1.04584, -0.0622192):
        assert math.isclose(ao.coefficient_from_hex(synth.encode_coefficient(v)), v, rel_tol=1e-7)


def test_header_round_trip():
    spec = synth.DropSpec(seed=3)
    frames = synth.header_frames(spec)
    bits = [1] * 40 + sum(frames, []) + [1, 0, 1] * 40
    md = ao.parse_header(bits)
    assert all(md["counter_found"])
    assert md["serial_no"] == "00123456" and md["probe_code"] == "a000" and md["max_depth"] == "1000"
    np.testing.assert_allclose(md["zcoeff"], spec.zcoeff, rtol=1e-7)
    np.testing.assert_allclose(md["tcoeff"], spec.tcoeff, rtol=1e-7)
    np.testing.assert_allclose(md["ccoeff"], spec.ccoeff, rtol=1e-7)


def test_trim_header_finds_pulse_end():
    spec = synth.DropSpec(seed=5)
    hdr = sum(synth.header_frames(spec), [])
    bits = [0, 1, 0] * 10 + [1] * 800 + hdr + [1, 0] * 300
    out = ao.trim_header(bits)
    assert len(out) == 2400
    md = ao.parse_header(out)
    assert sum(md["counter_found"]) == 72


def test_pss78_unesco_check_value():
    # UNESCO 1983: R = 1.888091, t68 = 40, p = 10000 dbar -> S = 40.00000
    c = 1.888091 / pss78.C3515_INV
    t90 = 40.0 / 1.00024
    assert abs(pss78.SP_from_C(c, t90, 10000.0) - 40.0) < 5e-5
    # standard seawater: C(35,15,0) = 42.914 mS/cm
    assert abs(pss78.SP_from_C(42.914, 15.0 / 1.00024, 0.0) - 35.0) < 1e-6


# gsw.SP_from_C check values: the six-point example of the GSW toolbox documentation for gsw_SP_from_C
GSW_C = [34.5487, 34.7275, 34.8605, 34.6810, 34.5680, 34.5600]
GSW_T = [28.7856, 28.4329, 22.8103, 10.2600, 6.8863, 4.4036]
GSW_P = [10.0, 50.0, 125.0, 250.0, 600.0, 1000.0]
GSW_SP = [20.009869599086951, 20.265511864874270, 22.981513062527689, 31.204503263727982, 34.032315787432829,
          36.400308494388170]


def test_pss78_matches_gsw_documented_check_values():
    got = pss78.SP_from_C(np.array(GSW_C), np.array(GSW_T), np.array(GSW_P))
    np.testing.assert_allclose(got, GSW_SP, rtol=1e-14, atol=0)
    for c, t, p, sp in zip(GSW_C, GSW_T, GSW_P, GSW_SP):          # scalar call, as parse.py:132 makes it
        assert abs(float(pss78.SP_from_C(c, t, p)) - sp) <= 1e-14 * sp


def test_pss78_hill_extension_continuous_and_nan():
    for t in (0.0, 10.0, 25.0):
        lo, hi = 0.5, 6.0
        for _ in range(200):                      # bisect conductivity where SP == 2
            mid = 0.5 * (lo + hi)
            if pss78.SP_from_C(mid, t, 0.0) < 2.0:
                lo = mid
            else:
                hi = mid
        assert abs(pss78.SP_from_C(lo, t, 0.0) - pss78.SP_from_C(hi, t, 0.0)) < 1e-7
    assert math.isnan(pss78.SP_from_C(-1.0, 10.0, 0.0))
    assert abs(pss78.SP_from_C(0.0, 10.0, 0.0)) < 1e-12
    v = pss78.SP_from_C(np.array([30.0, 55.0]), np.array([5.0, 25.0]), np.array([10.0, 500.0]))
    assert v.shape == (2,) and np.all((v > 20) & (v < 45))


def test_boxsmooth_lag_matches_definition():
    rng = np.random.default_rng(2)
    d = rng.random(40)
    d[7] = np.nan
    out = ao.boxsmooth_lag(d, 5, 3)
    ref = d.copy()
    for i in range(3, 40):
        w = d[0:i + 1] if i < 5 else d[i - 5:i + 1]
        ref[i] = np.nanmean(w)
    np.testing.assert_allclose(out, ref, rtol=1e-15, equal_nan=True)


def test_int16_abs_wrap_quirk():
    snd = np.array([-32768, 100, -200, 50], dtype=np.int16)
    pcm, fs = ao.normalise_pcm(snd, 44100)
    assert fs == 44100
    # np.abs(int16(-32768)) wraps, so the amplitude is 200 (reference AXCTDprocessor.py:56)
    assert abs(pcm[2] - (-200 - snd.mean()) / 200) < 1e-15
