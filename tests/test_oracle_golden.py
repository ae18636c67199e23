"""The numpy restatement (oracle/) against outputs of the UNMODIFIED reference
(tests/golden, made by oracle/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_util import DECIM_CASES, FULL_CASES, SMALL_CASES, WIDE_CASES, Golden
from oracle import axctd_oracle as ao


def _run_oracle(g):
    return ao.process_pcm(g.pcm(), g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange)


def _check(g, op, full=True):
    m = g.meta
    assert op.firstpulse400 == m["firstpulse400"]
    assert op.profstartind == m["profstartind"]
    assert op.numpoints == m["numpoints"] and float(op.f_s) == m["f_s"]
    assert abs(op.high_bit_scale - m["high_bit_scale"]) <= 1e-12 * m["high_bit_scale"]
    # discrete outputs: exact
    assert np.array_equal(np.asarray(op.all_bits, dtype=np.uint8), g.bits)
    assert np.array_equal(np.asarray(op.all_edges, dtype=np.int64), g.edges)
    assert op.hexframes == g.hexframes
    tr = g.trace()
    assert len(tr) == len(op.trace)
    for a, b in zip(op.trace, tr):
        for k in ("s", "e", "status", "n_power", "nrows", "nhex"):
            assert a[k] == b[k], (k, a, b)
        if b["nbits"] >= 0:
            for k in ("nbits", "first_edge", "last_edge"):
                assert a[k] == b[k], (k, a, b)
    md = {k: v for k, v in op.metadata.items()}
    for k, v in m["metadata"].items():
        assert md[k] == v, k
    # calibrated values: the reference rounds to 2 dp, so equality is expected
    for k in ("time", "depth", "temperature", "conductivity", "salinity", "r400_prof", "r7500_prof"):
        a = np.asarray(getattr(op, k), dtype=np.float64)
        assert a.shape == g.z[k].shape, k
        np.testing.assert_allclose(a, g.z[k], rtol=1e-6, atol=0, equal_nan=True, err_msg=k)
    np.testing.assert_allclose(np.asarray(op.r400), g.z["r400"].astype(np.float64), rtol=1e-4 if not full else 1e-9, atol=1e-6 if not full else 1e-11, equal_nan=True)
    if "conf" in g.z.files:
        np.testing.assert_allclose(np.asarray(op.all_conf), g.z["conf"], rtol=1e-10, equal_nan=True)
    if "output_text" in m:
        settings_cli = {"minR400": 2.0, "mindR7500": 1.5, "deadfreq": 3000.0, "pointsperloop": 100000,
                        "triggerrange": [30, -1], "mark_space_freqs": [400.0, 800.0], "use_bandpass": False}
        assert ao.format_output(op, g.name + ".wav", [0, -1], settings_cli) == m["output_text"]
    else:
        with pytest.raises(KeyError):
            ao.format_output(op, g.name + ".wav", [0, -1], {"minR400": 2.0, "mindR7500": 1.5, "deadfreq": 3000.0,
                                                            "pointsperloop": 1, "triggerrange": [30, -1]})


@pytest.mark.parametrize("name", SMALL_CASES + DECIM_CASES)
def test_oracle_matches_reference_small(name):
    g = Golden(name)
    _check(g, _run_oracle(g))


@pytest.mark.parametrize("name", WIDE_CASES)
def test_oracle_matches_reference_on_wide_samples(name):
    """24-bit PCM / IEEE float WAV files (int32 / float32 arrays from scipy.io.wavfile.read): the oracle's
    normalisation is the reference's numpy expression for any sample type (AXCTDprocessor.py:55-62)."""
    g = Golden(name)
    wide, _ = g.wide()
    _check(g, ao.process_pcm(wide, g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange))


@pytest.mark.parametrize("name", [FULL_CASES[0], FULL_CASES[2]])
def test_oracle_matches_reference_full_size(name):
    g = Golden(name)
    _check(g, _run_oracle(g), full=False)
