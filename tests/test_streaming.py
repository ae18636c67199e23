"""The incremental decoder (axctd_batch_stream_* behind axctdprocessor_b200.stream.StreamingDecoder) on the TEST-ONLY
host emulation of the kernel bodies: every poll is held to the oracle's per-iteration lists on the prefix-normalised
recording.  The GPU version of the same check is tests/test_gpu_parity.py::test_streaming_polls_match_oracle."""
import numpy as np
import pytest

import synth
from emu_util import emu_engine
from parity_util import check_streaming_against_oracle


@pytest.fixture(scope="module")
def eng():
    e = emu_engine()
    yield e
    e.close()


@pytest.mark.parametrize("kw,settings,trig", [
    (dict(fs=44100, duration_s=64.0, seed=77, snr_db=15.0), None, None),
    (dict(fs=48000, duration_s=56.0, seed=78, snr_db=25.0), {"usebandpass": True, "refreshrate": 1.0}, None),
    (dict(fs=44100, duration_s=64.0, seed=14, snr_db=25.0), None, [30, 41]),      # latest-trigger branch re-fires after status 2
])
def test_streaming_polls_match_oracle_emulated(eng, kw, settings, trig):
    spec = synth.DropSpec(**kw)
    check_streaming_against_oracle(eng, synth.generate_drop(spec), spec.fs, seed=spec.seed, settings=settings, triggerrange=trig)


def test_streaming_batch_refuses_whole_file_calls(eng):
    b = eng.batch([44100 * 10], [eng.config(44100)])
    b.stream_begin(0.0, 1000.0)
    with pytest.raises(RuntimeError):
        b.upload(0, np.zeros(44100 * 10, dtype=np.int16))
    with pytest.raises(RuntimeError):
        b.run()
    with pytest.raises(RuntimeError):
        b.stream_append(0, np.zeros(44100 * 11, dtype=np.int16))
    b.close()


def test_recordings_above_50khz_are_not_streamed(eng):
    from axctdprocessor_b200.stream import StreamingDecoder
    with pytest.raises(ValueError):
        StreamingDecoder(96000, engine=eng)
    b = eng.batch([96000 * 10], [eng.config(48000, decimate=2)])
    with pytest.raises(RuntimeError):
        b.stream_begin(0.0, 1000.0)
    b.close()


def test_streaming_recording_without_a_pulse(eng):
    """Noise only: the fixed grid grows poll by poll, no iteration is ever demodulated, and the finished result is the
    oracle's (status 0 throughout, no rows) -- AXCTDprocessor.py:332-333."""
    from parity_util import oracle_prefix_normalised
    from axctdprocessor_b200.stream import StreamingDecoder
    rng = np.random.default_rng(5)
    fs = 44100
    pcm = np.clip(rng.normal(0.0, 900.0, 20 * fs), -32768, 32767).astype(np.int16)
    op, _ = oracle_prefix_normalised(pcm, fs)
    sd = StreamingDecoder(fs, engine=eng, max_seconds=25.0)
    for a in range(0, len(pcm), fs // 3):
        sd.push(pcm[a:a + fs // 3])
        assert sd.poll() is None
    r = sd.finish()
    assert r.status == 0 and r.summary.firstpulse400 == -1 == op.firstpulse400 and len(r.rows) == 0 and len(op.time) == 0
    assert len(r.chunks) == len(op.trace)
    for c, t in zip(r.chunks, op.trace):
        assert (c["s"], c["e"], c["status"]) == (t["s"], t["e"], t["status"])
    p, r400, _ = sd.batch.power(0)
    assert np.array_equal(p, np.asarray(op.power_inds))
    np.testing.assert_allclose(r400, op.r400, rtol=1e-9, atol=1e-11)
    sd._own = False
    sd.close()


def test_streaming_in_tiny_pieces_and_early_finish(eng):
    """100 ms pieces (most polls close no iteration), and a recording that ends before the normalisation window is
    full: finish() takes the normalisation from what there is."""
    from parity_util import check_streaming_against_oracle, oracle_prefix_normalised
    from axctdprocessor_b200.stream import StreamingDecoder
    spec = synth.DropSpec(fs=44100, duration_s=50.0, seed=91, snr_db=20.0)
    pcm = synth.generate_drop(spec)
    check_streaming_against_oracle(eng, pcm, spec.fs, seed=1, piece_s=(0.08, 0.12))
    short = pcm[:int(1.2 * spec.fs)]
    sd = StreamingDecoder(spec.fs, engine=eng, max_seconds=10.0, norm_seconds=2.0)
    sd.push(short)
    assert sd.poll() is None
    r = sd.finish()
    op, _ = oracle_prefix_normalised(short, spec.fs, norm_seconds=2.0)
    assert r.status == 0 and len(r.chunks) == len(op.trace) and len(r.rows) == 0
    sd._own = False
    sd.close()
