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
