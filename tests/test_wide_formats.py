"""Recordings whose samples are not 16-bit integers (24 / 32-bit PCM, IEEE float): scipy.io.wavfile.read -- and so
the reference, AXCTDprocessor.py:41 -- accepts them.  The host normalises (and halves) them exactly as
AXCTDprocessor.py:55-62 does and the engine takes the double-precision signal (axctd_batch_upload_f64, config
decimate = 3).  Fixtures: the unmodified reference run on such WAV files (oracle/make_golden.py, WIDE)."""
import os

import numpy as np
import pytest

import synth
from golden_util import WIDE_CASES, Golden
from parity_util import check_against_golden


def run_engine_wide(eng, samples, fs, settings=None, triggerrange=None):
    from axctdprocessor_b200.AXCTDprocessor import normalised_signal
    pcm, fs2 = normalised_signal(samples, fs)
    cfg = eng.config(fs2, settings=settings, triggerrange=triggerrange, decimate=3)
    b = eng.batch([len(pcm)], [cfg])
    b.upload(0, pcm)
    b.run()
    out = dict(result=b.result(0), bits=b.bits(0), edges=b.edges(0), power=b.power(0), timing=b.timing())
    b.close()
    return out


@pytest.mark.parametrize("fmt", synth.WIDE_FORMATS)
@pytest.mark.parametrize("channels", [1, 2])
def test_reader_returns_what_scipy_returns(tmp_path, fmt, channels):
    from scipy.io import wavfile
    from axctdprocessor_b200.AXCTDprocessor import read_wav
    pcm = synth.generate_drop(synth.DropSpec(fs=44100, duration_s=2.0, seed=3, lead_in_s=0.2, channels=channels))
    wide = synth.widen(pcm, fmt, 3)
    path = str(tmp_path / "w.wav")
    synth.write_wav_wide(path, wide, 44100, fmt)
    fs_ref, a_ref = wavfile.read(path)
    fs, a = read_wav(path)
    assert fs == fs_ref and a.dtype == a_ref.dtype and a.shape == a_ref.shape
    assert np.array_equal(a, a_ref)


def test_reader_rejects_unknown_formats(tmp_path):
    import struct
    from axctdprocessor_b200.AXCTDprocessor import read_wav
    path = str(tmp_path / "x.wav")
    body = b"\0" * 64
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", 36 + len(body)) + b"WAVE")
        f.write(b"fmt " + struct.pack("<IHHIIHH", 16, 6, 1, 8000, 8000, 1, 8))      # A-law: scipy refuses it too
        f.write(b"data" + struct.pack("<I", len(body)) + body)
    with pytest.raises(ValueError):
        read_wav(path)


@pytest.mark.parametrize("name", WIDE_CASES)
def test_emulated_engine_matches_reference_on_wide_samples(name):
    from emu_util import emu_engine
    g = Golden(name)
    wide, fmt = g.wide()
    e = emu_engine()
    try:
        check_against_golden(run_engine_wide(e, wide, g.spec.fs), g)
    finally:
        e.close()


def _check_processor(ap, g):
    assert ap.hexframes == g.hexframes
    assert ap.numpoints == g.meta["numpoints"] and float(ap.f_s) == g.meta["f_s"]
    assert ap.firstpulse400 == g.meta["firstpulse400"] and ap.profstartind == g.meta["profstartind"]
    for k in ("time", "depth", "temperature", "conductivity", "salinity"):
        np.testing.assert_allclose(np.asarray(getattr(ap, k)), g.z[k], rtol=1e-6, atol=0, err_msg=k)


def test_processor_class_reads_a_24_bit_file_through_the_emulation(tmp_path):
    """The drop-in class end to end on a 24-bit WAV file (reader, host normalisation, upload, decode)."""
    from emu_util import emu_engine
    from axctdprocessor_b200 import AXCTDprocessor as axp
    g = Golden("g44_pcm24")
    wide, fmt = g.wide()
    path = str(tmp_path / "g44_pcm24.wav")
    synth.write_wav_wide(path, wide, g.spec.fs, fmt)
    e = emu_engine()
    try:
        ap = axp.AXCTD_Processor(path, engine=e)
        assert ap.audiostream.dtype == np.float64 and abs(np.max(np.abs(ap.audiostream)) - 1.0) < 1e-3
        ap.run()
        _check_processor(ap, g)
    finally:
        e.close()


def test_int16_upload_is_refused_for_a_double_precision_drop():
    from emu_util import emu_engine
    e = emu_engine()
    try:
        cfg = e.config(44100, decimate=3)
        b = e.batch([50000], [cfg])
        with pytest.raises(RuntimeError):
            b.upload(0, np.zeros(50000, dtype=np.int16))
        b.close()
    finally:
        e.close()


# ---------------------------------------------------------------- the same on the device
@pytest.fixture(scope="module")
def eng():
    from axctdprocessor_b200 import engine
    e = engine.Engine(0)
    e.set_option("pool_poison", 1)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("name", WIDE_CASES)
def test_gpu_engine_matches_reference_on_wide_samples(eng, name):
    g = Golden(name)
    wide, fmt = g.wide()
    check_against_golden(run_engine_wide(eng, wide, g.spec.fs), g)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["g48_float32", "g96_pcm24_decim"])
def test_gpu_processor_class_on_wide_files(eng, tmp_path, name):
    from axctdprocessor_b200 import AXCTDprocessor as axp
    g = Golden(name)
    wide, fmt = g.wide()
    path = str(tmp_path / (name + ".wav"))
    synth.write_wav_wide(path, wide, g.spec.fs, fmt)
    ap = axp.AXCTD_Processor(path, engine=eng)
    ap.run()
    _check_processor(ap, g)


@pytest.mark.gpu
def test_gpu_batch_mixes_int16_and_double_precision_drops(eng):
    """One batch, three kinds of drops: int16, int16 above 50 kHz (halved on the device), double precision."""
    from axctdprocessor_b200.AXCTDprocessor import normalised_signal
    from parity_util import run_engine
    ga, gb, gc = Golden("g44_40db"), Golden("g96_decim"), Golden("g48_float32")
    wide, _ = gc.wide()
    sig, fsc = normalised_signal(wide, gc.spec.fs)
    pa, pb = ga.pcm(), gb.pcm()
    cfgs = [eng.config(ga.spec.fs), eng.config(gb.spec.fs / 2, decimate=2), eng.config(fsc, decimate=3)]
    b = eng.batch([len(pa), len(pb), len(sig)], cfgs)
    b.upload(0, pa); b.upload(1, pb); b.upload(2, sig)
    b.run()
    for i, g in enumerate((ga, gb, gc)):
        out = dict(result=b.result(i), bits=b.bits(i), edges=b.edges(i), power=b.power(i))
        check_against_golden(out, g)
    b.close()


def test_engine_close_releases_batches_that_are_still_open():
    """A batch left open (an exception between batch() and close()) must not outlive its engine: Engine.close()
    destroys it, and the later Batch.close() / garbage collection is a no-op."""
    from emu_util import emu_engine
    e = emu_engine()
    b = e.batch([50000], [e.config(44100)])
    e.close()
    assert b.h is None
    b.close()


@pytest.mark.parametrize("bits,tag", [(16, 1), (24, 1), (32, 3)])
def test_reader_handles_extensible_headers_and_extra_chunks(tmp_path, bits, tag):
    """WAVE_FORMAT_EXTENSIBLE (the sub-format carries the real tag), a LIST chunk before the data and an odd-sized
    chunk with its pad byte: the reader walks the chunks as scipy does."""
    import struct
    from scipy.io import wavfile
    from axctdprocessor_b200.AXCTDprocessor import read_wav
    rng = np.random.default_rng(bits)
    nch, fs, n = 2, 48000, 1001
    bps = bits // 8
    if tag == 3:
        payload = rng.uniform(-1, 1, size=(n, nch)).astype("<f4").tobytes()
    elif bits == 16:
        payload = rng.integers(-32768, 32768, size=(n, nch)).astype("<i2").tobytes()
    else:
        payload = rng.integers(0, 256, size=n * nch * 3, dtype=np.uint8).tobytes()
    guid_tail = bytes.fromhex("000000001000800000aa00389b71")
    fmt = struct.pack("<HHIIHH", 0xFFFE, nch, fs, fs * nch * bps, nch * bps, bits) + struct.pack("<HHI", 22, bits, 3) + struct.pack("<H", tag) + guid_tail
    info = b"INFOISFT" + struct.pack("<I", 5) + b"test\0" + b"\0"      # odd-sized sub-chunk + pad byte
    body = (b"WAVE" + b"fmt " + struct.pack("<I", len(fmt)) + fmt + b"LIST" + struct.pack("<I", len(info)) + info
            + b"data" + struct.pack("<I", len(payload)) + payload)
    path = str(tmp_path / "ext.wav")
    with open(path, "wb") as f:
        f.write(b"RIFF" + struct.pack("<I", len(body)) + body)
    fs_ref, a_ref = wavfile.read(path)
    fs_got, a = read_wav(path)
    assert fs_got == fs_ref and a.dtype == a_ref.dtype and a.shape == a_ref.shape and np.array_equal(a, a_ref)
