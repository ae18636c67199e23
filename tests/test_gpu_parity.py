"""Parity tests proper: the CUDA engine, called through the C ABI, against
(a) fixtures produced by the unmodified reference (tests/golden), (b) the
oracle restatement on fresh seeded inputs, and (c) size-independent properties
at BASELINE.json's full sizes.  Bar: bits, bit edges, frames, CRC flags, chunk
chain and header metadata bit-exact; T/C/S/depth 1e-6 relative; signal levels
1e-4 relative (north_star)."""
import os

import numpy as np
import pytest

import synth
from golden_util import DECIM_CASES, FULL_CASES, SMALL_CASES, Golden
from parity_util import check_against_golden, check_against_oracle, check_rows_against_oracle, frames_view, mono, run_engine

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from axctdprocessor_b200 import engine
    e = engine.Engine(0)
    # recycled device blocks (the engine's block cache) come back filled with 0xA5 instead of whatever the last batch
    # left: no kernel may rely on freshly allocated memory being zero
    e.set_option("pool_poison", 1)
    yield e
    e.close()


def _engine(**opts):
    from axctdprocessor_b200 import engine
    e = engine.Engine(0)
    for k, v in opts.items():
        e.set_option(k, v)
    return e


@pytest.mark.parametrize("name", SMALL_CASES + DECIM_CASES)
def test_matches_reference_small(eng, name):
    g = Golden(name)
    out = run_engine(eng, mono(g.pcm()), g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange)
    check_against_golden(out, g)


@pytest.mark.parametrize("name", FULL_CASES)
def test_matches_reference_full_size(eng, name):
    """BASELINE configs 1 and 2 (720 s, 44.1 kHz, 40 dB / 10 dB) and the config-5 stand-in (1800 s at 8 dB, chunk
    4 x fs, dead frequency 2500 Hz, detuned mark / space)."""
    g = Golden(name)
    out = run_engine(eng, g.pcm(), g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange)
    check_against_golden(out, g)


@pytest.mark.parametrize("kw", [dict(fs=44100, duration_s=64.0, seed=101, snr_db=12.0),
                                dict(fs=48000, duration_s=58.0, seed=102, snr_db=18.0),
                                dict(fs=44100, duration_s=47.0, seed=103, snr_db=6.0, tone_after_pulse_s=31.0)])
def test_matches_oracle_on_fresh_seeds(eng, kw):
    from oracle import axctd_oracle as ao
    spec = synth.DropSpec(**kw)
    pcm = synth.generate_drop(spec)
    check_against_oracle(run_engine(eng, pcm, spec.fs), ao.process_pcm(pcm, spec.fs))


def test_bandpass_and_custom_settings_match_oracle(eng):
    from oracle import axctd_oracle as ao
    spec = synth.DropSpec(fs=48000, duration_s=55.0, seed=104, snr_db=35.0)
    pcm = synth.generate_drop(spec)
    st = {"usebandpass": True, "deadfreq": 2800.0, "refreshrate": 3.0}
    check_against_oracle(run_engine(eng, pcm, spec.fs, settings=st), ao.process_pcm(pcm, spec.fs, settings=st))


def test_reference_exceptions_are_reported(eng, tmp_path):
    """Digital silence after the pulse: the reference dies with IndexError at demodulate.py:85."""
    spec = synth.DropSpec(fs=44100, duration_s=20.0, seed=105, snr_db=40.0)
    pcm = synth.generate_drop(spec).copy()
    pcm[int(9.0 * 44100):] = 0
    out = run_engine(eng, pcm, spec.fs)
    from oracle import axctd_oracle as ao
    with pytest.raises(IndexError):
        ao.process_pcm(pcm, spec.fs)
    assert out["result"].status == 16
    with pytest.raises(IndexError):
        out["result"].raise_for_status()


def test_cli_output_file_is_byte_exact(tmp_path):
    from axctdprocessor_b200 import processAXCTD
    g = Golden("g44_40db")
    synth.write_wav(str(tmp_path / "g44_40db.wav"), g.pcm(), g.spec.fs)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        processAXCTD.main(["-i", "g44_40db.wav", "-o", "out.txt"])
    finally:
        os.chdir(cwd)
    assert (tmp_path / "out.txt").read_text() == g.meta["output_text"]


def test_cli_decimated_recording_is_byte_exact(tmp_path):
    """96 kHz recording: halved on the device (AXCTDprocessor.py:60-62); f_s prints as 48000.0."""
    from axctdprocessor_b200 import processAXCTD
    g = Golden("g96_decim")
    synth.write_wav(str(tmp_path / "g96_decim.wav"), g.pcm(), g.spec.fs)
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        processAXCTD.main(["-i", "g96_decim.wav", "-o", "out.txt"])
    finally:
        os.chdir(cwd)
    assert (tmp_path / "out.txt").read_text() == g.meta["output_text"]


def test_device_generator_is_bit_identical_to_numpy(eng):
    spec = synth.DropSpec(fs=44100, duration_s=30.0, seed=106, snr_db=10.0)
    ref = synth.generate_drop(spec)
    b = eng.batch([len(ref)], [eng.config(spec.fs)])
    b.synth_fill(0, spec)
    assert np.array_equal(b.download(0), ref)
    b.close()


def test_batch_equals_single_drops(eng):
    """Mixed 44.1 / 48 kHz batch: every drop decodes exactly as it does alone."""
    specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=46.0 + 3 * i, seed=200 + i, snr_db=10.0 + 5 * i) for i in range(6)]
    cfgs = [eng.config(s.fs) for s in specs]
    n = [int(round(s.duration_s * s.fs)) for s in specs]
    b = eng.batch(n, cfgs)
    for i, s in enumerate(specs):
        b.synth_fill(i, s)
    b.run()
    for i, s in enumerate(specs):
        pcm = b.download(i)
        single = run_engine(eng, pcm, s.fs)
        r = b.result(i)
        assert r.status == 0 and single["result"].status == 0
        assert np.array_equal(b.bits(i)[0], single["bits"][0])
        assert np.array_equal(b.edges(i)[0], single["edges"][0])
        assert np.array_equal(r.frames["word"], single["result"].frames["word"])
        assert np.array_equal(r.frames["keep"], single["result"].frames["keep"])
        np.testing.assert_array_equal(r.frames["temperature"], single["result"].frames["temperature"])
    b.close()


@pytest.mark.parametrize("opts", [dict(tone_direct=1), dict(force_exact=1), dict(inject_misspec=1), dict(tone_mma=10), dict(tone_mma=1), dict(tone_int8=0), dict(tone_complement=0), dict(fuse_bits=0), dict(fuse_bits=1, bit_tol=1e-2), dict(fuse_bits=1, bit_tol=0.9, expect_fallback=1),
                                  dict(segment_len=4096), dict(segment_len=32768), dict(filter_variant=1),
                                  dict(bitfix_all=1), dict(bit_tol=1e-3, hist_tol=1e-3), dict(ws=1), dict(ws=1, segment_len=4096), dict(fir_first=0), dict(fir_first=0, segment_len=4096), dict(tone_mma=0),
                                  dict(bulk=1), dict(bulk=1, fir_first=0), dict(bulk=1, segment_len=4096), dict(bulk=0)])
def test_kernel_variants_agree_with_reference(opts):
    g = Golden("g48_25db")
    opts = dict(opts)
    expect_fallback = opts.pop("expect_fallback", 0)
    e = _engine(**opts)
    out = run_engine(e, g.pcm(), g.spec.fs)
    if expect_fallback:             # nearly every bit asks for a double-precision window: the list overflows and the run is repeated in the two-step form
        assert out["result"].summary.n_recheck > g.meta["n_bits"] // 2
    if "inject_misspec" in opts:
        assert out["result"].summary.n_chain_fixups >= 1
    check_against_golden(out, g)
    if "bitfix_all" in opts:        # every window re-evaluated in double: conf agrees with the reference to fp64 round-off
        np.testing.assert_allclose(out["bits"][1], g.z["conf"], rtol=1e-9, equal_nan=True)
        assert out["result"].summary.n_recheck >= g.meta["n_bits"]
    e.close()


@pytest.mark.parametrize("pair", [0, 1])
def test_mixed_rate_batch_in_one_or_two_demodulation_launches(pair):
    """A batch holding both rate classes (44.1 and 48 kHz: window lengths 39 and 43): one launch per class, or both in
    one launch (k_demod_fused_pair, option pair_launch)."""
    ga, gb = Golden("g44_40db"), Golden("g48_25db")
    e = _engine(pair_launch=pair)
    pa, pb = ga.pcm(), gb.pcm()
    b = e.batch([len(pa), len(pb), len(pa)], [e.config(ga.spec.fs), e.config(gb.spec.fs), e.config(ga.spec.fs)])
    b.upload(0, pa); b.upload(1, pb); b.upload(2, pa)
    b.run()
    for i, g in enumerate((ga, gb, ga)):
        check_against_golden(dict(result=b.result(i), bits=b.bits(i), edges=b.edges(i), power=b.power(i)), g)
    b.close()
    e.close()


@pytest.mark.parametrize("opts", [dict(), dict(ws=1), dict(filter_variant=1)])
def test_guard_band_samples_are_settled_by_exact_recomputation(opts):
    """a filter output inside the guard band must not fail the drop.  With the guard widened to 3e-6 the
    fast passes flag dozens of samples; the ones inside demodulated iterations are re-signed in scipy's operation
    order from the iteration start and confirmed, the ones in the lead-in are ignored, and the decode stays exact."""
    g = Golden("g44_10db")
    e = _engine(guard=3e-6, **opts)
    out = run_engine(e, g.pcm(), g.spec.fs)
    s = out["result"].summary
    assert s.status == 0 and s.n_uncertain == 0
    assert s.n_guard_hits > 20 and 0 < s.n_guard_confirmed <= s.n_guard_hits
    check_against_golden(out, g)
    e.close()
    q = Golden("g44_nopulse")
    e = _engine(guard=1e-3, **opts)
    out = run_engine(e, q.pcm(), q.spec.fs)
    assert out["result"].summary.status == 0 and out["result"].summary.n_guard_hits > 0
    e.close()


@pytest.mark.parametrize("name", ["g44_bandpass", "g44_10db", "g48_chunk8", "g44_chunk05"])
@pytest.mark.parametrize("bulk", [0, 1])
def test_both_staging_forms_match_reference(name, bulk):
    """The continuous pass with its rows staged by cp.async.bulk (TMA unit, UBLKCP) and by LDGSTS: band-pass
    (six sections, reference-order cascade), low-pass at 10 dB, and both ends of the chunk-size sweep."""
    g = Golden(name)
    e = _engine(bulk=bulk)
    check_against_golden(run_engine(e, g.pcm(), g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange), g)
    e.close()


def test_multichannel_frames_are_deinterleaved_on_the_device(eng, tmp_path):
    """AXCTDprocessor.py:46-52 on the GPU: interleaved frames are uploaded as read and k_deinterleave picks the
    first channel (2 channels: the 16-byte path; 3 channels and a ragged tail: the generic path)."""
    from axctdprocessor_b200 import AXCTDprocessor
    g = Golden("g44_stereo")
    frames = g.pcm()
    check_against_golden(run_engine(eng, frames, g.spec.fs), g)
    three = np.ascontiguousarray(np.concatenate([frames, frames[:, 1:2] // 2], axis=1))
    check_against_golden(run_engine(eng, three, g.spec.fs), g)
    for cut in (1, 3, 8):
        b = eng.batch([len(frames) - cut], [eng.config(g.spec.fs)])
        b.upload(0, frames[:-cut])
        assert np.array_equal(b.download(0), frames[:-cut, 0])
        b.close()
    wav = tmp_path / "s.wav"
    synth.write_wav(str(wav), frames, g.spec.fs)
    ap = AXCTDprocessor.AXCTD_Processor(str(wav), engine=eng)
    assert ap.audiostream.ndim == 2
    ap.run()
    assert ap.hexframes == g.hexframes and ap.firstpulse400 == g.meta["firstpulse400"]


def test_full_size_round_trip_property(eng):
    """12-minute 48 kHz drop generated on the device (fresh seed, 30 dB): every
    decoded frame must be one of the transmitted frames, in order, with no gaps
    after the profile start (encode -> modulate -> demodulate -> decode)."""
    spec = synth.DropSpec(fs=48000, duration_s=720.0, seed=301, snr_db=30.0)
    n = int(round(spec.duration_s * spec.fs))
    b = eng.batch([n], [eng.config(spec.fs)])
    truth = b.synth_fill(0, spec)
    b.run()
    r = b.result(0)
    assert r.status == 0 and r.summary.n_uncertain == 0
    fr = r.frames
    assert len(fr) > 16000
    sent = (1 << 31) | (truth.data_frames[:, 0].astype(np.int64) << 18) | (truth.data_frames[:, 1].astype(np.int64) << 6)
    got = (fr["word"].astype(np.int64) >> 6) << 6
    # locate the first decoded frame in the transmitted sequence, then require a contiguous match
    start = int(np.flatnonzero(sent == got[0])[0])
    k = min(len(got), len(sent) - start)
    match = got[:k] == sent[start:start + k]
    assert match.mean() > 0.999, match.mean()
    assert abs(r.summary.firstpulse400 / spec.fs - spec.lead_in_s) < 0.2
    assert abs(r.summary.profstartind / spec.fs - (spec.lead_in_s + spec.tone_after_pulse_s)) < 0.2
    b.close()


def test_smoke_entry_point():
    import __graft_entry__ as ge
    ge.smoke()


def test_multi_drop_recording_is_cut_and_decoded_per_drop(eng):
    """BASELINE config 3 in miniature: several drops in one recording (also at 96 kHz, halved on the device).
    The segmentation driver cuts it from the engine's own 400 Hz level; every segment decodes exactly as
    the oracle decodes that segment on its own."""
    from axctdprocessor_b200 import segment
    from oracle import axctd_oracle as ao
    for fs, dec in ((44100, 1), (96000, 2)):
        specs = [synth.DropSpec(fs=fs, duration_s=52.0 + 3 * i, seed=500 + i, snr_db=30.0 - 8 * i) for i in range(3)]
        pcm = np.concatenate([synth.generate_drop(s) for s in specs])
        out = segment.process_recording(eng, pcm, fs / dec, decimate=dec)
        assert len(out) == 3
        for a, b, res in out:
            assert res.status == 0
            seg = pcm[a:b]
            full = run_engine(eng, seg, fs)
            op = ao.process_pcm(seg, fs)
            check_against_oracle(full, op)
            assert np.array_equal(res.rows["word"], full["result"].rows["word"])


@pytest.mark.parametrize("case", ["clipped", "too_short", "one_chunk"])
def test_edge_inputs_match_oracle(eng, case):
    """Ingest edge cases on the CUDA path: a sample at -32768 (np.abs wraps, AXCTDprocessor.py:56; exercises
    k_stats_wrap), a recording shorter than 4*N_power (:295) and one that holds a single iteration."""
    from oracle import axctd_oracle as ao
    spec = synth.DropSpec(fs=44100, duration_s=48.0, seed=77, snr_db=25.0)
    pcm = synth.generate_drop(spec).copy()
    if case == "clipped":
        pcm[1000] = -32768
        pcm[2000:2010] = -32768
    elif case == "too_short":
        pcm = pcm[:17000]
    else:
        pcm = pcm[:60000]
    out = run_engine(eng, pcm, spec.fs)
    op = ao.process_pcm(pcm, spec.fs)
    s = out["result"].summary
    assert s.pcm_ampl == int(np.max(np.abs(pcm)))
    assert s.pcm_sum == int(pcm.astype(np.int64).sum())
    assert s.n_chunks == len(op.trace)
    if case == "clipped":
        check_against_oracle(out, op)
    else:
        assert s.status == 0 and s.n_bits == 0 and s.n_frames == 0 and s.firstpulse400 == -1


def test_random_batch_matches_oracle(eng):
    """Twelve drops of mixed rate / SNR / length decoded as one batch, each compared with the oracle."""
    from oracle import axctd_oracle as ao
    rng = np.random.default_rng(2024)
    specs = [synth.DropSpec(fs=int(rng.choice([44100, 48000])), duration_s=float(rng.uniform(45.0, 70.0)), seed=700 + i,
                            snr_db=float(rng.uniform(6.0, 40.0)), tone_after_pulse_s=float(rng.uniform(30.5, 36.0)))
             for i in range(12)]
    pcms = [synth.generate_drop(s) for s in specs]
    b = eng.batch([len(p) for p in pcms], [eng.config(s.fs) for s in specs])
    for i, p in enumerate(pcms):
        b.upload(i, p)
    b.run()
    outs = [dict(result=b.result(i), bits=b.bits(i), edges=b.edges(i), power=b.power(i)) for i in range(len(specs))]
    b.close()
    for s, p, out in zip(specs, pcms, outs):
        check_against_oracle(out, ao.process_pcm(p, s.fs))


@pytest.mark.parametrize("kw,settings,trig", [
    (dict(fs=44100, duration_s=64.0, seed=77, snr_db=15.0), None, None),
    (dict(fs=48000, duration_s=75.0, seed=4242, snr_db=25.0), None, None),
    (dict(fs=48000, duration_s=56.0, seed=78, snr_db=25.0), {"usebandpass": True, "refreshrate": 1.0}, None),
    (dict(fs=44100, duration_s=64.0, seed=14, snr_db=25.0), None, [30, 41]),
])
def test_streaming_polls_match_oracle(eng, kw, settings, trig):
    """stream.StreamingDecoder on the device (axctd_batch_stream_*): the recording arrives in pieces of random
    length; every poll decodes only the iterations that became complete and its rows are held to the oracle's
    per-iteration lists on the prefix-normalised recording; the finished result is held to the oracle as a whole
    (bits, edges, chunk chain, frames, unrounded values)."""
    from parity_util import check_streaming_against_oracle
    spec = synth.DropSpec(**kw)
    check_streaming_against_oracle(eng, np.ascontiguousarray(synth.generate_drop(spec)), spec.fs, seed=spec.seed, settings=settings,
                                   triggerrange=trig)


def test_streaming_work_per_poll_does_not_grow_with_the_prefix(eng):
    """A 10-minute drop polled once a second: the filter pass of a poll covers the new second only, so its device
    time late in the recording stays what it was early on (a whole-prefix re-decode grows linearly), and the finished
    result equals the oracle's on the prefix-normalised recording."""
    from axctdprocessor_b200.stream import StreamingDecoder
    from parity_util import oracle_prefix_normalised
    spec = synth.DropSpec(fs=44100, duration_s=600.0, seed=4711, snr_db=25.0)
    g = eng.batch([int(round(spec.duration_s * spec.fs))], [eng.config(spec.fs)])
    g.synth_fill(0, spec)
    pcm = g.download(0)
    g.close()
    sd = StreamingDecoder(spec.fs, engine=eng, max_seconds=620.0, norm_seconds=2.0)
    step = spec.fs
    for a in range(0, len(pcm), step):
        sd.push(pcm[a:a + step])
        sd.poll()
    r = sd.finish()
    runs = sd.runs
    early = np.median([x["filter_ms"] for x in runs[60:120]])
    late = np.median([x["filter_ms"] for x in runs[-61:-1]])
    assert late < 2.0 * early + 0.05, (early, late)
    op, _ = oracle_prefix_normalised(pcm, spec.fs)
    check_rows_against_oracle(r, op)
    sd._own = False
    sd.close()


def test_concurrent_decoder_matches_single_batch(eng):
    """batch.ConcurrentDecoder (sub-batches on their own engines, streams and host threads) against one batch."""
    from axctdprocessor_b200 import batch as axbatch
    specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=50.0 + i, seed=7000 + i, snr_db=(40.0, 10.0)[i % 2]) for i in range(7)]
    pcms = [np.ascontiguousarray(synth.generate_drop(s)) for s in specs]
    ref = axbatch.process_drops(eng, pcms, [s.fs for s in specs])
    cd = axbatch.ConcurrentDecoder(0, [len(p) for p in pcms], [s.fs for s in specs], shards=3)
    for i, p in enumerate(pcms):
        cd.upload(i, p)
    cd.run(steps=3)
    got = cd.results(full=False)
    cd.close()
    for r, g in zip(ref, got):
        assert r.status == 0 and g.status == 0
        assert np.array_equal(r.rows, g.rows)


def _oracle_words_and_rows(args):
    from oracle import axctd_oracle as ao
    pcm, fs = args
    op = ao.process_pcm(pcm, fs)
    return (np.array([int(h, 16) for h in op.hexframes], dtype=np.uint32), np.asarray(op.time, dtype=np.float64),
            np.asarray(op.temperature, dtype=np.float64))


def test_config3_full_size_recording(eng):
    """BASELINE config 3 at full size: a 96 kHz, 1-hour recording holding five drops is halved on the device, cut by
    the segmentation driver and decoded as one batch; every segment's frames are the oracle's for that segment
    (frame words bit-exact, times and temperatures of the kept rows)."""
    from axctdprocessor_b200 import segment
    from oracle import axctd_oracle as ao
    fs = 96000
    specs = [synth.DropSpec(fs=fs, duration_s=720.0, seed=3300 + i, snr_db=(40.0, 25.0, 10.0)[i % 3]) for i in range(5)]
    n = [int(round(s.duration_s * s.fs)) for s in specs]
    gen = eng.batch(n, [eng.config(fs)] * len(n))          # the device twin of synth.generate_drop as a generator
    parts = []
    for i, s in enumerate(specs):
        gen.synth_fill(i, s)
        parts.append(gen.download(i))
    gen.close()
    pcm = np.concatenate(parts)
    del parts
    out = segment.process_recording(eng, pcm, fs / 2, decimate=2)
    assert len(out) == 5 and [a for a, _, _ in out][0] == 0
    for a, b, r in out:
        assert r.status == 0 and 16900 < int(r.summary.n_frames) < 17100
    # every segment against the oracle run on that segment as a stand-alone recording (one process per segment)
    import multiprocessing as mp
    with mp.get_context("fork").Pool(processes=len(out)) as pool:
        ops = pool.map(_oracle_words_and_rows, [(pcm[a:b], fs) for a, b, _ in out])
    for (a, b, r), (words, times, temps) in zip(out, ops):
        tab = r.table()
        assert np.array_equal(words, tab["word"][tab["hex_returned"] == 1]), (a, b)
        kept = tab[tab["keep"] == 1]
        np.testing.assert_allclose(kept["time_s"], times, rtol=1e-6, atol=0)
        np.testing.assert_allclose(kept["temperature"], temps, rtol=1e-6, atol=0)


def test_calibration_known_answers_on_the_device(eng):
    """k_calib's arithmetic (ax_sp_from_c, ax_dataconvert) on the GPU against the GSW documentation's check values
    for gsw_SP_from_C (the call parse.py:132 makes) and parse.dataconvert's summation order."""
    from test_oracle_units import GSW_C, GSW_P, GSW_SP, GSW_T
    from oracle import axctd_oracle as ao, pss78
    cf = [-0.0622192, 1.04584, 3.0e-5, -2.0e-7]
    sp, poly = eng.calib_eval(GSW_C, GSW_T, GSW_P, coeff=cf)
    np.testing.assert_allclose(sp, GSW_SP, rtol=1e-13, atol=0)
    np.testing.assert_allclose(poly, [ao.dataconvert(c, cf) for c in GSW_C], rtol=1e-15, atol=0)
    # a sweep over the oceanographic range incl. the Hill (SP < 2) branch and invalid input, against oracle/pss78.py
    rng = np.random.default_rng(7)
    c = np.concatenate([rng.uniform(0.0, 70.0, 4000), rng.uniform(0.0, 3.0, 1000), [-1.0, 0.0]])
    t = np.concatenate([rng.uniform(-2.0, 35.0, 5000), [10.0, 10.0]])
    p = np.concatenate([rng.uniform(0.0, 2000.0, 5000), [0.0, 0.0]])
    sp, _ = eng.calib_eval(c, t, p)
    np.testing.assert_allclose(sp, pss78.SP_from_C(c, t, p), rtol=1e-12, atol=1e-13, equal_nan=True)


def test_pipelined_decoder_matches_oracle(eng):
    """batch.PipelinedDecoder -- the ingest path bench.py's e2e figure goes through (pinned host PCM -> alternating
    engines -> compact rows) -- held to the oracle drop by drop, and to a reference fixture."""
    import torch
    from axctdprocessor_b200 import batch as axbatch
    from oracle import axctd_oracle as ao
    g = Golden("g44_10db")
    specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=47.0 + 2 * i, seed=8100 + i, snr_db=(35.0, 12.0, 22.0)[i % 3]) for i in range(5)]
    pcms = [np.ascontiguousarray(synth.generate_drop(s)) for s in specs] + [np.ascontiguousarray(g.pcm())]
    fss = [s.fs for s in specs] + [g.spec.fs]
    pinned = []
    for p in pcms:
        t = torch.empty(len(p), dtype=torch.int16).pin_memory()
        t.numpy()[:] = p
        pinned.append(t)
    pipe = axbatch.PipelinedDecoder(0, slots=2)
    groups = [[0, 1], [2, 3], [4, 5]]
    got = {}

    def submit(h):
        pipe.submit([pinned[i].data_ptr() for i in h], [len(pcms[i]) for i in h], [fss[i] for i in h])

    for rep in range(2):                                   # the second pass reuses the cached batches
        submit(groups[0])
        for q, h in enumerate(groups):
            if q + 1 < len(groups):
                submit(groups[q + 1])
            for i, r in zip(h, pipe.collect(full=False)):
                got[(rep, i)] = r
    pipe.close()
    ops = [ao.process_pcm(p, fs) for p, fs in zip(pcms, fss)]
    for rep in range(2):
        for i, op in enumerate(ops):
            check_rows_against_oracle(got[(rep, i)], op)
    r = got[(1, 5)]
    tab = r.table()
    assert ["%08x" % int(w) for w in tab["word"][tab["hex_returned"] == 1]] == g.hexframes
    np.testing.assert_allclose(tab["temperature"][tab["keep"] == 1], g.z["temperature"], rtol=1e-6)


def test_concurrent_decoder_matches_oracle(eng):
    """batch.ConcurrentDecoder (sub-batches in flight on their own engines / streams / host threads: the path
    bench.py's device-resident figure goes through) held to the oracle drop by drop."""
    from axctdprocessor_b200 import batch as axbatch
    from oracle import axctd_oracle as ao
    specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=50.0 + i, seed=7100 + i, snr_db=(40.0, 10.0, 20.0)[i % 3]) for i in range(7)]
    pcms = [np.ascontiguousarray(synth.generate_drop(s)) for s in specs]
    cd = axbatch.ConcurrentDecoder(0, [len(p) for p in pcms], [s.fs for s in specs], shards=4)
    for i, p in enumerate(pcms):
        cd.upload(i, p)
    cd.run(steps=2)
    got = cd.results(full=False)
    cd.close()
    for s, p, r in zip(specs, pcms, got):
        check_rows_against_oracle(r, ao.process_pcm(p, s.fs))


def _shard_worker(rank, world, port, q, devices):
    import torch.distributed as dist
    sys_path = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import sys
    sys.path.insert(0, sys_path)
    sys.path.insert(0, os.path.join(sys_path, "tests"))
    import synth as S
    from axctdprocessor_b200 import batch as B, engine as E
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    specs = _shard_specs()
    sizes = [int(round(s.duration_s * s.fs)) for s in specs]
    mine = B.partition_drops(sizes, world)[rank]
    eng = E.Engine(devices[rank])
    pcms = [np.ascontiguousarray(S.generate_drop(specs[i])) for i in mine]
    res = B.process_drops(eng, pcms, [specs[i].fs for i in mine])
    local = {i: dict(status=r.status, rows=r.rows.copy(), chunks=r.chunks.copy(), firstpulse400=int(r.summary.firstpulse400),
                     profstartind=int(r.summary.profstartind), scale=float(r.summary.high_bit_scale),
                     n_bits=int(r.summary.n_bits), n_edges=int(r.summary.n_edges)) for i, r in zip(mine, res)}
    merged = B.gather_results(local, world, rank)
    dist.barrier()
    dist.destroy_process_group()
    eng.close()
    if rank == 0:
        q.put(merged)


def _shard_specs():
    return [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=46.0 + 4 * (i % 3), seed=9100 + i, snr_db=(30.0, 14.0)[i % 2]) for i in range(6)]


def test_two_rank_sharded_decode_equals_oracle_and_single_gpu(eng):
    """SURVEY 4.3 item 5 on hardware: batch.partition_drops + gather_results with one process per rank (both on
    the visible GPUs; a second GPU is used when the box has one) -- every drop of the merged result equals the
    oracle's decode and the single-process decode of the same drop."""
    import torch
    import torch.multiprocessing as mp
    from axctdprocessor_b200 import batch as axbatch
    from oracle import axctd_oracle as ao
    ndev = torch.cuda.device_count()
    devices = [0, 1 if ndev > 1 else 0]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_shard_worker, args=(r, 2, port, q, devices)) for r in range(2)]
    for p in procs:
        p.start()
    merged = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    specs = _shard_specs()
    assert sorted(merged) == list(range(len(specs)))
    pcms = [np.ascontiguousarray(synth.generate_drop(s)) for s in specs]
    single = axbatch.process_drops(eng, pcms, [s.fs for s in specs])
    for i, (s, p) in enumerate(zip(specs, pcms)):
        m = merged[i]
        assert m["status"] == 0
        assert np.array_equal(m["rows"], single[i].rows) and np.array_equal(m["chunks"], single[i].chunks)
        check_rows_against_oracle(single[i], ao.process_pcm(p, s.fs))
        assert (m["firstpulse400"], m["profstartind"], m["n_bits"]) == (int(single[i].summary.firstpulse400), int(single[i].summary.profstartind), int(single[i].summary.n_bits))


def test_engines_on_two_devices_in_one_process():
    """Two engines on two GPUs of one process, used alternately (per-device shared-memory opt-in, device guards of
    every ABI entry point, per-device block cache): both decode to the oracle's result."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    from axctdprocessor_b200 import engine
    from oracle import axctd_oracle as ao
    spec = synth.DropSpec(fs=44100, duration_s=52.0, seed=616, snr_db=18.0)
    pcm = synth.generate_drop(spec)
    op = ao.process_pcm(pcm, spec.fs)
    e0, e1 = engine.Engine(0), engine.Engine(1)
    try:
        c0, c1 = e0.config(spec.fs), e1.config(spec.fs)          # interleaved on purpose: each call must land on its own device
        b1 = e1.batch([len(pcm)], [c1]); b0 = e0.batch([len(pcm)], [c0])
        b0.upload(0, pcm); b1.upload(0, pcm)
        b1.run(); b0.run()
        for b in (b0, b1):
            out = dict(result=b.result(0), bits=b.bits(0), edges=b.edges(0), power=b.power(0))
            check_against_oracle(out, op)
        b0.close(); b1.close()
        # recycled blocks stay on their device
        b0 = e0.batch([len(pcm)], [c0]); b0.upload(0, pcm); b0.run()
        check_rows_against_oracle(b0.result(0, full=False), op)
        b0.close()
    finally:
        e0.close(); e1.close()


def test_random_parameter_sweep_matches_oracle(eng):
    """BASELINE config 5 in miniature: sixteen drops decoded as ONE batch, every one with its own randomly drawn
    parameter point (chunk 0.5-8 s, low-pass / band-pass, dead frequency, mark / space pair with a matching
    transmitter, detector thresholds, trigger window) -- each compared with the oracle run with the same settings,
    including the cases in which the reference raises."""
    from axctdprocessor_b200 import engine as axengine
    from oracle import axctd_oracle as ao
    rng = np.random.default_rng(55)
    cases = []
    for i in range(16):
        ms = [(400, 800), (405, 795), (420, 780)][int(rng.integers(0, 3))]
        st = {"refreshrate": float(rng.choice([0.5, 0.75, 1.0, 2.0, 3.0, 4.0, 8.0])), "usebandpass": bool(rng.integers(0, 2)),
              "deadfreq": float(rng.choice([2500.0, 2800.0, 3000.0, 3500.0])), "mark_space_freqs": [float(ms[0]), float(ms[1])],
              "minr400": float(rng.choice([1.5, 2.0, 2.5])), "mindr7500": float(rng.choice([1.0, 1.5, 2.0]))}
        trig = [[30, -1], [30, 40], [32, -1]][int(rng.integers(0, 3))]
        spec = synth.DropSpec(fs=int(rng.choice([44100, 48000])), duration_s=float(rng.uniform(50.0, 72.0)), seed=5600 + i,
                              snr_db=float(rng.uniform(8.0, 35.0)), mark_hz=ms[0], space_hz=ms[1],
                              tone_after_pulse_s=float(rng.uniform(33.0, 37.0)))
        cases.append((spec, st, trig))
    pcms = [synth.generate_drop(s) for s, _, _ in cases]
    cfgs = [eng.config(s.fs, settings=st, triggerrange=trig) for s, st, trig in cases]
    b = eng.batch([len(p) for p in pcms], cfgs)
    for i, p in enumerate(pcms):
        b.upload(i, p)
    b.run()
    outs = [dict(result=b.result(i), bits=b.bits(i), edges=b.edges(i), power=b.power(i)) for i in range(len(cases))]
    b.close()
    n_ok = 0
    for (spec, st, trig), p, out in zip(cases, pcms, outs):
        try:
            op = ao.process_pcm(p, spec.fs, settings=st, triggerrange=trig)
        except Exception as exc:                       # the reference crashes on this point: the engine reports the same exception
            code = out["result"].status
            assert code != 0, (st, trig, type(exc))
            assert issubclass(axengine.STATUS_EXCEPTIONS[code][0], type(exc)) or issubclass(type(exc), axengine.STATUS_EXCEPTIONS[code][0]), (code, type(exc))
            continue
        check_against_oracle(out, op)
        n_ok += 1
    assert n_ok >= 10
