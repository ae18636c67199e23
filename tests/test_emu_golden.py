"""Host logic + kernel bodies (TEST-ONLY host emulation, tests/emu/README.md)
against fixtures produced by the unmodified reference.  CPU only; the GPU
versions of these checks are in test_gpu_parity.py."""
import os

import numpy as np
import pytest

from emu_util import emu_engine
from golden_util import DECIM_CASES, SMALL_CASES, Golden
from parity_util import check_against_golden, mono, run_engine


@pytest.fixture(scope="module")
def eng():
    e = emu_engine()
    yield e
    e.close()


@pytest.mark.parametrize("name", SMALL_CASES + DECIM_CASES)
def test_emulated_engine_matches_reference(eng, name):
    g = Golden(name)
    out = run_engine(eng, mono(g.pcm()), g.spec.fs, settings=g.user_settings, triggerrange=g.triggerrange)
    check_against_golden(out, g)


def test_emulated_engine_full_size_config1(eng):
    g = Golden("config1_720s")
    out = run_engine(eng, g.pcm(), g.spec.fs)
    check_against_golden(out, g)


def test_chain_repair_loop_converges_to_same_answer():
    g = Golden("g44_10db")
    e = emu_engine(inject_misspec=1)
    out = run_engine(e, g.pcm(), g.spec.fs)
    assert out["result"].summary.n_chain_fixups >= 1
    check_against_golden(out, g)
    e.close()


@pytest.mark.parametrize("opts", [dict(force_exact=1), dict(segment_len=4096), dict(exact_head=2048, segment_len=8192),
                                  dict(bit_tol=1e-3, hist_tol=1e-3)])
def test_decomposition_invariance(opts):
    g = Golden("g48_25db")
    e = emu_engine(**opts)
    check_against_golden(run_engine(e, g.pcm(), g.spec.fs), g)
    e.close()


def test_double_precision_windows_reproduce_reference_confidence():
    """bitfix_all: every mark/space window comes from ax_gwin_* (double, straight from the int16
    samples, truncated at the chunk start): conf then agrees with the reference to fp64 round-off."""
    g = Golden("g44_10db")
    e = emu_engine(bitfix_all=1)
    out = run_engine(e, g.pcm(), g.spec.fs)
    check_against_golden(out, g)
    np.testing.assert_allclose(out["bits"][1], g.z["conf"], rtol=1e-9, equal_nan=True)
    assert out["result"].summary.n_recheck >= g.meta["n_bits"]
    e.close()


def test_device_generator_twin_matches_numpy(eng):
    import synth
    spec = synth.DropSpec(fs=48000, duration_s=20.0, seed=31, snr_db=15.0)
    ref = synth.generate_drop(spec)
    b = eng.batch([len(ref)], [eng.config(spec.fs)])
    b.synth_fill(0, spec)
    assert np.array_equal(b.download(0), ref)
    b.close()


def test_cli_mirror_writes_reference_output_file(eng, tmp_path, capsys):
    import synth
    from axctdprocessor_b200 import processAXCTD
    g = Golden("g44_40db")
    wav = tmp_path / "g44_40db.wav"
    synth.write_wav(str(wav), g.pcm(), g.spec.fs)
    out = tmp_path / "out.txt"
    settings = {"triggerrange": [30, -1], "minR400": 2.0, "mindR7500": 1.5, "deadfreq": 3000.0, "pointsperloop": 100000,
                "mark_space_freqs": [400.0, 800.0], "use_bandpass": False}
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        processAXCTD.processAXCTD("g44_40db.wav", str(out), [0, -1], settings, engine=eng)
    finally:
        os.chdir(cwd)
    assert out.read_text() == g.meta["output_text"]
    assert "Processing profile" in capsys.readouterr().out


def test_cli_mirror_decimates_recordings_above_50khz(eng, tmp_path):
    import synth
    from axctdprocessor_b200 import processAXCTD
    g = Golden("g96_decim")
    synth.write_wav(str(tmp_path / "g96_decim.wav"), g.pcm(), g.spec.fs)
    settings = {"triggerrange": [30, -1], "minR400": 2.0, "mindR7500": 1.5, "deadfreq": 3000.0, "pointsperloop": 100000,
                "mark_space_freqs": [400.0, 800.0], "use_bandpass": False}
    cwd = os.getcwd()
    os.chdir(tmp_path)
    try:
        processAXCTD.processAXCTD("g96_decim.wav", "out.txt", [0, -1], settings, engine=eng)
    finally:
        os.chdir(cwd)
    assert (tmp_path / "out.txt").read_text() == g.meta["output_text"]


def test_cli_mirror_reproduces_reference_crashes(eng, tmp_path):
    import synth
    from axctdprocessor_b200 import AXCTDprocessor, processAXCTD
    g = Golden("g44_nopulse")
    wav = tmp_path / "n.wav"
    synth.write_wav(str(wav), g.pcm(), g.spec.fs)
    settings = {"triggerrange": [30, -1], "minR400": 2.0, "mindR7500": 1.5, "deadfreq": 3000.0, "pointsperloop": 100000,
                "mark_space_freqs": [400.0, 800.0], "use_bandpass": False}
    with pytest.raises(KeyError, match="zcoeff_default"):          # processAXCTD.py:165 as shipped
        processAXCTD.processAXCTD(str(wav), str(tmp_path / "o.txt"), [0, -1], settings, engine=eng)
    assert (tmp_path / "o.txt").read_text().endswith("Conversion equations:\n")      # partial file left behind
    with pytest.raises(NameError):                                   # AXCTDprocessor.py:66 as shipped
        AXCTDprocessor.AXCTD_Processor(str(wav), timerange=[5, -1], engine=eng)
    # wired mode: defaults are printed, flags act
    processAXCTD.processAXCTD(str(wav), str(tmp_path / "w.txt"), [0, -1], settings, mode="wired", engine=eng)
    assert "(default)" in (tmp_path / "w.txt").read_text()


def test_wired_cli_equals_reference_driven_by_internal_keys(eng, tmp_path):
    import synth
    from axctdprocessor_b200 import processAXCTD
    g = Golden("g44_wired")
    wav = tmp_path / "w.wav"
    synth.write_wav(str(wav), g.pcm(), g.spec.fs)
    us = g.user_settings
    argv = ["-i", str(wav), "-o", str(tmp_path / "w.txt"), "--wired", "-p", str(us["minr400"]), "-t", str(us["mindr7500"]),
            "-d", str(us["deadfreq"]), "-l", str(int(us["refreshrate"] * g.spec.fs)), "-a", str(g.triggerrange[0])]
    import axctdprocessor_b200.AXCTDprocessor as A
    old = A._default_engines.get(0)
    A._default_engines[0] = eng
    try:
        processAXCTD.main(argv)
    finally:
        if old is None:
            A._default_engines.pop(0, None)
        else:
            A._default_engines[0] = old
    rows = [ln for ln in (tmp_path / "w.txt").read_text().splitlines() if ln[:8].strip().replace(".", "").isdigit() and "," in ln]
    assert len(rows) == min(g.meta["n_rows"], g.meta["n_hexframes"])
    assert [r.split(",")[1].strip() for r in rows] == g.hexframes[:len(rows)]


def test_multi_drop_recording_is_cut_and_decoded_per_drop(eng):
    """Three drops back to back in one recording: the segmentation driver finds them from the engine's
    400 Hz level and every segment decodes exactly as the oracle decodes that segment on its own."""
    import synth
    from axctdprocessor_b200 import segment
    from oracle import axctd_oracle as ao
    from parity_util import check_against_oracle
    specs = [synth.DropSpec(fs=44100, duration_s=52.0 + 3 * i, seed=400 + i, snr_db=30.0 - 8 * i) for i in range(3)]
    pcm = np.concatenate([synth.generate_drop(s) for s in specs])
    out = segment.process_recording(eng, pcm, 44100)
    assert len(out) == 3
    bounds = np.cumsum([0] + [int(round(s.duration_s * s.fs)) for s in specs])
    for i, (a, b, res) in enumerate(out):
        assert abs(a - bounds[i]) < 0.3 * 44100 or i == 0          # cut ~5 s before the pulse = start of the drop's lead-in
        assert res.status == 0
        seg = pcm[a:b]
        cfg = eng.config(44100)
        bt = eng.batch([len(seg)], [cfg])
        bt.upload(0, seg)
        bt.run()
        full = dict(result=bt.result(0), bits=bt.bits(0), edges=bt.edges(0), power=bt.power(0))
        bt.close()
        check_against_oracle(full, ao.process_pcm(seg, 44100))
        assert np.array_equal(res.rows["word"], full["result"].rows["word"])


@pytest.mark.parametrize("case", ["clipped", "too_short", "one_chunk"])
def test_edge_inputs_match_oracle(eng, case):
    """Ingest edge cases: a sample at -32768 (np.abs wraps, AXCTDprocessor.py:56), a recording shorter than
    4*N_power (run() never iterates, :295) and one that holds a single iteration."""
    import synth
    from oracle import axctd_oracle as ao
    from parity_util import check_against_oracle
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
    assert s.pcm_ampl == int(np.max(np.abs(pcm)))          # (wraps exactly like the reference's int16 abs)
    assert s.n_chunks == len(op.trace)
    if case == "clipped":
        check_against_oracle(out, op)
    else:
        assert s.status == 0 and s.n_bits == 0 and s.n_frames == 0 and s.firstpulse400 == -1


def test_random_batch_matches_oracle(eng):
    """Four drops of mixed rate / SNR / length decoded as one batch, each compared with the oracle."""
    import synth
    from oracle import axctd_oracle as ao
    from parity_util import check_against_oracle
    rng = np.random.default_rng(2025)
    specs = [synth.DropSpec(fs=int(rng.choice([44100, 48000])), duration_s=float(rng.uniform(45.0, 60.0)), seed=800 + i,
                            snr_db=float(rng.uniform(6.0, 40.0)), tone_after_pulse_s=float(rng.uniform(30.5, 36.0)))
             for i in range(4)]
    pcms = [synth.generate_drop(s) for s in specs]
    b = eng.batch([len(p) for p in pcms], [eng.config(s.fs) for s in specs])
    for i, p in enumerate(pcms):
        b.upload(i, p)
    b.run()
    outs = [dict(result=b.result(i), bits=b.bits(i), edges=b.edges(i), power=b.power(i)) for i in range(len(specs))]
    b.close()
    for s, p, out in zip(specs, pcms, outs):
        check_against_oracle(out, ao.process_pcm(p, s.fs))


def test_pipelined_decoder_returns_the_same_results(eng):
    """batch.PipelinedDecoder (two engines taking alternate batches) against a plain batch."""
    import synth
    from axctdprocessor_b200 import batch as axbatch
    specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=46.0, seed=900 + i, snr_db=20.0) for i in range(4)]
    pcms = [np.ascontiguousarray(synth.generate_drop(s)) for s in specs]
    ref = axbatch.process_drops(eng, pcms, [s.fs for s in specs])
    pipe = axbatch.PipelinedDecoder(slots=2, engine_factory=emu_engine)
    groups = [[0, 1], [2, 3]]
    pipe.submit([pcms[i].ctypes.data for i in groups[0]], [len(pcms[i]) for i in groups[0]], [specs[i].fs for i in groups[0]])
    pipe.submit([pcms[i].ctypes.data for i in groups[1]], [len(pcms[i]) for i in groups[1]], [specs[i].fs for i in groups[1]])
    got = pipe.collect() + pipe.collect()
    pipe.close()
    for r, g in zip(ref, got):
        assert r.status == 0 and g.status == 0
        assert np.array_equal(r.rows, g.rows)


def test_concurrent_decoder_returns_the_same_results(eng):
    """batch.ConcurrentDecoder (sub-batches on their own engines and host threads) against a plain batch."""
    import synth
    from axctdprocessor_b200 import batch as axbatch
    specs = [synth.DropSpec(fs=(44100, 48000)[i % 2], duration_s=46.0, seed=920 + i, snr_db=20.0) for i in range(5)]
    pcms = [np.ascontiguousarray(synth.generate_drop(s)) for s in specs]
    ref = axbatch.process_drops(eng, pcms, [s.fs for s in specs])
    for shards in (2, 3):
        cd = axbatch.ConcurrentDecoder(0, [len(p) for p in pcms], [s.fs for s in specs], shards=shards, engine_factory=emu_engine)
        assert sorted(i for part in cd.parts for i in part) == list(range(5))
        for i, p in enumerate(pcms):
            cd.upload(i, p)
        timings = cd.run(steps=2)
        assert len(timings) == shards and all(len(t) == 2 for t in timings)
        got = cd.results(full=False)
        cd.close()
        for r, g in zip(ref, got):
            assert r.status == 0 and g.status == 0
            assert np.array_equal(r.rows, g.rows)


def test_streaming_decoder_ends_at_the_batch_result(eng):
    """stream.StreamingDecoder: pushes + polls, then finish() equals the batch decode of the same samples."""
    import synth
    from axctdprocessor_b200 import batch as axbatch, stream as axstream
    spec = synth.DropSpec(fs=44100, duration_s=52.0, seed=930, snr_db=25.0)
    pcm = np.ascontiguousarray(synth.generate_drop(spec))
    ref = axbatch.process_drops(eng, [pcm], [spec.fs])[0]
    sd = axstream.StreamingDecoder(spec.fs, engine=eng, min_new_seconds=1.0)
    cuts = [0, int(3.0 * spec.fs), int(44.0 * spec.fs), int(44.5 * spec.fs), len(pcm)]
    polled = []
    for a, b in zip(cuts[:-1], cuts[1:]):
        sd.push(pcm[a:b])
        polled.append(sd.poll())
    assert polled[0] is None                       # 3 s: no pulse yet, nothing decodable
    assert polled[2] is None                       # less than min_new_seconds of new audio: no decode
    got = sd.finish()
    assert got.status == 0 and np.array_equal(got.rows, ref.rows)
    rows = [p for p in polled if p is not None]
    assert rows and all((np.diff(p["time_s"]) >= 0).all() for p in rows)
    assert sum(len(p) for p in rows) <= int((got.table()["keep"] == 1).sum()) + 8
    sd.stop()
    assert sd.push(pcm[:10]) == len(pcm)           # ignored after stop()


def test_values_outside_the_compact_row_range_fall_back_to_full_records(eng):
    """A header that announces an absurd depth slope: depth * 100 no longer fits the compact row's int32, the
    row is flagged AXCTD_ROW_WIDE and table() takes the values from the full frame records instead."""
    import synth
    from axctdprocessor_b200 import _lib, engine
    from oracle import axctd_oracle as ao
    spec = synth.DropSpec(fs=44100, duration_s=47.0, seed=950, snr_db=30.0, zcoeff=(0.72, 5.0e8, 0.0, 0.0))
    pcm = synth.generate_drop(spec)
    cfg = eng.config(spec.fs)
    b = eng.batch([len(pcm)], [cfg])
    b.upload(0, pcm)
    b.run()
    full, lean = b.result(0, full=True), b.result(0, full=False)
    b.close()
    assert (full.rows["flags"] & _lib.ROW_WIDE).any()
    tab = full.table()
    op = ao.process_pcm(pcm, spec.fs)
    kept = tab[tab["keep"] == 1]
    np.testing.assert_allclose(kept["depth"], np.asarray(op.depth), rtol=1e-6)
    assert kept["depth"].max() > 2.2e7
    with pytest.raises(RuntimeError):
        lean.table()


def test_calibration_known_answers_on_the_kernel_bodies(eng):
    """ax_sp_from_c / ax_dataconvert (the bodies k_calib runs) against the GSW documentation's check values for
    gsw_SP_from_C and against parse.dataconvert's summation order."""
    from test_oracle_units import GSW_C, GSW_P, GSW_SP, GSW_T
    from oracle import axctd_oracle as ao
    cf = [-0.0622192, 1.04584, 3.0e-5, -2.0e-7]
    sp, poly = eng.calib_eval(GSW_C, GSW_T, GSW_P, coeff=cf)
    np.testing.assert_allclose(sp, GSW_SP, rtol=1e-13, atol=0)
    np.testing.assert_array_equal(poly, [ao.dataconvert(c, cf) for c in GSW_C])


def test_guard_band_samples_are_settled_by_exact_recomputation():
    """Filter outputs inside the guard band no longer fail the drop: those inside demodulated iterations are
    re-signed with scipy-order arithmetic from the iteration start (ax_unc_resolve_item) and only a disagreement
    (or more hits than can be enumerated) raises AXCTD_DROP_UNCERTAIN; hits in the lead-in are ignored.  The
    guard is widened here so that the path runs at all (the product's 1e-12 is hit about once per 500 drops)."""
    g = Golden("g44_10db")
    e = emu_engine(guard=3e-6)
    out = run_engine(e, g.pcm(), g.spec.fs)
    s = out["result"].summary
    assert s.status == 0 and s.n_uncertain == 0
    assert s.n_guard_hits > 20 and 0 < s.n_guard_confirmed <= s.n_guard_hits
    check_against_golden(out, g)
    e.close()
    e = emu_engine(guard=1e-3)                      # far more hits than the list holds: cannot be enumerated
    out = run_engine(e, g.pcm(), g.spec.fs)
    assert out["result"].summary.status == 33 and out["result"].summary.n_guard_hits > 256
    e.close()
    q = Golden("g44_nopulse")                       # no pulse, nothing demodulated: guard hits cannot matter
    e = emu_engine(guard=1e-3)
    out = run_engine(e, q.pcm(), q.spec.fs)
    assert out["result"].summary.status == 0 and out["result"].summary.n_guard_hits > 0
    e.close()


def test_multichannel_frames_are_deinterleaved_by_the_engine(eng, tmp_path):
    """AXCTDprocessor.py:46-52: only the first channel of a multi-channel file is used.  The frames travel as read
    (axctd_batch_upload_interleaved) and the engine picks channel 0; checked for 2 and 3 channels (odd frame
    count: vector and tail paths) against the reference fixture, and through the AXCTD_Processor mirror."""
    import synth
    from axctdprocessor_b200 import AXCTDprocessor
    g = Golden("g44_stereo")
    frames = g.pcm()
    assert frames.ndim == 2 and frames.shape[1] == 2
    check_against_golden(run_engine(eng, frames, g.spec.fs), g)
    three = np.ascontiguousarray(np.concatenate([frames, frames[:, 1:2] // 2], axis=1))
    check_against_golden(run_engine(eng, three, g.spec.fs), g)
    b = eng.batch([len(frames) - 3], [eng.config(g.spec.fs)])
    b.upload(0, frames[:-3])
    assert np.array_equal(b.download(0), frames[:-3, 0])
    b.close()
    wav = tmp_path / "s.wav"
    synth.write_wav(str(wav), frames, g.spec.fs)
    ap = AXCTDprocessor.AXCTD_Processor(str(wav), engine=eng)
    assert ap.audiostream.ndim == 2
    ap.run()
    assert ap.hexframes == g.hexframes and ap.firstpulse400 == g.meta["firstpulse400"]
