"""Shared comparison of an engine result with a golden fixture or an oracle run."""
import numpy as np


CONF_RTOL = 5e-6


def mono(pcm):
    return np.ascontiguousarray(pcm[:, 0]) if pcm.ndim == 2 else pcm


def run_engine(eng, pcm, fs, settings=None, triggerrange=None):
    if fs > 50000:          # AXCTDprocessor.py:60-62: the engine halves the recording on the device
        cfg = eng.config(fs / 2, settings=settings, triggerrange=triggerrange, decimate=2)
    else:
        cfg = eng.config(fs, settings=settings, triggerrange=triggerrange)
    b = eng.batch([len(pcm)], [cfg])
    b.upload(0, pcm)
    b.run()
    out = dict(result=b.result(0), bits=b.bits(0), edges=b.edges(0), power=b.power(0), timing=b.timing())
    b.close()
    return out


def frames_view(res):
    fr = res.frames
    tab = res.table()                       # compact rows (integer hundredths) must reproduce the rounded doubles exactly
    assert len(tab) == len(fr)
    for k in tab.dtype.names:
        np.testing.assert_array_equal(tab[k], fr[k], err_msg=k)
    kept = fr[fr["keep"] == 1]
    hexr = ["%08x" % int(w) for w in fr["word"][fr["hex_returned"] == 1]]
    return kept, hexr


RAW_RTOL = 1e-6          # north_star: calibrated T / C / S / depth within 1e-6 relative, on the UNROUNDED values


def check_raw(fr, raw):
    """Every CRC-valid profile frame before rounding and QC (parse.py:92 / axctd_frame.*_raw): time, depth,
    temperature, conductivity and salinity at 1e-6 relative, signal levels at 1e-4 relative (north_star)."""
    if not raw:
        return
    assert len(fr) == len(raw["time"]), (len(fr), len(raw["time"]))
    for key, name in (("time_raw", "time"), ("depth_raw", "depth"), ("temperature_raw", "temperature"),
                      ("conductivity_raw", "conductivity"), ("salinity_raw", "salinity")):
        np.testing.assert_allclose(fr[key], raw[name], rtol=RAW_RTOL, atol=0, equal_nan=True, err_msg=key)
    for key, name in (("r400_raw", "r400"), ("r7500_raw", "r7500")):
        np.testing.assert_allclose(fr[key], raw[name], rtol=1e-4, atol=1e-9, equal_nan=True, err_msg=key)


def check_against_golden(out, g, level_rtol=1e-9):
    """Discrete outputs exact; calibrated values 1e-6 relative (north_star)."""
    res, m = out["result"], g.meta
    s = res.summary
    assert s.status == 0, (s.status, s.status_chunk)
    assert s.n_uncertain == 0
    assert s.firstpulse400 == m["firstpulse400"] and s.profstartind == m["profstartind"]
    assert s.numpoints == m["numpoints"]
    assert abs(s.high_bit_scale - m["high_bit_scale"]) <= 1e-12 * m["high_bit_scale"]
    bits, conf = out["bits"]
    edges = out["edges"][0]
    assert len(bits) == m["n_bits"] and np.array_equal(bits, g.bits), "bitstream"
    assert len(edges) == m["n_edges"] and np.array_equal(edges, g.edges), "bit edges"
    tr = g.trace()
    assert len(tr) == len(res.chunks)
    for k, (c, t) in enumerate(zip(res.chunks, tr)):
        got = (c["s"], c["e"], c["status"], c["n_power_total"], c["n_bits"], c["first_edge"], c["last_edge"], c["n_rows"], c["n_hex"])
        exp = (t["s"], t["e"], t["status"], t["n_power"], t["nbits"], t["first_edge"], t["last_edge"], t["nrows"], t["nhex"])
        assert got == exp, (k, got, exp)
        if "profstart" in t:
            assert c["profstartind"] == t["profstart"], (k, c["profstartind"], t["profstart"])
    kept, hexr = frames_view(res)
    assert hexr == g.hexframes, "hex frames"
    check_raw(res.frames, {k[4:]: g.z[k] for k in g.z.files if k.startswith("raw_")})
    for key, gk in (("time_s", "time"), ("depth", "depth"), ("temperature", "temperature"),
                    ("conductivity", "conductivity"), ("salinity", "salinity")):
        assert len(kept[key]) == len(g.z[gk]), gk
        np.testing.assert_allclose(kept[key], g.z[gk], rtol=1e-6, atol=0, equal_nan=True, err_msg=gk)
    for key, gk in (("r400", "r400_prof"), ("r7500", "r7500_prof")):
        np.testing.assert_allclose(kept[key], g.z[gk], rtol=1e-4, atol=0, equal_nan=True, err_msg=gk)
    p, r400, r7500 = out["power"]
    assert np.array_equal(p, g.z["power_inds"])
    np.testing.assert_allclose(r400, g.z["r400"].astype(np.float64), rtol=max(level_rtol, 1e-6 if g.z["r400"].dtype == np.float32 else 0), atol=1e-6 if g.z["r400"].dtype == np.float32 else 1e-11, equal_nan=True)
    np.testing.assert_allclose(r7500, g.z["r7500"].astype(np.float64), rtol=max(level_rtol, 1e-6 if g.z["r7500"].dtype == np.float32 else 0), atol=1e-6 if g.z["r7500"].dtype == np.float32 else 1e-11, equal_nan=True)
    if "conf" in g.z.files:
        # the mark / space windows are summed in fp32 (decisions near a boundary are re-made in double)
        # (fp32 error is relative to the stronger tone: conf = a2*s/a1 moves by ~1e-6 * (s + conf) * max(a1,a2)/a1)
        np.testing.assert_allclose(conf, g.z["conf"], rtol=CONF_RTOL, atol=CONF_RTOL, equal_nan=True)
        assert s.win32_max_rel_err < 2e-6, s.win32_max_rel_err
    # header metadata (exact)
    from axctdprocessor_b200.AXCTDprocessor import header_metadata
    for slot in range(2):
        key = f"frame_data_{slot + 2}"
        if key in m["metadata"]:
            assert s.header_parsed[slot]
            md = header_metadata(list(s.frame_data[slot]), list(s.counter_found[slot]))
            assert md["frame_data"] == m["metadata"][key]
            assert md["counter_found"] == m["metadata"][f"counter_found_{slot + 2}"]
        else:
            assert not s.header_parsed[slot]
    assert list(s.tcoeff_used) == [float(x) for x in m["tcoeff"]]
    assert list(s.ccoeff_used) == [float(x) for x in m["ccoeff"]]
    assert list(s.zcoeff_used) == [float(x) for x in m["zcoeff"]]


def check_against_oracle(out, op):
    res = out["result"]
    s = res.summary
    assert s.status == 0, (s.status, s.status_chunk)
    assert s.firstpulse400 == op.firstpulse400 and s.profstartind == op.profstartind
    assert abs(s.high_bit_scale - op.high_bit_scale) <= 1e-12 * op.high_bit_scale
    bits, conf = out["bits"]
    assert np.array_equal(bits, np.asarray(op.all_bits, dtype=np.uint8)), "bitstream"
    assert np.array_equal(out["edges"][0], np.asarray(op.all_edges, dtype=np.int64)), "bit edges"
    assert len(res.chunks) == len(op.trace)
    for c, t in zip(res.chunks, op.trace):
        assert (c["s"], c["e"], c["status"], c["n_rows"], c["n_hex"]) == (t["s"], t["e"], t["status"], t["nrows"], t["nhex"])
    kept, hexr = frames_view(res)
    assert hexr == op.hexframes
    if getattr(op, "all_raw", None) is not None and len(op.all_raw) == len(res.frames):
        a = np.asarray(op.all_raw, dtype=np.float64).reshape(-1, 7)
        check_raw(res.frames, dict(time=a[:, 0], depth=a[:, 1], temperature=a[:, 2], conductivity=a[:, 3],
                                   salinity=a[:, 4], r400=a[:, 5], r7500=a[:, 6]))
    else:
        assert len(res.frames) == 0 or not op.keep_trace
    for key, name in (("time_s", "time"), ("depth", "depth"), ("temperature", "temperature"),
                      ("conductivity", "conductivity"), ("salinity", "salinity")):
        np.testing.assert_allclose(kept[key], np.asarray(getattr(op, name), dtype=np.float64), rtol=1e-6, atol=0, equal_nan=True)
    for key, name in (("r400", "r400_prof"), ("r7500", "r7500_prof")):
        np.testing.assert_allclose(kept[key], np.asarray(getattr(op, name), dtype=np.float64), rtol=1e-4, atol=0, equal_nan=True)


def check_rows_against_oracle(res, op):
    """What the reference's API exposes per drop (compact rows + summary, no full frame records): exact hex words,
    row selection, detector indices and scale; rounded values equal to the oracle's."""
    s = res.summary
    assert s.status == 0, (s.status, s.status_chunk)
    assert s.firstpulse400 == op.firstpulse400 and s.profstartind == op.profstartind
    assert abs(s.high_bit_scale - op.high_bit_scale) <= 1e-12 * op.high_bit_scale
    assert s.n_bits == len(op.all_bits) and s.n_edges == len(op.all_edges)
    tab = res.table()
    assert ["%08x" % int(w) for w in tab["word"][tab["hex_returned"] == 1]] == op.hexframes
    kept = tab[tab["keep"] == 1]
    for key, name in (("time_s", "time"), ("depth", "depth"), ("temperature", "temperature"),
                      ("conductivity", "conductivity"), ("salinity", "salinity")):
        assert len(kept) == len(getattr(op, name))
        np.testing.assert_allclose(kept[key], np.asarray(getattr(op, name), dtype=np.float64), rtol=1e-6, atol=0, equal_nan=True)
    assert len(res.chunks) == len(op.trace)
    for c, t in zip(res.chunks, op.trace):
        assert (c["s"], c["e"], c["status"], c["n_rows"], c["n_hex"]) == (t["s"], t["e"], t["status"], t["nrows"], t["nhex"])


def oracle_prefix_normalised(pcm, fs, norm_seconds=2.0, settings=None, triggerrange=None):
    """The oracle over a recording normalised with the mean and the peak of its first norm_seconds (the streaming
    decoder's fixed normalisation) instead of the whole file's (AXCTDprocessor.py:55-57)."""
    from axctdprocessor_b200.stream import prefix_normalisation
    from oracle import axctd_oracle as ao
    dc, ampl = prefix_normalisation(pcm[:int(norm_seconds * fs)])
    x = (np.asarray(pcm, dtype=np.float64) - dc) / ampl
    return ao.OracleProcessor(x, fs, settings=settings, triggerrange=triggerrange).run(), (dc, ampl)


def check_streaming_against_oracle(eng, pcm, fs, seed=0, settings=None, triggerrange=None, piece_s=(0.3, 3.1), norm_seconds=2.0):
    """Feed the recording to StreamingDecoder in pieces of random length, poll after every piece and hold
    (a) every poll's rows to the oracle's per-iteration lists (AXCTDprocessor.py:612) of exactly the iterations that
    poll closed, (b) the finished result to the oracle as a whole.  Returns the decoder's per-run records."""
    from axctdprocessor_b200.stream import StreamingDecoder
    op, _ = oracle_prefix_normalised(pcm, fs, norm_seconds, settings, triggerrange)
    cum = np.concatenate([[0], np.cumsum([t["nrows"] for t in op.trace])]).astype(int)     # kept rows before iteration k
    sd = StreamingDecoder(fs, settings=settings, triggerrange=triggerrange, engine=eng, max_seconds=len(pcm) / fs + 5.0,
                          norm_seconds=norm_seconds)
    rng = np.random.default_rng(seed)
    pos, k_prev, polls_with_rows = 0, 0, 0
    try:
        while pos < len(pcm):
            n = int(rng.integers(int(piece_s[0] * fs), int(piece_s[1] * fs)))
            sd.push(pcm[pos:pos + n]); pos += n
            new = sd.poll()
            if sd.last is None:
                assert new is None
                continue
            k_now = int(sd.last.summary.n_chunks)
            assert k_prev <= k_now <= len(op.trace)
            # only complete iterations are decoded, and they are the oracle's
            for k in range(k_prev, k_now):
                c, t = sd.last.chunks[k], op.trace[k]
                assert (c["s"], c["e"], c["status"], c["n_rows"], c["n_hex"]) == (t["s"], t["e"], t["status"], t["nrows"], t["nhex"]), k
                assert c["e"] < sd.batch.summary(0).numpoints
            lo, hi = cum[k_prev], cum[k_now]
            assert (0 if new is None else len(new)) == hi - lo, (k_prev, k_now)
            if new is not None:
                polls_with_rows += 1
                for key, name in (("time_s", "time"), ("depth", "depth"), ("temperature", "temperature"),
                                  ("conductivity", "conductivity"), ("salinity", "salinity")):
                    np.testing.assert_allclose(new[key], np.asarray(getattr(op, name)[lo:hi], dtype=np.float64), rtol=1e-6, atol=0, equal_nan=True)
            k_prev = k_now
        r, rest = sd.finish_rows()
        assert len(rest) == cum[-1] - cum[k_prev]
        out = dict(result=r, bits=sd.batch.bits(0), edges=sd.batch.edges(0), power=sd.batch.power(0))
        check_against_oracle(out, op)
        assert polls_with_rows >= 3
        return sd.runs
    finally:
        eng_keep = sd.eng
        sd._own = False
        sd.close()
        assert eng_keep is eng
