"""Long recordings that hold several drops (BASELINE config 3).

The reference's state machine is one-way: it decodes the first drop's headers and
treats the rest of the file as that drop's profile (AXCTDprocessor.py:375-406).
This module is the driver the survey calls for: find where each drop starts, cut
the recording into one segment per drop (each with a quiet lead-in, like a
stand-alone recording) and decode all segments as one batch of independent
drops.  Every segment is decoded exactly as the reference decodes that segment
saved as its own WAV file.

Finding the drops uses the engine's own 400 Hz signal level (reference
AXCTDprocessor.py:355-371, the smoothed log10 ratio to the dead frequency) over
the whole recording ("scan_only": statistics + tone pass, no demodulation).  A
drop starts with three 400 Hz pulses 9.68 s apart, each after a quiet gap, and
its profile data (400/800 Hz FSK, level about minR400 - 0.5) keeps the level up
until the probe ends, so: a *rising edge* is a level >= minR400 right after
quiet_s seconds in which the level never reached minR400 / 2 (true noise sits
near 0), and a drop starts at a rising edge that has no other rising edge in the
preceding regroup_s seconds.
"""
from __future__ import annotations

import numpy as np

from . import engine as _engine


def scan_batch(eng: _engine.Engine, pcm, fs_raw: float, settings=None):
    """Upload the recording once and take its tone levels on the fixed chunk grid at the rate it was recorded at
    (no demodulation, no decimation: the 400 Hz level is the same signal at either rate).  Returns the batch -- which
    keeps the recording on the device for the segment decode -- and (power_inds, r400, r7500) in raw sample units.
    ``pcm``: int16 ndarray, mono or (n, channels) frames."""
    st = dict(settings or {})
    st["minr400"] = 1e300                      # never leaves status 0: every chunk stays on the fixed grid
    eng.set_option("scan_only", 1)
    try:
        cfg = eng.config(fs_raw, settings=st)
        b = eng.batch([len(pcm)], [cfg])
        try:
            b.upload(0, pcm)
            b.run()
            return b, b.power(0)
        except Exception:
            b.close()
            raise
    finally:
        eng.set_option("scan_only", 0)


def scan_levels(eng: _engine.Engine, pcm: np.ndarray, fs_raw: float, settings=None):
    """(power_inds, r400, r7500) over the whole recording at its own rate (see scan_batch)."""
    b, lv = scan_batch(eng, pcm, fs_raw, settings=settings)
    b.close()
    return lv


def find_drops(power_inds, r400, fs: float, n_total: int, min_r400: float = 2.0, quiet_s: float = 2.0,
               regroup_s: float = 15.0, lead_s: float = 5.0):
    """[(start, end)] sample ranges (effective rate), one per drop."""
    p = np.asarray(power_inds, dtype=np.int64)
    lvl = np.nan_to_num(np.asarray(r400, dtype=np.float64), nan=-np.inf)
    hot = lvl >= min_r400
    if len(p) == 0 or not hot.any():
        return []
    # a rising edge: the level leaves the quiet band (< minR400 / 2) after >= quiet_s in it and reaches minR400 within 1 s
    loud = lvl >= 0.5 * min_r400
    last_loud = np.maximum.accumulate(np.where(loud, p, -1))
    prev_loud = np.concatenate([[-1], last_loud[:-1]])
    onset = loud & (((prev_loud < 0) & (p >= quiet_s * fs)) | ((prev_loud >= 0) & (p - prev_loud >= quiet_s * fs)))
    rising = np.zeros(len(p), dtype=bool)
    for i in np.flatnonzero(onset):
        j = np.searchsorted(p, p[i] + fs, side="right")
        rising[i] = hot[i:j].any()
    edges = p[rising]
    if len(edges) == 0:
        return []
    starts = [int(edges[0])]
    for a, b in zip(edges[:-1], edges[1:]):
        if b - a > regroup_s * fs:
            starts.append(int(b))
    cuts = [max(0, s - int(lead_s * fs)) for s in starts]
    out = []
    for i, c in enumerate(cuts):
        end = cuts[i + 1] if i + 1 < len(cuts) else n_total
        if end > c:
            out.append((c, end))
    return out


def process_recording(eng: _engine.Engine, pcm: np.ndarray, fs: float, settings=None, triggerrange=None, decimate: int = 1,
                      **find_kw):
    """Decode every drop of a long recording.  ``fs`` is the effective rate (after any /2 decimation), ``pcm`` the raw
    samples.  Returns [(start_raw, end_raw, DropResult)] with sample ranges in the units of ``pcm``.

    The recording crosses the PCIe link once: the scan batch keeps it on the device and every segment is filled from
    there (axctd_batch_copy_from).  Each segment is then normalised, halved (recordings above 50 kHz) and decoded as
    the stand-alone recording it would be on disk, so its result is the reference's for that file."""
    pcm = np.ascontiguousarray(pcm, dtype=np.int16)
    fs_raw = fs * decimate
    scan, (p, r400, _) = scan_batch(eng, pcm, fs_raw, settings=settings)
    try:
        min_r400 = float((settings or {}).get("minr400", _engine.DEFAULT_SETTINGS["minr400"]))
        raw = find_drops(p, r400, fs_raw, len(pcm), min_r400=min_r400, **find_kw)
        if not raw:
            return []
        cfg = eng.config(fs, settings=settings, triggerrange=triggerrange, decimate=decimate)
        b = eng.batch([hi - lo for lo, hi in raw], [cfg] * len(raw))
        try:
            for i, (lo, hi) in enumerate(raw):
                b.copy_from(i, scan, 0, lo, hi - lo)
            b.run()
            res = [b.result(i) for i in range(len(raw))]
        finally:
            b.close()
    finally:
        scan.close()
    return [(lo, hi, r) for (lo, hi), r in zip(raw, res)]
